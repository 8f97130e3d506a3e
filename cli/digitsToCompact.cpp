// digitsToCompact [gz=0] [term=0] [outputfilename=output.compact] < digits
// (/root/reference/src/digitsToCompact.cpp:26-92): decimal digits on stdin -> 4 bit per symbol
// compact container.  term=0: digit d -> symbol d; term=1: digit d -> symbol d+1 and one symbol 0
// appended.  Any other input byte is an error.  Host only: no part of the hot path.
#include "../bwtb3m_b200/csrc/formats.h"
#include "fastain.h"
#include <stdlib.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		if (arg.help) {
			std::cerr << "usage: " << arg.progname << " [options] < input" << std::endl << std::endl;
			std::cerr << "options:" << std::endl;
			std::cerr << "gz=[0|1] (input file is plain (0)/gzip compressed (1), default is uncompressed)" << std::endl;
			std::cerr << "outputfilename=<output.compact> (name of output file, default is output.compact)" << std::endl;
			std::cerr << "term=[0|1] (0: map digits to symbols 0-9, 1: map digits to symbols 1-10 and append 0 at the end)" << std::endl;
			return EXIT_SUCCESS;
		}
		bool const gz = arg.getu("gz", 0) != 0;
		bool const addterm = arg.getu("term", 0) != 0;
		b3m::CompactWriter compactout(arg.get("outputfilename", "output.compact"), 4);
		b3mcli::ByteSource in(std::string(), gz);
		std::vector<uint8_t> B(8 * 1024);
		uint8_t const termadd = addterm ? 1 : 0;
		size_t num;
		while ((num = in.read(B.data(), B.size())) != 0) {
			bool err = false;
			for (size_t i = 0; i < num; ++i) {
				err |= B[i] < '0' || B[i] > '9';
				B[i] = (uint8_t)(B[i] - '0' + termadd);
			}
			if (err) throw std::runtime_error("Input file contains non decimal digit symbols.");
			compactout.write(B.data(), num);
		}
		if (addterm) { uint8_t const c = 0; compactout.write(&c, 1); }
		compactout.flush();
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
