// bwtb3mtobwa <in.bwt> <out.bwt> <out.sa>  (/root/reference/src/bwtb3mtobwa.cpp:23-58)
#include "../include/b3m.h"
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdlib.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		if (arg.help || arg.rest.size() < 3) {
			std::cerr << "usage: " << arg.progname << " <in.bwt> <out.bwt> <out.sa>" << std::endl;
			return EXIT_FAILURE;
		}
		char err[2048] = "";
		if (b3m_to_bwa(arg.rest[0].c_str(), arg.rest[1].c_str(), arg.rest[2].c_str(), err, sizeof(err)) != 0) throw std::runtime_error(err);
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
