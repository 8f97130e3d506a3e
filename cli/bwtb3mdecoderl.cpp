// bwtb3mdecoderl [inputtype=bytestream] <in.bwt> ...  -- decode .bwt files to raw symbols on stdout
// (/root/reference/src/bwtb3mdecoderl.cpp:23-68): RLDecoder over all files, decodeRun() until
// sym < 0; symbols >= 256 are rejected for byte output.
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		if (arg.help || arg.rest.empty()) {
			std::cerr << "usage: " << arg.progname << " [inputtype=bytestream] <in.bwt> ..." << std::endl;
			return EXIT_FAILURE;
		}
		b3m::RlDecoder dec(arg.rest, 0, 1);
		std::vector<unsigned char> buf(1 << 16);
		while (true) {
			std::pair<int64_t, uint64_t> const r = dec.decodeRun();
			if (r.first < 0) break;
			if (r.first >= 256) throw std::runtime_error("symbol does not fit a byte");
			uint64_t left = r.second;
			size_t const fill = (size_t)std::min<uint64_t>(left, buf.size());
			memset(buf.data(), (int)r.first, fill);
			while (left) {
				size_t const w = (size_t)std::min<uint64_t>(left, buf.size());
				if (fwrite(buf.data(), 1, w, stdout) != w) throw std::runtime_error("write failed");
				left -= w;
			}
		}
		fflush(stdout);
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
