// decodecompact <in.compact> ...  -- symbols of compact containers as raw bytes on stdout
// (/root/reference/src/decodecompact.cpp:21-44: CompactDecoderWrapper over each argument, 8 KiB
// reads copied to std::cout).  Host only: no part of the hot path.
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdio.h>
#include <stdlib.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		std::vector<uint8_t> B(8 * 1024);
		for (int i = 1; i < argc; ++i) {
			b3m::CompactReader cdw(argv[i]);
			size_t got;
			while ((got = cdw.read(B.data(), B.size())) != 0)
				if (fwrite(B.data(), 1, got, stdout) != got) throw std::runtime_error("write failed");
		}
		fflush(stdout);
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
