// sasubsample [-s <power of two>] < in.sa|in.isa > out   (/root/reference/src/sasubsample.cpp:23-61)
// Keeps every s-th value of a sampled SA / ISA file: native uint64 [rate][count][values] in,
// [rate*s][ceil(count/s)][values 0, s, 2s, ...] out.  Host only: no part of the hot path.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>
#include <stdexcept>
#include <vector>
#include <algorithm>

int main(int argc, char ** argv) {
	try {
		uint64_t sub = 1;
		for (int i = 1; i < argc; ++i) {
			if (!strcmp(argv[i], "-s") && i + 1 < argc) sub = strtoull(argv[++i], nullptr, 10);
			else if (!strncmp(argv[i], "-s", 2) && argv[i][2]) sub = strtoull(argv[i] + 2, nullptr, 10);
			else if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
				std::cerr << "usage: " << argv[0] << " [-s <power of two>] < in.sa > out.sa" << std::endl;
				return EXIT_FAILURE;
			}
		}
		if (!sub || (sub & (sub - 1))) throw std::runtime_error("sasubsample: the subsampling factor must be a power of two"); // sasubsample.cpp:30
		uint64_t hdr[2];
		if (fread(hdr, 8, 2, stdin) != 2) throw std::runtime_error("sasubsample: input too short for [rate][count]");
		uint64_t const inrate = hdr[0], incnt = hdr[1];
		uint64_t const out[2] = {inrate * sub, (incnt + sub - 1) / sub};
		if (fwrite(out, 8, 2, stdout) != 2) throw std::runtime_error("sasubsample: write failed");
		std::vector<uint64_t> buf(1 << 16), keep;
		keep.reserve(buf.size());
		uint64_t i = 0;
		while (i < incnt) {
			size_t const want = (size_t)std::min<uint64_t>(buf.size(), incnt - i);
			if (fread(buf.data(), 8, want, stdin) != want) throw std::runtime_error("sasubsample: input shorter than its count says");
			keep.clear();
			for (size_t k = 0; k < want; ++k, ++i) if (!(i & (sub - 1))) keep.push_back(buf[k]);
			if (!keep.empty() && fwrite(keep.data(), 8, keep.size(), stdout) != keep.size()) throw std::runtime_error("sasubsample: write failed");
		}
		fflush(stdout);
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
