// bwttestdecodespeed <in.bwt> [chains=...] [steps=...]   (/root/reference/src/bwttestdecodespeed.cpp:27-129)
// LF-steps/s instrument: dependent LF chains from evenly spaced <prefix>.isa samples.  The reference
// interleaves 1..8 chains on one CPU thread; a GPU needs tens of thousands in flight, so the table runs
// over chain counts 2^10 .. 2^22 (or the one given).  Output lines as in the reference:
// "<chains>\t<seconds per chain> <steps/s>".
#include "../include/b3m.h"
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdlib.h>
#include <unistd.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		if (arg.help || arg.rest.empty()) {
			std::cerr << "usage: " << arg.progname << " <in.bwt> [chains=<n>] [steps=<per chain>] [device=0]" << std::endl;
			return EXIT_FAILURE;
		}
		long const nc = sysconf(_SC_NPROCESSORS_ONLN);
		uint64_t const one = arg.getu("chains", 0), steps = arg.getu("steps", 0);
		for (uint64_t chains = one ? one : 1024; chains <= (one ? one : (1ull << 22)); chains <<= 2) {
			double sps = 0, sec = 0; char err[2048] = "";
			if (b3m_lf_speed(arg.rest[0].c_str(), chains, steps, nc > 0 ? (uint64_t)nc : 1, (int)arg.getu("device", 0), &sps, &sec, err, sizeof(err)) != 0)
				throw std::runtime_error(err);
			std::cout << chains << "\t" << sec / (double)chains << " " << sps << std::endl;
		}
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
