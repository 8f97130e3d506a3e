// bwtcomputessa <in.bwt> [sasamplingrate= isasamplingrate= tmpprefix= copyinputtomemory= threads=
// maxsortmem= maxtmpfiles= ref_isa= ref_sa=]   (/root/reference/src/bwtcomputessa.cpp:20-57)
#include "../include/b3m.h"
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdlib.h>
#include <unistd.h>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		if (arg.help || arg.rest.empty()) {
			std::cerr << "usage: " << arg.progname << " <in.bwt> [sasamplingrate=32] [isasamplingrate=32] [threads=#cpus] [ref_isa=] [ref_sa=] [device=0] [verbose=1]" << std::endl;
			return EXIT_FAILURE;
		}
		long const nc = sysconf(_SC_NPROCESSORS_ONLN);
		std::string const tmp = arg.get("tmpprefix", "");
		std::string const refisa = arg.get("ref_isa", ""), refsa = arg.get("ref_sa", "");
		char err[2048] = "";
		int const rc = b3m_compute_ssa(arg.rest[0].c_str(), arg.getu("sasamplingrate", 32), arg.getu("isasamplingrate", 32), tmp.c_str(),
		                               (int)arg.getu("copyinputtomemory", 0), arg.getu("threads", nc > 0 ? (uint64_t)nc : 1),
		                               arg.getu("maxsortmem", 2ull << 30), arg.getu("maxtmpfiles", 1024), (int)arg.getu("verbose", 1),
		                               refisa.c_str(), refsa.c_str(), (int)arg.getu("device", 0), err, sizeof(err));
		if (rc != 0) throw std::runtime_error(err);
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
