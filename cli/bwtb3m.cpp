// bwtb3m [key=value ...] <inputfile>  -- same command line as /root/reference/src/bwtb3m.cpp:25-72:
// ArgInfo key=value parsing, help on -h or without a positional argument, one library call,
// "[M] ... runtime ..." on stderr, what() + EXIT_FAILURE on error.
#include "../include/b3m.h"
#include "../bwtb3m_b200/csrc/formats.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <time.h>
#include <unistd.h>
#include <iostream>
#include <sstream>

static std::string default_tmp(std::string const & prog) {
	// libmaus2::util::ArgInfo::getDefaultTmpFileName: program name + host + pid + time
	char host[256] = "localhost";
	gethostname(host, sizeof(host) - 1);
	std::string base = prog;
	size_t const sl = base.find_last_of('/');
	if (sl != std::string::npos) base = base.substr(sl + 1);
	std::ostringstream o;
	o << base << "_" << host << "_" << getpid() << "_" << time(nullptr);
	return o.str();
}

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
		b3m_options o;
		b3m_options_init(&o);
		std::string const deftmp = default_tmp(arg.progname);
		if (arg.help || arg.rest.empty()) {
			std::ostringstream str;
			str << "This is bwtb3m (" << b3m_version() << ")\n\n";
			str << "usage: " << arg.progname << " [options] <inputfile>\n\n";
			str << "options:\n";
			str << "inputtype=[<bytestream>] (bytestream,compactstream,pac,pacterm)\n";
			str << "outputfilename=[<" << deftmp << ".bwt>] (name of output .bwt file)\n";
			str << "sasamplingrate=[" << o.sasamplingrate << "] sampling rate for sampled suffix array\n";
			str << "isasamplingrate=[" << o.isasamplingrate << "] sampling rate for sampled inverse suffix array\n";
			str << "mem=[0] memory target (0: free device memory; bounds the block size)\n";
			str << "numthreads=[" << o.numthreads << "] number of host threads\n";
			str << "bwtonly=[" << o.bwtonly << "] compute BWT only (no sampled suffix array and reverse)\n";
			str << "tmpprefix=[" << deftmp << "] (prefix for tmp files)\n";
			str << "sparsetmpprefix=[tmpprefix] (accepted; gap arrays live in device memory)\n";
			str << "copyinputtomemory=[0] (accepted; the input is always staged in device memory)\n";
			str << "largelcpthres=[" << o.largelcpthres << "] (large LCP value threshold)\n";
			str << "verbose=[" << o.verbose << "] (verbosity level)\n";
			str << "device=[0] (CUDA device)\n";
			str << "numblocks=[0] (0: derive from mem; >0: force the number of blocks)\n";
			str << "ngpus=[1] (GPUs of this box sharing the build: devices device .. device+ngpus-1)\n";
			str << "\nfile formats: .sa / .isa / .preisa are the reference's native uint64 layouts; the .bwt / .hist containers are this\n";
			str << "library's own (same content, different bytes than libmaus2's RLEncoder / NumberMapSerialisation): read them with\n";
			str << "the tools built from this tree (bwtb3mdecoderl, bwtb3mtobwa, bwtcomputessa, checkbwt), see README.md\n";
			throw std::runtime_error(str.str());
		}
		std::string const fn = arg.rest[0];
		std::string const inputtype = arg.get("inputtype", "bytestream");
		std::string const tmpprefix = arg.get("tmpprefix", deftmp);
		std::string const outfn = arg.get("outputfilename", tmpprefix + ".bwt");
		std::string const sparsetmp = arg.get("sparsetmpprefix", tmpprefix);
		o.fn = fn.c_str();
		o.inputtype = inputtype.c_str();
		o.outputfilename = outfn.c_str();
		o.tmpprefix = tmpprefix.c_str();
		o.sparsetmpprefix = sparsetmp.c_str();
		o.sasamplingrate = arg.getu("sasamplingrate", o.sasamplingrate);
		o.isasamplingrate = arg.getu("isasamplingrate", o.isasamplingrate);
		o.mem = arg.getu("mem", o.mem);
		o.numthreads = arg.getu("numthreads", o.numthreads);
		o.bwtonly = (int)arg.getu("bwtonly", 0);
		o.copyinputtomemory = (int)arg.getu("copyinputtomemory", 0);
		o.largelcpthres = arg.getu("largelcpthres", o.largelcpthres);
		o.verbose = (int)arg.getu("verbose", 0);
		o.device = (int)arg.getu("device", 0);
		o.numblocks = arg.getu("numblocks", 0);
		o.ngpus = (int)arg.getu("ngpus", 1);
		b3m_result res;
		char err[2048] = "";
		if (b3m_compute_bwt(&o, &res, err, sizeof(err)) != 0) throw std::runtime_error(err);
		struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
		struct rusage ru; getrusage(RUSAGE_SELF, &ru);
		double const sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
		std::cerr << "[M] MemUsage(maxrss=" << ru.ru_maxrss / 1024 << "MiB) runtime " << sec << "s (device " << res.seconds_device << "s, n=" << res.n
		          << ", blocks=" << res.numblocks << ")" << std::endl;
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
