// checkbwt [-i inputtype] [-V] <in.bwt> <text>   (/root/reference/src/checkbwt.cpp:248-279)
// Verifies a .bwt (+ <prefix>.preisa anchors) against the text by LF walk on the GPU and prints the
// reference's verdict line "[V] gok=<0|1>".  Like the reference it exits with EXIT_SUCCESS whenever the
// check could be carried out; use the printed verdict (or b3m_check_bwt's *ok) for the result.
#include "../include/b3m.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

int main(int argc, char ** argv) {
	try {
		std::string inputtype = "bytestream";
		std::vector<std::string> rest;
		int verbose = 1, device = 0;
		for (int i = 1; i < argc; ++i) {
			std::string const a = argv[i];
			if (a == "-i" && i + 1 < argc) inputtype = argv[++i];
			else if (a.rfind("-i", 0) == 0 && a.size() > 2) inputtype = a.substr(2);
			else if (a == "-V" && i + 1 < argc) ++i;               // progress interval of the reference: accepted, unused
			else if (a == "-d" && i + 1 < argc) device = atoi(argv[++i]);
			else if (a == "-q") verbose = 0;
			else if (a == "-h" || a == "--help") rest.clear(), i = argc;
			else rest.push_back(a);
		}
		if (rest.size() < 2) {
			std::cerr << "usage: " << argv[0] << " [-i bytestream|compactstream|pac|pacterm] [-d device] <in.bwt> <text>" << std::endl;
			return EXIT_FAILURE;
		}
		long const nc = sysconf(_SC_NPROCESSORS_ONLN);
		int ok = 0; uint64_t bad = 0; char err[2048] = "";
		if (b3m_check_bwt(rest[0].c_str(), rest[1].c_str(), inputtype.c_str(), nc > 0 ? (uint64_t)nc : 1, device, verbose, &ok, &bad, err, sizeof(err)) != 0)
			throw std::runtime_error(err);
		std::cerr << "[V] gok=" << ok << std::endl;
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
