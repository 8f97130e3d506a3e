// fagzToCompact [rc=1] [gz=1] [limit=<bases>] [outputfilename=output.compact] [inputfilenames=<list file>] [verbose=1] <in.fa[.gz]> ...
// (/root/reference/src/fagzToCompact.cpp:76-183): FASTA -> 3 bit per symbol compact container for
// `bwtb3m inputtype=compactstream`: A,C,G,T (any case) = 1..4, every other letter = 5, and a symbol 0 behind
// every sequence; with rc=1 each record is followed by its reverse complement (5 stays 5) and another 0.
// Files are read while the accumulated input size is below limit=.  Host only: no part of the hot path.
#include "../bwtb3m_b200/csrc/formats.h"
#include "fastain.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <iostream>

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		bool const rc = arg.getu("rc", 1) != 0;
		bool const gz = arg.getu("gz", 1) != 0;
		uint64_t const limit = arg.getu("limit", ~0ull);
		int const verbose = (int)arg.getu("verbose", 1);
		std::vector<std::string> const inputfilenames = b3mcli::input_names(arg.rest, arg.get("inputfilenames", ""));
		if (arg.help || inputfilenames.empty()) {
			std::cerr << "usage: " << arg.progname << " [rc=1] [gz=1] [limit=<bases>] [outputfilename=output.compact] [inputfilenames=<file of names>] [verbose=1] <in.fa[.gz]> ..." << std::endl;
			return EXIT_FAILURE;
		}
		b3m::CompactWriter compactout(arg.get("outputfilename", "output.compact"), 3);
		if (!rc) std::cerr << "[V] not storing reverse complements" << std::endl;
		uint8_t ftable[256], ctable[6] = {5, 4, 3, 2, 1, 5}; // complement of the mapped symbols: A<->T, C<->G
		memset(ftable, 5, sizeof(ftable));
		ftable['a'] = ftable['A'] = 1;
		ftable['c'] = ftable['C'] = 2;
		ftable['g'] = ftable['G'] = 3;
		ftable['t'] = ftable['T'] = 4;
		uint8_t const zero = 0;
		uint64_t insize = 0;
		b3mcli::FastaRecord pat;
		for (size_t i = 0; i < inputfilenames.size() && insize < limit; ++i) {
			std::string const & fn = inputfilenames[i];
			b3mcli::ByteSource src(fn, gz);
			b3mcli::FastaReader fain(src);
			while (fain.next(pat)) {
				if (verbose) std::cerr << (i + 1) << " " << b3mcli::strip_after_dot(b3mcli::basename_of(fn)) << " " << pat.sid << "...";
				std::string & s = pat.spattern;
				for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ftable[(uint8_t)s[j]];
				compactout.write((const uint8_t *)s.data(), s.size());
				compactout.write(&zero, 1);
				if (rc) {
					std::reverse(s.begin(), s.end());
					for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ctable[(uint8_t)s[j]];
					compactout.write((const uint8_t *)s.data(), s.size());
					compactout.write(&zero, 1);
				}
				insize += s.size() + 1;
				if (verbose) std::cerr << "done, input size " << b3mcli::format_bytes(s.size() + 1) << " acc " << b3mcli::format_bytes(insize) << std::endl;
			}
		}
		std::cerr << "Done, total input size " << insize << std::endl;
		compactout.flush();
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
