// fagzToCompact4 [rc=1] [gz=1] [outputfilename=<out.compact>] [inputfilenames=<list file>] [verbose=1] <in.fa[.gz]> ...
// (/root/reference/src/fagzToCompact4.cpp:80-264): FASTA -> 2 bit per base compact container for
// `bwtb3m inputtype=compactstream`, plus <out>.meta.  Every record is written forward and, with
// rc=1, followed by its reverse complement.  Runs of symbols other than ACGT (any case) are replaced
// by random bases; their [from,to) intervals go to the meta file:
//   big-endian uint64: #sequences, then per sequence: length, #replaced intervals, (from, to)...
// The replacement bases are not reproducible against the reference either (it draws them from
// libmaus2::random::Random::rand8, seeded per process); here they come from a fixed-seed splitmix64
// so that two runs over the same input give the same file.  Host only: no part of the hot path.
#include "../bwtb3m_b200/csrc/formats.h"
#include "fastain.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <iostream>
#include <sstream>

struct SplitMix64 {
	uint64_t s;
	uint64_t next() {
		uint64_t z = (s += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
};

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		bool const rc = arg.getu("rc", 1) != 0;
		bool const gz = arg.getu("gz", 1) != 0;
		int const verbose = (int)arg.getu("verbose", 1);
		std::vector<std::string> const inputfilenames = b3mcli::input_names(arg.rest, arg.get("inputfilenames", ""));
		if (arg.help || inputfilenames.empty()) {
			std::cerr << "usage: " << arg.progname << " [rc=1] [gz=1] [outputfilename=<prefix.compact>] [inputfilenames=<file of names>] [verbose=1] <in.fa[.gz]> ..." << std::endl;
			return EXIT_FAILURE;
		}
		std::string defout = b3mcli::common_prefix(inputfilenames);
		defout = b3mcli::clip_off(defout, ".gz");
		defout = b3mcli::clip_off(defout, ".fasta");
		defout = b3mcli::clip_off(defout, ".fa");
		std::string const outputfilename = arg.get("outputfilename", defout + ".compact");
		std::string const metaoutputfilename = outputfilename + ".meta";
		b3m::CompactWriter compactout(outputfilename, 2);
		if (!rc) std::cerr << "[V] not storing reverse complements" << std::endl;

		uint8_t ftable[256], ctable[4] = {3, 2, 1, 0};
		memset(ftable, 4, sizeof(ftable));
		ftable['a'] = ftable['A'] = 0;
		ftable['c'] = ftable['C'] = 1;
		ftable['g'] = ftable['G'] = 2;
		ftable['t'] = ftable['T'] = 3;

		std::vector<uint8_t> meta;
		b3m::put_be64(meta, 0); // #sequences, patched below
		uint64_t nseq = 0, insize = 0;
		SplitMix64 rng{0x6233746f43343a31ull};
		b3mcli::FastaRecord pat;
		for (size_t i = 0; i < inputfilenames.size(); ++i) {
			std::string const & fn = inputfilenames[i];
			b3mcli::ByteSource src(fn, gz);
			b3mcli::FastaReader fain(src);
			while (fain.next(pat)) {
				if (verbose) std::cerr << (i + 1) << " " << b3mcli::strip_after_dot(b3mcli::basename_of(fn)) << " " << pat.sid << "...";
				std::string & s = pat.spattern;
				b3m::put_be64(meta, s.size());
				size_t const nrpos = meta.size();
				b3m::put_be64(meta, 0);
				for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ftable[(uint8_t)s[j]];
				uint64_t nr = 0;
				size_t l = 0;
				while (l < s.size()) {
					while (l < s.size() && s[l] < 4) ++l;
					size_t h = l;
					while (h < s.size() && s[h] == 4) ++h;
					if (h > l) {
						for (size_t j = l; j < h; ++j) s[j] = (char)(rng.next() >> 62);
						b3m::put_be64(meta, l);
						b3m::put_be64(meta, h);
						++nr;
					}
					l = h;
				}
				for (int k = 0; k < 8; ++k) meta[nrpos + k] = (uint8_t)(nr >> (8 * (7 - k)));
				compactout.write((const uint8_t *)s.data(), s.size());
				if (rc) {
					std::reverse(s.begin(), s.end());
					for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ctable[(uint8_t)s[j]];
					compactout.write((const uint8_t *)s.data(), s.size());
				}
				insize += s.size() + 1;
				++nseq;
				if (verbose) std::cerr << "done, input size " << b3mcli::format_bytes(s.size() + 1) << " acc " << b3mcli::format_bytes(insize) << std::endl;
			}
		}
		for (int k = 0; k < 8; ++k) meta[k] = (uint8_t)(nseq >> (8 * (7 - k)));
		b3m::write_file(metaoutputfilename, meta.data(), meta.size());
		std::cerr << "Done, total input size " << insize << std::endl;
		compactout.flush();
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
