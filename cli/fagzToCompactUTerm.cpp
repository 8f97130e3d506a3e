// fagzToCompactUTerm [rc=1] [gz=1] [outputfilename=output.compact] [inputfilenames=<list file>] [verbose=1] <in.fa[.gz]> ...
// (/root/reference/src/fagzToCompactUTerm.cpp:78-222): FASTA -> 3 bit per symbol compact container in which every
// sequence ends in its own terminator: A,C,G,T (any case) = 2..5, every other letter = 6, and behind sequence k
// (reverse complements count as sequences of their own) the number k written MSB first as `seqbits` symbols 0/1,
// seqbits = bits needed for the largest id.  Two passes over the input: the first counts the sequences.
// Host only: no part of the hot path.
#include "../bwtb3m_b200/csrc/formats.h"
#include "fastain.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <iostream>

static void put_seq_id(b3m::CompactWriter & w, uint64_t seqid, unsigned seqbits) {
	for (unsigned i = 0; i < seqbits; ++i) {
		uint8_t const v = (uint8_t)((seqid >> (seqbits - i - 1)) & 1);
		w.write(&v, 1);
	}
}

int main(int argc, char ** argv) {
	try {
		b3m::ArgInfo const arg(argc, argv);
		bool const rc = arg.getu("rc", 1) != 0;
		bool const gz = arg.getu("gz", 1) != 0;
		int const verbose = (int)arg.getu("verbose", 1);
		std::vector<std::string> const inputfilenames = b3mcli::input_names(arg.rest, arg.get("inputfilenames", ""));
		if (arg.help || inputfilenames.empty()) {
			std::cerr << "usage: " << arg.progname << " [rc=1] [gz=1] [outputfilename=output.compact] [inputfilenames=<file of names>] [verbose=1] <in.fa[.gz]> ..." << std::endl;
			return EXIT_FAILURE;
		}
		b3m::CompactWriter compactout(arg.get("outputfilename", "output.compact"), 3);
		if (!rc) std::cerr << "[V] not storing reverse complements" << std::endl;
		b3mcli::FastaRecord pat;
		uint64_t numseq = 0;
		for (size_t i = 0; i < inputfilenames.size(); ++i) {
			b3mcli::ByteSource src(inputfilenames[i], gz);
			b3mcli::FastaReader fain(src);
			while (fain.next(pat)) ++numseq;
		}
		if (rc) numseq *= 2;
		unsigned seqbits = 0;
		if (numseq) for (uint64_t v = numseq - 1; v; v >>= 1) ++seqbits; // libmaus2::math::numbits(numseq-1)
		std::cerr << "[V] numseq=" << numseq << std::endl;
		uint8_t ftable[256], ctable[7] = {6, 6, 5, 4, 3, 2, 6};
		memset(ftable, 6, sizeof(ftable));
		ftable['a'] = ftable['A'] = 2;
		ftable['c'] = ftable['C'] = 3;
		ftable['g'] = ftable['G'] = 4;
		ftable['t'] = ftable['T'] = 5;
		uint64_t seqid = 0;
		for (size_t i = 0; i < inputfilenames.size(); ++i) {
			std::string const & fn = inputfilenames[i];
			b3mcli::ByteSource src(fn, gz);
			b3mcli::FastaReader fain(src);
			while (fain.next(pat)) {
				if (verbose) std::cerr << (i + 1) << " " << b3mcli::strip_after_dot(b3mcli::basename_of(fn)) << " " << pat.sid << "...";
				std::string & s = pat.spattern;
				for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ftable[(uint8_t)s[j]];
				compactout.write((const uint8_t *)s.data(), s.size());
				put_seq_id(compactout, seqid++, seqbits);
				if (rc) {
					std::reverse(s.begin(), s.end());
					for (size_t j = 0; j < s.size(); ++j) s[j] = (char)ctable[(uint8_t)s[j]];
					compactout.write((const uint8_t *)s.data(), s.size());
					put_seq_id(compactout, seqid++, seqbits);
				}
				if (verbose) std::cerr << "done, input size " << b3mcli::format_bytes(s.size()) << std::endl;
			}
		}
		compactout.flush();
		if (seqid != numseq) throw std::runtime_error("the input changed between the two passes");
		return EXIT_SUCCESS;
	} catch (std::exception const & ex) {
		std::cerr << ex.what() << std::endl;
		return EXIT_FAILURE;
	}
}
