// Host-side file formats of the bwtb3m surface: .hist, .preisa(+.meta), .sa, .isa,
// the run-length Huffman .bwt container, and the key=value argument convention of the CLIs.
//
// Pinned by the reference's own readers/writers:
//   .sa/.isa   native uint64 [rate][count][values...]  (/root/reference/src/sasubsample.cpp:34-58)
//   .preisa    native uint64 (rank,pos) pairs          (/root/reference/src/hwtPreIsaToIsa.cpp:55-77)
//   .preisa.meta  one big-endian number = rate         (/root/reference/src/hwtPreIsaToIsa.cpp:41,45-51)
// Parity unpinned (libmaus2 internals, no golden bytes in the reference):
//   .compact   CompactArrayWriterFile container: four numbers + 64-bit words, symbols MSB first (either byte order is read)
//   .hist      NumberMapSerialisation: big-endian uint64 count, then (symbol,count) pairs
//   .bwt       run-length Huffman container: only the decoded run/symbol sequence is pinned
//              (/root/reference/src/bwtb3mdecoderl.cpp:27-34); the byte layout here is this
//              implementation's own ("RL container").
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace b3m {

struct IoError : std::runtime_error {
	explicit IoError(std::string const & s) : std::runtime_error(s) {}
};

// ---- plain numbers -----------------------------------------------------------------------
void put_be64(std::vector<uint8_t> & o, uint64_t v);
uint64_t get_be64(const uint8_t * p);
std::vector<uint8_t> read_file(std::string const & fn);
void write_file(std::string const & fn, const void * data, size_t bytes);
uint64_t file_size(std::string const & fn);
bool file_exists(std::string const & fn);

// ---- .hist -------------------------------------------------------------------------------
void write_hist(std::string const & fn, const uint64_t hist[256]);
std::map<int64_t, uint64_t> read_hist(std::string const & fn);

// ---- .sa / .isa --------------------------------------------------------------------------
void write_sampled(std::string const & fn, uint64_t rate, const uint64_t * v, uint64_t count);
void read_sampled(std::string const & fn, uint64_t * rate, std::vector<uint64_t> * v);

// ---- .preisa -----------------------------------------------------------------------------
void write_preisa(std::string const & fn, const uint64_t * pairs, uint64_t npairs, uint64_t rate);
std::vector<uint64_t> read_preisa(std::string const & fn); // flat (rank,pos,rank,pos,...)

// ---- canonical Huffman -------------------------------------------------------------------
struct HuffCode {
	// index = symbol; len 0 = unused
	std::vector<uint8_t> len;
	std::vector<uint32_t> code;
};
constexpr unsigned RL_MAXCODELEN = 24;
// code lengths limited to RL_MAXCODELEN; canonical codes assigned by (length, symbol)
HuffCode huff_build(const uint64_t * freq, size_t nsyms);
void huff_assign(HuffCode & h);

// ---- run-length Huffman .bwt container ---------------------------------------------------
// Layout (all numbers big-endian, bit stream MSB first):
//   "B3MRL01\0"  n  nruns  runs_per_block  nblocks
//   symbol table : u16 count, then count x (u16 symbol, u8 code length)
//   length table : u16 count, then count x (u16 run-length value, u8 code length); value 0 is
//                  the escape code: it is followed by a 6-bit field b-1 and the run length in b bits
//   payload      : nblocks blocks, each starting on a 64-bit boundary relative to the payload
//                  start; a block holds runs_per_block runs (the last block fewer), each run =
//                  symbol code, length code [, escape payload]
//   index        : nblocks x (u64 payload word offset, u64 symbols before the block)
//   last 8 bytes : file offset of the index
constexpr uint32_t RL_LENBINS = 256;   // run lengths 1..255 have their own code, longer ones escape
constexpr uint32_t RL_RUNS_PER_BLOCK = 4096;

struct RlHeader {
	uint64_t n = 0, nruns = 0, runs_per_block = RL_RUNS_PER_BLOCK, nblocks = 0;
	HuffCode sym, len;
};

// writer used by the device encoder (K8) and the host encoder: header, then the already encoded
// payload in any number of pieces, then the block index
class RlContainerWriter {
public:
	RlContainerWriter(std::string const & fn, RlHeader const & h);
	~RlContainerWriter();
	void payload(const void * bytes, size_t nbytes);
	void finish(const uint64_t * block_word_off, const uint64_t * block_sym_off);
private:
	struct Impl;
	std::unique_ptr<Impl> impl;
};

// host encoder (tools and tests; same bytes as the device encoder)
void rl_encode_host(std::string const & fn, const uint8_t * syms, uint64_t n);

// reader: libmaus2::huffman::RLDecoder's surface as used by the reference's tools
class RlDecoder {
public:
	// several files addressed as one sequence; offset = symbol index to start at
	RlDecoder(std::vector<std::string> const & files, uint64_t offset = 0, uint64_t numthreads = 1);
	~RlDecoder();
	// next run: (symbol, length); symbol < 0 at the end of the data
	std::pair<int64_t, uint64_t> decodeRun();
	// next symbol, -1 at the end
	int64_t decode();
	static uint64_t getLength(std::vector<std::string> const & files, uint64_t numthreads = 1);
	static uint64_t getLength(std::string const & file, uint64_t numthreads = 1) { return getLength(std::vector<std::string>(1, file), numthreads); }
	// whole sequence into memory, blocks decoded by numthreads threads
	static std::vector<uint8_t> decodeAll(std::vector<std::string> const & files, uint64_t numthreads);
	// RLDecoder::getBlockSymHistograms(bwt, out, tmp, minsym, maxsym, numthreads, log)
	// (/root/reference/src/bwtdecodeblock.cpp:216-223,356-365): for every block of the run-length container the number of
	// occurrences of each symbol minsym..maxsym in the blocks BEFORE it, (maxsym - minsym + 1) big-endian uint64 per block
	// -- the `.sparserank` file that SparseRank::rankm seeks into.  Returns the number of blocks.
	static uint64_t getBlockSymHistograms(std::string const & bwtfn, std::string const & outfn, int64_t minsym, int64_t maxsym, uint64_t numthreads);
	// rank_sym(L, i) = occurrences of sym in L[0..i): one seek into the .sparserank file plus the runs of one block
	// (what SparseRank::rankm does, /root/reference/src/bwtdecodeblock.cpp:210-242)
	static uint64_t rankm(std::string const & bwtfn, std::string const & sparserankfn, int64_t minsym, int64_t maxsym, int64_t sym, uint64_t i);
private:
	struct Impl;
	std::unique_ptr<Impl> impl;
};

// ---- compactstream container ----------------------------------------------------------------
// What libmaus2::bitio::CompactArrayWriterFile writes and CompactDecoderWrapper reads
// (/root/reference/src/fagzToCompact4.cpp:105,232,265; /root/reference/src/decodecompact.cpp:30-36).
// Layout [parity unpinned, SURVEY 8c]: the serialised CompactArray (/root/reference/src/digitsToCompact.cpp:122 reads
// the writer's file back as one): four uint64 (bits per symbol, number of symbols, number of 64-bit words, the word
// count again as the array header), then the words; symbols MSB first inside a word, the last word zero padded.
// Byte order: the writer emits native little-endian numbers and words (libmaus2's Serialize<uint64_t>, the facility
// behind .sa/.isa, /root/reference/src/sasubsample.cpp:34-58); the reader and K1 (`k_unpack_compact`) also accept
// big-endian numbers with a big-endian bit stream -- bits per symbol is 1..8 in exactly one of the two readings.
class CompactWriter {
public:
	CompactWriter(std::string const & fn, unsigned bits);
	~CompactWriter();
	// values must be < 2^bits
	void write(const uint8_t * syms, size_t n);
	// pads the last word, patches the counts into the header and closes the file
	void flush();
	uint64_t size() const;
private:
	struct Impl;
	std::unique_ptr<Impl> impl;
};

class CompactReader {
public:
	explicit CompactReader(std::string const & fn);
	~CompactReader();
	uint64_t size() const;
	unsigned bits() const;
	// next (at most n) symbols, one per byte; returns how many were delivered (0 at the end)
	size_t read(uint8_t * out, size_t n);
private:
	struct Impl;
	std::unique_ptr<Impl> impl;
};

// ---- key=value arguments (libmaus2::util::ArgInfo convention, /root/reference/src/bwtb3m.cpp:29) --
struct ArgInfo {
	std::string progname;
	std::map<std::string, std::string> kv;
	std::vector<std::string> rest;
	bool help = false;
	ArgInfo(int argc, char ** argv);
	bool has(std::string const & k) const { return kv.count(k) != 0; }
	std::string get(std::string const & k, std::string const & def) const;
	// numbers accept k/m/g/t suffixes (powers of 1024), README.md:42
	uint64_t getu(std::string const & k, uint64_t def) const;
	static uint64_t parse_unit_number(std::string const & s);
};

} // namespace b3m
