// Internal (C++) interface between the engine and the CUDA stages K1..K7.
// Replaces, stage by stage, the libmaus2 code reached from
// /root/reference/src/bwtb3m.cpp:62-63 (BwtMergeSort::computeBwt); see DESIGN.md.
#pragma once
#include "common.cuh"

namespace b3m {

// ---- device text -------------------------------------------------------------------------
// Dense symbol codes 0..sigma-1, one byte per symbol.  For pacterm the unique terminator is
// implicit: it sits at position ntext (n = ntext+1) and is smaller than every code.
struct DevText {
	const uint8_t * codes = nullptr;
	const uint64_t * packed = nullptr; // sigma <= 4: 2 bit/symbol, 32 per word, first symbol in the top bits (textview.cuh)
	uint64_t ntext = 0;     // stored symbols
	uint64_t n = 0;         // BWT length (ntext, or ntext+1 with implicit terminator)
	uint32_t sigma = 0;     // number of distinct codes
	int has_term = 0;       // pacterm
	unsigned keybits = 8;   // bits per code inside sort keys (2, 4 or 8)
};

// what the block-level kernels need to read the circular text
struct TextRef {
	const uint8_t * codes;
	uint64_t ntext, n;
	int has_term;
	const uint64_t * packed; // 2-bit packed copy (textview.cuh) or nullptr
};

struct CTab { uint32_t c[257]; }; // C[code] passed to kernels by value

// ---- K1 ---------------------------------------------------------------------------------
void k1_hist_bytes(Stream & st, const uint8_t * d_in, uint64_t nbytes, uint64_t * d_hist256);
void k1_map_bytes(Stream & st, const uint8_t * d_in, uint64_t n, const uint8_t * d_lut256, uint8_t * d_out);
void k1_unpack_compact(Stream & st, const uint8_t * d_words, uint64_t n, unsigned b, bool le_words, uint8_t * d_out, uint64_t * d_hist256);
// packed copy of n codes < 4; d_out holds n/32 + 3 words, the tail is zero
void k1_pack2(Stream & st, const uint8_t * d_codes, uint64_t n, uint64_t * d_out);
// pac payload (l bases, BWA layout) -> packed text (l/32 + 3 words, zero tail), histogram of the four codes, code of the last base
void k1_pac_to_packed(Stream & st, const uint8_t * d_pac, uint64_t l, uint64_t * d_out, uint64_t * d_hist256, uint8_t * d_lastcode);
// packed text -> one byte per symbol (n codes + 16 zero bytes)
void k1_unpack_packed(Stream & st, const uint64_t * d_packed, uint64_t n, uint8_t * d_out);

// ---- K2 ---------------------------------------------------------------------------------
struct SortStats {
	uint64_t rounds = 0;          // prefix doubling rounds including round 0
	uint64_t radix_passes = 0;
	uint64_t radix_bytes = 0;     // algorithmic bytes of all radix passes
	uint64_t active_sum = 0;      // sum over rounds of records entering the round
	uint64_t other_bytes = 0;     // key extraction, flagging, rank scatter, compaction
	uint64_t tied0 = 0;           // suffixes sharing their first k0 symbols with another one (after round 0)
	uint64_t unresolved0 = 0;     // suffixes still tied after the in-CTA group sort (enter prefix doubling)
};

// What a whole-text sort (wstart == 0, W == ntext) can emit straight from the sorted order, fused
// with the last sorting step (K3 + the sampling the reference does by LF walk): the BWT, the
// (rank,pos) anchors and the sampled SA / ISA.  rank of window suffix k = k + shift.
struct FusedOut {
	uint8_t * bwt = nullptr;       // bwt[k + shift] = code preceding suffix sa[k]
	uint64_t shift = 0;
	int has_term = 0;              // the suffix at text position 0 is preceded by the terminator: code 0, row -> special[0]
	uint32_t * special = nullptr;  // [0] row of the terminator symbol, [1] row of the suffix at position 0
	uint32_t * prerank = nullptr;  // prerank[p >> prelog] = rank of position p, p multiple of 2^prelog
	uint32_t prelog = 0;
	unsigned long long * sa_s = nullptr;  // sa_s[r >> salog] = position, r multiple of 2^salog (nullptr: bwtonly)
	unsigned long long * sa_s2 = nullptr; // optional second copy (a sharded build: sa_s is another GPU's memory, sa_s2 this GPU's, from where the samples are sent to the host)
	uint32_t salog = 0;
	unsigned long long * isa_s = nullptr; // isa_s[p >> isalog] = rank
	uint32_t isalog = 0;
};

// Optional early delivery: a pinned host buffer that receives the rank-sampled SA in chunks while
// the last sorting step is still running (PCIe D2H overlaps the kernel).
struct StreamOut {
	unsigned long long * host_sa = nullptr; // nsa values
	uint64_t nsa = 0;
	bool delivered = false;                 // set when host_sa holds every value but [0] of a terminated text
	// BWA's packed BWT (terminated texts): needs the row of the suffix at position 0 (primary) before the
	// rows behind it can be packed, so the sort resolves that suffix's tile first
	uint32_t * host_bwa = nullptr;          // ceil((n-1)/16) words
	uint32_t * d_bwa = nullptr;             // device staging buffer of the same size
	bool bwa_delivered = false;
};
void k9_pack_bwa_range(cudaStream_t s, const uint8_t * bwt, uint64_t seq_len, uint64_t primary, uint32_t * words, uint64_t w_lo, uint64_t w_hi);

// Sorts the W suffixes that start at text positions wstart+i, 0 <= i < W.
// circular != 0: W == ntext, wstart == 0, indices wrap (terminator-free whole text).
// circular == 0: the end of the window is a sentinel smaller than every symbol; text positions
//                wrap modulo ntext when text_wraps != 0.
// sa (allocated here, W entries): sa[k] = window-relative start of the k-th smallest suffix;
// rank (caller's, W entries, may be nullptr): the inverse permutation.
// fo (may be nullptr; whole-text windows only): fused outputs, see FusedOut.
void k2_suffix_sort(Stream & st, DevText const & T, uint64_t wstart, uint64_t W, int circular, int text_wraps,
                    DevBuf<uint32_t> & sa, uint32_t * rank, SortStats * stats, const FusedOut * fo, StreamOut * so = nullptr);

// Suffix-range sharding (multi-GPU, sufsort.cu): part p holds the suffixes whose first-key bin lies
// in [bin_lo[p], bin_lo[p+1]); base[p] = number of suffixes of the window in smaller bins.
struct KeyRangePlan {
	uint32_t nparts = 0;
	uint32_t binshift = 20;       // bin of a suffix = its first key >> binshift
	std::vector<uint32_t> bin_lo;
	std::vector<uint64_t> base;
	// 2-bit alphabets: the bins are the level-1 buckets of the MSD path (msd.cuh) and hist their sizes
	unsigned msd_b1 = 0, msd_b2 = 0;
	std::vector<unsigned long long> hist;
};
void k2_keyrange_plan(Stream & st, DevText const & T, int circular, uint32_t nparts, KeyRangePlan & plan);
// sorts the suffixes of one part of the whole text and emits their fused outputs at global ranks
// fo.shift + base[part] + k; returns the number of suffixes left unresolved (0: the slice is final)
uint64_t k2_sort_keyrange(Stream & st, DevText const & T, int circular, KeyRangePlan const & plan, uint32_t part, FusedOut const & fo, SortStats * stats);

// Position sharding of the MSD sorter's first level (multi-GPU, msd.cuh): part p counts and scatters the text
// positions of ITS range of tiles for all bins, writing every record straight into the array of the part that
// owns the record's bin (the other GPUs' arrays through CUDA IPC peer mappings: the exchange is the stores of the
// scatter kernel); level 2 and the finish then run on the part's own bins.
struct XShard {
	unsigned b1 = 0, b2 = 0;
	uint32_t part = 0, nparts = 0;
	uint32_t t_lo = 0, t_hi = 0;                 // tiles of the text this part scatters
	DevBuf<uint32_t> toff;                       // column-scanned counts of those tiles
	std::vector<unsigned long long> total;       // global size of every level-1 bin (after the exchange of the counts)
	std::vector<uint32_t> bnd;                   // part p owns the bins [bnd[p], bnd[p+1])
	uint64_t rank_base = 0, records = 0;         // suffixes in the bins before this part's, and in them
};
// false: the sorter does not apply to this text (more than four codes, too short) -- use k2_sort_keyrange.
// d_totals (device, 2^b1 values): sizes of the bins over this part's tiles; *nbins = 2^b1
bool k2_xshard_count(Stream & st, DevText const & T, int circular, uint32_t part, uint32_t nparts, XShard & X, unsigned long long * d_totals, uint32_t * nbins);
// h_alltot: [nparts][2^b1] the d_totals of every part (host); recs[p] / cap[p]: part p's record array (device pointer valid here) and its
// capacity in records.  Throws when a bin or a part is too large for the path.
void k2_xshard_scatter(Stream & st, DevText const & T, int circular, XShard & X, const unsigned long long * h_alltot, unsigned long long * const * recs,
                       const uint64_t * cap, SortStats * stats);
// after every part has scattered: level 2 + finish on this part's bins; returns the number of suffixes left unresolved
uint64_t k2_xshard_finish(Stream & st, DevText const & T, int circular, XShard & X, unsigned long long * recs_own, FusedOut const & fo0, SortStats * stats,
                          StreamOut * so = nullptr);

// ---- K4 / K7 ----------------------------------------------------------------------------
// Rank dictionary, 2-bit flavour: 64-byte lines = 4 x uint32 cumulative counts + 48 bytes
// (192 symbols x 2 bit).  Byte flavour: 128 symbols per block, 256 x uint32 counts + 128 bytes.
struct DevDict {
	int flavour = 0;              // 2 or 8
	uint64_t n = 0;               // symbols
	const void * lines = nullptr; // flavour 2: uint4[4*nlines]; flavour 8: see rankdict.cuh
	uint64_t nlines = 0;
	uint32_t sigma = 0;           // number of codes
	uint32_t C[257];              // C[c] = # symbols with code < c (exception position excluded)
	uint32_t exc_pos = 0xffffffffu; // position whose stored symbol must not be counted
	uint32_t exc_code = 0;        // code stored at exc_pos
	uint32_t exc_lf = 0;          // LF target when a walk stands on exc_pos (terminator row -> 0)
};
size_t dict_bytes(int flavour, uint64_t n, uint32_t sigma);
void k4_build_dict(Stream & st, const uint8_t * bwt, uint64_t n, int flavour, uint32_t sigma, void * lines);

struct WalkStats { uint64_t steps = 0; uint64_t chains = 0; };
// K7: from every anchor (rank, pos) walk LF towards smaller positions for `len` steps,
// sampling SA by rank and ISA by position.  pos_off: value added to a text position before it
// is reported (0), n: BWT length.
void k7_walk(Stream & st, DevDict const & D, const uint32_t * anchor_rank, uint64_t nanchors, uint64_t arate,
             uint64_t n, uint64_t sarate, uint64_t isarate, uint64_t * sa_out, uint64_t * isa_out, WalkStats * ws,
             uint64_t q_lo, uint64_t q_hi /* anchors [q_lo,q_hi) only: the multi-GPU driver splits them over the ranks */);
// the same from anchors at arbitrary positions: anchor_steps[q] = distance to the previous anchor
void k7_walk_anchors(Stream & st, DevDict const & D, const uint32_t * anchor_rank, const uint64_t * anchor_pos, const uint64_t * anchor_steps,
                     uint64_t nanchors, uint64_t n, uint64_t sarate, uint64_t isarate, uint64_t * sa_out, uint64_t * isa_out, WalkStats * ws);
// checkbwt: the same walk, comparing the BWT symbol at every rank with the text (d_result: [0] mismatches, [1] a bad rank)
void k7_check_walk(Stream & st, DevDict const & D, const uint8_t * codes, uint64_t ntext, int has_term, const uint32_t * anchor_rank,
                   const uint64_t * anchor_pos, const uint64_t * anchor_steps, uint64_t nanchors, uint64_t n, uint64_t * d_result);
// LF-steps/s instrument (restates bwttestdecodespeed.cpp:82-96 for many chains)
void k7_lfbench(Stream & st, DevDict const & D, const uint32_t * start_rank, uint64_t nchains, uint64_t steps, uint32_t * out_rank);

} // namespace b3m
