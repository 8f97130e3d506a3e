/*
 * b3m.h -- C ABI of the B200-native BWT-by-balanced-block-merging engine (libb3m.so).
 *
 * This is the drop-in boundary for the hot path of gt1/bwtb3m.  The reference has no FFI layer:
 * its seam is one C++ static call into libmaus2,
 *     libmaus2::suffixsort::bwtb3m::BwtMergeSort::computeBwt(options,&std::cerr)
 *         (/root/reference/src/bwtb3m.cpp:62-63)
 * plus the process-level contract (key=value command line, output files).  Every entry point
 * below names the reference interface it replaces.  Plain C types only: pointers, sizes, fixed
 * width integers.  No exception crosses this boundary; every call returns 0 on success and a
 * non-zero code on failure with a message in the caller's buffer (file-level calls) or in
 * b3m_engine_last_error() (engine calls).  There is no CPU fallback: without a CUDA device the
 * calls fail.
 */
#ifndef B3M_H
#define B3M_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* input types kept from the reference's option `inputtype`
 * (/root/reference/src/bwtb3m.cpp:43; BwtMergeSortOptions::parseInputType,
 *  /root/reference/src/checkbwt.cpp:254-270); lz4 and utf-8 are out of scope. */
#define B3M_INPUT_BYTESTREAM 0
#define B3M_INPUT_COMPACTSTREAM 1
#define B3M_INPUT_PAC 2
#define B3M_INPUT_PACTERM 3

const char * b3m_version(void);

/* returns B3M_INPUT_* or -1 (replaces BwtMergeSortOptions::parseInputType) */
int b3m_parse_inputtype(const char * name);

/* ------------------------------------------------------------------------------------------
 * File level: the reference's three library calls on this path.
 * ---------------------------------------------------------------------------------------- */

/* Mirrors libmaus2::suffixsort::bwtb3m::BwtMergeSortOptions as constructed from the command
 * line (/root/reference/src/bwtb3m.cpp:43-56,62): same names, meaning and defaults. */
typedef struct b3m_options {
	const char * fn;              /* input file (positional argument 0) */
	const char * inputtype;       /* "bytestream" (default) | "compactstream" | "pac" | "pacterm" */
	const char * outputfilename;  /* default: tmpprefix + ".bwt" */
	uint64_t sasamplingrate;      /* default 32 */
	uint64_t isasamplingrate;     /* default 262144 */
	uint64_t mem;                 /* memory target, default 2 GiB; bounds the block size */
	uint64_t numthreads;          /* host threads (file encoding); default: logical CPUs */
	int bwtonly;                  /* 1: only .bwt + .preisa (+.preisa.meta) */
	const char * tmpprefix;       /* prefix for temporary files */
	const char * sparsetmpprefix; /* accepted for compatibility (gap arrays live in HBM) */
	int copyinputtomemory;        /* accepted for compatibility (input is always staged in HBM) */
	uint64_t largelcpthres;       /* default 16384 */
	int verbose;
	/* additions of this implementation */
	int device;                   /* CUDA device ordinal, default 0 */
	uint64_t numblocks;           /* 0: derive from mem (one block if it fits); >0: force */
	int ngpus;                    /* GPUs of this box the build spreads over (devices device .. device+ngpus-1, which must be
	                               * NVLink / PCIe peers); 0 or 1: one GPU.  The reference scales the same call with numthreads
	                               * (/root/reference/src/bwtb3m.cpp:48-50); see b3m_multi_* below */
} b3m_options;

/* Mirrors libmaus2::suffixsort::bwtb3m::BwtMergeSortResult
 * (/root/reference/src/checkbwt.cpp:39-53): the names of the files written. */
typedef struct b3m_result {
	char textfn[1024];
	char bwtfn[1024];
	char histfn[1024];
	char preisafn[1024];
	char metafn[1024];
	char safn[1024];   /* empty when bwtonly */
	char isafn[1024];  /* empty when bwtonly */
	uint64_t n;        /* BWT length (pacterm: bases + 1) */
	uint64_t numblocks;
	double seconds_total;
	double seconds_device;
} b3m_result;

void b3m_options_init(b3m_options * o);

/* replaces BwtMergeSort::computeBwt(options, logstr)  (/root/reference/src/bwtb3m.cpp:63) */
int b3m_compute_bwt(const b3m_options * o, b3m_result * res, char * err, size_t errlen);

/* replaces BwtComputeSSA::computeSSA(bwt,sasamplingrate,isasamplingrate,tmpfilenamebase,
 * copyinputtomemory,numthreads,maxsortmem,maxtmpfiles,logstr,ref_isa_fn,ref_sa_fn)
 * (/root/reference/src/bwtcomputessa.cpp:39-51): .bwt + .preisa -> .sa + .isa */
int b3m_compute_ssa(const char * bwtfn, uint64_t sasamplingrate, uint64_t isasamplingrate,
                    const char * tmpprefix, int copyinputtomemory, uint64_t numthreads,
                    uint64_t maxsortmem, uint64_t maxtmpfiles, int verbose,
                    const char * ref_isa_fn, const char * ref_sa_fn, int device,
                    char * err, size_t errlen);

/* replaces libmaus2::fm::MausFmToBwaConversion::rewrite(in,outbwt,outsa,numthreads)
 * (/root/reference/src/bwtb3mtobwa.cpp:29) */
int b3m_to_bwa(const char * inbwt, const char * outbwt, const char * outsa, char * err, size_t errlen);

/* replaces the reference's verifier checkBwt<io_type>(arg) (/root/reference/src/checkbwt.cpp:26-246): LF-walks
 * the BWT from the (rank,pos) anchors of <prefix>.preisa and compares every symbol with the text in front of the
 * walk's position; every text position is compared exactly once, the symbol counts must agree too.
 * *ok = 1 iff nothing differs (the reference prints this as "[V] gok=1"); *mismatches = how many symbols differ */
int b3m_check_bwt(const char * bwtfn, const char * textfn, const char * inputtype, uint64_t numthreads, int device, int verbose,
                  int * ok, uint64_t * mismatches, char * err, size_t errlen);

/* replaces the LF-steps/s instrument bwttestdecodespeed (/root/reference/src/bwttestdecodespeed.cpp:27-97):
 * `nchains` dependent LF chains of `steps` steps (0: min(ceil(n/nchains), 2^27) as in the reference) over a rank
 * dictionary of the BWT, started at evenly spaced samples of <prefix>.isa */
int b3m_lf_speed(const char * bwtfn, uint64_t nchains, uint64_t steps, uint64_t numthreads, int device, double * steps_per_s, double * seconds,
                 char * err, size_t errlen);

/* Reader of the .bwt container for bindings that cannot link C++: replaces
 * libmaus2::huffman::RLDecoder::getLength (/root/reference/src/hwtPreIsaToIsa.cpp:53) and a full
 * RLDecoder::decode() loop (/root/reference/src/bwtb3mdecoderl.cpp:27-46). */
int b3m_bwt_length(const char * bwtfn, uint64_t * n, char * err, size_t errlen);
int b3m_bwt_decode(const char * bwtfn, uint8_t * out, uint64_t cap, uint64_t numthreads, char * err, size_t errlen);
/* Host writer of the same container from a symbol array (replaces RLEncoderStd::encode + flush,
 * /root/reference/src/lcpbit.cpp:3677-3681); a tool for tests and converters -- the build path
 * b3m_compute_bwt encodes on the device and never calls it. */
/* RLDecoder::getBlockSymHistograms (/root/reference/src/bwtdecodeblock.cpp:356-365): writes the `.sparserank` file of a
 * .bwt -- per block of the container the occurrences of the symbols minsym..maxsym BEFORE the block, big-endian uint64
 * -- and b3m_bwt_rank answers rank_sym(L, i) from it with one seek and the runs of one block
 * (SparseRank::rankm, /root/reference/src/bwtdecodeblock.cpp:210-242).  Host only. */
int b3m_bwt_block_sym_histograms(const char * bwtfn, const char * outfn, int64_t minsym, int64_t maxsym, uint64_t numthreads, uint64_t * nblocks,
                                 char * err, size_t errlen);
int b3m_bwt_rank(const char * bwtfn, const char * sparserankfn, int64_t minsym, int64_t maxsym, int64_t sym, uint64_t i, uint64_t * rank,
                 char * err, size_t errlen);
int b3m_bwt_encode_host(const char * bwtfn, const uint8_t * syms, uint64_t n, char * err, size_t errlen);

/* The compactstream container (`inputtype=compactstream`) for bindings that cannot link C++: replaces
 * libmaus2::bitio::CompactArrayWriterFile(fn,b) + write + flush (/root/reference/src/fagzToCompact4.cpp:105,232,265;
 * /root/reference/src/digitsToCompact.cpp:35,79,89) and libmaus2::bitio::CompactDecoderWrapper
 * (/root/reference/src/decodecompact.cpp:30-36).  bits = 1..8, symbols one per byte.  The file is a serialised
 * CompactArray (/root/reference/src/digitsToCompact.cpp:122): four uint64 (bits, n, words, words) and the 64-bit words,
 * symbols MSB first in a word; written with native little-endian numbers and words, read in either byte order
 * (the byte order is not pinned by anything in the reference). */
int b3m_compact_write(const char * fn, unsigned bits, const uint8_t * syms, uint64_t n, char * err, size_t errlen);
int b3m_compact_info(const char * fn, unsigned * bits, uint64_t * n, char * err, size_t errlen);
int b3m_compact_read(const char * fn, uint8_t * out, uint64_t cap, char * err, size_t errlen);

/* ------------------------------------------------------------------------------------------
 * Engine level: the same path on caller-owned host or device buffers.  One engine per GPU and
 * per host thread; the engine owns its device memory and enqueues everything on one stream.
 * ---------------------------------------------------------------------------------------- */
typedef struct b3m_engine b3m_engine;

/* cuda_stream: a cudaStream_t to enqueue on, or NULL for an engine-owned stream */
int b3m_engine_create(int device, void * cuda_stream, b3m_engine ** out, char * err, size_t errlen);
void b3m_engine_destroy(b3m_engine * e);
const char * b3m_engine_last_error(const b3m_engine * e);

/* K1: stage + decode the input (the bytes of the input file) and build the symbol histogram.
 * Replaces the {Byte,Compact,Pac,PacTerm}InputTypes readers (/root/reference/src/checkbwt.cpp:260-266). */
int b3m_engine_load_host(b3m_engine * e, const void * input, uint64_t nbytes, int inputtype);
int b3m_engine_load_device(b3m_engine * e, const void * d_input, uint64_t nbytes, int inputtype);

typedef struct b3m_build_params {
	uint64_t numblocks;        /* >= 1 */
	uint64_t preisarate;       /* spacing of the (rank,pos) anchors, power of two; 0: choose */
	uint64_t sasamplingrate;   /* power of two */
	uint64_t isasamplingrate;  /* power of two */
	int bwtonly;
	uint64_t largelcpthres;
	int sampling;              /* B3M_SAMPLING_*: how the sampled SA/ISA are obtained */
	uint64_t * host_sa;        /* optional PINNED host buffer of ceil(n/sasamplingrate) values: the sampled SA is copied
	                            * there while the build is still running (one block, direct sampling); a later
	                            * b3m_engine_fetch with the same pointer does not copy again.  NULL: off */
	uint32_t * host_bwa;       /* likewise (pacterm, one block): PINNED host buffer of ceil((n-1)/16) words that receives
	                            * BWA's packed BWT (see b3m_engine_fetch_bwa) while the build runs.  NULL: off */
	int sortpath;              /* B3M_SORT_*: suffix sorter of a one-block build over an alphabet of at most four codes */
	int gapmode;               /* B3M_GAP_*: how K5 counts the gap array of a merge (builds of two or more blocks) */
} b3m_build_params;
/* AUTO: gap arrays that fit the L2 cache are counted by the chains themselves (atomic adds), larger ones from a
 * list of the chains' ranks partitioned by one radix pass, so that the counters in use stay in L2.  ATOMIC / LIST
 * force one of them (tests, measurements).  Same results. */
#define B3M_GAP_AUTO 0
#define B3M_GAP_ATOMIC 1
#define B3M_GAP_LIST 2
/* AUTO: the MSD bucket sort (two global levels + a finish in shared memory) for texts of 2^16 symbols or more,
 * the LSD radix sort otherwise.  LSD / MSD force one of them (tests, measurements).  Same results. */
#define B3M_SORT_AUTO 0
#define B3M_SORT_LSD 1
#define B3M_SORT_MSD 2
/* AUTO: straight from the suffix array when the build holds all of it (one block), by the LF walk
 * from the anchors (the reference's method, /root/reference/src/hwtPreIsaToIsa.cpp:114-161)
 * otherwise.  WALK: always by the LF walk.  The results are identical. */
#define B3M_SAMPLING_AUTO 0
#define B3M_SAMPLING_WALK 1

/* K2..K7: block sort -> gap arrays -> merge -> final BWT, anchors, sampled SA/ISA; results stay
 * in HBM until fetched.  Replaces BwtMergeSortTemplate<InputTypes>::computeBwt. */
int b3m_engine_build(b3m_engine * e, const b3m_build_params * p);

typedef struct b3m_info {
	uint64_t n;            /* BWT length */
	uint64_t sigma;        /* distinct symbols (terminator included) */
	uint64_t numblocks;
	uint64_t preisarate, npreisa;
	uint64_t sasamplingrate, nsa;
	uint64_t isasamplingrate, nisa;
	uint64_t hist[256];    /* symbol -> count, reference symbol space */
	/* counters the roofline report is computed from */
	uint64_t sort_rounds, radix_passes, radix_bytes, sort_active_sum, sort_other_bytes;
	uint64_t gap_lf_steps, walk_lf_steps, walk_chains, gap_chains;
	uint64_t merge_bytes, extract_bytes, dict_bytes, decode_bytes;
	uint64_t launches;     /* kernels launched since the engine was created */
	uint64_t max_lcpnext;
	/* device time per phase of the last build, milliseconds (CUDA events on the engine stream) */
	float ms_decode, ms_sort, ms_extract, ms_dict, ms_gap, ms_merge, ms_walk, ms_total;
	/* K2: suffixes that shared their whole first sort key with another suffix, and those still
	 * tied after the in-CTA group sort (they enter the prefix-doubling rounds) */
	uint64_t sort_tied0, sort_unresolved0;
	/* device memory the engine holds (text slab + working-set slab of the last build) and the most it had in use */
	uint64_t arena_capacity, arena_peak;
} b3m_info;
int b3m_engine_info(b3m_engine * e, b3m_info * info);

/* D2H of the results; any pointer may be NULL.  bwt: n symbols in the reference's symbol space;
 * preisa_pairs: 2*npreisa uint64 (rank,pos); sa: nsa uint64 (SA[k*sasamplingrate]);
 * isa: nisa uint64 (ISA[k*isasamplingrate]). */
int b3m_engine_fetch(b3m_engine * e, uint8_t * bwt, uint64_t * preisa_pairs, uint64_t * sa, uint64_t * isa);

/* device pointers of the results (valid until the next load/build): bwt codes are dense ranks,
 * see b3m_engine_info; for the multi-GPU orchestration and device-resident benchmarks. */
int b3m_engine_device_results(b3m_engine * e, const void ** d_bwt_codes, const void ** d_preisa_rank,
                              const void ** d_sa, const void ** d_isa);

/* K8: run-length + Huffman encode the BWT of the last build on the device and write the .bwt
 * container (replaces the RLEncoder output stage of computeBwt). */
int b3m_engine_write_bwt(b3m_engine * e, const char * bwtfn);
/* The run stream (symbol, length) of the BWT, for a reference-side binding that feeds
 * libmaus2's own RLEncoderStd::encodeRun (/root/reference/src/lcpbit.cpp:728-745).  With both
 * pointers NULL only *nruns is returned. */
int b3m_engine_fetch_runs(b3m_engine * e, uint8_t * syms, uint64_t * lens, uint64_t cap, uint64_t * nruns);
/* K9: the BWA index of the last build (pacterm input only), packed on the device; engine half of
 * MausFmToBwaConversion::rewrite (/root/reference/src/bwtb3mtobwa.cpp:29) for callers that hold the
 * engine.  bwt_words: ceil(seq_len/16) uint32 exactly as in BWA's .bwt file behind its 40-byte
 * header (16 symbols per word, symbol k at bits (15-(k&15))*2, terminator row removed);
 * primary = rank of the suffix at text position 0; L2[0..4] = cumulative base counts; seq_len = n-1.
 * BWA's .sa payload is sa[1..] of b3m_engine_fetch.  With bwt_words NULL only the scalars are returned. */
int b3m_engine_fetch_bwa(b3m_engine * e, uint32_t * bwt_words, uint64_t cap_words, uint64_t * primary, uint64_t * L2, uint64_t * seq_len);
/* K9 into a caller-owned device buffer (cudaMalloc'ed, possibly shared through CUDA IPC): BWA's words
 * [w_lo, w_hi) of the last pacterm build, on the engine's stream, no host copy. */
int b3m_engine_pack_bwa(b3m_engine * e, void * d_words, uint64_t w_lo, uint64_t w_hi);
/* K4 + K7 on an existing BWT: sampled SA/ISA from n symbols and npairs (rank,pos) anchors
 * (engine half of b3m_compute_ssa); fetch with b3m_engine_fetch(e, NULL, NULL, sa, isa). */
int b3m_engine_ssa_from_bwt(b3m_engine * e, const uint8_t * bwt, uint64_t n, const uint64_t * preisa_pairs, uint64_t npairs,
                            uint64_t sasamplingrate, uint64_t isasamplingrate);

/* ---- multi-GPU driver: the steps of one merge on caller-owned DEVICE buffers -----------------
 * One engine per GPU, text replicated (every rank calls b3m_engine_load_*), rank i sorts the text
 * range [i*bs,(i+1)*bs); block BWTs, gap arrays, gt bits and anchors travel between the ranks
 * with NCCL (bwtb3m_b200/multigpu.py is the driver).  Replaces the merge tree of
 * BwtMergeSortTemplate::computeBwt (libmaus2; reached from /root/reference/src/bwtb3m.cpp:63).
 *   d_gt      n bytes          gt[i] = [rot(i) > rot(start of i's node)]        (Appendix A.3)
 *   d_prerank ceil(n/rate) u32 rank of position q*rate inside its current node
 *   d_rsamp   ceil(n/rate) u32 r(j) at the anchors of the right part (written by blk_gap)     */
/* anchor spacing b3m_engine_build would choose for the loaded text (64 when bwtonly, as the reference) */
int b3m_engine_default_preisarate(b3m_engine * e, int bwtonly, uint64_t * rate);
int b3m_engine_blk_begin(b3m_engine * e, uint64_t preisarate, uint64_t largelcpthres, void * d_gt, void * d_prerank, void * d_rsamp);
/* sort the range [a0,a1) (numblocks leaves merged locally): d_L_out receives a1-a0 codes, the
 * placeholder row (the reference's bwtterm) is returned in *term_out */
int b3m_engine_blk_build_range(b3m_engine * e, uint64_t a0, uint64_t a1, uint64_t numblocks, void * d_L_out, uint32_t * term_out);
/* chain geometry for a right part of nr symbols (identical on every rank) */
int b3m_engine_blk_chains(b3m_engine * e, uint64_t nr, uint64_t * chl, uint64_t * nch);
/* d_r0[c] (u32) += z-rank of chain c's start over this engine's leaves inside [a0,a1) */
int b3m_engine_blk_zranks(b3m_engine * e, uint64_t a0, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, void * d_r0);
/* K5 for the chains [c_lo,c_hi) of the merge A=[a0,a0+na) | R=[a0+na,r1): d_G (u32[na+1]) += gaps,
 * d_gtnew[p-a1] = new gt bits, d_rsamp[p/rate] = r(p) */
int b3m_engine_blk_gap(b3m_engine * e, const void * d_LA, uint64_t a0, uint64_t na, uint32_t termA, uint64_t r1, uint64_t chl, uint64_t nch,
                       uint64_t c_lo, uint64_t c_hi, const void * d_r0, void * d_gtnew, void * d_G);
/* K6: d_LM (na+nr codes) = merge by d_G, which is left holding its inclusive prefix sums */
int b3m_engine_blk_merge(b3m_engine * e, const void * d_LA, uint64_t na, uint32_t termA, void * d_LR, uint64_t nr, uint32_t termR, uint64_t a1,
                         void * d_G, void * d_LM, uint32_t * termM);
/* anchors of [a0,r1) move by the rank maps of the merge (d_prerank, d_rsamp of blk_begin) */
int b3m_engine_blk_merge_samples(b3m_engine * e, uint64_t a0, uint64_t a1, uint64_t r1, const void * d_Sincl);
/* install the root BWT (n codes) and the anchors in d_prerank, build the dictionary and walk the
 * anchors [q_lo,q_hi); unset SA/ISA samples hold ~0 so that the ranks' partial arrays combine */
int b3m_engine_blk_finish(b3m_engine * e, const void * d_L_root, uint32_t term_root, uint64_t q_lo, uint64_t q_hi, uint64_t sasamplingrate,
                          uint64_t isasamplingrate, int bwtonly, uint64_t numblocks);

/* ---- multi-GPU driver, suffix-range sharding ---------------------------------------------------
 * The text is replicated; the suffixes are split by the leading symbols of their first sort key into
 * nparts key ranges of about equal size (every rank derives the same split from the text), and
 * engine `part` sorts range `part` only.  Because the ranges are contiguous in suffix-array order,
 * each rank produces a contiguous slice of the BWT and of the rank-sampled SA, and its share of
 * the anchors / position-sampled ISA, at their GLOBAL places in caller-owned device buffers that
 * must be zero on entry: d_bwt n+16 bytes (codes), d_prerank ceil(n/preisarate) u32, d_sa / d_isa
 * u64 (NULL when bwtonly), d_special 4 u32.  The ranks' buffers combine by a sum (ncclSum); no gap
 * array and no merge is needed.  *unresolved != 0 means the text has repeats this path does not
 * sort (longer than the two sort keys); the caller then uses the block merge tree (blk_* above).
 * b3m_engine_shard_finish installs the combined buffers as the engine's results.  This path is an
 * addition of this implementation; it replaces the same computeBwt call as b3m_engine_build. */
int b3m_engine_shard_build(b3m_engine * e, uint32_t part, uint32_t nparts, const b3m_build_params * p, void * d_bwt, void * d_prerank,
                           void * d_sa, void * d_isa, void * d_special, uint64_t * unresolved);
/* first_row[p] (p = 0..nparts) = first BWT row of key range p: range p produces the rows
 * [first_row[p], first_row[p+1]) of the BWT and the SA samples of the ranks in that interval, so the
 * dense results can travel as slices instead of a reduction over the whole arrays */
int b3m_engine_shard_rows(b3m_engine * e, uint32_t nparts, uint64_t * first_row);
/* transport packing of BWT rows for the slice exchange (alphabets of at most four codes): nrows
 * codes <-> ceil(nrows/4) bytes, row 4q+j in bits [2j,2j+1] of byte q; fails for larger alphabets */
int b3m_engine_pack_rows(b3m_engine * e, const void * d_rows, uint64_t nrows, void * d_packed);
int b3m_engine_unpack_rows(b3m_engine * e, const void * d_packed, uint64_t nrows, void * d_rows);
int b3m_engine_shard_finish(b3m_engine * e, const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa,
                            const void * d_special, uint32_t nparts);

/* ---- multi-GPU driver, position sharding of the first sorting level (alphabets of at most four codes) -------
 * The scalable form of the sharded build: nothing is computed on the whole text by every GPU.  Part p
 *   1. b3m_engine_xshard_count: counts, for ITS range of text positions, how many suffixes start with each value of
 *      their leading bits (*nbins values in d_totals, device; *nbins = 0: the sorter does not apply to this text, use
 *      b3m_engine_shard_build).  The caller exchanges the counts (all-gather, nparts * nbins values);
 *   2. b3m_engine_xshard_scatter: builds the sort records of its positions and stores every record straight into the
 *      record array of the part that owns the record's key range -- d_recs[q] is part q's array as mapped HERE (for
 *      q != p a CUDA IPC peer mapping: the records cross NVLink as the stores of the scatter kernel), caps[q] its
 *      capacity in 8-byte records (about n / nparts * 1.25).  all_totals: host, [nparts][nbins].  Fails (on every
 *      part alike) when a key range does not fit.  The caller then synchronises the parts (any collective);
 *   3. b3m_engine_xshard_finish: sorts its own key range (d_recs_own = d_recs[p]) and emits BWT rows, anchors and
 *      samples at their global places as b3m_engine_shard_build does.
 * b3m_engine_shard_rows / _shard_finish / _shard_adopt apply afterwards.  Build parameters are those of step 1. */
int b3m_engine_xshard_count(b3m_engine * e, uint32_t part, uint32_t nparts, const b3m_build_params * p, void * d_totals, uint32_t * nbins);
int b3m_engine_xshard_scatter(b3m_engine * e, const uint64_t * all_totals, void * const * d_recs, const uint64_t * caps);
/* Optional, before b3m_engine_xshard_finish: the samples of the rank-sampled SA that this rank produces are also
 * stored into d_sa_local (this GPU's memory, ceil(n/sasamplingrate) uint64 like d_sa) and copied from there into the
 * page-locked host buffer host_sa -- at their global places, chunk by chunk while the finish kernel still runs, over
 * THIS GPU's PCIe link.  When the call returns they are in host memory (sample 0 of a terminated text excepted: the
 * owner of d_sa holds it).  Applies to the next b3m_engine_xshard_finish only; needs sasamplingrate >= 32. */
int b3m_engine_xshard_stream_sa(b3m_engine * e, void * d_sa_local, uint64_t * host_sa);
/* after b3m_engine_xshard_finish: 1 if this rank's samples were sent that way (key ranges of fewer than 64 sub-buckets
 * are not: the caller then copies the samples out of d_sa as usual) */
int b3m_engine_xshard_sa_delivered(b3m_engine * e, int * delivered);
int b3m_engine_xshard_finish(b3m_engine * e, void * d_recs_own, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa, void * d_special,
                             uint64_t * unresolved);

/* the same without the copies: the engine refers to the caller's buffers, which must stay valid (and unchanged)
 * until the engine's next load or build */
int b3m_engine_shard_adopt(b3m_engine * e, const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa,
                           const void * d_special, uint32_t nparts);

/* Result buffers of a multi-GPU build that every rank writes DIRECTLY (one process per GPU): the owner allocates
 * them with b3m_dev_alloc and exports a 64-byte CUDA IPC handle, the other processes open it and pass the mapped
 * pointer to b3m_engine_shard_build, whose kernels then store their BWT rows, anchors and samples into the
 * owner's HBM over NVLink (peer stores inside the sorting kernels: no gather, no reduction, no collective on the
 * data path).  Every entry is written by exactly one rank, so the buffers need no initialisation. */
int b3m_dev_alloc(int device, uint64_t bytes, void ** dptr, char * err, size_t errlen);
int b3m_dev_free(int device, void * dptr, char * err, size_t errlen);
int b3m_ipc_export(int device, const void * dptr, void * handle64, char * err, size_t errlen);
int b3m_ipc_open(int device, const void * handle64, void ** dptr, char * err, size_t errlen);
int b3m_ipc_close(int device, void * dptr, char * err, size_t errlen);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `cuda_stream` (NULL: the default stream) of `device`.
 * With it every rank of a multi-GPU build pulls ITS slice of the owner's result buffers over NVLink and sends
 * it to the host over its own PCIe link (bwtb3m_b200/multigpu.py fetch_distributed), instead of all results
 * leaving through the owner's link. */
int b3m_dev_copy(int device, void * dst, const void * src, uint64_t bytes, void * cuda_stream, char * err, size_t errlen);

/* ---- multi-GPU build inside ONE process (what `ngpus` of b3m_options / of the bwtb3m command line selects) ----
 * ngpus engines, one host thread per GPU, peer access between all of them; no NCCL and no second process.
 * Replaces the same BwtMergeSort::computeBwt call (/root/reference/src/bwtb3m.cpp:62-63), whose own scaling knob is
 * numthreads (/root/reference/src/bwtb3m.cpp:48-50).
 *   b3m_multi_load_host  every GPU uploads 1/ngpus of the input over its own PCIe link, the pieces are exchanged by
 *                        peer copies (NVLink), every GPU decodes the text (K1);
 *   b3m_multi_build      position-sharded sort (b3m_engine_xshard_*; alphabets of more than four codes: key ranges,
 *                        b3m_engine_shard_build) with every GPU storing its BWT rows, anchors and samples into GPU 0's
 *                        buffers; numblocks > 1, sampling = WALK and texts with repeats beyond the sort keys are built
 *                        by GPU 0 alone (the general path).  Same results as b3m_engine_build, bit for bit.
 * Afterwards b3m_multi_engine(m, 0) holds the results: fetch / fetch_bwa / write_bwt / info as after b3m_engine_build.
 * devices: ngpus ordinals, or NULL for 0 .. ngpus-1. */
typedef struct b3m_multi b3m_multi;
int b3m_multi_create(int ngpus, const int * devices, b3m_multi ** out, char * err, size_t errlen);
void b3m_multi_destroy(b3m_multi * m);
const char * b3m_multi_last_error(const b3m_multi * m);
int b3m_multi_load_host(b3m_multi * m, const void * input, uint64_t nbytes, int inputtype);
int b3m_multi_build(b3m_multi * m, const b3m_build_params * p);
b3m_engine * b3m_multi_engine(b3m_multi * m, int i);
/* what the last load / build did: strategy = "xshard" | "shard" | "single" | "single (text with long repeats)";
 * host wall-clock milliseconds of the two calls */
int b3m_multi_stats(b3m_multi * m, char * strategy, size_t len, double * ms_load, double * ms_build);

/* LF-steps/s instrument on the dictionary of the last build: nchains dependent LF chains of
 * `steps` steps each, started at evenly spaced sampled ranks; returns elapsed device ms.
 * Restates /root/reference/src/bwttestdecodespeed.cpp:82-96 for thousands of chains. */
int b3m_engine_lf_bench(b3m_engine * e, uint64_t nchains, uint64_t steps, float * ms, uint64_t * checksum);

/* wait for the engine's stream */
int b3m_engine_sync(b3m_engine * e);

/* Opt-in per-kernel timing for the roofline report: when on, the heavy kernels of the next
 * builds are bracketed with CUDA events on the engine's stream.  b3m_engine_kernel_times()
 * writes one line per kernel name, "name launches total_ms algorithmic_bytes\n", and clears
 * the records. */
int b3m_engine_set_profile(b3m_engine * e, int on);
int b3m_engine_kernel_times(b3m_engine * e, char * buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif
