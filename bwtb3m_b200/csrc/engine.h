// Engine class behind the C ABI (include/b3m.h).
#pragma once
#include "../../include/b3m.h"
#include "kernels.h"
#include "rankdict.cuh"

namespace b3m {

struct PhaseTimer {
	Stream & st;
	cudaEvent_t ev[16];
	int n = 0;
	explicit PhaseTimer(Stream & s) : st(s) { for (auto & e : ev) B3M_CUDA(cudaEventCreate(&e)); }
	~PhaseTimer() { for (auto & e : ev) cudaEventDestroy(e); }
	void mark() { B3M_CUDA(cudaEventRecord(ev[n++], st.s)); }
	float ms(int a, int b) { float t = 0; cudaEventElapsedTime(&t, ev[a], ev[b]); return t; }
};
struct EventAccum;

// what K5 needs from the left part of a merge: rank dictionary over L_A and C_A
struct GapCtx {
	uint64_t na = 0;           // |A|: the gap array has na + 1 counters
	DevBuf<uint8_t> lines;
	DictView D;
	CTab C;
};

// one leaf of the merge tree: the block's own suffixes in sorted order (kept for z-ranks)
struct BlockLeaf {
	uint64_t s = 0, mt = 0;
	bool keep = false;
	DevBuf<uint32_t> sa; // block-relative start positions
};
// a node of the merge tree over the text range [a0,a1): its BWT with a placeholder (code 0) at
// row `term`, the rank of the suffix starting at a0 (the reference's bwtterm row)
struct BlockNode {
	uint64_t a0 = 0, a1 = 0;
	uint32_t term = 0;
	DevBuf<uint8_t> L;
};

struct Engine {
	int device = 0;
	Arena arena;
	Stream st;
	bool own_stream = true;
	uint8_t * pinned = nullptr; // 4 KiB pinned staging area for small copies

	// input
	bool loaded = false;
	int inputtype = 0;
	DevBuf<uint8_t> raw, codes;  // codes: pac / pacterm inputs create them on demand (ensure_codes)
	DevBuf<uint8_t> lastcode;    // code of the last stored symbol (the seam of a circular text, the row behind the terminator)
	DevBuf<uint64_t> packed;   // 2-bit packed text when sigma <= 4 (textview.cuh)
	DevBuf<uint64_t> d_hist;
	DevText T;
	uint64_t hist[256];      // reference symbol -> count
	uint64_t codehist[256];  // dense code -> count (terminator excluded)
	uint8_t code2sym[256];

	// results of the last build
	bool have_results = false;
	bool ssa_only = false;     // results come from ssa_from_bwt: only sa/isa can be fetched
	b3m_build_params params{};
	uint64_t numblocks = 1;
	uint64_t prerate = 0, npre = 0, nsa = 0, nisa = 0;
	DevBuf<uint8_t> bwt;        // n dense codes (terminator row holds code 0, see root_exc_pos)
	DevBuf<uint32_t> prerank;   // rank of positions 0, prerate, 2*prerate, ...
	DevBuf<uint64_t> sa, isa;
	uint32_t * bwa_on_host = nullptr; // likewise BWA's packed BWT words (b3m_build_params.host_bwa)
	DevBuf<uint32_t> bwa_words;
	uint64_t * sa_on_host = nullptr; // host buffer that already holds the sampled SA of the last build (b3m_build_params.host_sa)
	DevBuf<uint8_t> dict;
	DevBuf<uint32_t> d_special;
	DevBuf<uint8_t> gt;         // multi-block: gt[i] = [rot(i) > rot(start of i's current node)]
	DevBuf<uint32_t> rsamp;     // multi-block: r(j) at the anchors of the right part of a merge
	DevDict D;
	uint32_t root_exc_pos = 0xffffffffu;

	// statistics
	SortStats sortstats;
	WalkStats walkstats;
	uint64_t gap_lf_steps = 0, gap_chains = 0, merge_bytes = 0, extract_bytes = 0, dict_bytes_moved = 0, decode_bytes = 0;
	uint64_t max_lcpnext = 0, large_lcp_blocks = 0;
	float ms_decode = 0, ms_sort = 0, ms_extract = 0, ms_dict = 0, ms_gap = 0, ms_merge = 0, ms_walk = 0, ms_total = 0;

	Engine(int dev, void * stream);
	~Engine();
	void reset_results();
	uint64_t work_bytes(uint64_t nblocks, uint64_t window, bool with_results) const;
	void load(const void * input, uint64_t nbytes, int itype, bool on_device);
	void build(b3m_build_params const & p);
	void build_blocks(PhaseTimer & pt, uint32_t * exc_pos);
	void build_tree(std::vector<BlockLeaf> & leaves, uint64_t lo, uint64_t hi, uint64_t base, uint64_t end, uint64_t bs, BlockNode & out,
	                EventAccum & tsort, EventAccum & tgap, EventAccum & tmerge);
	void leaf_build(BlockLeaf & leaf, uint64_t s, uint64_t m, uint8_t * L, uint32_t * term_pos, SortStats * ss);
	void node_merge(BlockNode & A, BlockNode & R, std::vector<BlockLeaf> & leaves, BlockNode & M, EventAccum & tgap, EventAccum & tmerge);
	uint32_t fetch_special(int slot);
	// block-level buffers in use (the engine's own, or the caller's in the multi-GPU driver)
	uint8_t * gtp = nullptr;
	uint32_t * prep = nullptr;
	uint32_t * rsp = nullptr;
	std::vector<BlockLeaf> dist_leaves;
	uint64_t choose_preisarate_pub(uint64_t n) const;
	void blk_begin(uint64_t preisarate, uint64_t largelcpthres, void * d_gt, void * d_prerank, void * d_rsamp);
	void blk_build_range(uint64_t a0, uint64_t a1, uint64_t nb, void * d_L_out, uint32_t * term_out);
	void blk_finish(const void * d_L_root, uint32_t term_root, uint64_t q_lo, uint64_t q_hi, uint64_t sarate, uint64_t isarate, int bwtonly, uint64_t nblocks);
	void gap_prepare(GapCtx & ctx, const uint8_t * LA, uint64_t a0, uint64_t na, uint32_t termA);
	void chain_geometry(uint64_t nr, uint64_t * chl, uint64_t * nch) const;
	void zranks_add(std::vector<BlockLeaf> & lv, uint64_t a0, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, uint32_t * r0);
	void gap_run(GapCtx & ctx, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, uint64_t c_lo, uint64_t c_hi, const uint32_t * r0,
	             const uint8_t * gt_in, uint8_t * gtnew, uint32_t termA, uint32_t * G, uint32_t * rs);
	void merge_run(const uint8_t * LA, uint64_t na, uint32_t termA, uint8_t * LR, uint64_t nr, uint32_t termR, uint64_t a1, uint32_t * G,
	               uint8_t * LM, uint32_t * termM);
	void merge_samples(uint64_t a0, uint64_t a1, uint64_t r1, uint32_t * pre, const uint32_t * Sincl, const uint32_t * rs);
	// multi-GPU, suffix-range sharding (sufsort.cu): one key range per engine, caller-owned output buffers
	KeyRangePlan kr_plan;
	void kr_build_part(uint32_t part, uint32_t nparts, b3m_build_params const & p, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa,
	                   void * d_special, uint64_t * unresolved);
	void kr_rows(uint32_t nparts, uint64_t * first);
	// multi-GPU, position sharding of the MSD sorter's first level (XShard, kernels.h)
	XShard xs;
	void prepare_shard_params(b3m_build_params const & p, uint32_t nparts);
	void xs_count(uint32_t part, uint32_t nparts, b3m_build_params const & p, void * d_totals, uint32_t * nbins);
	void xs_scatter(const uint64_t * h_alltot, void * const * d_recs, const uint64_t * caps);
	void xs_finish(void * d_recs_own, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa, void * d_special, uint64_t * unresolved);
	PhaseTimer * xs_pt = nullptr;
	unsigned long long * xs_sa_local = nullptr; // armed by xs_stream_sa for the next xs_finish: this GPU's copy of the sampled SA ..
	unsigned long long * xs_sa_host = nullptr;  // .. and the page-locked host buffer its part of the samples is sent to while the finish runs
	bool xs_sa_delivered = false;               // the last xs_finish did send its samples (the path needs 64 sub-buckets or more)
	void xs_stream_sa(void * d_sa_local, uint64_t * host_sa);
	void pack_rows(const void * d_rows, uint64_t nrows, void * d_packed, bool unpack);
	void kr_finish(const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa, const void * d_special, uint32_t nparts, bool adopt);
	// K8 / output side
	uint64_t rl_bytes = 0, rl_nruns = 0;
	void symbols_device(DevBuf<uint8_t> & out);
	uint64_t rl_runs(const uint8_t * s, DevBuf<uint32_t> & start);
	void write_bwt(const char * fn);
	void fetch_runs(uint8_t * h_sym, uint64_t * h_len, uint64_t cap, uint64_t * nruns_out);
	void fetch_bwa(uint32_t * h_words, uint64_t cap, uint64_t * primary, uint64_t * L2, uint64_t * seq_len_out);
	void ensure_codes();       // T.codes for the paths that read one byte per symbol (block merge tree, checkbwt)
	void pack_bwa_device(uint32_t * d_words, uint64_t w_lo, uint64_t w_hi);
	// sampled SA/ISA from an existing BWT and (rank,pos) anchors (bwtcomputessa path)
	void install_bwt_symbols(const uint8_t * h_bwt, uint64_t n, uint64_t extra_bytes);
	void lf_speed(const uint64_t * h_start, uint64_t nstart, uint64_t nchains, uint64_t steps, float * ms);
	void ssa_from_bwt(const uint8_t * h_bwt, uint64_t n, const uint64_t * h_pairs, uint64_t npairs, uint64_t sarate, uint64_t isarate);
	// checkbwt: verifies BWT symbols (reference symbol space) + anchors against the loaded text; returns mismatches
	uint64_t check_bwt(const uint8_t * h_bwt, uint64_t n, const uint64_t * h_pairs, uint64_t npairs, uint64_t * badrank);
	void make_dict(uint32_t exc_pos, uint32_t exc_code, uint32_t exc_lf);
	void fetch(uint8_t * h_bwt, uint64_t * h_pairs, uint64_t * h_sa, uint64_t * h_isa);
	void info(b3m_info * o);
	void lf_bench(uint64_t nchains, uint64_t steps, float * ms, uint64_t * checksum);
};

} // namespace b3m

// the C ABI's engine handle (include/b3m.h)
struct b3m_engine { b3m::Engine * e; std::string err; };
