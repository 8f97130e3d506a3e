// Multi-block path: block partition + bounded LCP (A4), leaf sorting with look-ahead, gt bits,
// anchors (A5), z-ranks, K5 gap arrays by backward search (A7), K6 gap-driven merge (A8).
// Replaces BwtMergeSortTemplate<InputTypes>::computeBwt's block / merge-tree stages reached from
// /root/reference/src/bwtb3m.cpp:62-63 (libmaus2, absent); the identities every kernel
// implements are SURVEY.md Appendix A.1-A.4 (brute-force checked there, restated on the CPU in
// oracle/b3m_oracle.c: leaf_build, node_merge).
#include "engine.h"
#include "rankdict.cuh"
#include "scan.cuh"
#include "textview.cuh"
#include "radix.cuh"
#include <string.h>
#include <algorithm>
#include <string>

namespace b3m {

// ------------------------------------------------------------------------------------------
// circular text access; the implicit terminator of pacterm (position ntext) is a unique symbol
// smaller than every code
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int text_sym(TextRef const & t, uint64_t p) {
	return (t.has_term && p == t.ntext) ? -1 : (int)t.codes[p];
}

// lexicographic comparison of the rotations starting at a and b
__device__ int cmp_rot(TextRef const & t, uint64_t a, uint64_t b) {
	if (a == b) return 0;
	for (uint64_t k = 0; k < t.n; ++k) {
		int const x = text_sym(t, a), y = text_sym(t, b);
		if (x != y) return x < y ? -1 : 1;
		a = (a + 1 == t.n) ? 0 : a + 1;
		b = (b + 1 == t.n) ? 0 : b + 1;
	}
	return 0;
}

// LCP of the rotations starting at a and b, bounded by cap, and the sign of the comparison at the first
// difference (0: none within cap).  2-bit packed texts are compared 32 symbols per step while both
// cursors are away from the end of the text.
__device__ __forceinline__ void cmp_lcp(TextRef const & t, uint64_t a, uint64_t b, uint64_t cap, uint64_t & lcp, int & sign) {
	uint64_t l = 0;
	sign = 0;
	while (l < cap) {
		if (t.packed && a + 32 <= t.ntext && b + 32 <= t.ntext) {
			uint64_t const x = pk_window(t.packed, a), y = pk_window(t.packed, b);
			if (x == y) {
				uint64_t const step = cap - l < 32 ? cap - l : 32;
				l += step; a += step; b += step;
				if (a == t.n) a = 0;
				if (b == t.n) b = 0;
				continue;
			}
			uint64_t const d = (uint64_t)__clzll((long long)(x ^ y)) >> 1; // first differing symbol
			if (d >= cap - l) { l = cap; break; }
			l += d;
			sign = x > y ? 1 : -1;
			break;
		}
		int const sx = text_sym(t, a), sy = text_sym(t, b);
		if (sx != sy) { sign = sx < sy ? -1 : 1; break; }
		if (sx < 0) break; // both cursors on the terminator: the same position
		++l;
		a = (a + 1 == t.n) ? 0 : a + 1;
		b = (b + 1 == t.n) ? 0 : b + 1;
	}
	lcp = l;
}

// A4: lcpnext = max over i in [s,e), i != ep, of LCP(rot(i), rot(ep)), each LCP bounded by cap
__global__ void __launch_bounds__(256)
k_lcpnext(TextRef t, uint64_t s, uint64_t e, uint64_t ep, uint64_t cap, uint64_t skip_below, unsigned long long * __restrict__ out) {
	uint64_t const i = s + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t l = 0;
	if (i < e && i != ep) {
		int sign;
		cmp_lcp(t, i, ep, cap, l, sign);
		if (l < skip_below) l = 0;
	}
	// warp maximum, one atomic per warp
	uint32_t lo = (uint32_t)l, hi = (uint32_t)(l >> 32);
	uint32_t const mhi = __reduce_max_sync(0xffffffffu, hi);
	lo = (hi == mhi) ? lo : 0u;
	uint32_t const mlo = __reduce_max_sync(0xffffffffu, lo);
	if ((threadIdx.x & 31) == 0) {
		unsigned long long const m = ((unsigned long long)mhi << 32) | mlo;
		if (m) atomicMax(out, m);
	}
}

// leaf: block BWT from the block's own suffixes in sorted order; the block-start suffix gets
// the placeholder code 0 (bwtterm of the reference) and its row is recorded; the (rank,pos) anchors
// of the block are picked up on the way
__global__ void __launch_bounds__(256)
k_leaf_emit(const uint8_t * __restrict__ codes, uint64_t s, const uint32_t * __restrict__ bsa, uint64_t mt, uint32_t shift,
            uint64_t ratemask, uint32_t rateshift, uint8_t * __restrict__ L, uint32_t * __restrict__ special, uint32_t * __restrict__ prerank) {
	uint64_t const k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= mt) return;
	uint32_t const i = bsa[k];
	uint8_t c = 0;
	if (i == 0) special[0] = (uint32_t)k + shift; else c = codes[s + i - 1];
	L[k + shift] = c;
	uint64_t const p = s + i;
	if ((p & ratemask) == 0) prerank[p >> rateshift] = (uint32_t)k + shift;
}

// leaf: gt bits relative to the block start (Appendix A.3): gt[i] = [rot(i) > rot(s)], decided on the text
// itself -- a streaming pass with coalesced writes instead of a rank-by-position scatter
__global__ void __launch_bounds__(256)
k_leaf_gt(TextRef t, uint64_t s, uint64_t mt, uint8_t * __restrict__ gt) {
	uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= mt) return;
	int sign = 0;
	if (i) { uint64_t l; cmp_lcp(t, s + i, s, t.n, l, sign); }
	gt[s + i] = sign > 0;
}

// number of suffixes of one leaf that are smaller than rot(z), for the start point z of every chain
__global__ void __launch_bounds__(128)
k_zrank(TextRef t, const uint32_t * __restrict__ bsa, uint64_t mt, uint64_t s, uint64_t a1, uint64_t chl, uint64_t r1, uint64_t nch,
        uint32_t * __restrict__ out) {
	uint64_t const c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nch) return;
	uint64_t zhi = a1 + (c + 1) * chl;
	if (zhi > r1) zhi = r1;
	uint64_t const z = zhi == t.n ? 0 : zhi;
	uint64_t lo = 0, hi = mt;
	while (lo < hi) {
		uint64_t const mid = (lo + hi) >> 1;
		if (cmp_rot(t, s + bsa[mid], z) < 0) lo = mid + 1; else hi = mid;
	}
	out[c] += (uint32_t)lo;
}

// gt_R[r1] = [rot(r1) > rot(a1)] decided directly (oracle node_merge)
__global__ void k_gt_top(TextRef t, uint64_t r1, uint64_t a1, uint32_t * __restrict__ special) {
	special[2] = cmp_rot(t, r1 == t.n ? 0 : r1, a1) > 0 ? 1u : 0u;
}

// histogram of a text range (any alignment)
__global__ void __launch_bounds__(256)
k_hist_range(const uint8_t * __restrict__ in, uint64_t n, unsigned long long * __restrict__ hist) {
	__shared__ uint32_t sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) atomicAdd(&sh[in[i]], 1u);
	__syncthreads();
	if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// K5: gap array.  One backward-search chain per thread over R's text, right to left
// (Appendix A.2):  r <- C_A[c] + rank_c(L_A, r) + [c == T[a1-1] and gt_R[j]],  G[r]++ .
// ------------------------------------------------------------------------------------------

// byte k (0..15) of a 16-byte block held in registers
__device__ __forceinline__ uint32_t byte_of(uint4 const & v, uint32_t k) {
	uint32_t const w = (k & 8u) ? ((k & 4u) ? v.w : v.z) : ((k & 4u) ? v.y : v.x);
	return (w >> ((k & 3u) * 8u)) & 255u;
}
__device__ __forceinline__ void or_byte(uint4 & v, uint32_t k, uint32_t val) {
	uint32_t const b = val << ((k & 3u) * 8u);
	uint32_t const wi = k >> 2;
	v.x |= wi == 0 ? b : 0u; v.y |= wi == 1 ? b : 0u; v.z |= wi == 2 ? b : 0u; v.w |= wi == 3 ? b : 0u;
}

// Every chain streams through its own stretch of the text, of gt_in and of gt_out, one byte per
// step.  With half a million chains in flight those sectors do not survive in L1/L2 between two
// steps (an ncu capture showed 370 B of DRAM reads per step, profiles/r01s), so each stream is
// read / written in aligned 16-byte blocks held in registers: per step that leaves the dictionary
// line (64 B) and the counter sector of G.
__global__ void __launch_bounds__(256)
k_gap(DictView D, CTab C, TextRef t, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t c_lo, uint64_t c_hi, const uint32_t * __restrict__ r0,
      const uint8_t * __restrict__ gt_in /* indexed by text position */, uint8_t * __restrict__ gt_out /* indexed by position - a1 */,
      const uint32_t * __restrict__ special, uint32_t isa_a0, uint64_t ratemask, uint32_t rateshift,
      uint32_t * __restrict__ G, uint32_t * __restrict__ rsamp /* indexed by position / rate */,
      uint32_t * __restrict__ rlist /* nullptr: count in G at once; else step k of chain c writes rlist[k * rl_stride + (c - c_lo)] */, uint64_t rl_stride) {
	uint64_t const c = c_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c_hi) return;
	uint64_t const zlo = a1 + c * chl;
	uint64_t zhi = zlo + chl;
	if (zhi > r1) zhi = r1;
	uint32_t const lastA = t.codes[a1 - 1];
	uint32_t r = r0[c];
	uint4 cb = make_uint4(0, 0, 0, 0), gb = make_uint4(0, 0, 0, 0), ob = make_uint4(0, 0, 0, 0);
	uint64_t cblk = ~0ull, gblk = ~0ull, oblk = ~0ull;
	uint32_t omask = 0;
	uint64_t const gt_blocks = t.n >> 4; // whole 16-byte blocks of gt_in (the tail is read bytewise)
	auto flush = [&]() {
		if (omask == 0xffffu) *reinterpret_cast<uint4 *>(gt_out + (oblk << 4)) = ob;
		else for (uint32_t k = 0; k < 16; ++k) if ((omask >> k) & 1u) gt_out[(oblk << 4) + k] = (uint8_t)byte_of(ob, k);
	};
	for (uint64_t j = zhi; j > zlo; --j) {
		uint64_t const p = j - 1;
		if (t.has_term && p == t.ntext) r = 0; // the terminator suffix is smaller than every suffix of A
		else {
			uint32_t g;
			if (j < r1) {
				if ((j >> 4) < gt_blocks) {
					if ((j >> 4) != gblk) { gblk = j >> 4; gb = __ldg(reinterpret_cast<const uint4 *>(gt_in) + gblk); }
					g = byte_of(gb, (uint32_t)(j & 15u));
				} else g = gt_in[j];
			} else g = special[2];
			if ((p >> 4) != cblk) { cblk = p >> 4; cb = __ldg(reinterpret_cast<const uint4 *>(t.codes) + cblk); } // codes are padded by 16 bytes
			uint32_t const c0 = byte_of(cb, (uint32_t)(p & 15u));
			r = C.c[c0] + dict_rank(D, c0, r) + ((c0 == lastA && g) ? 1u : 0u);
		}
		if (rlist) rlist[(zhi - j) * rl_stride + (c - c_lo)] = r; // coalesced: the threads of a warp are consecutive chains
		else atomicAdd(&G[r], 1u);
		uint64_t const q = p - a1;
		if ((q >> 4) != oblk) { if (omask) flush(); oblk = q >> 4; ob = make_uint4(0, 0, 0, 0); omask = 0; }
		or_byte(ob, (uint32_t)(q & 15u), r > isa_a0 ? 1u : 0u); // Appendix A.3
		omask |= 1u << (uint32_t)(q & 15u);
		if ((p & ratemask) == 0) rsamp[p >> rateshift] = r;
	}
	if (omask) flush();
}

// G[r]++ for the ranks of a list (0xffffffff: no entry).  The list has been partitioned by the top 8 bits of the rank
// (one radix pass), and CTAs run in list order, so the counters touched at any time span 1/256 of G: they stay in L2
// and the read-modify-write of a counter never reaches DRAM one sector at a time.
__global__ void __launch_bounds__(256)
k_gap_hist(const uint32_t * __restrict__ list, uint64_t n, uint32_t * __restrict__ G) {
	uint64_t const i0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
	if (i0 + 4 <= n) {
		uint4 const v = ld_stream_u4(reinterpret_cast<const uint4 *>(list + i0));
		if (v.x != 0xffffffffu) atomicAdd(&G[v.x], 1u);
		if (v.y != 0xffffffffu) atomicAdd(&G[v.y], 1u);
		if (v.z != 0xffffffffu) atomicAdd(&G[v.z], 1u);
		if (v.w != 0xffffffffu) atomicAdd(&G[v.w], 1u);
	} else
		for (uint64_t i = i0; i < n; ++i) { uint32_t const r = list[i]; if (r != 0xffffffffu) atomicAdd(&G[r], 1u); }
}

// ------------------------------------------------------------------------------------------
// K6: merge by the gap array (Appendix A.4): exclusive scan of G, then for every k the G[k]
// symbols of L_R followed by L_A[k]; G is left holding its inclusive prefix sums.
// ------------------------------------------------------------------------------------------
constexpr uint32_t BIG_GAP = 512;
struct BigGap { uint32_t src, dst, len; };

__global__ void __launch_bounds__(256)
k_copy_big(const BigGap * __restrict__ list, const uint32_t * __restrict__ count, const uint8_t * __restrict__ LR, uint8_t * __restrict__ M) {
	uint32_t const n = *count;
	for (uint32_t e = blockIdx.y; e < n; e += gridDim.y) {
		BigGap const g = list[e];
		for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.len; i += (uint64_t)gridDim.x * blockDim.x)
			M[(uint64_t)g.dst + i] = LR[(uint64_t)g.src + i];
	}
}

__global__ void __launch_bounds__(256)
k_merge_samples(uint32_t * __restrict__ prerank, uint64_t qA0, uint64_t qA1, uint64_t qR1, const uint32_t * __restrict__ Sincl,
                const uint32_t * __restrict__ rsamp) {
	uint64_t const q = qA0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= qR1) return;
	uint32_t const k = prerank[q];
	prerank[q] = q < qA1 ? k + Sincl[k] : k + rsamp[q];
}

// ------------------------------------------------------------------------------------------
struct EventAccum {
	Stream & st;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
	explicit EventAccum(Stream & s) : st(s) {}
	~EventAccum() { for (auto & p : ev) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } }
	void begin() { cudaEvent_t a, b; B3M_CUDA(cudaEventCreate(&a)); B3M_CUDA(cudaEventCreate(&b)); B3M_CUDA(cudaEventRecord(a, st.s)); ev.push_back({a, b}); }
	void end() { B3M_CUDA(cudaEventRecord(ev.back().second, st.s)); }
	float total() { float t = 0; for (auto & p : ev) { float x = 0; cudaEventElapsedTime(&x, p.first, p.second); t += x; } return t; }
};

static TextRef text_ref(DevText const & T) { return TextRef{T.codes, T.ntext, T.n, T.has_term, T.keybits == 2 ? T.packed : nullptr}; }

static unsigned ilog2u(uint64_t v) { unsigned s = 0; while ((1ull << s) < v) ++s; return s; }

uint32_t Engine::fetch_special(int slot) {
	B3M_CUDA(cudaMemcpyAsync(pinned, d_special.get() + slot, 4, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	return *(uint32_t *)pinned;
}

// A4 + A5: one leaf.  Sorts the block's suffixes with lcpnext+1 look-ahead symbols and emits
// the block BWT into L (m codes), the gt bits and the anchors of the block.
void Engine::leaf_build(BlockLeaf & leaf, uint64_t s, uint64_t m, uint8_t * L, uint32_t * term_pos, SortStats * ss) {
	TextRef const t = text_ref(T);
	uint64_t const e = s + m;
	bool const has_termsuffix = T.has_term && e == T.n;     // the block holds the terminator suffix
	uint64_t const mt = has_termsuffix ? m - 1 : m;         // suffixes that start with a stored symbol
	uint32_t const shift = has_termsuffix ? 1 : 0;
	uint64_t W = mt;
	if (!has_termsuffix) {
		// bounded first, exact only when the bound was hit (large-LCP escape, largelcpthres)
		uint64_t const ep = e == T.n ? 0 : e;
		uint64_t const cap = params.largelcpthres ? params.largelcpthres : 16384;
		DevBuf<unsigned long long> dl(st, 1);
		uint64_t lcpnext = 0;
		for (int pass = 0; pass < 2; ++pass) {
			B3M_CUDA(cudaMemsetAsync(dl.get(), 0, 8, st.s));
			B3M_LAUNCH(st, k_lcpnext, (unsigned)div_up(m, 256), 256, 0, t, s, e, ep, pass ? T.n : cap, pass ? cap : 0, dl.get());
			B3M_CUDA(cudaMemcpyAsync(pinned, dl.get(), 8, cudaMemcpyDeviceToHost, st.s));
			B3M_CUDA(cudaStreamSynchronize(st.s));
			lcpnext = std::max<uint64_t>(lcpnext, *(unsigned long long *)pinned);
			if (lcpnext < cap) break;
			++large_lcp_blocks;
		}
		max_lcpnext = std::max(max_lcpnext, lcpnext);
		W = m + lcpnext + 1;
		if (T.has_term) W = std::min(W, T.ntext - s); // the terminator ends every comparison
	}
	leaf.s = s; leaf.mt = mt;
	if (mt) {
		{
			DevBuf<uint32_t> wsa;
			// no rank by position is needed: gt bits come from the text, anchors from the sorted order
			k2_suffix_sort(st, T, s, W, 0, T.has_term ? 0 : 1, wsa, nullptr, ss, nullptr);
			leaf.sa.alloc(st, mt);
			// keep the block's own suffixes, in order
			const uint32_t * sa = wsa.get();
			uint32_t * bsa = leaf.sa.get();
			uint64_t const mtl = mt;
			if (W == mt) {
				B3M_CUDA(cudaMemcpyAsync(bsa, sa, 4 * mt, cudaMemcpyDeviceToDevice, st.s));
			} else {
				scan_apply<OpSum>(st, W,
					[=] __device__(uint64_t k) -> uint32_t { return sa[k] < mtl ? 1u : 0u; },
					[=] __device__(uint64_t k, uint32_t excl, uint32_t v) { if (v) bsa[excl] = sa[k]; });
			}
		}
		B3M_LAUNCH(st, k_leaf_emit, (unsigned)div_up(mt, 256), 256, 0, T.codes, s, (const uint32_t *)leaf.sa.get(), mt, shift, prerate - 1, ilog2u(prerate),
		           L, d_special.get(), prep);
		B3M_LAUNCH(st, k_leaf_gt, (unsigned)div_up(mt, 256), 256, 0, t, s, mt, gtp);
		extract_bytes += mt * (4 + 32 + 1 + 4 + 1);
	}
	if (has_termsuffix) {
		// rank 0 is the terminator suffix; its predecessor is the last base (or it is the block start)
		uint32_t const zero = 0;
		if (mt) B3M_CUDA(cudaMemcpyAsync(L, T.codes + T.ntext - 1, 1, cudaMemcpyDeviceToDevice, st.s));
		else {
			B3M_CUDA(cudaMemsetAsync(L, 0, 1, st.s));
			B3M_CUDA(cudaMemsetAsync(d_special.get(), 0, 4, st.s));
		}
		B3M_CUDA(cudaMemsetAsync(gtp + T.ntext, 0, 1, st.s));
		if ((T.ntext & (prerate - 1)) == 0) B3M_CUDA(cudaMemcpyAsync(prep + T.ntext / prerate, &zero, 4, cudaMemcpyHostToDevice, st.s));
	}
	*term_pos = fetch_special(0);
	if (!has_termsuffix) leaf.keep = true; else leaf.sa.release(); // the last block is never a left part
}

// ---- primitives of one merge (shared by the single-GPU tree and the multi-GPU driver) -------
// dictionary over L_A (the placeholder of A's block-start row excluded) and C_A over A's text
void Engine::gap_prepare(GapCtx & ctx, const uint8_t * LA, uint64_t a0, uint64_t na, uint32_t termA) {
	int const flavour = T.sigma <= 4 ? 2 : 8;
	ctx.na = na;
	ctx.lines.alloc(st, dict_bytes(flavour, na, T.sigma));
	k4_build_dict(st, LA, na, flavour, T.sigma, ctx.lines.get());
	DictView & D = ctx.D;
	D.base = ctx.lines.get(); D.flavour = (uint32_t)flavour; D.spad = d8_spad(T.sigma);
	D.stride = flavour == 2 ? 64u : 4u * D.spad + D8_SYMS;
	D.exc_pos = termA; D.exc_code = 0;
	DevBuf<unsigned long long> dh(st, 256);
	B3M_CUDA(cudaMemsetAsync(dh.get(), 0, 256 * 8, st.s));
	unsigned const grid = (unsigned)std::min<uint64_t>(div_up(na, 256 * 64), (uint64_t)st.sms * 8);
	B3M_LAUNCH(st, k_hist_range, grid ? grid : 1, 256, 0, T.codes + a0, na, dh.get());
	B3M_CUDA(cudaMemcpyAsync(pinned, dh.get(), 256 * 8, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	uint64_t acc = 0;
	for (int c = 0; c < 257; ++c) { ctx.C.c[c] = (uint32_t)acc; if (c < 256) acc += ((unsigned long long *)pinned)[c]; }
}

// R's text is split into nch chains of chl positions (the same rule on every rank)
void Engine::chain_geometry(uint64_t nr, uint64_t * chl, uint64_t * nch) const {
	uint64_t c = std::min<uint64_t>(std::max<uint64_t>(nr / 64, 1), (uint64_t)st.sms * 2048 * 2);
	*chl = div_up(nr, c);
	*nch = div_up(nr, *chl);
}

// r0[c] += number of suffixes of this engine's kept leaves inside [a0,a1) smaller than the
// start point of chain c
void Engine::zranks_add(std::vector<BlockLeaf> & lv, uint64_t a0, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, uint32_t * r0) {
	TextRef const t = text_ref(T);
	for (auto & lf : lv)
		if (lf.keep && lf.s >= a0 && lf.s < a1 && lf.mt)
			B3M_LAUNCH(st, k_zrank, (unsigned)div_up(nch, 128), 128, 0, t, (const uint32_t *)lf.sa.get(), lf.mt, lf.s, a1, chl, r1, nch, r0);
}

// K5 for the chains [c_lo,c_hi): G += gaps, gtnew[p-a1], rsamp[p/rate]
void Engine::gap_run(GapCtx & ctx, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, uint64_t c_lo, uint64_t c_hi, const uint32_t * r0,
                     const uint8_t * gt_in, uint8_t * gtnew, uint32_t termA, uint32_t * G, uint32_t * rs) {
	TextRef const t = text_ref(T);
	if (c_hi > nch) c_hi = nch;
	if (c_lo >= c_hi) return;
	if (!(T.has_term && r1 == T.n)) B3M_LAUNCH(st, k_gt_top, 1, 1, 0, t, r1, a1, d_special.get());
	uint64_t const steps = std::min(a1 + c_hi * chl, r1) - (a1 + c_lo * chl);
	// Counting in G straight from the chains is a random read-modify-write of one 32-byte sector per LF step; once G
	// outgrows the L2 that is what bounds K5 (16 G steps/s against 59 G of the same walk without counters, K7).  Then the
	// chains write their ranks as a list (coalesced, 4 B per step), one radix pass partitions the list by the top 8 bits
	// of the rank, and k_gap_hist counts partition after partition with its counters resident in L2.
	uint64_t const gbytes = 4 * (ctx.na + 1);
	bool const use_list = params.gapmode == B3M_GAP_LIST || (params.gapmode != B3M_GAP_ATOMIC && gbytes > (48ull << 20));
	if (!use_list) {
		B3M_LAUNCH_T(st, "gap_chains", steps * 128ull, k_gap, (unsigned)div_up(c_hi - c_lo, 256), 256, 0, ctx.D, ctx.C, t, a1, r1, chl, c_lo, c_hi, r0,
		             gt_in, gtnew, (const uint32_t *)d_special.get(), termA, prerate - 1, ilog2u(prerate), G, rs, (uint32_t *)nullptr, (uint64_t)0);
	} else {
		uint64_t const nl = c_hi - c_lo, len = chl * nl;
		B3M_REQUIRE(len < 0xFFFFFF00ull, "internal: gap list too long");
		int const bw = (int)ceil_log2_u64(ctx.na + 2);
		bool const part = bw > 8 && (gbytes > (32ull << 20) || params.gapmode == B3M_GAP_LIST);
		DevBuf<uint32_t> rl(st, len + 4), rl2;
		if (part) rl2.alloc(st, len + 4);
		B3M_CUDA(cudaMemsetAsync(rl.get(), 0xff, 4 * (len + 4), st.s)); // the last chain may be shorter than the others
		B3M_LAUNCH_T(st, "gap_chains", steps * 68ull, k_gap, (unsigned)div_up(nl, 256), 256, 0, ctx.D, ctx.C, t, a1, r1, chl, c_lo, c_hi, r0,
		             gt_in, gtnew, (const uint32_t *)d_special.get(), termA, prerate - 1, ilog2u(prerate), G, rs, rl.get(), nl);
		const uint32_t * list = rl.get();
		if (part) {
			RadixRec<1> cur{{rl.get()}}, alt{{rl2.get()}};
			radix_sort_bits<1>(st, cur, alt, 0, len, bw - 8, bw, nullptr);
			list = cur.a[0];
		}
		B3M_LAUNCH_T(st, "gap_hist", len * 4ull + steps * 8ull, k_gap_hist, (unsigned)div_up(div_up(len, 4), 256), 256, 0, list, len, G);
		B3M_CUDA(cudaStreamSynchronize(st.s)); // rl / rl2 are released here
	}
	gap_lf_steps += steps; gap_chains += c_hi - c_lo;
}

// K6: LM = merge of LA and LR by G; G is left holding its inclusive prefix sums
void Engine::merge_run(const uint8_t * LA, uint64_t na, uint32_t termA, uint8_t * LR, uint64_t nr, uint32_t termR, uint64_t a1, uint32_t * G,
                       uint8_t * LM, uint32_t * termM) {
	// the stale placeholder of R becomes the true seam symbol T[a1-1] (Appendix A.4)
	B3M_CUDA(cudaMemcpyAsync(LR + termR, T.codes + a1 - 1, 1, cudaMemcpyDeviceToDevice, st.s));
	DevBuf<BigGap> big(st, nr / BIG_GAP + 2);
	uint32_t * bigcount = d_special.get() + 3;
	B3M_CUDA(cudaMemsetAsync(bigcount, 0, 4, st.s));
	{
		uint32_t * g = G;
		const uint8_t * LRc = LR;
		BigGap * bl = big.get();
		uint64_t const nal = na;
		scan_apply<OpSum>(st, na + 1,
			[=] __device__(uint64_t k) -> uint32_t { return g[k]; },
			[=] __device__(uint64_t k, uint32_t excl, uint32_t v) {
				uint64_t const o = k + excl;
				if (v <= BIG_GAP) { for (uint32_t x = 0; x < v; ++x) LM[o + x] = LRc[excl + x]; }
				else { uint32_t const e = atomicAdd(bigcount, 1u); bl[e] = BigGap{excl, (uint32_t)o, v}; }
				if (k < nal) LM[o + v] = LA[k];
				g[k] = excl + v;
			}, "merge_scatter", 8ull * (na + 1) + 2ull * (na + nr));
		B3M_LAUNCH(st, k_copy_big, dim3((unsigned)st.sms, 64), 256, 0, (const BigGap *)bl, (const uint32_t *)bigcount, LRc, LM);
	}
	B3M_CUDA(cudaMemcpyAsync(pinned, G + termA, 4, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	*termM = termA + *(uint32_t *)pinned;
	merge_bytes += 8ull * (na + 1) + 2ull * (na + nr);
}

// anchors move by the same rank maps (Appendix A.4)
void Engine::merge_samples(uint64_t a0, uint64_t a1, uint64_t r1, uint32_t * pre, const uint32_t * Sincl, const uint32_t * rs) {
	uint64_t const qA0 = div_up(a0, prerate), qA1 = div_up(a1, prerate), qR1 = div_up(r1, prerate);
	if (qR1 > qA0) B3M_LAUNCH(st, k_merge_samples, (unsigned)div_up(qR1 - qA0, 256), 256, 0, pre, qA0, qA1, qR1, Sincl, rs);
}

// A7 + A8: merge node A = [a0,a1) (left) with R = [a1,r1) (right) into M.
void Engine::node_merge(BlockNode & A, BlockNode & R, std::vector<BlockLeaf> & leaves, BlockNode & M, EventAccum & tgap, EventAccum & tmerge) {
	uint64_t const a0 = A.a0, a1 = A.a1, r1 = R.a1;
	uint64_t const na = a1 - a0, nr = r1 - a1;
	B3M_REQUIRE(R.a0 == a1, "internal: merge of non-adjacent nodes");
	tgap.begin();
	DevBuf<uint32_t> G(st, na + 1);
	{
		GapCtx ctx;
		gap_prepare(ctx, A.L.get(), a0, na, A.term);
		uint64_t chl, nch;
		chain_geometry(nr, &chl, &nch);
		DevBuf<uint32_t> r0(st, nch);
		B3M_CUDA(cudaMemsetAsync(r0.get(), 0, 4 * nch, st.s));
		zranks_add(leaves, a0, a1, r1, chl, nch, r0.get());
		DevBuf<uint8_t> gtnew(st, nr);
		B3M_CUDA(cudaMemsetAsync(G.get(), 0, 4 * (na + 1), st.s));
		gap_run(ctx, a1, r1, chl, nch, 0, nch, r0.get(), gtp, gtnew.get(), A.term, G.get(), rsp);
		B3M_CUDA(cudaMemcpyAsync(gtp + a1, gtnew.get(), nr, cudaMemcpyDeviceToDevice, st.s));
	}
	tgap.end();
	tmerge.begin();
	M.a0 = a0; M.a1 = r1;
	M.L.alloc(st, na + nr + 16);
	merge_run(A.L.get(), na, A.term, R.L.get(), nr, R.term, a1, G.get(), M.L.get(), &M.term);
	merge_samples(a0, a1, r1, prep, G.get(), rsp);
	tmerge.end();
	A.L.release(); R.L.release();
}

void Engine::build_tree(std::vector<BlockLeaf> & leaves, uint64_t lo, uint64_t hi, uint64_t base, uint64_t end, uint64_t bs, BlockNode & out,
                        EventAccum & tsort, EventAccum & tgap, EventAccum & tmerge) {
	if (hi - lo == 1) {
		uint64_t const s = base + lo * bs, e = std::min(s + bs, end);
		out.a0 = s; out.a1 = e;
		out.L.alloc(st, e - s + 16);
		tsort.begin();
		leaf_build(leaves[lo], s, e - s, out.L.get(), &out.term, &sortstats);
		tsort.end();
		return;
	}
	uint64_t const mid = (lo + hi) / 2;
	BlockNode A, R;
	build_tree(leaves, lo, mid, base, end, bs, A, tsort, tgap, tmerge);
	build_tree(leaves, mid, hi, base, end, bs, R, tsort, tgap, tmerge);
	node_merge(A, R, leaves, out, tgap, tmerge);
}

void Engine::build_blocks(PhaseTimer & pt, uint32_t * exc_pos) {
	uint64_t const bs = div_up(T.n, numblocks);
	numblocks = div_up(T.n, bs);
	B3M_REQUIRE(numblocks >= 2, "internal: build_blocks needs at least two blocks");
	ensure_codes(); // the block kernels read one byte per symbol
	gt.alloc(st, T.n);
	rsamp.alloc(st, npre);
	gtp = gt.get(); prep = prerank.get(); rsp = rsamp.get();
	std::vector<BlockLeaf> leaves(numblocks);
	EventAccum tsort(st), tgap(st), tmerge(st);
	BlockNode root;
	sortstats = SortStats();
	build_tree(leaves, 0, numblocks, 0, T.n, bs, root, tsort, tgap, tmerge);
	leaves.clear();
	gt.release(); rsamp.release();
	gtp = nullptr; rsp = nullptr;
	// root: the remaining placeholder is the row of suffix 0 (Appendix A.4)
	if (T.has_term) *exc_pos = root.term;
	else {
		B3M_CUDA(cudaMemcpyAsync(root.L.get() + root.term, T.codes + T.n - 1, 1, cudaMemcpyDeviceToDevice, st.s));
		*exc_pos = 0xffffffffu;
	}
	bwt = std::move(root.L);
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_sort = tsort.total(); ms_gap = tgap.total(); ms_merge = tmerge.total(); ms_extract = 0;
	(void)pt;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU driver entry points (C ABI b3m_engine_blk_*): the same primitives on caller-owned
// device buffers, so that block BWTs, gap arrays, gt bits and anchors can travel between the
// ranks with NCCL.
// ------------------------------------------------------------------------------------------
void Engine::blk_begin(uint64_t preisarate, uint64_t largelcpthres, void * d_gt, void * d_prerank, void * d_rsamp) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no input loaded");
	B3M_REQUIRE(d_gt && d_prerank && d_rsamp, "null block buffers");
	ensure_codes(); // the block kernels read one byte per symbol
	reset_results();
	params = b3m_build_params();
	params.largelcpthres = largelcpthres ? largelcpthres : 16384;
	prerate = preisarate ? preisarate : choose_preisarate_pub(T.n);
	B3M_REQUIRE(prerate && !(prerate & (prerate - 1)), "preisarate must be a power of two");
	npre = div_up(T.n, prerate);
	gtp = (uint8_t *)d_gt; prep = (uint32_t *)d_prerank; rsp = (uint32_t *)d_rsamp;
	dist_leaves.clear();
	sortstats = SortStats(); walkstats = WalkStats();
	gap_lf_steps = gap_chains = merge_bytes = extract_bytes = 0; max_lcpnext = large_lcp_blocks = 0;
	ms_sort = ms_extract = ms_dict = ms_gap = ms_merge = ms_walk = ms_total = 0;
	d_special.alloc(st, 8);
	B3M_CUDA(cudaMemsetAsync(d_special.get(), 0xff, 16, st.s));
}

void Engine::blk_build_range(uint64_t a0, uint64_t a1, uint64_t nb, void * d_L_out, uint32_t * term_out) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(gtp, "blk_begin was not called");
	B3M_REQUIRE(a0 < a1 && a1 <= T.n, "bad text range");
	uint64_t const m = a1 - a0;
	if (nb < 1) nb = 1;
	if (nb > m) nb = m;
	uint64_t const bs = div_up(m, nb);
	nb = div_up(m, bs);
	arena.ensure_slab(1, (size_t)work_bytes(2, bs, false));
	EventAccum tsort(st), tgap(st), tmerge(st);
	size_t const first = dist_leaves.size();
	dist_leaves.resize(first + nb);
	// build_tree indexes leaves from 0: work on a temporary vector, then move
	std::vector<BlockLeaf> lv(nb);
	BlockNode root;
	build_tree(lv, 0, nb, a0, a1, bs, root, tsort, tgap, tmerge);
	for (uint64_t b = 0; b < nb; ++b) dist_leaves[first + b] = std::move(lv[b]);
	B3M_CUDA(cudaMemcpyAsync(d_L_out, root.L.get(), m, cudaMemcpyDeviceToDevice, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	*term_out = root.term;
	ms_sort += tsort.total(); ms_gap += tgap.total(); ms_merge += tmerge.total();
}

void Engine::blk_finish(const void * d_L_root, uint32_t term_root, uint64_t q_lo, uint64_t q_hi, uint64_t sarate, uint64_t isarate, int bwtonly, uint64_t nblocks) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(gtp, "blk_begin was not called");
	auto pow2 = [](uint64_t v) { return v && !(v & (v - 1)); };
	B3M_REQUIRE(pow2(sarate) && pow2(isarate), "sampling rates must be powers of two");
	params.sasamplingrate = sarate; params.isasamplingrate = isarate; params.bwtonly = bwtonly; params.preisarate = prerate;
	numblocks = nblocks;
	bwt.alloc(st, T.n + 16);
	B3M_CUDA(cudaMemcpyAsync(bwt.get(), d_L_root, T.n, cudaMemcpyDeviceToDevice, st.s));
	prerank.alloc(st, npre);
	B3M_CUDA(cudaMemcpyAsync(prerank.get(), prep, 4 * npre, cudaMemcpyDeviceToDevice, st.s));
	uint32_t exc_pos = 0xffffffffu;
	if (T.has_term) exc_pos = term_root;
	else B3M_CUDA(cudaMemcpyAsync(bwt.get() + term_root, T.codes + T.n - 1, 1, cudaMemcpyDeviceToDevice, st.s));
	root_exc_pos = exc_pos;
	PhaseTimer pt(st);
	pt.mark();
	make_dict(exc_pos, 0, 0);
	pt.mark();
	nsa = nisa = 0;
	if (!bwtonly) {
		nsa = div_up(T.n, sarate); nisa = div_up(T.n, isarate);
		sa.alloc(st, nsa); isa.alloc(st, nisa);
		B3M_CUDA(cudaMemsetAsync(sa.get(), 0xff, 8 * nsa, st.s));
		B3M_CUDA(cudaMemsetAsync(isa.get(), 0xff, 8 * nisa, st.s));
		if (q_hi > npre) q_hi = npre;
		if (q_lo < q_hi)
			k7_walk(st, D, prerank.get(), npre, prerate, T.n, sarate, isarate, sa.get(), isa.get(), &walkstats, q_lo, q_hi);
	}
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_dict = pt.ms(0, 1); ms_walk = pt.ms(1, 2);
	gtp = nullptr; rsp = nullptr; prep = nullptr;
	dist_leaves.clear();
	have_results = true;
}

} // namespace b3m

// ---- C ABI ----------------------------------------------------------------------------------
#define B3M_GUARD(h, ...)                                                    \
	if (!(h)) return 1;                                                      \
	try { __VA_ARGS__; (h)->err.clear(); return 0; }                                \
	catch (std::exception const & ex) { (h)->err = ex.what(); return 2; }    \
	catch (...) { (h)->err = "unknown error"; return 3; }

extern "C" {

int b3m_engine_default_preisarate(b3m_engine * h, int bwtonly, uint64_t * rate) {
	B3M_GUARD(h, { if (!rate) throw b3m::Error("null argument"); B3M_REQUIRE(h->e->loaded, "no input loaded"); *rate = bwtonly ? 64 : h->e->choose_preisarate_pub(h->e->T.n); });
}
int b3m_engine_blk_begin(b3m_engine * h, uint64_t preisarate, uint64_t largelcpthres, void * d_gt, void * d_prerank, void * d_rsamp) {
	B3M_GUARD(h, h->e->blk_begin(preisarate, largelcpthres, d_gt, d_prerank, d_rsamp));
}
int b3m_engine_blk_build_range(b3m_engine * h, uint64_t a0, uint64_t a1, uint64_t numblocks, void * d_L_out, uint32_t * term_out) {
	B3M_GUARD(h, { if (!d_L_out || !term_out) throw b3m::Error("null argument"); h->e->blk_build_range(a0, a1, numblocks, d_L_out, term_out); });
}
int b3m_engine_blk_chains(b3m_engine * h, uint64_t nr, uint64_t * chl, uint64_t * nch) {
	B3M_GUARD(h, { if (!chl || !nch) throw b3m::Error("null argument"); h->e->chain_geometry(nr, chl, nch); });
}
int b3m_engine_blk_zranks(b3m_engine * h, uint64_t a0, uint64_t a1, uint64_t r1, uint64_t chl, uint64_t nch, void * d_r0) {
	B3M_GUARD(h, { B3M_CUDA(cudaSetDevice(h->e->device)); h->e->zranks_add(h->e->dist_leaves, a0, a1, r1, chl, nch, (uint32_t *)d_r0); });
}
int b3m_engine_blk_gap(b3m_engine * h, const void * d_LA, uint64_t a0, uint64_t na, uint32_t termA, uint64_t r1, uint64_t chl, uint64_t nch,
                       uint64_t c_lo, uint64_t c_hi, const void * d_r0, void * d_gtnew, void * d_G) {
	B3M_GUARD(h, {
		b3m::Engine & e = *h->e;
		B3M_CUDA(cudaSetDevice(e.device));
		B3M_REQUIRE(e.gtp, "blk_begin was not called");
		b3m::EventAccum tg(e.st);
		tg.begin();
		b3m::GapCtx ctx;
		e.gap_prepare(ctx, (const uint8_t *)d_LA, a0, na, termA);
		e.gap_run(ctx, a0 + na, r1, chl, nch, c_lo, c_hi, (const uint32_t *)d_r0, e.gtp, (uint8_t *)d_gtnew, termA, (uint32_t *)d_G, e.rsp);
		tg.end();
		B3M_CUDA(cudaStreamSynchronize(e.st.s));
		e.ms_gap += tg.total();
	});
}
int b3m_engine_blk_merge(b3m_engine * h, const void * d_LA, uint64_t na, uint32_t termA, void * d_LR, uint64_t nr, uint32_t termR, uint64_t a1,
                         void * d_G, void * d_LM, uint32_t * termM) {
	B3M_GUARD(h, {
		b3m::Engine & e = *h->e;
		B3M_CUDA(cudaSetDevice(e.device));
		if (!termM) throw b3m::Error("null argument");
		b3m::EventAccum tm(e.st);
		tm.begin();
		e.merge_run((const uint8_t *)d_LA, na, termA, (uint8_t *)d_LR, nr, termR, a1, (uint32_t *)d_G, (uint8_t *)d_LM, termM);
		tm.end();
		B3M_CUDA(cudaStreamSynchronize(e.st.s));
		e.ms_merge += tm.total();
	});
}
int b3m_engine_blk_merge_samples(b3m_engine * h, uint64_t a0, uint64_t a1, uint64_t r1, const void * d_Sincl) {
	B3M_GUARD(h, { B3M_CUDA(cudaSetDevice(h->e->device)); B3M_REQUIRE(h->e->gtp, "blk_begin was not called"); h->e->merge_samples(a0, a1, r1, h->e->prep, (const uint32_t *)d_Sincl, h->e->rsp); });
}
int b3m_engine_blk_finish(b3m_engine * h, const void * d_L_root, uint32_t term_root, uint64_t q_lo, uint64_t q_hi, uint64_t sasamplingrate,
                          uint64_t isasamplingrate, int bwtonly, uint64_t numblocks) {
	B3M_GUARD(h, { if (!d_L_root) throw b3m::Error("null argument"); h->e->blk_finish(d_L_root, term_root, q_lo, q_hi, sasamplingrate, isasamplingrate, bwtonly, numblocks); });
}

} // extern "C"
