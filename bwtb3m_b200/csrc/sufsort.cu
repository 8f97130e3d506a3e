// K2: per-block suffix sorting by prefix doubling on top of the LSD radix sort (radix.cuh).
// Replaces libmaus2's CPU block sorter reached through BwtMergeBlockSortRequest::dispatch
// (/root/reference/src/checkbwt.cpp:24); comparisons are circular exactly as in
// the reference's definition BWT[i] = s[(SA[i]+n-1)%n] (/root/reference/src/lcpbit.cpp:3668-3669).
//
// Round 0 sorts all W suffixes by a 32-bit key holding their first k0 = 32/keybits symbols.
// Round r >= 1 touches only suffixes whose group is still tied: it gathers the rank of the
// suffix h symbols ahead, radix sorts (group, rank-ahead) and splits the groups; h doubles.
#include "kernels.h"
#include "radix.cuh"
#include "scan.cuh"
#include <time.h>
#include <stdlib.h>

namespace b3m {

struct TextView {
	const uint8_t * codes;
	uint64_t ntext;
	uint64_t wstart;
	uint64_t W;
	int circular;    // window == whole text, indices wrap modulo W
	int text_wraps;  // text positions wrap modulo ntext (linear window over a circular text)
};

__device__ __forceinline__ uint32_t tv_symbol(TextView const & v, uint64_t i /* window index, may exceed W */) {
	if (v.circular) {
		if (i >= v.W) i -= v.W;
		return v.codes[i];
	}
	if (i >= v.W) return 0u; // past the window: padding (order fixed by the short-suffix rule)
	uint64_t p = v.wstart + i;
	if (v.text_wraps) p %= v.ntext;
	return v.codes[p];
}

// Input order of round 0: in linear mode the nshort suffixes that run past the window end come
// first, shortest first, so that the stable sort leaves them in front of equal padded keys.
__global__ void __launch_bounds__(256)
k_make_keys(TextView v, unsigned bits, unsigned k0, uint64_t nshort, uint32_t * __restrict__ key, uint32_t * __restrict__ idx) {
	uint64_t const t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= v.W) return;
	uint64_t const i = (t < nshort) ? (v.W - 1 - t) : (t - nshort);
	uint32_t k = 0;
	for (unsigned s = 0; s < k0; ++s) k = (k << bits) | tv_symbol(v, i + s);
	key[t] = k;
	idx[t] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
k_gather_ahead(const uint32_t * __restrict__ aidx, uint64_t na, const uint32_t * __restrict__ rank, uint64_t h,
               uint64_t W, int circular, uint32_t * __restrict__ key2) {
	uint64_t const a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (a >= na) return;
	uint64_t j = (uint64_t)aidx[a] + h;
	uint32_t k;
	if (circular) { j %= W; k = rank[j]; }
	else k = (j < W) ? rank[j] + 1u : 0u;
	key2[a] = k;
}

static double wall_ms() {
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static bool trace_on() { static int t = -1; if (t < 0) t = getenv("B3M_TRACE") ? 1 : 0; return t == 1; }
#define TRACE(msg) do { if (trace_on()) { cudaStreamSynchronize(st.s); double t_ = wall_ms(); fprintf(stderr, "[T] %-28s %9.3f ms\n", msg, t_ - t_last); t_last = wall_ms(); } } while (0)

static uint32_t fetch_u32(Stream & st, const uint32_t * d) {
	uint32_t h = 0;
	B3M_CUDA(cudaMemcpyAsync(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	return h;
}

void k2_suffix_sort(Stream & st, DevText const & T, uint64_t wstart, uint64_t W, int circular, int text_wraps,
                    uint32_t * sa, uint32_t * rank, SortStats * stats) {
	if (W == 0) return;
	B3M_REQUIRE(W < 0xFFFFFFF0ull, "window too large for 32-bit suffix indices");
	unsigned const bits = T.keybits;
	unsigned const k0 = 32 / bits;
	uint64_t const nshort = circular ? 0 : (W < (uint64_t)(k0 - 1) ? W : (uint64_t)(k0 - 1));
	TextView v{T.codes, T.ntext, wstart, W, circular, text_wraps};
	SortStats S;
	double t_last = wall_ms();

	DevBuf<uint32_t> scalar(st, 4);
	uint32_t * d_total = scalar.get();
	uint64_t na = 0;
	DevBuf<uint32_t> pool[6];
	{
		// ---------------- round 0 ----------------
		DevBuf<uint32_t> key0(st, W), key1(st, W), idx1(st, W);
		// sa doubles as the first index buffer
		RadixRec<2> cur{{key0.get(), sa}}, alt{{key1.get(), idx1.get()}};
		unsigned const grid = (unsigned)div_up(W, 256);
		TRACE("r0 alloc");
		B3M_LAUNCH_T(st, "make_keys", W * 9ull, k_make_keys, grid, 256, 0, v, bits, k0, nshort, cur.a[0], cur.a[1]);
		S.other_bytes += W * (1ull + 8ull);
		RadixStats rs;
		TRACE("r0 make_keys");
		radix_sort_bits<2>(st, cur, alt, 0, W, 0, 32, &rs);
		TRACE("r0 radix");
		S.radix_passes += rs.passes; S.radix_bytes += rs.bytes; S.active_sum += W; S.rounds = 1;
		const uint32_t * skey = cur.a[0];
		const uint32_t * sidx = cur.a[1];
		uint32_t * grp = alt.a[0];
		uint64_t const Wm = W, k0m = k0;
		int const lin = !circular;
		// head flags -> group head position (max-scan); rank scatter; final place of every index
		scan_apply<OpMax>(st, W,
			[=] __device__(uint64_t k) -> uint32_t {
				if (k == 0) return 0u;
				bool head = skey[k] != skey[k - 1];
				if (lin) head = head || ((uint64_t)sidx[k] + k0m > Wm) || ((uint64_t)sidx[k - 1] + k0m > Wm);
				return head ? (uint32_t)k : 0u;
			},
			[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) {
				uint32_t const head = excl > v0 ? excl : v0;
				grp[k] = head;
				uint32_t const i = sidx[k];
				rank[i] = head;
				if (sa != sidx) sa[k] = i;
			}, "heads_rank_scatter", W * 48ull);
		S.other_bytes += W * (8ull + 8ull + 4ull + 4ull + 4ull);
		TRACE("r0 heads+rank scatter");
		// compaction of suffixes whose group has more than one member
		B3M_CUDA(cudaMemsetAsync(d_total, 0, 4, st.s));
		// two passes: count, then allocate and fill
		auto active = [=] __device__(uint64_t k) -> uint32_t {
			bool const hk = grp[k] == (uint32_t)k;
			bool const hn = (k + 1 == Wm) || (grp[k + 1] == (uint32_t)(k + 1));
			return (hk && hn) ? 0u : 1u;
		};
		scan_apply<OpSum>(st, W, active,
			[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) { if (k + 1 == Wm) *d_total = excl + v0; });
		na = fetch_u32(st, d_total);
		S.other_bytes += W * 8ull;
		TRACE("r0 count active");
		if (na) {
			for (int b = 0; b < 6; ++b) pool[b].alloc(st, na);
			uint32_t * agrp = pool[0].get();
			uint32_t * aidx = pool[1].get();
			const uint32_t * saf = sa;
			scan_apply<OpSum>(st, W, active,
				[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) {
					if (v0) { agrp[excl] = grp[k]; aidx[excl] = saf[k]; }
				});
			S.other_bytes += W * 8ull + na * 8ull;
		}
		TRACE("r0 compact");
	}
	TRACE("r0 free");

	// ---------------- doubling rounds ----------------
	uint32_t * bufs[6];
	for (int b = 0; b < 6; ++b) bufs[b] = pool[b].get();
	// roles: bufs[0]=grp, bufs[1]=idx, bufs[2]=key2, bufs[3..5]=ping-pong partners
	uint64_t h = k0;
	int const bw = (int)ceil_log2_u64(W + 2);
	while (na) {
		if (circular && h >= W) break; // non-primitive text: ties stay in current order (unpinned)
		unsigned const grid = (unsigned)div_up(na, 256);
		B3M_LAUNCH_T(st, "gather_ahead", na * 40ull, k_gather_ahead, grid, 256, 0, (const uint32_t *)bufs[1], na, (const uint32_t *)rank, h, W, circular, bufs[2]);
		S.other_bytes += na * (4ull + 32ull + 4ull);
		TRACE("rN gather");
		RadixRec<3> cur{{bufs[0], bufs[2], bufs[1]}}, alt{{bufs[3], bufs[4], bufs[5]}};
		RadixStats rs;
		radix_sort_bits<3>(st, cur, alt, 1, na, 0, bw, &rs); // rank ahead (minor key)
		radix_sort_bits<3>(st, cur, alt, 0, na, 0, bw, &rs); // group (major key)
		S.radix_passes += rs.passes; S.radix_bytes += rs.bytes; S.active_sum += na; S.rounds++;
		TRACE("rN radix");
		const uint32_t * sg = cur.a[0];
		const uint32_t * sk = cur.a[1];
		const uint32_t * si = cur.a[2];
		uint32_t * ngrp = alt.a[0];
		uint64_t const nam = na;
		scan_apply<OpMaxMax>(st, na,
			[=] __device__(uint64_t a) -> uint2 {
				if (a == 0) return make_uint2(0u, 0u);
				bool const seg = sg[a] != sg[a - 1];
				bool const head = seg || (sk[a] != sk[a - 1]);
				return make_uint2(seg ? (uint32_t)a : 0u, head ? (uint32_t)a : 0u);
			},
			[=] __device__(uint64_t a, uint2 excl, uint2 v0) {
				uint32_t const segstart = excl.x > v0.x ? excl.x : v0.x;
				uint32_t const newhead = excl.y > v0.y ? excl.y : v0.y;
				uint32_t const g = sg[a];
				uint32_t const i = si[a];
				uint32_t const ng = g + (newhead - segstart);
				sa[g + ((uint32_t)a - segstart)] = i;
				rank[i] = ng;
				ngrp[a] = ng;
			});
		S.other_bytes += na * (2 * 12ull + 4ull + 32ull + 4ull);
		auto active = [=] __device__(uint64_t a) -> uint32_t {
			bool const hk = (a == 0) || (ngrp[a] != ngrp[a - 1]);
			bool const hn = (a + 1 == nam) || (ngrp[a + 1] != ngrp[a]);
			return (hk && hn) ? 0u : 1u;
		};
		uint32_t * ogrp = alt.a[1];
		uint32_t * oidx = alt.a[2];
		scan_apply<OpSum>(st, na, active,
			[=] __device__(uint64_t a, uint32_t excl, uint32_t v0) {
				if (v0) { ogrp[excl] = ngrp[a]; oidx[excl] = si[a]; }
				if (a + 1 == nam) *d_total = excl + v0;
			});
		S.other_bytes += na * (2 * 8ull + 8ull);
		uint64_t const nn = fetch_u32(st, d_total);
		TRACE("rN split+compact");
		// next round: grp = alt[1], idx = alt[2]; everything else is free
		uint32_t * nb[6] = {alt.a[1], alt.a[2], cur.a[0], cur.a[1], cur.a[2], alt.a[0]};
		for (int b = 0; b < 6; ++b) bufs[b] = nb[b];
		na = nn;
		h *= 2;
	}
	if (stats) { // accumulated over the leaves of a multi-block build
		stats->rounds = stats->rounds > S.rounds ? stats->rounds : S.rounds;
		stats->radix_passes += S.radix_passes; stats->radix_bytes += S.radix_bytes;
		stats->active_sum += S.active_sum; stats->other_bytes += S.other_bytes;
	}
}

} // namespace b3m
