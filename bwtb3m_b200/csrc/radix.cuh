// LSD radix sort on records held as NA parallel uint32 arrays (structure of arrays).
// One pass = upsweep (per-tile digit histogram) -> exclusive scan -> downsweep (stable rank in
// shared memory with warp match/ballot, reorder through shared memory, coalesced scatter).
// No CTA waits on another CTA, so a pass cannot hang.
//
// Algorithmic HBM bytes per pass and record: 4 (upsweep key read)
// + 4*NA (downsweep read) + 4*NA (downsweep write).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace b3m {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;
constexpr int RADIX_ITEMS = 16;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS; // 4096 records
constexpr int RADIX_BINS = 256;

template <int NA>
struct RadixRec {
	uint32_t * a[NA];
};

// warp-aggregated shared-memory histogram update; returns rank of this lane among equal digits
// that precede it in the warp plus the running count before this round.
__device__ __forceinline__ uint32_t warp_rank_digit(uint32_t * wcnt, uint32_t d, bool valid) {
	unsigned const peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
	uint32_t before = 0;
	if (valid) before = wcnt[d];
	__syncwarp();
	unsigned const lt = lanemask_lt();
	if (valid && (peers & lt) == 0) wcnt[d] = before + __popc(peers);
	__syncwarp();
	return before + __popc(peers & lt);
}

__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_upsweep(const uint32_t * __restrict__ key, uint64_t n, int shift, uint32_t mask,
                uint32_t * __restrict__ counts, uint32_t ntiles) {
	__shared__ uint32_t wcnt[RADIX_WARPS][RADIX_BINS];
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS; i += RADIX_THREADS) (&wcnt[0][0])[i] = 0;
	__syncthreads();
	uint64_t const chunk = (uint64_t)blockIdx.x * RADIX_TILE + (uint64_t)w * (32 * RADIX_ITEMS);
	uint32_t k[RADIX_ITEMS];
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		k[j] = (i < n) ? key[i] : 0u;
	}
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		warp_rank_digit(wcnt[w], (k[j] >> shift) & mask, i < n);
	}
	__syncthreads();
	for (int d = threadIdx.x; d < RADIX_BINS; d += RADIX_THREADS) {
		uint32_t s = 0;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) s += wcnt[ww][d];
		counts[(uint64_t)d * ntiles + blockIdx.x] = s;
	}
}

template <int NA>
__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_downsweep(RadixRec<NA> in, RadixRec<NA> out, int ka, uint64_t n, int shift, uint32_t mask,
                  const uint32_t * __restrict__ offsets, uint32_t ntiles) {
	__shared__ uint32_t wcnt[RADIX_WARPS][RADIX_BINS];
	__shared__ uint32_t gbase[RADIX_BINS];
	__shared__ uint32_t stage[RADIX_TILE];
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS; i += RADIX_THREADS) (&wcnt[0][0])[i] = 0;
	__syncthreads();
	uint64_t const tbase = (uint64_t)blockIdx.x * RADIX_TILE;
	uint64_t const chunk = tbase + (uint64_t)w * (32 * RADIX_ITEMS);
	uint32_t const nvalid = (n - tbase) < (uint64_t)RADIX_TILE ? (uint32_t)(n - tbase) : (uint32_t)RADIX_TILE;

	uint32_t k[RADIX_ITEMS];
	uint32_t slot[RADIX_ITEMS];
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		k[j] = (i < n) ? in.a[ka][i] : 0xffffffffu;
	}
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		// records past n take the last bin; being last in tile order they rank after every real record
		uint32_t const d = (i < n) ? ((k[j] >> shift) & mask) : (RADIX_BINS - 1);
		slot[j] = warp_rank_digit(wcnt[w], d, true);
	}
	__syncthreads();
	// per digit: exclusive scan over warps, then over digits
	{
		uint32_t const d = threadIdx.x; // RADIX_THREADS == RADIX_BINS
		uint32_t s = 0;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) { uint32_t const t = wcnt[ww][d]; wcnt[ww][d] = s; s += t; }
		uint32_t total;
		uint32_t const incl = block_scan_inclusive<OpSum>(s, &total);
		uint32_t const dstart = incl - s;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) wcnt[ww][d] += dstart;
		gbase[d] = offsets[(uint64_t)d * ntiles + blockIdx.x] - dstart;
	}
	__syncthreads();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		uint32_t const d = (i < n) ? ((k[j] >> shift) & mask) : (RADIX_BINS - 1);
		slot[j] += wcnt[w][d];
		stage[slot[j]] = k[j];
	}
	__syncthreads();
	uint32_t gpos[RADIX_ITEMS];
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint32_t const s = j * RADIX_THREADS + threadIdx.x;
		if (s < nvalid) {
			uint32_t const kk = stage[s];
			gpos[j] = gbase[(kk >> shift) & mask] + s;
			out.a[ka][gpos[j]] = kk;
		}
	}
	#pragma unroll
	for (int a = 0; a < NA; ++a) {
		if (a == ka) continue;
		__syncthreads();
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			if (i < n) stage[slot[j]] = in.a[a][i];
		}
		__syncthreads();
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint32_t const s = j * RADIX_THREADS + threadIdx.x;
			if (s < nvalid) out.a[a][gpos[j]] = stage[s];
		}
	}
}

struct RadixStats {
	uint64_t passes = 0;
	uint64_t bytes = 0; // algorithmic bytes moved by all passes
};

// Sorts records by bits [bit_lo, bit_hi) of array ka (stable).  `cur` and `alt` are ping-pong
// buffers; on return *swapped tells whether the result lives in alt.
template <int NA>
void radix_sort_bits(Stream & st, RadixRec<NA> & cur, RadixRec<NA> & alt, int ka, uint64_t n,
                     int bit_lo, int bit_hi, RadixStats * rs) {
	if (n == 0 || bit_hi <= bit_lo) return;
	uint32_t const ntiles = (uint32_t)div_up(n, RADIX_TILE);
	DevBuf<uint32_t> counts(st, (size_t)ntiles * RADIX_BINS);
	for (int shift = bit_lo; shift < bit_hi; shift += 8) {
		int const bits = (bit_hi - shift) < 8 ? (bit_hi - shift) : 8;
		uint32_t const mask = (1u << bits) - 1u;
		B3M_LAUNCH(st, k_radix_upsweep, ntiles, RADIX_THREADS, 0, cur.a[ka], n, shift, mask, counts.get(), ntiles);
		scan_exclusive_inplace<OpSum>(st, counts.get(), (uint64_t)ntiles * RADIX_BINS);
		B3M_LAUNCH(st, (k_radix_downsweep<NA>), ntiles, RADIX_THREADS, 0, cur, alt, ka, n, shift, mask,
		           (const uint32_t *)counts.get(), ntiles);
		RadixRec<NA> t = cur; cur = alt; alt = t;
		if (rs) { rs->passes++; rs->bytes += n * (4ull + 8ull * NA); }
	}
}

} // namespace b3m
