// K2, 2-bit alphabets, whole text in one window: MSD bucket sort of the suffixes in two global
// levels and one CTA-local finish.  Replaces libmaus2's CPU block sorter reached through
// BwtMergeBlockSortRequest::dispatch (/root/reference/src/checkbwt.cpp:24) on the
// one-block path; the order produced is the one of the reference's definition
// BWT[i] = s[(SA[i]+n-1)%n] (/root/reference/src/lcpbit.cpp:3668-3669).
//
// A record is ONE 64-bit word: [63:62] code preceding the suffix, [61:32] key30 = bits
// [b1, b1+30) of the suffix (2 bits per symbol), [31:0] suffix index.
//   count    k_msd_count     per tile of 8192 text positions: how many suffixes start with each value of
//                            their first b1 bits (read off the packed text; 16-bit counts)
//            k_msd_col*      exclusive scan of those counts down every column: where the records of tile t
//                            and bin b go inside bin b, and the size of every bin (no look-back chain, no
//                            spinning between CTAs)
//   level 1  k_msd_scatter   builds the records of a tile from the packed text (a thread takes 16 consecutive
//                            positions out of three 64-bit words), ranks them by their first b1 bits
//                            with shared-memory atomics (an MSD pass need not be stable) and writes runs per bin
//   level 2  k_msd_local     sorts every tile of a level-1 bucket by its next b2 bits IN PLACE (loads
//                            to registers, ranks with shared-memory atomics, one bulk async store
//                            shared -> global per tile) and records where each of the 2^b2 runs starts
//   totals   k_msd_subtotals size of every (b1+b2)-bit sub-bucket from the run tables; scan -> ranks
//   finish   k_msd_finish    one CTA per sub-bucket: gathers its runs from the tiles of the parent
//                            bucket (16-byte cp.async, no register staging), sorts them in shared memory
//                            (local digit by atomics, the few records per digit by comparison: rest of
//                            key30, then 32 more symbols read from the text, then remaining length), and
//                            emits BWT, anchors and sampled SA/ISA at the final ranks (or the order + head flags)
// Algorithmic HBM bytes per suffix: 0.25 + 0.25 (count) + 3 * 0.5 (column scan: 16-bit counts read
// twice, 32-bit offsets written) + 0.25 + 0.5 + 8 (level 1) + 8 + 8 (level 2) + 8 + 1.25 (finish)
// = 36.5, against 73.75 of the LSD path (radix.cuh) -- and a fraction of its instructions, which is
// what bounded that path.  Sub-buckets larger than MSD_CAP records and local digits shared by more
// than MSD_MAXRUN records are left as unresolved groups for the prefix-doubling rounds of sufsort.cu.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "textview.cuh"
#include "kernels.h"

namespace b3m {

constexpr int MSD_THREADS = 512;
constexpr int MSD_ITEMS = 16;
constexpr int MSD_TILE = MSD_THREADS * MSD_ITEMS;   // 8192 records per tile of both levels
constexpr int MSD_MAXBINS = 2048;                   // bins of a global level
constexpr int MSD_CAP = 8192;                       // records (incl. alignment padding) one finish CTA holds
constexpr int MSD_LBITS_MIN = 8;                    // local digit of the finish: 256 ..
constexpr int MSD_LBITS_MAX = 13;                   // .. 8192 bins, about two per record
constexpr int MSD_MAXRUN = 32;                      // records per local digit sorted by comparison
constexpr int MSD_CSLOTS = 256;
constexpr int MSD_COLCHUNK = 256;                   // tiles per CTA of the column scan
constexpr uint32_t MSD_KEYMASK = 0x3fffffffu;
constexpr unsigned long long MSD_PAD = ~0ull;       // not a record: indices stay below 2^32 - 256

// exclusive scan of cnt[0..nb) in place (nb <= PERMAX * THREADS); mine[] = the counts of this thread's
// bins b0 .. b0+per-1 (b0 = threadIdx.x * per); returns the total.  Ends with a barrier.
template <int THREADS, int PERMAX>
__device__ __forceinline__ uint32_t msd_scan_bins(uint32_t * cnt, unsigned nb, uint32_t * wsum, uint32_t (&mine)[PERMAX], unsigned & b0, unsigned & per) {
	unsigned const lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	per = (nb + THREADS - 1) / THREADS;
	b0 = threadIdx.x * per;
	uint32_t s = 0;
	#pragma unroll
	for (unsigned q = 0; q < (unsigned)PERMAX; ++q) { mine[q] = (q < per && b0 + q < nb) ? cnt[b0 + q] : 0u; s += mine[q]; }
	uint32_t incl = s;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	if (w == 0) {
		uint32_t x = lane < THREADS / 32 ? wsum[lane] : 0u;
		#pragma unroll
		for (int o = 1; o < THREADS / 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += t; }
		if (lane < THREADS / 32) wsum[lane] = x;
	}
	__syncthreads();
	uint32_t run = (w ? wsum[w - 1] : 0u) + incl - s;
	uint32_t const total = wsum[THREADS / 32 - 1];
	#pragma unroll
	for (unsigned q = 0; q < (unsigned)PERMAX; ++q) if (q < per && b0 + q < nb) { cnt[b0 + q] = run; run += mine[q]; }
	__syncthreads();
	return total;
}

// the same for 2^k bins, keeping nothing in registers: every thread scans nb / THREADS consecutive bins (four at a
// time with 128-bit accesses when it has that many; cnt is 16-byte aligned)
template <int THREADS>
__device__ __forceinline__ void msd_scan_bins_wide(uint32_t * cnt, unsigned nb, uint32_t * wsum) {
	unsigned const lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	unsigned const per = (nb + THREADS - 1) / THREADS, b0 = threadIdx.x * per;
	unsigned const b1 = b0 + per < nb ? b0 + per : nb;
	uint32_t s = 0;
	if ((per & 3u) == 0) {
		const uint4 * c4 = reinterpret_cast<const uint4 *>(cnt + b0);
		#pragma unroll 4
		for (unsigned q = 0; q < per / 4; ++q) { uint4 const x = c4[q]; s += x.x + x.y + x.z + x.w; }
	} else
		for (unsigned b = b0; b < b1; ++b) s += cnt[b];
	uint32_t incl = s;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	if (w == 0) {
		uint32_t x = lane < THREADS / 32 ? wsum[lane] : 0u;
		#pragma unroll
		for (int o = 1; o < THREADS / 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += t; }
		if (lane < THREADS / 32) wsum[lane] = x;
	}
	__syncthreads();
	uint32_t run = (w ? wsum[w - 1] : 0u) + incl - s;
	if ((per & 3u) == 0) {
		uint4 * c4 = reinterpret_cast<uint4 *>(cnt + b0);
		#pragma unroll 4
		for (unsigned q = 0; q < per / 4; ++q) {
			uint4 const x = c4[q];
			uint4 y;
			y.x = run; y.y = y.x + x.x; y.z = y.y + x.y; y.w = y.z + x.z; run = y.w + x.w;
			c4[q] = y;
		}
	} else
		for (unsigned b = b0; b < b1; ++b) { uint32_t const c = cnt[b]; cnt[b] = run; run += c; }
	__syncthreads();
}

// first 64 bits of the suffix at window index i (zero behind the end of a linear window, wrapping
// in a circular one) and the code in front of it
__device__ __forceinline__ void msd_record(TextView const & v, uint64_t i, unsigned b1, uint32_t & d, uint32_t & hi32) {
	uint64_t const sb = tv_symbols(v, i, 32, 2);
	d = (uint32_t)(sb >> (64u - b1));
	hi32 = (tv_pred(v, i) << 30) | ((uint32_t)(sb >> (34u - b1)) & MSD_KEYMASK);
}

// The MSD_ITEMS records of the positions pos0 .. pos0+15 (one thread): d[j] = first b1 bits (~0: no such
// position), hi32[j] = upper half of the record.  Fast path: three packed words hold the symbols from one
// before pos0 on; every field is a constant-distance bit field of that 128-bit window.
template <bool WANT_HI>
__device__ __forceinline__ void msd_records16(TextView const & v, unsigned b1, uint64_t pos0, uint32_t (&d)[MSD_ITEMS], uint32_t (&hi32)[MSD_ITEMS]) {
	bool const fast = pos0 >= 1 && (v.circular ? pos0 + MSD_ITEMS + 32 <= v.W : pos0 + MSD_ITEMS <= v.W);
	if (fast) {
		uint64_t const q = pos0 - 1; // wstart == 0: window index == text position
		const uint64_t * wp = v.packed + (q >> 5);
		unsigned const sh = (unsigned)(q & 31u) << 1;
		uint64_t const w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
		uint64_t const hi = sh ? ((w0 << sh) | (w1 >> (64u - sh))) : w0; // symbols q .. q+31
		uint64_t const lo = sh ? ((w1 << sh) | (w2 >> (64u - sh))) : w1; // symbols q+32 .. q+63
		#pragma unroll
		for (int j = 0; j < MSD_ITEMS; ++j) {
			uint64_t const sb = (hi << (2 * j + 2)) | (lo >> (62 - 2 * j)); // suffix bits from symbol j+1 on
			d[j] = (uint32_t)(sb >> (64u - b1));
			if (WANT_HI) hi32[j] = ((uint32_t)(hi >> (62 - 2 * j)) << 30) | ((uint32_t)(sb >> (34u - b1)) & MSD_KEYMASK);
		}
	} else {
		#pragma unroll
		for (int j = 0; j < MSD_ITEMS; ++j) {
			uint64_t const i = pos0 + j;
			uint32_t dd = 0xffffffffu, h = 0;
			if (i < v.W) msd_record(v, i, b1, dd, h);
			d[j] = dd;
			if (WANT_HI) hi32[j] = h;
		}
	}
}

// ---- histogram of the first b1 bits over the whole text (the plan of a sharded build) --------
__global__ void __launch_bounds__(256)
k_msd_hist(TextView v, unsigned b1, unsigned long long * __restrict__ ghist) {
	__shared__ uint32_t sh[MSD_MAXBINS];
	unsigned const nb = 1u << b1;
	for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	uint64_t const nchunks = div_up(v.W, MSD_ITEMS);
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nchunks; q += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t d[MSD_ITEMS], h[MSD_ITEMS];
		msd_records16<false>(v, b1, q * MSD_ITEMS, d, h);
		#pragma unroll
		for (int j = 0; j < MSD_ITEMS; ++j) if (d[j] != 0xffffffffu) atomicAdd(&sh[d[j]], 1u);
	}
	__syncthreads();
	for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) if (sh[i]) atomicAdd(&ghist[i], (unsigned long long)sh[i]);
}

// ---- per-tile counts of the kept bins -------------------------------------------------------
__global__ void __launch_bounds__(MSD_THREADS)
k_msd_count(TextView v, unsigned b1, uint32_t d_lo, uint32_t nkeep, uint32_t t_lo, uint16_t * __restrict__ tcount /* [tiles from t_lo][nkeep] */) {
	__shared__ uint32_t cnt[MSD_MAXBINS];
	for (unsigned i = threadIdx.x; i < nkeep; i += MSD_THREADS) cnt[i] = 0;
	__syncthreads();
	uint32_t d[MSD_ITEMS], h[MSD_ITEMS];
	msd_records16<false>(v, b1, (uint64_t)(blockIdx.x + t_lo) * MSD_TILE + (uint64_t)MSD_ITEMS * threadIdx.x, d, h);
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) { uint32_t const b = d[j] - d_lo; if (b < nkeep) atomicAdd(&cnt[b], 1u); }
	__syncthreads();
	uint16_t * const row = tcount + (uint64_t)blockIdx.x * nkeep;
	for (unsigned i = threadIdx.x; i < nkeep; i += MSD_THREADS) row[i] = (uint16_t)cnt[i]; // at most 8192 per tile and bin
}

// ---- column scan of the tile counts ------------------------------------------------------------
// partial[c][b] = sum of tcount[t][b] over the tiles t of chunk c
__global__ void __launch_bounds__(256)
k_msd_colsum(const uint16_t * __restrict__ tcount, uint32_t ntiles, uint32_t nkeep, uint32_t * __restrict__ partial) {
	uint32_t const t_lo = blockIdx.x * MSD_COLCHUNK, t_hi = t_lo + MSD_COLCHUNK < ntiles ? t_lo + MSD_COLCHUNK : ntiles;
	for (uint32_t b = threadIdx.x; b < nkeep; b += 256) {
		uint32_t s = 0;
		for (uint32_t t = t_lo; t < t_hi; ++t) s += tcount[(uint64_t)t * nkeep + b];
		partial[(uint64_t)blockIdx.x * nkeep + b] = s;
	}
}
// exclusive scan of partial down every column, in place; total[b] = size of bin b
__global__ void __launch_bounds__(256)
k_msd_colscan(uint32_t * __restrict__ partial, uint32_t nchunks, uint32_t nkeep, unsigned long long * __restrict__ total) {
	uint32_t const b = blockIdx.x * 256 + threadIdx.x;
	if (b >= nkeep) return;
	unsigned long long run = 0;
	for (uint32_t c = 0; c < nchunks; ++c) {
		uint32_t const x = partial[(uint64_t)c * nkeep + b];
		partial[(uint64_t)c * nkeep + b] = (uint32_t)run;
		run += x;
	}
	total[b] = run;
}
// toff[t][b] = number of records of bin b in the tiles before t
__global__ void __launch_bounds__(256)
k_msd_colapply(const uint16_t * __restrict__ tcount, uint32_t ntiles, uint32_t nkeep, const uint32_t * __restrict__ partial, uint32_t * __restrict__ toff) {
	uint32_t const t_lo = blockIdx.x * MSD_COLCHUNK, t_hi = t_lo + MSD_COLCHUNK < ntiles ? t_lo + MSD_COLCHUNK : ntiles;
	for (uint32_t b = threadIdx.x; b < nkeep; b += 256) {
		uint32_t run = partial[(uint64_t)blockIdx.x * nkeep + b];
		for (uint32_t t = t_lo; t < t_hi; ++t) {
			toff[(uint64_t)t * nkeep + b] = run;
			run += tcount[(uint64_t)t * nkeep + b];
		}
	}
}

// ---- level 1 ---------------------------------------------------------------------------------
constexpr int MSD_MAXPARTS = 16;
struct MsdP1 {
	TextView v;
	unsigned b1;
	uint32_t d_lo, nkeep;           // bins [d_lo, d_lo + nkeep) are kept (one key range of a bin-sharded build, or all)
	uint32_t t_lo;                  // first tile of the text this launch covers (position-sharded builds: one range of tiles per GPU)
	uint32_t nt;                    // tiles of the launch
	const uint32_t * base;          // [nkeep] where this launch's first record of kept bin b goes in its destination array
	const uint32_t * toff;          // [tiles of the launch][nkeep] records of bin b in the earlier tiles of the launch
	// destination arrays: bin b belongs to part p with bnd[p] <= b < bnd[p+1]; out[p] may be another GPU's memory
	// (CUDA IPC peer mapping): the records then cross NVLink as the stores of this kernel
	unsigned nparts;
	uint32_t bnd[MSD_MAXPARTS + 1];
	unsigned long long * out[MSD_MAXPARTS];
};

// SUB = 1: one tile per CTA of 512 threads (two CTAs per SM).  SUB = 2: two consecutive tiles per CTA of 1024 threads:
// the column scan puts the records of consecutive tiles next to each other inside a bin, so the CTA stores runs of
// twice the length -- what a position-sharded build wants, whose runs cross NVLink (A.nt = tiles of the launch).
template <int SUB>
__global__ void __launch_bounds__(MSD_THREADS * SUB, 2 / SUB)
k_msd_scatter(MsdP1 A) {
	constexpr int THREADS = MSD_THREADS * SUB;
	extern __shared__ __align__(16) uint8_t msd_dyn[];
	unsigned long long * const stage = reinterpret_cast<unsigned long long *>(msd_dyn); // SUB * MSD_TILE records
	__shared__ uint32_t cnt[MSD_MAXBINS];
	__shared__ uint32_t gdel[MSD_MAXBINS];
	__shared__ uint8_t gown[MSD_MAXBINS];
	__shared__ unsigned long long * s_out[MSD_MAXPARTS];
	__shared__ uint32_t wsum[THREADS / 32];
	unsigned const nkeep = A.nkeep;
	for (unsigned i = threadIdx.x; i < nkeep; i += THREADS) cnt[i] = 0;
	if (threadIdx.x < A.nparts) s_out[threadIdx.x] = A.out[threadIdx.x];
	__syncthreads();
	uint32_t const tile = blockIdx.x * SUB;                 // first tile of this CTA
	uint32_t const sub = threadIdx.x / MSD_THREADS, ltid = threadIdx.x % MSD_THREADS;
	uint64_t const t0 = (uint64_t)(tile + A.t_lo) * MSD_TILE;
	uint32_t const lpos = sub * (uint32_t)MSD_TILE + (uint32_t)MSD_ITEMS * ltid; // first position of this thread, relative to t0

	uint32_t hi32[MSD_ITEMS], dr[MSD_ITEMS]; // dr = (kept bin << 16) | rank inside the CTA's bin; ~0: not kept
	if (tile + sub < A.nt) msd_records16<true>(A.v, A.b1, t0 + lpos, dr, hi32);
	else {
		#pragma unroll
		for (int j = 0; j < MSD_ITEMS; ++j) { dr[j] = 0xffffffffu; hi32[j] = 0; }
	}
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) {
		uint32_t const d = dr[j] - A.d_lo;
		dr[j] = d < nkeep ? ((d << 16) | atomicAdd(&cnt[d], 1u)) : 0xffffffffu; // "no position" (~0) minus d_lo (< 2048) is no bin either
	}
	__syncthreads();
	constexpr int PERMAX = MSD_MAXBINS / THREADS;
	uint32_t mine[PERMAX];
	unsigned b0, per;
	uint32_t const nvalid = msd_scan_bins<THREADS, PERMAX>(cnt, nkeep, wsum, mine, b0, per);
	const uint32_t * const orow = A.toff + (uint64_t)tile * nkeep;
	#pragma unroll
	for (unsigned q = 0; q < (unsigned)PERMAX; ++q)
		if (q < per && b0 + q < nkeep) {
			unsigned const b = b0 + q;
			gdel[b] = __ldg(A.base + b) + __ldg(orow + b) - cnt[b];
			unsigned own = 0;
			for (unsigned p = 1; p < A.nparts; ++p) own += A.bnd[p] <= b ? 1u : 0u;
			gown[b] = (uint8_t)own;
		}
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) {
		if (dr[j] != 0xffffffffu) {
			uint32_t const d = dr[j] >> 16;
			uint32_t const slot = cnt[d] + (dr[j] & 0xffffu);
			stage[slot] = ((unsigned long long)hi32[j] << 32) | (d << 15) | (lpos + (uint32_t)j);
		}
	}
	__syncthreads();
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) {
		uint32_t const s = j * THREADS + threadIdx.x;
		if (s < nvalid) {
			unsigned long long const w = stage[s];
			uint32_t const lo = (uint32_t)w, d = lo >> 15;
			s_out[gown[d]][gdel[d] + s] = (w & 0xffffffff00000000ull) | (uint32_t)(t0 + (lo & 0x7fffu));
		}
	}
}

// ---- level 2 ---------------------------------------------------------------------------------
struct MsdP2 {
	unsigned b2, nkeep;
	const uint32_t * base;          // [nkeep + 1]
	const uint32_t * tpre;          // [nkeep + 1] first tile of kept bin d
	unsigned long long * recs;      // sorted in place
	uint16_t * table;               // bin d: (2^b2 + 1) rows of ntiles(d) columns at tpre[d] * (2^b2 + 1); [r][k] = start of run r in tile k
};

__global__ void __launch_bounds__(MSD_THREADS, 2)
k_msd_local(MsdP2 A) {
	extern __shared__ __align__(16) uint8_t msd_dyn[];
	unsigned long long * const stage = reinterpret_cast<unsigned long long *>(msd_dyn); // MSD_TILE + 2 words
	__shared__ uint32_t cnt[MSD_MAXBINS];
	__shared__ uint32_t wsum[MSD_THREADS / 32];
	__shared__ uint32_t s_bin;
	unsigned const nb2 = 1u << A.b2;
	// tile g of the grid -> (kept bin d, tile k of that bin): tpre[d] <= g < tpre[d+1]; every thread tests a few bins
	for (unsigned b = threadIdx.x; b < A.nkeep; b += MSD_THREADS)
		if (__ldg(A.tpre + b) <= blockIdx.x && blockIdx.x < __ldg(A.tpre + b + 1)) s_bin = b;
	for (unsigned i = threadIdx.x; i < nb2; i += MSD_THREADS) cnt[i] = 0;
	__syncthreads();
	unsigned const d = s_bin;
	uint32_t const tp = __ldg(A.tpre + d), ntp = __ldg(A.tpre + d + 1) - tp, k = blockIdx.x - tp;
	uint32_t const start = __ldg(A.base + d) + k * (uint32_t)MSD_TILE;
	uint32_t const left = __ldg(A.base + d + 1) - start;
	uint32_t const m = left < (uint32_t)MSD_TILE ? left : (uint32_t)MSD_TILE;
	unsigned long long * const g = A.recs + start;
	unsigned long long r[MSD_ITEMS];
	uint32_t dr[MSD_ITEMS];
	unsigned const sh = 62u - A.b2; // key30 sits in bits 61..32
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) {
		uint32_t const s = j * MSD_THREADS + threadIdx.x;
		r[j] = s < m ? g[s] : 0ull;
	}
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j) {
		uint32_t const s = j * MSD_THREADS + threadIdx.x;
		uint32_t const d2 = (uint32_t)((r[j] << 2) >> (sh + 2));
		dr[j] = s < m ? ((d2 << 16) | atomicAdd(&cnt[d2], 1u)) : 0xffffffffu;
	}
	__syncthreads();
	uint32_t mine[4];
	unsigned b0, per;
	msd_scan_bins<MSD_THREADS, 4>(cnt, nb2, wsum, mine, b0, per);
	uint16_t * const tab = A.table + (uint64_t)tp * (nb2 + 1) + k;
	for (unsigned b = threadIdx.x; b <= nb2; b += MSD_THREADS) tab[(uint64_t)b * ntp] = (uint16_t)(b < nb2 ? cnt[b] : m);
	// sorted tile in shared memory, shifted by the parity of `start` so that 16-byte aligned global
	// words are 16-byte aligned in shared memory (bulk copies need both)
	unsigned const par = start & 1u;
	#pragma unroll
	for (int j = 0; j < MSD_ITEMS; ++j)
		if (dr[j] != 0xffffffffu) stage[par + cnt[dr[j] >> 16] + (dr[j] & 0xffffu)] = r[j];
	// make the generic-proxy writes visible to the async proxy, then one thread stores the tile
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	uint32_t const s_lo = par, s_hi = m - ((start + m) & 1u); // slots [s_lo, s_hi) start and end on 16-byte boundaries
	if (threadIdx.x == 0) {
		if (s_hi > s_lo) {
			uint32_t const bytes = (s_hi - s_lo) * 8u;
			uint32_t const src = (uint32_t)__cvta_generic_to_shared(stage + par + s_lo);
			unsigned long long * const dst = g + s_lo;
			for (uint32_t off = 0; off < bytes; off += 16384u) {
				uint32_t const len = bytes - off < 16384u ? bytes - off : 16384u;
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"((const char *)dst + off), "r"(src + off), "r"(len) : "memory");
			}
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		if (par && m) g[0] = stage[par];
		if (s_hi < m && s_hi >= s_lo) g[s_hi] = stage[par + s_hi];
		if (s_hi > s_lo) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}

// ---- sizes of the sub-buckets ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_msd_subtotals(unsigned b2, const uint32_t * __restrict__ tpre, const uint16_t * __restrict__ table, uint32_t * __restrict__ subcount,
                uint32_t * __restrict__ maxcount) {
	__shared__ uint32_t rowsum[MSD_MAXBINS + 1];
	unsigned const nb2 = 1u << b2, d = blockIdx.x;
	uint32_t const tp = tpre[d], ntp = tpre[d + 1] - tp;
	const uint16_t * const tab = table + (uint64_t)tp * (nb2 + 1);
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (unsigned r = w; r <= nb2; r += 8) {
		const uint16_t * row = tab + (uint64_t)r * ntp;
		uint32_t s = 0;
		for (uint32_t k = lane; k < ntp; k += 32) s += row[k];
		s = __reduce_add_sync(0xffffffffu, s);
		if (lane == 0) rowsum[r] = s;
	}
	__syncthreads();
	uint32_t mx = 0;
	for (unsigned r = threadIdx.x; r < nb2; r += 256) {
		uint32_t const c = rowsum[r + 1] - rowsum[r];
		subcount[(uint64_t)d * nb2 + r] = c;
		mx = c > mx ? c : mx;
	}
	mx = __reduce_max_sync(0xffffffffu, mx);
	if (lane == 0 && mx) atomicMax(maxcount, mx);
}

// ---- finish ------------------------------------------------------------------------------------
struct MsdFin {
	TextView v;
	int lin;
	unsigned b1, b2, glog;          // 2^glog lanes copy one run
	unsigned nkeep;
	const unsigned long long * recs;
	const uint32_t * base, * tpre;
	const uint16_t * table;
	const uint32_t * subbase;       // [nkeep * 2^b2 + 1] exclusive scan of the sub-bucket sizes
	uint32_t sb0;                   // first sub-bucket of this launch
	uint32_t * sa_out;              // ORDER
	uint8_t * hflag;
	FusedOut fo;                    // FUSED (shift includes the rank of the first kept record)
	uint32_t imask, rmask;          // a suffix at position i / of rank r is sampled only if (i & imask) == 0 or (r & rmask) == 0
	unsigned long long * counters;  // [MSD_CSLOTS][4]: unresolved, tied, second keys read, flags (1: run too long, 2: sub-bucket too large)
	const uint8_t * shortflag;      // linear windows: [sub-bucket] != 0 where one of the last 18 suffixes of the window lies (nullptr: none)
};

// marks the sub-buckets that hold one of the last 18 suffixes of a linear window: only there can a suffix be shorter
// than the prefix an unresolved group shares
__global__ void k_msd_shortflags(TextView v, unsigned b1, unsigned b2, uint32_t d_lo, uint32_t nkeep, uint8_t * __restrict__ flags) {
	uint64_t const first = v.W > 18 ? v.W - 18 : 0;
	uint64_t const i = first + threadIdx.x;
	if (i >= v.W) return;
	uint32_t d, hi32;
	msd_record(v, i, b1, d, hi32);
	uint32_t const b = d - d_lo;
	if (b < nkeep) flags[((uint64_t)b << b2) | ((hi32 & MSD_KEYMASK) >> (30u - b2))] = 1;
}

__device__ __forceinline__ void msd_second_key(TextView const & v, unsigned skip, int lin, uint32_t i, unsigned long long & k2, uint32_t & rem) {
	k2 = tv_symbols(v, (uint64_t)i + skip, 32, 2);
	uint64_t const left = v.W - i;
	rem = (uint32_t)((lin && left < skip + 32u) ? left : skip + 32u);
}

// everything fo_emit (sufsort.cu) writes except the BWT code (SA: the sampled SA as well)
template <bool SA>
__device__ __forceinline__ void msd_emit_samples(FusedOut const & fo, uint32_t i, uint32_t r) {
	if (i == 0) { if (fo.has_term) fo.special[0] = r; fo.special[1] = r; }
	if (fo.prelog >= 32 ? i == 0 : (i & ((1u << fo.prelog) - 1u)) == 0) fo.prerank[fo.prelog >= 32 ? 0 : (i >> fo.prelog)] = r;
	if (fo.isa_s && (fo.isalog >= 32 ? i == 0 : (i & ((1u << fo.isalog) - 1u)) == 0)) fo.isa_s[fo.isalog >= 32 ? 0 : (i >> fo.isalog)] = r;
	if (SA && fo.sa_s && (fo.salog >= 32 ? r == 0 : (r & ((1u << fo.salog) - 1u)) == 0)) fo.sa_s[fo.salog >= 32 ? 0 : (r >> fo.salog)] = i;
}

// one chunk of THREADS tiles of the parent bucket: thread <-> tile k0 + threadIdx.x; returns this
// thread's run widened to 16-byte boundaries as (first record a0, padded length plen, place off in rec[])
// and adds the chunk's padded total to `done`.  g0/g1 = the run itself.
template <int THREADS>
__device__ __forceinline__ void msd_run_place(const uint16_t * __restrict__ row0, const uint16_t * __restrict__ row1, uint32_t pstart, uint32_t ntp,
                                              uint32_t k0, uint32_t * wsum, uint32_t & done, uint32_t & g0, uint32_t & g1, uint32_t & a0, uint32_t & plen, uint32_t & off) {
	unsigned const lane = threadIdx.x & 31;
	uint32_t const k = k0 + threadIdx.x;
	g0 = g1 = 0;
	if (k < ntp) { uint32_t const s = row0[k]; g0 = pstart + k * (uint32_t)MSD_TILE + s; g1 = g0 + (row1[k] - s); }
	a0 = g0 & ~1u;
	plen = g1 > g0 ? ((g1 + 1u) & ~1u) - a0 : 0u;
	uint32_t incl = plen;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
	__syncthreads(); // wsum (and the caller's descriptors) of the previous chunk have been consumed
	if (lane == 31) wsum[threadIdx.x >> 5] = incl;
	__syncthreads();
	uint32_t before = 0, ctot = 0;
	#pragma unroll
	for (int ww = 0; ww < THREADS / 32; ++ww) { uint32_t const x = wsum[ww]; before += ww < (int)(threadIdx.x >> 5) ? x : 0u; ctot += x; }
	off = done + before + incl - plen;
	done += ctot;
}

// ---- mbarrier + bulk async copy (TMA, 1-D) --------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
	uint32_t ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
	} while (!ok);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void * src, uint32_t bytes, uint32_t mbar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

constexpr int MSD_WIN = 3;                          // sentinel slots on either side of the sorted records
constexpr int MSD_QCAP = 1024;                      // records of a CTA whose general path is deferred behind the main loop
constexpr int MSD_FIN_SMEM = 32 + (MSD_CAP + 4) * 8 + MSD_CAP + 16 + (1 << MSD_LBITS_MAX) * 4;

// Shared memory: 32 B (front sentinels) | rec[MSD_CAP + 4] | s_bwt[MSD_CAP] | 16 B (cnt[-1] = 0) | cnt[2^lb]
// Steps: (1) every run of the sub-bucket arrives by ONE bulk async copy issued by the thread that owns its
// descriptor, all counted on one mbarrier; (2) local digit: count with shared-memory atomics, scan, then a
// second atomic on the scanned table hands every record its slot (no rank kept in registers); afterwards bin d
// is rec[cnt[d-1] .. cnt[d]); (3) order inside a bin: the bins hold 0.7 records on average, so a record
// compares itself with its neighbours at distance 1 and 2 (the array is sorted by digit: a later record with a
// smaller key, or an earlier one with a larger key, is in the same bin) and checks that the bin ends within
// distance 2 on both sides; records with an equal neighbour or a longer bin take the general path (bin bounds
// from the table, second keys from the text, unresolved groups) -- queued and worked off densely behind the main
// loop, so that one such record does not hold up its warp; (4) emit.
template <bool FUSED, bool ORDER, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
k_msd_finish(const __grid_constant__ MsdFin A) {
	constexpr int ITEMS = MSD_CAP / THREADS;
	extern __shared__ __align__(16) uint8_t msd_dyn[];
	unsigned long long * const rec = reinterpret_cast<unsigned long long *>(msd_dyn + 32);             // MSD_CAP + 4
	uint8_t * const s_bwt = msd_dyn + 32 + (size_t)(MSD_CAP + 4) * 8;                                  // MSD_CAP
	uint32_t * const cnt = reinterpret_cast<uint32_t *>(s_bwt + MSD_CAP + 16);                         // cnt[-1 .. 2^lb)
	__shared__ uint32_t wsum[THREADS / 32];
	__shared__ uint32_t s_sa[512];                                // sampled SA of this CTA's rows (at most MSD_CAP / 32)
	__shared__ uint16_t s_q[MSD_QCAP];                            // deferred records
	__shared__ uint32_t s_cnt[4], s_nq;
	__shared__ __align__(8) unsigned long long s_mbar;
	// linear windows: the suffixes shorter than the prefix an unresolved group shares (at most 16 in the whole text)
	__shared__ uint32_t s_sh_e[16], s_sh_L[16], s_nshort;
	unsigned const nb2 = 1u << A.b2;
	uint32_t const sb = blockIdx.x + A.sb0;
	uint32_t const o0 = A.subbase[sb], m = A.subbase[sb + 1] - o0;
	if (m == 0) return;
	uint32_t const W32 = (uint32_t)A.v.W;
	unsigned const d = sb >> A.b2, d2 = sb & (nb2 - 1u);
	uint32_t const tp = __ldg(A.tpre + d), ntp = __ldg(A.tpre + d + 1) - tp;
	uint32_t const pstart = __ldg(A.base + d);
	const uint16_t * const row0 = A.table + (uint64_t)tp * (nb2 + 1) + (uint64_t)d2 * ntp;
	const uint16_t * const row1 = row0 + ntp;
	unsigned const lane = threadIdx.x & 31;
	if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
	if (threadIdx.x == 0) { s_nshort = 0; s_nq = 0; }

	if ((uint64_t)m + 2ull * ntp > (uint64_t)MSD_CAP) {
		__syncthreads();
		// too large for one CTA: the whole sub-bucket stays one unresolved group sharing h2 symbols.  The
		// suffixes of a linear window that end inside those symbols are no members of it: they are smaller
		// than the rest, shorter first, and are placed in front (a group must share REAL symbols, the
		// doubling rounds compare what follows them).
		if (ORDER) {
			uint32_t const h2 = (A.b1 + A.b2) >> 1;
			if (A.lin) {
				uint32_t done = 0;
				for (uint32_t k = 0; k < ntp; ++k) {
					uint32_t const s = row0[k], len = row1[k] - s;
					const unsigned long long * src = A.recs + pstart + (uint64_t)k * MSD_TILE + s;
					for (uint32_t x = threadIdx.x; x < len; x += THREADS) {
						uint32_t const L = W32 - (uint32_t)src[x];
						if (L < h2) { uint32_t const q = atomicAdd(&s_nshort, 1u); s_sh_e[q] = done + x; s_sh_L[q] = L; }
					}
					done += len;
				}
				__syncthreads();
			}
			uint32_t const ns = s_nshort;
			uint32_t done = 0; // runs are copied one after the other
			for (uint32_t k = 0; k < ntp; ++k) {
				uint32_t const s = row0[k], len = row1[k] - s;
				const unsigned long long * src = A.recs + pstart + (uint64_t)k * MSD_TILE + s;
				for (uint32_t x = threadIdx.x; x < len; x += THREADS) {
					uint32_t const i = (uint32_t)src[x], e = done + x, L = W32 - i;
					uint32_t pos = e, hf = e == 0 ? 1u : 0u;
					if (ns) {
						uint32_t before = 0, smaller = 0;
						for (uint32_t q = 0; q < ns; ++q) { before += s_sh_e[q] < e ? 1u : 0u; smaller += s_sh_L[q] < L ? 1u : 0u; }
						if (L < h2) { pos = smaller; hf = 1; }
						else { pos = ns + e - before; hf = pos == ns ? 1u : 0u; }
					}
					A.sa_out[o0 + pos] = i;
					A.hflag[o0 + pos] = (uint8_t)hf;
				}
				done += len;
			}
		}
		if (threadIdx.x == 0) {
			atomicAdd(&A.counters[(sb % MSD_CSLOTS) * 4 + 0], (unsigned long long)m);
			atomicOr(&A.counters[(sb % MSD_CSLOTS) * 4 + 3], 2ull);
		}
		return;
	}

	// local digit: about two bins per record
	unsigned lb = 32u - (unsigned)__clz((int)(2u * m - 1u)); // ceil(log2(2m))
	lb = lb < (unsigned)MSD_LBITS_MIN ? (unsigned)MSD_LBITS_MIN : (lb > (unsigned)MSD_LBITS_MAX ? (unsigned)MSD_LBITS_MAX : lb);
	unsigned const nlb = 1u << lb;
	unsigned const dsh = 30u - A.b2 - lb;  // the local digit follows the b2 bits of level 2 inside key30: bits [dsh, dsh + lb) of a record's upper word
	uint32_t const dmask = nlb - 1u;

	// ---- (1) gather: one bulk copy per run (a run is widened to 16-byte boundaries, the at most two records of
	//      neighbouring runs this drags in are overwritten with MSD_PAD afterwards); the table is cleared meanwhile ----
	uint32_t const rec_s = (uint32_t)__cvta_generic_to_shared(rec), mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
	if (threadIdx.x == 0) mbar_init(mbar, (ntp + THREADS - 1) / THREADS);
	__syncthreads();
	uint32_t mpad = 0; // records in rec[], padding included
	uint32_t g0 = 0, g1 = 0, a0, plen = 0, off = 0;
	for (uint32_t k0 = 0; k0 < ntp; k0 += THREADS) {
		uint32_t const before = mpad;
		msd_run_place<THREADS>(row0, row1, pstart, ntp, k0, wsum, mpad, g0, g1, a0, plen, off);
		if (plen) bulk_g2s(rec_s + 8u * off, A.recs + a0, plen * 8u, mbar);
		if (threadIdx.x == 0) mbar_arrive_expect_tx(mbar, (mpad - before) * 8u);
	}
	for (unsigned i = threadIdx.x; i < nlb; i += THREADS) cnt[i] = 0;
	if (threadIdx.x < 4) cnt[(int)threadIdx.x - 4] = 0;
	if (threadIdx.x < (unsigned)MSD_WIN) rec[-1 - (int)threadIdx.x] = 0ull; // front sentinels: smaller than every key
	bool const shortsb = A.shortflag && A.shortflag[sb]; // uniform
	mbar_wait(mbar, 0);
	// the records dragged in from neighbouring runs become padding (several chunks of tiles: one more walk over the descriptors)
	if (ntp <= (uint32_t)THREADS) {
		if (plen) {
			if (g0 & 1u) rec[off] = MSD_PAD;
			if (g1 & 1u) rec[off + plen - 1] = MSD_PAD;
		}
	} else {
		uint32_t done = 0;
		for (uint32_t k0 = 0; k0 < ntp; k0 += THREADS) {
			msd_run_place<THREADS>(row0, row1, pstart, ntp, k0, wsum, done, g0, g1, a0, plen, off);
			if (plen) {
				if (g0 & 1u) rec[off] = MSD_PAD;
				if (g1 & 1u) rec[off + plen - 1] = MSD_PAD;
			}
		}
	}
	__syncthreads();

	// ---- (2) local digit: count, scan, hand out the slots; the records pass through registers ----
	{
		uint32_t rlo[ITEMS], rhi[ITEMS]; // lower word ~0: padding (a record's index is below 2^32 - 256) or no record
		#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			uint32_t const s = j * THREADS + threadIdx.x;
			rlo[j] = 0xffffffffu; rhi[j] = 0;
			if (s < mpad) { uint2 const x = *reinterpret_cast<const uint2 *>(rec + s); rlo[j] = x.x; rhi[j] = x.y; }
			if (rlo[j] != 0xffffffffu) atomicAdd(&cnt[(rhi[j] >> dsh) & dmask], 1u);
		}
		__syncthreads();
		msd_scan_bins_wide<THREADS>(cnt, nlb, wsum);
		#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (rlo[j] != 0xffffffffu) {
				uint32_t const pos = atomicAdd(&cnt[(rhi[j] >> dsh) & dmask], 1u);
				*reinterpret_cast<uint2 *>(rec + pos) = make_uint2(rlo[j], rhi[j]);
			}
		if (threadIdx.x < (unsigned)MSD_WIN) rec[m + threadIdx.x] = (unsigned long long)MSD_KEYMASK << 32; // back sentinels: no key is larger
		__syncthreads();
	}
	// now bin dg = rec[cnt[dg - 1] .. cnt[dg]), cnt[-1] = 0

	// crowded local digits stay unresolved groups sharing hbig symbols: list the suffixes too short for that
	uint32_t const hbig = (A.b1 + A.b2 + lb) >> 1; // at most 16
	if (shortsb) { // only the CTAs whose sub-bucket holds one of the last suffixes of a linear window
		for (uint32_t s = threadIdx.x; s < m; s += THREADS) {
			uint32_t const L = W32 - (uint32_t)rec[s];
			if (L < hbig) { uint32_t const q = atomicAdd(&s_nshort, 1u); s_sh_e[q] = s; s_sh_L[q] = L; }
		}
		__syncthreads();
	}
	// ---- (3) order inside a bin, (4) emit ----
	// sampled SA through shared memory when a CTA holds at most 512 samples
	bool const stage_sa = FUSED && A.fo.sa_s && A.fo.salog >= 5 && A.fo.salog < 32;
	uint32_t const rshift = (uint32_t)A.fo.shift + o0;
	uint32_t const r_first = stage_sa ? ((rshift + A.rmask) & ~A.rmask) : 0u; // first sampled rank of this CTA
	unsigned const skip = (A.b1 + 30u) >> 1; // symbols covered by b1 + key30
	uint32_t ntied = 0, nunres = 0, ngather = 0, flags = 0;
	const uint32_t * const rh = reinterpret_cast<const uint32_t *>(rec) + 1; // upper word of slot y: rh[2 * y]

	// final place f (relative to the sub-bucket) and head flag of the record in slot s, bin bounds from the table
	auto general_path = [&](uint32_t s, unsigned long long me, uint32_t & f, uint32_t & hf) {
		uint32_t const mh = (uint32_t)(me >> 32), mk = mh & MSD_KEYMASK;
		uint32_t const dg = (mh >> dsh) & dmask;
		uint32_t const a = cnt[(int)dg - 1], b = cnt[dg];
		f = a; hf = 1;
		if (b - a > (uint32_t)MSD_MAXRUN) {
			// one unresolved group; the listed short suffixes of this digit go in front of it, shorter first
			uint32_t const ns = s_nshort, myL = W32 - (uint32_t)me;
			uint32_t nsb = 0, before = 0, smaller = 0;
			for (uint32_t q = 0; q < ns; ++q) {
				uint32_t const e = s_sh_e[q];
				if (e >= a && e < b) { ++nsb; before += e < s ? 1u : 0u; smaller += s_sh_L[q] < myL ? 1u : 0u; }
			}
			if (A.lin && myL < hbig) { f = a + smaller; hf = 1; }
			else { f = s + nsb - before; hf = (s - a == before) ? 1u : 0u; ++nunres; flags |= 1u; }
		} else if (b - a > 1) {
			uint32_t less = 0, eq = 0;
			#pragma unroll 1
			for (uint32_t y = a; y < b; ++y) {
				uint32_t const k = rh[2 * y] & MSD_KEYMASK;
				less += k < mk ? 1u : 0u;
				eq += k == mk ? 1u : 0u;
			}
			f = a + less;
			if (eq > 1) {
				// equal on all the bits a record carries: compare the next 32 symbols, read from the text
				++ntied;
				unsigned long long mk2; uint32_t mr;
				msd_second_key(A.v, skip, A.lin, (uint32_t)me, mk2, mr);
				uint32_t eqb = 0, eqa = 0;
				#pragma unroll 1
				for (uint32_t y = a; y < b; ++y) {
					unsigned long long const o = rec[y];
					if (y == s || ((uint32_t)(o >> 32) & MSD_KEYMASK) != mk) continue;
					unsigned long long ok2; uint32_t orr;
					msd_second_key(A.v, skip, A.lin, (uint32_t)o, ok2, orr);
					bool const same = ok2 == mk2 && orr == mr;
					f += (ok2 < mk2 || (ok2 == mk2 && orr < mr) || (same && y < s)) ? 1u : 0u;
					eqb += (same && y < s) ? 1u : 0u;
					eqa += same ? 1u : 0u;
				}
				hf = eqb == 0 ? 1u : 0u;
				if (eqa) ++nunres;
				++ngather;
			}
		}
	};
	auto emit = [&](unsigned long long me, uint32_t f, uint32_t hf) {
		uint32_t const i = (uint32_t)me;
		if (ORDER) { A.sa_out[o0 + f] = i; A.hflag[o0 + f] = (uint8_t)hf; }
		if (FUSED) {
			s_bwt[f] = (uint8_t)(me >> 62);
			uint32_t const r = rshift + f;
			if (stage_sa) {
				// sampled SA: the position goes to shared memory and leaves with a coalesced store below
				if ((r & A.rmask) == 0) s_sa[(r - r_first) >> A.fo.salog] = i;
				if ((i & A.imask) == 0) msd_emit_samples<false>(A.fo, i, r);
			} else if ((i & A.imask) == 0 || (r & A.rmask) == 0) msd_emit_samples<true>(A.fo, i, r);
		}
	};

	#pragma unroll 1
	for (uint32_t s = threadIdx.x; s < m; s += THREADS) {
		unsigned long long const me = rec[s];
		uint32_t const mh = (uint32_t)(me >> 32), mk = mh & MSD_KEYMASK;
		const uint32_t * const w = rh + 2 * s;
		uint32_t const p1 = w[-2] & MSD_KEYMASK, p2 = w[-4] & MSD_KEYMASK, n1 = w[2] & MSD_KEYMASK, n2 = w[4] & MSD_KEYMASK;
		uint32_t const p3 = w[-6], n3 = w[6];
		// keys are below 2^30: the sign bit of a difference is the comparison
		uint32_t f = s + ((n1 - mk) >> 31) + ((n2 - mk) >> 31) - ((mk - p1) >> 31) - ((mk - p2) >> 31);
		uint32_t hf = 1;
		// an equal neighbour, or a bin that reaches distance 3 (the digit and everything above it are equal)
		bool const general = p1 == mk || n1 == mk || p2 == mk || n2 == mk || (((p3 ^ mh) & MSD_KEYMASK) >> dsh) == 0 || (((n3 ^ mh) & MSD_KEYMASK) >> dsh) == 0;
		if (general) {
			uint32_t const q = atomicAdd(&s_nq, 1u);
			if (q < (uint32_t)MSD_QCAP) { s_q[q] = (uint16_t)s; continue; }
			general_path(s, me, f, hf);
		}
		emit(me, f, hf);
	}
	__syncthreads();
	{
		uint32_t const nq = s_nq < (uint32_t)MSD_QCAP ? s_nq : (uint32_t)MSD_QCAP;
		#pragma unroll 1
		for (uint32_t q = threadIdx.x; q < nq; q += THREADS) {
			uint32_t const s = s_q[q];
			unsigned long long const me = rec[s];
			uint32_t f, hf;
			general_path(s, me, f, hf);
			emit(me, f, hf);
		}
	}
	if (FUSED) {
		__syncthreads();
		if (stage_sa) {
			uint32_t const r_end = rshift + m;
			uint32_t const ns = r_first < r_end ? ((r_end - 1u - r_first) >> A.fo.salog) + 1u : 0u;
			for (uint32_t x = threadIdx.x; x < ns; x += THREADS) {
				A.fo.sa_s[(r_first >> A.fo.salog) + x] = s_sa[x];
				if (A.fo.sa_s2) A.fo.sa_s2[(r_first >> A.fo.salog) + x] = s_sa[x];
			}
		}
		uint8_t * const out = A.fo.bwt + A.fo.shift + o0;
		// head bytes up to a 4-byte boundary, words, tail bytes
		uint32_t const mis = (uint32_t)((4u - ((uintptr_t)out & 3u)) & 3u);
		uint32_t const head = mis < m ? mis : m;
		if (threadIdx.x < head) out[threadIdx.x] = s_bwt[threadIdx.x];
		uint32_t const nw = (m - head) >> 2;
		for (uint32_t x = threadIdx.x; x < nw; x += THREADS) {
			uint32_t const p = head + 4 * x;
			reinterpret_cast<uint32_t *>(out + head)[x] = (uint32_t)s_bwt[p] | ((uint32_t)s_bwt[p + 1] << 8) | ((uint32_t)s_bwt[p + 2] << 16) | ((uint32_t)s_bwt[p + 3] << 24);
		}
		uint32_t const tail0 = head + 4 * nw;
		if (threadIdx.x < m - tail0) out[tail0 + threadIdx.x] = s_bwt[tail0 + threadIdx.x];
	}
	ntied = __reduce_add_sync(0xffffffffu, ntied);
	nunres = __reduce_add_sync(0xffffffffu, nunres);
	ngather = __reduce_add_sync(0xffffffffu, ngather);
	flags = __reduce_or_sync(0xffffffffu, flags);
	if (lane == 0) {
		if (nunres) atomicAdd(&s_cnt[0], nunres);
		if (ntied) atomicAdd(&s_cnt[1], ntied);
		if (ngather) atomicAdd(&s_cnt[2], ngather);
		if (flags) atomicOr(&s_cnt[3], flags);
	}
	__syncthreads();
	if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&A.counters[(sb % MSD_CSLOTS) * 4 + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
	if (threadIdx.x == 3 && s_cnt[3]) atomicOr(&A.counters[(sb % MSD_CSLOTS) * 4 + 3], (unsigned long long)s_cnt[3]);
}

} // namespace b3m
