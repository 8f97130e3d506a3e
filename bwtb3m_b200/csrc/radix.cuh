// LSD radix sort on records held as NA parallel uint32 arrays (structure of arrays) plus an
// optional aux byte, 8-bit digits, one kernel per digit ("onesweep"): each CTA (512 threads,
// 16 records per thread) takes the next 8192-record tile from a ticket counter, ranks its records
// stably (eight ballots per record, their predicates from one R2P, find the lanes with an equal
// digit; 16-bit per-warp counters in shared memory), obtains its global offsets by decoupled
// look-back over the tile status words of its predecessors, reorders through shared memory and
// writes runs of records that share a digit.  The only partial tile is a launch of its own, so the
// main kernel carries no bounds checks.  For suffix sorting the first pass builds its records
// straight from the 2-bit packed text (radix_sort_suffix_keys).
//
// Measured at 3.1 G records on a B200: 20.2 ms per pass = 42 % of HBM peak,
// 16.4 ms on already sorted input, 25 ms with 4096-record tiles, 21.8 ms with 16384-record tiles
// at one CTA per SM, 27.5 ms with __match_any_sync instead of the ballots (MATCH runs on the ADU
// pipe).  What bounds it is instruction issue and shared-memory wavefronts (45 % of them bank
// conflicts of the scatter into sorted order), not DRAM.
// Forward progress: tickets are handed out in launch order, so every predecessor of a tile is
// already resident or finished when the tile starts to look back.
//
// Algorithmic HBM bytes per pass and record: 4*NA (+1 aux) read and the same
// written, plus one 4-byte key read per radix_sort_bits() call for the digit histograms (the
// suffix-key sort gets all four histograms from one pass over the packed text, k_hist_4mers).
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "textview.cuh"

namespace b3m {

constexpr int RADIX_THREADS = 512;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;
constexpr int RADIX_ITEMS = 16;
constexpr int RADIX_CTAS_PER_SM = 2;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS; // 8192 records: long runs per digit keep the scattered writes DRAM-friendly
constexpr int RADIX_BINS = 256;
constexpr int RADIX_MAXDIG = 4;

template <int NA>
struct RadixRec {
	uint32_t * a[NA];
	uint8_t * aux = nullptr; // optional one-byte payload that travels with the record
};

// lanes holding the same 8-bit digit
__device__ __forceinline__ unsigned warp_peers8(uint32_t d) {
	unsigned mism = 0u; // lanes whose digit differs from mine in some bit
	#pragma unroll
	for (int b = 0; b < 8; ++b) {
		// m = ballot(bit b of d); mism |= bit ? ~m : m   (predicate straight from the AND, no shifts or selects)
		asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t;\n\t"
		             "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
		             "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
		             "@p not.b32 m, m;\n\tor.b32 %0, %0, m;\n\t}"
		             : "+r"(mism) : "r"(d), "r"(1u << b));
	}
	return ~mism;
}

// ---- digit histograms of one key array: up to 4 digits in one read -----------------------
static __global__ void __launch_bounds__(256)
k_radix_hist(const uint32_t * __restrict__ key, uint64_t n, int bit_lo, int ndig, uint32_t lastmask,
             unsigned long long * __restrict__ ghist /* [ndig][256] */) {
	__shared__ uint32_t sh[RADIX_MAXDIG][RADIX_BINS];
	for (int i = threadIdx.x; i < RADIX_MAXDIG * RADIX_BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
	__syncthreads();
	uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned const lane = threadIdx.x & 31;
	for (uint64_t b0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); b0 < n; b0 += stride) {
		uint64_t const i = b0 + lane;
		bool const valid = i < n;
		uint32_t const k = valid ? (key[i] >> bit_lo) : 0u;
		unsigned const vmask = __ballot_sync(0xffffffffu, valid);
		for (int d = 0; d < ndig; ++d) {
			uint32_t const dg = (k >> (8 * d)) & (d == ndig - 1 ? lastmask : 255u);
			// a digit that is constant across the warp (high bits of small keys) is aggregated
			uint32_t const d0 = __shfl_sync(0xffffffffu, dg, 0);
			if (__all_sync(0xffffffffu, !valid || dg == d0)) { if (lane == 0) atomicAdd(&sh[d][d0], (uint32_t)__popc(vmask)); }
			else if (valid) atomicAdd(&sh[d][dg], 1u);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < ndig * RADIX_BINS; i += blockDim.x)
		if ((&sh[0][0])[i]) atomicAdd(&ghist[i], (unsigned long long)(&sh[0][0])[i]);
}

// exclusive scan of each digit's histogram; skip[d] = 1 when one bin holds every record
static __global__ void __launch_bounds__(256)
k_radix_hist_scan(const unsigned long long * __restrict__ ghist, int ndig, uint64_t n, uint32_t * __restrict__ base /* [ndig][256] */,
                  uint32_t * __restrict__ skip) {
	for (int d = 0; d < ndig; ++d) {
		uint32_t const c = (uint32_t)ghist[d * RADIX_BINS + threadIdx.x];
		uint32_t total;
		uint32_t const incl = block_scan_inclusive<OpSum>(c, &total);
		base[d * RADIX_BINS + threadIdx.x] = incl - c;
		if ((uint64_t)c == n) skip[d] = 1;
		__syncthreads();
	}
}

// status word: bits 63..62 = flag (0 empty, 1 aggregate of this tile, 2 inclusive prefix), low 32 bits = count
constexpr unsigned long long RADIX_FLAG_AGG = 1ull << 62;
constexpr unsigned long long RADIX_FLAG_INC = 2ull << 62;

template <int NA>
struct RadixPassArgs {
	const uint32_t * in[NA];   // in[0] is the array that holds the digit
	uint32_t * out[NA];
	const uint8_t * aux_in;    // optional byte payload (nullptr: none)
	uint8_t * aux_out;
};

// The first pass of a suffix sort can take its records straight from the text (TEXT): record t
// is the suffix at window index i(t) (the short suffixes of a linear window first, see
// sufsort.cu), key = its first k0 symbols, payload = i, aux byte = preceding code | next symbols.
struct RadixTextSrc {
	TextView v;
	uint64_t nshort;
	unsigned bits, k0;
};

__device__ __forceinline__ void radix_text_record(RadixTextSrc const & S, uint64_t t, uint32_t & key, uint32_t & idx, uint32_t & aux) {
	uint64_t const i = (t < S.nshort) ? (S.v.W - 1 - t) : (t - S.nshort);
	unsigned const nx = 8u / S.bits - 1u;
	uint64_t const ks = tv_symbols(S.v, i, S.k0 + nx, S.bits);
	key = (uint32_t)(ks >> (nx * S.bits));
	idx = (uint32_t)i;
	aux = ((tv_pred(S.v, i) << (nx * S.bits)) | (uint32_t)(ks & ((1u << (nx * S.bits)) - 1u))) & 255u;
}

// Keys and aux bytes of the records chunk + 32*j + lane, j < ITEMS (one warp, `chunk` warp-uniform).
// Fast path (2-bit packed text, the warp's records away from both ends of the window and of the
// text): lane l reads the words under positions p0+l, p0+l+32, ... -- its bit offset inside a word
// never changes, and every word is the second half of the previous record's window.
template <int ITEMS>
__device__ __forceinline__ void radix_text_load(RadixTextSrc const & S, uint64_t chunk, unsigned lane, uint32_t (&k)[ITEMS], uint32_t (&aux)[ITEMS]) {
	uint64_t const i0 = chunk - S.nshort;
	uint64_t p0 = S.v.wstart + i0;
	if (S.v.text_wraps && p0 >= S.v.ntext) p0 -= S.v.ntext;
	bool const fast = S.bits == 2 && S.v.packed && chunk >= S.nshort && i0 + 32 * ITEMS + 35 <= S.v.W &&
	                  p0 >= 1 && p0 + 32 * ITEMS + 35 <= S.v.ntext;
	if (fast) {
		uint64_t const pl = p0 + lane;
		const uint64_t * wp = S.v.packed + (pl >> 5);
		unsigned const sh = (unsigned)(pl & 31u) << 1;
		uint64_t prevw = (sh == 0) ? __ldg(wp - 1) : 0ull; // pl >= 32 whenever sh == 0 (p0 >= 1)
		uint64_t cw = __ldg(wp);
		#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			uint64_t const nw = __ldg(wp + j + 1);
			uint64_t const win = sh ? ((cw << sh) | (nw >> (64u - sh))) : cw;
			uint32_t const pred = sh ? (uint32_t)(cw >> (64u - sh)) & 3u : (uint32_t)prevw & 3u;
			k[j] = (uint32_t)(win >> 32);
			aux[j] = (pred << 6) | ((uint32_t)(win >> 26) & 63u);
			prevw = cw; cw = nw;
		}
	} else {
		#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			uint64_t const t = chunk + j * 32 + lane;
			uint32_t ii;
			k[j] = 0xffffffffu; aux[j] = 0;
			if (t < S.v.W) radix_text_record(S, t, k[j], ii, aux[j]);
		}
	}
}
// window index of record t of the text source
__device__ __forceinline__ uint32_t radix_text_index(RadixTextSrc const & S, uint64_t t) {
	return (uint32_t)((t < S.nshort) ? (S.v.W - 1 - t) : (t - S.nshort));
}

template <int NA, bool AUX, bool TEXT, bool FULL>
__global__ void __launch_bounds__(RADIX_THREADS, RADIX_CTAS_PER_SM)
k_radix_onesweep(RadixPassArgs<NA> A, RadixTextSrc S, uint64_t n, int shift, uint32_t mask, const uint32_t * __restrict__ base /* [256] */,
                 unsigned long long * __restrict__ status /* [ntiles][256] */, uint32_t * __restrict__ ticket) {
	__shared__ uint16_t wcnt[RADIX_WARPS][RADIX_BINS]; // counts, then tile-local offsets: all below RADIX_TILE
	__shared__ uint32_t gbase[RADIX_BINS];
	__shared__ uint32_t wsum[RADIX_BINS / 32];
	// dynamic part (radix_smem_bytes): keys, one payload array, aux bytes of the tile in sorted order
	extern __shared__ __align__(16) uint8_t radix_dyn[];
	uint32_t * const skey = reinterpret_cast<uint32_t *>(radix_dyn);
	uint32_t * const sval = skey + RADIX_TILE; // skey|sval together hold the (key, first payload) words of the tile
	uint8_t * const saux = reinterpret_cast<uint8_t *>(sval + (NA > 1 ? RADIX_TILE : 0));
	uint8_t * const stash = saux + RADIX_TILE; // TEXT: a thread parks the aux bytes it made until they are scattered (frees registers)
	__shared__ uint32_t s_tile;
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS / 2; i += RADIX_THREADS) reinterpret_cast<uint32_t *>(&wcnt[0][0])[i] = 0;
	__syncthreads();
	uint32_t const tile = s_tile;
	uint64_t const tbase = (uint64_t)tile * RADIX_TILE;
	uint64_t const chunk = tbase + (uint64_t)w * (32 * RADIX_ITEMS);
	// FULL: every record of the tile exists (the one partial tile at the end is a launch of its own)
	uint32_t const nvalid = FULL ? (uint32_t)RADIX_TILE : ((n - tbase) < (uint64_t)RADIX_TILE ? (uint32_t)(n - tbase) : (uint32_t)RADIX_TILE);
	#define RADIX_VALID(i) (FULL || (i) < n)

	uint32_t k[RADIX_ITEMS];
	bool tfast = false; // TEXT: this warp's records are taken lane-blocked (see below)
	if (TEXT) {
		static_assert(!TEXT || (NA == 2 && AUX), "text source: (key, index) records with an aux byte");
		// Fast path (2-bit packed text, the warp's records away from both ends of the window and of
		// the text): lane l takes the RADIX_ITEMS CONSECUTIVE positions p0 + ITEMS*l + j.  Three words
		// give it the 64 symbols from one before its first position; every key, carried symbol and
		// preceding code is then a constant-distance bit field of that 128-bit value.  (The first pass
		// may take the records of a warp in any fixed order: nothing has been sorted yet, and the short
		// suffixes that must stay in front sit in a warp that takes the general path.)
		uint64_t const i0 = chunk - S.nshort;
		uint64_t p0 = S.v.wstart + i0;
		if (S.v.text_wraps && p0 >= S.v.ntext) p0 -= S.v.ntext;
		bool const fast = S.bits == 2 && S.v.packed && chunk >= S.nshort && i0 + 32 * RADIX_ITEMS + 35 <= S.v.W &&
		                  p0 >= 1 && p0 + 32 * RADIX_ITEMS + 35 <= S.v.ntext;
		tfast = fast;
		if (fast) {
			static_assert(RADIX_ITEMS <= 16, "a lane's records must fit one 64-symbol window");
			uint64_t const q = p0 + (uint64_t)RADIX_ITEMS * lane - 1; // symbol before the lane's first record
			const uint64_t * wp = S.v.packed + (q >> 5);
			unsigned const sh = (unsigned)(q & 31u) << 1;
			uint64_t const w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
			uint64_t const hi = sh ? ((w0 << sh) | (w1 >> (64u - sh))) : w0; // symbols q .. q+31
			uint64_t const lo = sh ? ((w1 << sh) | (w2 >> (64u - sh))) : w1; // symbols q+32 .. q+63
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				// record j: preceding code = symbol j, key = symbols j+1 .. j+16, carried = symbols j+17 .. j+19
				uint32_t const pred = (uint32_t)(hi >> (62 - 2 * j)) & 3u;
				k[j] = (uint32_t)(((hi << (2 * j + 2)) | (lo >> (62 - 2 * j))) >> 32);
				int const xo = 2 * j + 34; // bit offset of the carried symbols from the top of hi:lo
				uint32_t const k19 = xo + 6 <= 64 ? (uint32_t)(hi >> (58 - xo)) & 63u
				                   : (xo >= 64 ? (uint32_t)(lo >> (122 - xo)) & 63u
				                               : (uint32_t)((hi << (xo - 58)) | (lo >> (122 - xo))) & 63u);
				uint32_t const aa = (pred << 6) | k19;
				stash[j * RADIX_THREADS + threadIdx.x] = (uint8_t)aa;
			}
		} else {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				uint32_t kk = 0xffffffffu, ii = 0, aa = 0;
				if (RADIX_VALID(i)) radix_text_record(S, i, kk, ii, aa);
				k[j] = kk;
				stash[j * RADIX_THREADS + threadIdx.x] = (uint8_t)aa;
			}
		}
	} else {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			k[j] = RADIX_VALID(i) ? A.in[0][i] : 0xffffffffu;
		}
	}
	// stable rank inside the warp; records past n take the last bin and, being last in tile
	// order, rank behind every real record
	uint16_t slot[RADIX_ITEMS];
	uint16_t * mycnt = wcnt[w];
	unsigned const lt = lanemask_lt();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		uint32_t const d = RADIX_VALID(i) ? ((k[j] >> shift) & mask) : (uint32_t)(RADIX_BINS - 1);
		unsigned const peers = warp_peers8(d);
		uint32_t const before = mycnt[d];
		__syncwarp();
		if ((peers & lt) == 0) mycnt[d] = (uint16_t)(before + __popc(peers));
		__syncwarp();
		slot[j] = (uint16_t)(before + __popc(peers & lt));
	}
	__syncthreads();
	// per digit (thread d <-> bin d): scan over warps, publish the tile aggregate, scan over digits
	bool const binthread = threadIdx.x < RADIX_BINS;
	uint32_t bs = 0, bcnt = 0, bincl = 0;
	volatile unsigned long long * stw = status + (uint64_t)tile * RADIX_BINS + (threadIdx.x & (RADIX_BINS - 1));
	if (binthread) {
		uint32_t const d = threadIdx.x;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) { uint32_t const t = wcnt[ww][d]; wcnt[ww][d] = (uint16_t)bs; bs += t; }
		// invalid records were counted in the last bin: they are not part of the global count
		bcnt = (!FULL && d == RADIX_BINS - 1) ? bs - (RADIX_TILE - nvalid) : bs;
		*stw = (tile == 0 ? RADIX_FLAG_INC : RADIX_FLAG_AGG) | bcnt;
		bincl = bs;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t const t = __shfl_up_sync(0xffffffffu, bincl, o);
			if (lane >= (unsigned)o) bincl += t;
		}
		if (lane == 31) wsum[w] = bincl;
	}
	__syncthreads();
	if (binthread) {
		uint32_t const d = threadIdx.x;
		uint32_t add = 0;
		#pragma unroll
		for (int ww = 0; ww < RADIX_BINS / 32; ++ww) add += (ww < (int)w) ? wsum[ww] : 0u;
		uint32_t const dstart = bincl - bs + add;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) wcnt[ww][d] = (uint16_t)(wcnt[ww][d] + dstart);
		// decoupled look-back
		uint32_t excl = 0;
		if (tile > 0) {
			int64_t t = (int64_t)tile - 1;
			while (true) {
				unsigned long long const v = *(volatile unsigned long long *)(status + (uint64_t)t * RADIX_BINS + d);
				if ((v >> 62) == 0) continue;
				excl += (uint32_t)v;
				if ((v >> 62) == 2) break;
				--t;
			}
			*stw = RADIX_FLAG_INC | (unsigned long long)(excl + bcnt);
		}
		gbase[d] = base[d] + excl - dstart;
	}
	__syncthreads();
	if (TEXT) {
		// (the text pass computes its payload, so nothing is in flight: key and payload share one 64-bit scatter)
		// the first payload: loaded right before it is scattered together with its key
		uint32_t v[NA > 1 ? RADIX_ITEMS : 1];
		if (NA > 1) {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				if (TEXT) { // the record's window index
					uint64_t const t = tfast ? chunk + (uint64_t)RADIX_ITEMS * lane + j : i;
					v[j] = (uint32_t)((t < S.nshort) ? (S.v.W - 1 - t) : (t - S.nshort));
				}
				else v[j] = RADIX_VALID(i) ? A.in[1][i] : 0u;
			}
		}
		// key and first payload go to their sorted slot as ONE 64-bit word (half the scatter instructions, fewer
		// bank conflicts than two 32-bit scatters, and one barrier less than staging the arrays one after the other)
		unsigned long long * const srec = reinterpret_cast<unsigned long long *>(radix_dyn);
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			uint32_t const d = RADIX_VALID(i) ? ((k[j] >> shift) & mask) : (uint32_t)(RADIX_BINS - 1);
			slot[j] = (uint16_t)(slot[j] + wcnt[w][d]);
			if (NA > 1) srec[slot[j]] = ((unsigned long long)k[j] << 32) | v[j];
			else skey[slot[j]] = k[j];
		}
		__syncthreads();
		// the global place of sorted slot s is computed once (k[] is free now) and reused by every array of the record
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint32_t const s = j * RADIX_THREADS + threadIdx.x;
			if (FULL || s < nvalid) {
				if (NA > 1) {
					unsigned long long const r = srec[s];
					uint32_t const kk = (uint32_t)(r >> 32);
					uint32_t const o = gbase[(kk >> shift) & mask] + s;
					A.out[0][o] = kk;
					A.out[1][o] = (uint32_t)r;
					k[j] = o;
				} else {
					uint32_t const kk = skey[s];
					uint32_t const o = gbase[(kk >> shift) & mask] + s;
					A.out[0][o] = kk;
					k[j] = o;
				}
			}
		}
		#pragma unroll
		for (int a = 2; a < NA; ++a) {
			// further payloads reuse the (now dead) record buffer
			uint32_t * const sv = reinterpret_cast<uint32_t *>(radix_dyn);
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				v[j] = RADIX_VALID(i) ? A.in[a][i] : 0u;
			}
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) sv[slot[j]] = v[j];
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint32_t const s = j * RADIX_THREADS + threadIdx.x;
				if (FULL || s < nvalid) A.out[a][k[j]] = sv[s];
			}
		}
		if (AUX) {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				if (TEXT) saux[slot[j]] = stash[j * RADIX_THREADS + threadIdx.x];
				else saux[slot[j]] = RADIX_VALID(i) ? A.aux_in[i] : (uint8_t)0;
			}
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint32_t const s = j * RADIX_THREADS + threadIdx.x;
				if (FULL || s < nvalid) A.aux_out[k[j]] = saux[s];
			}
		}
	} else {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			uint32_t const d = RADIX_VALID(i) ? ((k[j] >> shift) & mask) : (uint32_t)(RADIX_BINS - 1);
			slot[j] = (uint16_t)(slot[j] + wcnt[w][d]);
			skey[slot[j]] = k[j];
		}
		// payload loads are issued before the barrier so that they overlap the key scatter
		uint32_t v[NA > 1 ? RADIX_ITEMS : 1];
		if (NA > 1) {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				if (TEXT) { // the record's window index
					uint64_t const t = tfast ? chunk + (uint64_t)RADIX_ITEMS * lane + j : i;
					v[j] = (uint32_t)((t < S.nshort) ? (S.v.W - 1 - t) : (t - S.nshort));
				}
				else v[j] = RADIX_VALID(i) ? A.in[1][i] : 0u;
			}
		}
		__syncthreads();
		// the global place of sorted slot s is computed once (k[] is free now) and reused by every array of the record
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint32_t const s = j * RADIX_THREADS + threadIdx.x;
			if (FULL || s < nvalid) {
				uint32_t const kk = skey[s];
				uint32_t const o = gbase[(kk >> shift) & mask] + s;
				A.out[0][o] = kk;
				k[j] = o;
			}
		}
		#pragma unroll
		for (int a = 1; a < NA; ++a) {
			if (a > 1) {
				__syncthreads();
				#pragma unroll
				for (int j = 0; j < RADIX_ITEMS; ++j) {
					uint64_t const i = chunk + j * 32 + lane;
					v[j] = RADIX_VALID(i) ? A.in[a][i] : 0u;
				}
			}
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) sval[slot[j]] = v[j];
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint32_t const s = j * RADIX_THREADS + threadIdx.x;
				if (FULL || s < nvalid) A.out[a][k[j]] = sval[s];
			}
		}
		if (AUX) {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				saux[slot[j]] = RADIX_VALID(i) ? A.aux_in[i] : (uint8_t)0;
			}
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint32_t const s = j * RADIX_THREADS + threadIdx.x;
				if (FULL || s < nvalid) A.aux_out[k[j]] = saux[s];
			}
		}
	}
	#undef RADIX_VALID
}

// radix_text_record with a short cut for the common case (2-bit packed text, the record away from
// both ends): one 64-bit window starting one symbol before the suffix holds the preceding code
// (2 bits), the key (32 bits) and the carried symbols (6 bits).
__device__ __forceinline__ void radix_text_record_any(RadixTextSrc const & S, uint64_t t, uint32_t & key, uint32_t & idx, uint32_t & aux) {
	if (S.bits == 2 && S.v.packed && t >= S.nshort) {
		uint64_t const i = t - S.nshort;
		uint64_t p = S.v.wstart + i;
		if (S.v.text_wraps && p >= S.v.ntext) p -= S.v.ntext;
		if (p >= 1 && i + 35 <= S.v.W && p + 35 <= S.v.ntext) {
			uint64_t const w = pk_window(S.v.packed, p - 1);
			key = (uint32_t)(w >> 30);
			idx = (uint32_t)i;
			aux = ((uint32_t)(w >> 62) << 6) | ((uint32_t)(w >> 24) & 63u);
			return;
		}
	}
	radix_text_record(S, t, key, idx, aux);
}

template <int NA, bool AUX>
constexpr size_t radix_smem_bytes() { return (size_t)RADIX_TILE * (4 + (NA > 1 ? 4 : 0) + (AUX ? 1 : 0)); }

// one pass: the full tiles, then the partial tile at the end (same ticket counter, so it is the last tile)
template <int NA, bool AUX, bool TEXT>
void radix_launch_pass(Stream & st, const char * label, uint64_t pbytes, RadixPassArgs<NA> const & A, RadixTextSrc const & S, uint64_t n,
                       int shift, uint32_t mask, const uint32_t * base, unsigned long long * status, uint32_t * ticket) {
	uint32_t const nfull = (uint32_t)(n / RADIX_TILE);
	bool const partial = (n % RADIX_TILE) != 0;
	size_t const smem = radix_smem_bytes<NA, AUX>() + (TEXT ? (size_t)RADIX_TILE : 0);
	static std::atomic<uint64_t> configured{0}; // per template instance and device
	if (first_on_device(configured)) {
		B3M_CUDA(cudaFuncSetAttribute(k_radix_onesweep<NA, AUX, TEXT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		B3M_CUDA(cudaFuncSetAttribute(k_radix_onesweep<NA, AUX, TEXT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	}
	if (st.kt.on) {
		KernelTimes::Rec r{label, pbytes, st.kt.get(), st.kt.get()};
		cudaEventRecord(r.a, st.s);
		if (nfull) k_radix_onesweep<NA, AUX, TEXT, true><<<nfull, RADIX_THREADS, smem, st.s>>>(A, S, n, shift, mask, base, status, ticket);
		if (partial) k_radix_onesweep<NA, AUX, TEXT, false><<<1, RADIX_THREADS, smem, st.s>>>(A, S, n, shift, mask, base, status, ticket);
		cudaEventRecord(r.b, st.s);
		st.kt.recs.push_back(r);
	} else {
		if (nfull) k_radix_onesweep<NA, AUX, TEXT, true><<<nfull, RADIX_THREADS, smem, st.s>>>(A, S, n, shift, mask, base, status, ticket);
		if (partial) k_radix_onesweep<NA, AUX, TEXT, false><<<1, RADIX_THREADS, smem, st.s>>>(A, S, n, shift, mask, base, status, ticket);
	}
	st.launches += (nfull ? 1 : 0) + (partial ? 1 : 0);
	B3M_CUDA(cudaGetLastError());
}

struct RadixStats {
	uint64_t passes = 0;
	uint64_t bytes = 0; // algorithmic bytes moved by all passes
};

// Sorts records by bits [bit_lo, bit_hi) of array ka (stable).  `cur` and `alt` are ping-pong
// buffers; on return `cur` names the arrays that hold the result.
template <int NA>
void radix_sort_bits(Stream & st, RadixRec<NA> & cur, RadixRec<NA> & alt, int ka, uint64_t n,
                     int bit_lo, int bit_hi, RadixStats * rs, const unsigned long long * pre_hist = nullptr /* [4][256] digit histograms of bits 0..31, if the caller has them */) {
	if (n == 0 || bit_hi <= bit_lo) return;
	uint32_t const ntiles = (uint32_t)div_up(n, RADIX_TILE);
	DevBuf<unsigned long long> status(st, (size_t)ntiles * RADIX_BINS);
	for (int lo = bit_lo; lo < bit_hi; lo += 8 * RADIX_MAXDIG) {
		int const hi = (bit_hi - lo) > 8 * RADIX_MAXDIG ? lo + 8 * RADIX_MAXDIG : bit_hi;
		int const ndig = (hi - lo + 7) / 8;
		int const lastbits = (hi - lo) - 8 * (ndig - 1);
		uint32_t const lastmask = (1u << lastbits) - 1u;
		DevBuf<unsigned long long> ghist(st, RADIX_MAXDIG * RADIX_BINS);
		DevBuf<uint32_t> base(st, RADIX_MAXDIG * RADIX_BINS + 8 + RADIX_MAXDIG);
		uint32_t * skip = base.get() + RADIX_MAXDIG * RADIX_BINS;
		uint32_t * ticket = skip + 4;
		B3M_CUDA(cudaMemsetAsync(ghist.get(), 0, ghist.bytes(), st.s));
		B3M_CUDA(cudaMemsetAsync(skip, 0, (8 + RADIX_MAXDIG) * sizeof(uint32_t), st.s));
		uint64_t const want = div_up(n, 256 * 16);
		unsigned const hgrid = (unsigned)(want < (uint64_t)st.sms * 8 ? want : (uint64_t)st.sms * 8);
		bool const have_hist = pre_hist && lo == 0 && hi == 32;
		if (!have_hist) {
			B3M_LAUNCH_T(st, "radix_hist", 4ull * n, k_radix_hist, hgrid, 256, 0, (const uint32_t *)cur.a[ka], n, lo, ndig, lastmask, ghist.get());
			if (rs) rs->bytes += 4ull * n;
		}
		B3M_LAUNCH(st, k_radix_hist_scan, 1, 256, 0, have_hist ? pre_hist : (const unsigned long long *)ghist.get(), ndig, n, base.get(), skip);
		uint32_t hskip[4];
		B3M_CUDA(cudaMemcpyAsync(hskip, skip, sizeof(hskip), cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		for (int d = 0; d < ndig; ++d) {
			if (hskip[d]) continue; // every record has the same digit: the pass would be the identity
			B3M_CUDA(cudaMemsetAsync(status.get(), 0, status.bytes(), st.s));
			RadixPassArgs<NA> A;
			A.in[0] = cur.a[ka]; A.out[0] = alt.a[ka];
			for (int a = 0, o = 1; a < NA; ++a) if (a != ka && o < NA) { A.in[o] = cur.a[a]; A.out[o] = alt.a[a]; ++o; }
			A.aux_in = cur.aux; A.aux_out = alt.aux;
			uint64_t const pbytes = n * (8ull * NA + (cur.aux ? 2ull : 0ull));
			if (cur.aux)
				radix_launch_pass<NA, true, false>(st, NA == 2 ? "radix_onesweep<2+aux>" : "radix_onesweep<+aux>", pbytes, A, RadixTextSrc(), n, lo + 8 * d,
				                                   (d == ndig - 1 ? lastmask : 255u), (const uint32_t *)(base.get() + d * RADIX_BINS), status.get(), ticket + d);
			else
				radix_launch_pass<NA, false, false>(st, NA == 2 ? "radix_onesweep<2>" : (NA == 3 ? "radix_onesweep<3>" : "radix_onesweep"), pbytes, A, RadixTextSrc(), n, lo + 8 * d,
				                                    (d == ndig - 1 ? lastmask : 255u), (const uint32_t *)(base.get() + d * RADIX_BINS), status.get(), ticket + d);
			RadixRec<NA> t = cur; cur = alt; alt = t;
			if (rs) { rs->passes++; rs->bytes += pbytes; }
		}
	}
}

// ---- first key straight from 2-bit packed text ----------------------------------------------
// Digit d of the key of suffix i is the 4-mer at window index i + 4*(3-d), so all four digit
// histograms come from ONE histogram of the window's 4-mers plus at most 12 corrections at
// either end (exactly equal for a circular window).
static __global__ void __launch_bounds__(256)
k_hist_4mers(TextView v, unsigned long long * __restrict__ ghist /* [256] */) {
	__shared__ uint32_t sh[RADIX_WARPS][RADIX_BINS];
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
	__syncthreads();
	uint32_t * my = sh[threadIdx.x >> 5];
	uint64_t const nchunks = div_up(v.W, 32);
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nchunks; q += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t const j0 = q * 32;
		uint64_t const p = tv_interior(v, j0, 35);
		if (p != ~0ull) {
			uint64_t const a = pk_window(v.packed, p), b = pk_window(v.packed, p + 32);
			#pragma unroll
			for (int s = 0; s < 29; ++s) atomicAdd(&my[(uint32_t)(a >> (56 - 2 * s)) & 255u], 1u);
			#pragma unroll
			for (int s = 29; s < 32; ++s) atomicAdd(&my[(uint32_t)((a << (2 * s - 56)) | (b >> (120 - 2 * s))) & 255u], 1u);
		} else {
			for (uint64_t j = j0; j < j0 + 32 && j < v.W; ++j) atomicAdd(&my[(uint32_t)tv_symbols(v, j, 4, 2)], 1u);
		}
	}
	__syncthreads();
	for (int d = threadIdx.x; d < RADIX_BINS; d += blockDim.x) {
		uint32_t c = 0;
		#pragma unroll
		for (int w = 0; w < RADIX_WARPS; ++w) c += sh[w][d];
		if (c) atomicAdd(&ghist[d], (unsigned long long)c);
	}
}

// ghist[0][*] = histogram of the 4-mers at window indices [0, W)  ->  ghist[d][*], d = 0..3:
// digit d counts the 4-mers at indices [off, off + W), off = 4*(3-d)
static __global__ void __launch_bounds__(256) k_hist_4mers_fix(TextView v, unsigned long long * __restrict__ ghist) {
	unsigned long long const h0 = ghist[threadIdx.x];
	for (int d = 0; d < 3; ++d) ghist[d * RADIX_BINS + threadIdx.x] = h0;
	ghist[3 * RADIX_BINS + threadIdx.x] = h0;
	__syncthreads();
	if (threadIdx.x >= 3 || v.circular) return; // circular: every range of W indices holds the same 4-mers
	unsigned const d = threadIdx.x;
	unsigned long long * out = ghist + d * RADIX_BINS;
	uint64_t const off = 4 * (3 - d);
	uint64_t const cut = off < v.W ? off : v.W;
	for (uint64_t j = 0; j < cut; ++j) out[tv_symbols(v, j, 4, 2)] -= 1;
	for (uint64_t j = (off > v.W ? off : v.W); j < off + v.W; ++j) out[tv_symbols(v, j, 4, 2)] += 1;
}

// (key, index, aux) records of the W suffixes of window v, sorted by key.  `cur`/`alt` as in
// radix_sort_bits; the records are created by the first pass, so `cur` needs no initialisation.
inline void radix_sort_suffix_keys(Stream & st, TextView const & v, uint64_t nshort, unsigned k0, RadixRec<2> & cur, RadixRec<2> & alt, RadixStats * rs) {
	uint64_t const n = v.W;
	uint32_t const ntiles = (uint32_t)div_up(n, RADIX_TILE);
	DevBuf<unsigned long long> status(st, (size_t)ntiles * RADIX_BINS);
	DevBuf<unsigned long long> ghist(st, RADIX_MAXDIG * RADIX_BINS);
	DevBuf<uint32_t> base(st, RADIX_MAXDIG * RADIX_BINS + 8 + RADIX_MAXDIG);
	uint32_t * skip = base.get() + RADIX_MAXDIG * RADIX_BINS;
	uint32_t * ticket = skip + 4;
	B3M_CUDA(cudaMemsetAsync(ghist.get(), 0, ghist.bytes(), st.s));
	B3M_CUDA(cudaMemsetAsync(skip, 0, (8 + RADIX_MAXDIG) * sizeof(uint32_t), st.s));
	uint64_t const want = div_up(div_up(n, 32), 256 * 4);
	unsigned const hgrid = (unsigned)(want < (uint64_t)st.sms * 8 ? (want ? want : 1) : (uint64_t)st.sms * 8);
	B3M_LAUNCH_T(st, "hist_4mers", n / 4, k_hist_4mers, hgrid, 256, 0, v, ghist.get());
	B3M_LAUNCH(st, k_hist_4mers_fix, 1, 256, 0, v, ghist.get());
	B3M_LAUNCH(st, k_radix_hist_scan, 1, 256, 0, (const unsigned long long *)ghist.get(), 4, n, base.get(), skip);
	if (rs) rs->bytes += n / 4;
	uint32_t hskip[4];
	B3M_CUDA(cudaMemcpyAsync(hskip, skip, sizeof(hskip), cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	RadixTextSrc S{v, nshort, 2u, k0};
	for (int d = 0; d < 4; ++d) {
		if (d && hskip[d]) continue; // the first pass creates the records and always runs
		B3M_CUDA(cudaMemsetAsync(status.get(), 0, status.bytes(), st.s));
		RadixPassArgs<2> A;
		A.in[0] = cur.a[0]; A.out[0] = alt.a[0]; A.in[1] = cur.a[1]; A.out[1] = alt.a[1];
		A.aux_in = cur.aux; A.aux_out = alt.aux;
		uint64_t const pbytes = d ? n * 18ull : n * 9ull + n / 4;
		if (d == 0)
			radix_launch_pass<2, true, true>(st, "radix_onesweep<text>", pbytes, A, S, n, 0, 255u, (const uint32_t *)base.get(), status.get(), ticket);
		else
			radix_launch_pass<2, true, false>(st, "radix_onesweep<2+aux>", pbytes, A, S, n, 8 * d, 255u, (const uint32_t *)(base.get() + d * RADIX_BINS), status.get(), ticket + d);
		RadixRec<2> t = cur; cur = alt; alt = t;
		if (rs) { rs->passes++; rs->bytes += pbytes; }
	}
}

} // namespace b3m
