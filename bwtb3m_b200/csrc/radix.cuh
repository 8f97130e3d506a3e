// LSD radix sort on records held as NA parallel uint32 arrays (structure of arrays), 8-bit
// digits, one kernel per digit ("onesweep"): each CTA takes the next tile from a ticket counter,
// ranks its records stably in shared memory (8 ballots per record find equal digits inside a
// warp), obtains its global offsets by decoupled look-back over the tile status words of its
// predecessors, reorders through shared memory and writes coalesced runs.
//
// Why not __match_any_sync: MATCH runs on the ADU pipe; a first version measured 98% ADU
// utilisation and 7% of HBM peak (profiles/r01_radix_match_any.txt).
// Forward progress: tickets are handed out in launch order, so every predecessor of a tile is
// already resident or finished when the tile starts to look back.
//
// Algorithmic HBM bytes per pass and record: 4*NA read + 4*NA write, plus one
// 4-byte key read per radix_sort_bits() call for the digit histograms.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "textview.cuh"

namespace b3m {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;
constexpr int RADIX_ITEMS = 16;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS; // 4096 records
constexpr int RADIX_BINS = 256;
constexpr int RADIX_MAXDIG = 4;

template <int NA>
struct RadixRec {
	uint32_t * a[NA];
	uint8_t * aux = nullptr; // optional one-byte payload that travels with the record
};

// lanes holding the same 8-bit digit
__device__ __forceinline__ unsigned warp_peers8(uint32_t d) {
	unsigned peers = 0xffffffffu;
	#pragma unroll
	for (int b = 0; b < 8; ++b) {
		bool const bit = (d >> b) & 1u;
		unsigned const m = __ballot_sync(0xffffffffu, bit);
		peers &= bit ? m : ~m;
	}
	return peers;
}

// ---- digit histograms of one key array: up to 4 digits in one read -----------------------
__global__ void __launch_bounds__(256)
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
__global__ void __launch_bounds__(256)
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

template <int NA, bool AUX, bool TEXT>
__global__ void __launch_bounds__(RADIX_THREADS, 3)
k_radix_onesweep(RadixPassArgs<NA> A, RadixTextSrc S, uint64_t n, int shift, uint32_t mask, const uint32_t * __restrict__ base /* [256] */,
                 unsigned long long * __restrict__ status /* [ntiles][256] */, uint32_t * __restrict__ ticket) {
	__shared__ uint32_t wcnt[RADIX_WARPS][RADIX_BINS];
	__shared__ uint32_t gbase[RADIX_BINS];
	__shared__ uint32_t skey[RADIX_TILE];
	__shared__ uint32_t sval[RADIX_TILE];
	__shared__ uint8_t saux[AUX ? RADIX_TILE : 4];
	__shared__ uint32_t s_tile;
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS; i += RADIX_THREADS) (&wcnt[0][0])[i] = 0;
	__syncthreads();
	uint32_t const tile = s_tile;
	uint64_t const tbase = (uint64_t)tile * RADIX_TILE;
	uint64_t const chunk = tbase + (uint64_t)w * (32 * RADIX_ITEMS);
	uint32_t const nvalid = (n - tbase) < (uint64_t)RADIX_TILE ? (uint32_t)(n - tbase) : (uint32_t)RADIX_TILE;

	uint32_t k[RADIX_ITEMS];
	uint32_t taux[TEXT ? RADIX_ITEMS / 4 : 1];
	if (TEXT) {
		static_assert(!TEXT || (NA == 2 && AUX), "text source: (key, index) records with an aux byte");
		// Fast path (2-bit packed text, the warp's 512 records away from both ends of the window and
		// of the text): lane l reads the words under positions p0+l, p0+l+32, ... -- its bit offset
		// inside a word never changes, and every word is the second half of the previous record's window.
		uint64_t const i0 = chunk - S.nshort;
		uint64_t p0 = S.v.wstart + i0;
		if (S.v.text_wraps && p0 >= S.v.ntext) p0 -= S.v.ntext;
		bool const fast = S.bits == 2 && S.v.packed && chunk >= S.nshort && i0 + 32 * RADIX_ITEMS + 35 <= S.v.W &&
		                  p0 >= 1 && p0 + 32 * RADIX_ITEMS + 35 <= S.v.ntext;
		if (fast) {
			uint64_t const pl = p0 + lane;
			const uint64_t * wp = S.v.packed + (pl >> 5);
			unsigned const sh = (unsigned)(pl & 31u) << 1;
			uint64_t prevw = (sh == 0) ? __ldg(wp - 1) : 0ull; // pl >= 32 whenever sh == 0 (p0 >= 1)
			uint64_t cw = __ldg(wp);
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const nw = __ldg(wp + j + 1);
				uint64_t const win = sh ? ((cw << sh) | (nw >> (64u - sh))) : cw;
				uint32_t const k19 = (uint32_t)(win >> 26) & 63u;
				uint32_t const pred = sh ? (uint32_t)(cw >> (64u - sh)) & 3u : (uint32_t)prevw & 3u;
				k[j] = (uint32_t)(win >> 32);
				uint32_t const aa = (pred << 6) | k19;
				if ((j & 3) == 0) taux[TEXT ? j / 4 : 0] = aa; else taux[TEXT ? j / 4 : 0] |= aa << (8 * (j & 3));
				prevw = cw; cw = nw;
			}
		} else {
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				uint32_t kk = 0xffffffffu, ii = 0, aa = 0;
				if (i < n) radix_text_record(S, i, kk, ii, aa);
				k[j] = kk;
				if ((j & 3) == 0) taux[TEXT ? j / 4 : 0] = aa; else taux[TEXT ? j / 4 : 0] |= aa << (8 * (j & 3));
			}
		}
	} else {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			k[j] = (i < n) ? A.in[0][i] : 0xffffffffu;
		}
	}
	// stable rank inside the warp; records past n take the last bin and, being last in tile
	// order, rank behind every real record
	uint16_t slot[RADIX_ITEMS];
	uint32_t * mycnt = wcnt[w];
	unsigned const lt = lanemask_lt();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		uint32_t const d = (i < n) ? ((k[j] >> shift) & mask) : (uint32_t)(RADIX_BINS - 1);
		unsigned const peers = warp_peers8(d);
		uint32_t const before = mycnt[d];
		__syncwarp();
		if ((peers & lt) == 0) mycnt[d] = before + __popc(peers);
		__syncwarp();
		slot[j] = (uint16_t)(before + __popc(peers & lt));
	}
	__syncthreads();
	// per digit (thread d <-> bin d): scan over warps, publish the tile aggregate, scan over digits
	{
		uint32_t const d = threadIdx.x;
		uint32_t s = 0;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) { uint32_t const t = wcnt[ww][d]; wcnt[ww][d] = s; s += t; }
		// invalid records were counted in the last bin: they are not part of the global count
		uint32_t const cnt = (d == RADIX_BINS - 1) ? s - (RADIX_TILE - nvalid) : s;
		volatile unsigned long long * st = status + (uint64_t)tile * RADIX_BINS + d;
		*st = (tile == 0 ? RADIX_FLAG_INC : RADIX_FLAG_AGG) | cnt;
		uint32_t total;
		uint32_t const incl = block_scan_inclusive<OpSum>(s, &total);
		uint32_t const dstart = incl - s;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) wcnt[ww][d] += dstart;
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
			*st = RADIX_FLAG_INC | (unsigned long long)(excl + cnt);
		}
		gbase[d] = base[d] + excl - dstart;
	}
	__syncthreads();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint64_t const i = chunk + j * 32 + lane;
		uint32_t const d = (i < n) ? ((k[j] >> shift) & mask) : (uint32_t)(RADIX_BINS - 1);
		slot[j] = (uint16_t)(slot[j] + wcnt[w][d]);
		skey[slot[j]] = k[j];
	}
	// payload loads are issued before the barrier so that they overlap the key scatter
	uint32_t v[NA > 1 ? RADIX_ITEMS : 1];
	if (NA > 1) {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			if (TEXT) v[j] = (uint32_t)((i < S.nshort) ? (S.v.W - 1 - i) : (i - S.nshort)); // the record's window index
			else v[j] = (i < n) ? A.in[1][i] : 0u;
		}
	}
	__syncthreads();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint32_t const s = j * RADIX_THREADS + threadIdx.x;
		if (s < nvalid) {
			uint32_t const kk = skey[s];
			A.out[0][gbase[(kk >> shift) & mask] + s] = kk;
		}
	}
	#pragma unroll
	for (int a = 1; a < NA; ++a) {
		if (a > 1) {
			__syncthreads();
			#pragma unroll
			for (int j = 0; j < RADIX_ITEMS; ++j) {
				uint64_t const i = chunk + j * 32 + lane;
				v[j] = (i < n) ? A.in[a][i] : 0u;
			}
		}
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) sval[slot[j]] = v[j];
		__syncthreads();
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint32_t const s = j * RADIX_THREADS + threadIdx.x;
			if (s < nvalid) A.out[a][gbase[(skey[s] >> shift) & mask] + s] = sval[s];
		}
	}
	if (AUX) {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint64_t const i = chunk + j * 32 + lane;
			if (TEXT) saux[slot[j]] = (uint8_t)(taux[TEXT ? j / 4 : 0] >> (8 * (j & 3)));
			else saux[slot[j]] = (i < n) ? A.aux_in[i] : (uint8_t)0;
		}
		__syncthreads();
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) {
			uint32_t const s = j * RADIX_THREADS + threadIdx.x;
			if (s < nvalid) A.aux_out[gbase[(skey[s] >> shift) & mask] + s] = saux[s];
		}
	}
}

struct RadixStats {
	uint64_t passes = 0;
	uint64_t bytes = 0; // algorithmic bytes moved by all passes
};

// Sorts records by bits [bit_lo, bit_hi) of array ka (stable).  `cur` and `alt` are ping-pong
// buffers; on return `cur` names the arrays that hold the result.
template <int NA>
void radix_sort_bits(Stream & st, RadixRec<NA> & cur, RadixRec<NA> & alt, int ka, uint64_t n,
                     int bit_lo, int bit_hi, RadixStats * rs) {
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
		B3M_LAUNCH_T(st, "radix_hist", 4ull * n, k_radix_hist, hgrid, 256, 0, (const uint32_t *)cur.a[ka], n, lo, ndig, lastmask, ghist.get());
		B3M_LAUNCH(st, k_radix_hist_scan, 1, 256, 0, (const unsigned long long *)ghist.get(), ndig, n, base.get(), skip);
		if (rs) rs->bytes += 4ull * n;
		uint32_t hskip[4];
		B3M_CUDA(cudaMemcpyAsync(hskip, skip, sizeof(hskip), cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		for (int d = 0; d < ndig; ++d) {
			if (hskip[d]) continue; // every record has the same digit: the pass would be the identity
			B3M_CUDA(cudaMemsetAsync(status.get(), 0, status.bytes(), st.s));
			RadixPassArgs<NA> A;
			A.in[0] = cur.a[ka]; A.out[0] = alt.a[ka];
			for (int a = 0, o = 1; a < NA; ++a) if (a != ka) { A.in[o] = cur.a[a]; A.out[o] = alt.a[a]; ++o; }
			A.aux_in = cur.aux; A.aux_out = alt.aux;
			uint64_t const pbytes = n * (8ull * NA + (cur.aux ? 2ull : 0ull));
			if (cur.aux)
				B3M_LAUNCH_T(st, NA == 2 ? "radix_onesweep<2+aux>" : "radix_onesweep<+aux>", pbytes,
				           (k_radix_onesweep<NA, true, false>), ntiles, RADIX_THREADS, 0, A, RadixTextSrc(), n, lo + 8 * d,
				           (d == ndig - 1 ? lastmask : 255u), (const uint32_t *)(base.get() + d * RADIX_BINS), status.get(), ticket + d);
			else
				B3M_LAUNCH_T(st, NA == 2 ? "radix_onesweep<2>" : (NA == 3 ? "radix_onesweep<3>" : "radix_onesweep"), pbytes,
				           (k_radix_onesweep<NA, false, false>), ntiles, RADIX_THREADS, 0, A, RadixTextSrc(), n, lo + 8 * d,
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
__global__ void __launch_bounds__(256)
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
__global__ void __launch_bounds__(256) k_hist_4mers_fix(TextView v, unsigned long long * __restrict__ ghist) {
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
			B3M_LAUNCH_T(st, "radix_onesweep<text>", pbytes, (k_radix_onesweep<2, true, true>), ntiles, RADIX_THREADS, 0, A, S, n, 0, 255u,
			           (const uint32_t *)base.get(), status.get(), ticket);
		else
			B3M_LAUNCH_T(st, "radix_onesweep<2+aux>", pbytes, (k_radix_onesweep<2, true, false>), ntiles, RADIX_THREADS, 0, A, S, n, 8 * d, 255u,
			           (const uint32_t *)(base.get() + d * RADIX_BINS), status.get(), ticket + d);
		RadixRec<2> t = cur; cur = alt; alt = t;
		if (rs) { rs->passes++; rs->bytes += pbytes; }
	}
}

} // namespace b3m
