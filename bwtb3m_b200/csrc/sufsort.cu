// K2: per-block suffix sorting.  Replaces libmaus2's CPU block sorter reached through
// BwtMergeBlockSortRequest::dispatch (/root/reference/src/checkbwt.cpp:24);
// comparisons are circular exactly as in the reference's definition
// BWT[i] = s[(SA[i]+n-1)%n] (/root/reference/src/lcpbit.cpp:3668-3669).
//
// Round 0   LSD radix sort (radix.cuh) of all W suffixes by a 32-bit key holding their first
//           k0 = 32/keybits symbols; every record also carries its index and an aux byte (the
//           symbol before the suffix + the symbols right behind the key).  For 2-bit alphabets
//           the first pass makes the records from the packed text.
// Resolve   k_resolve: one pass over the sorted records.  Runs of equal keys of up to RS_EXT
//           suffixes are sorted inside the CTA by the symbols carried in the aux byte and, where
//           those tie, by their next 64 key bits read from the text.  With the whole text in one
//           window the same pass emits the BWT, the anchors and the sampled SA/ISA (FusedOut),
//           in chunks, so that finished rows can leave for the host while later chunks run
//           (StreamOut).  On random DNA this finishes the sort: no rank-by-position array is
//           ever written.
// Doubling  only if some suffixes are still tied (repeats longer than k0 + nx + 64/keybits
//           symbols, or runs longer than RS_EXT): classic prefix doubling on the tied suffixes
//           only -- gather the rank of the suffix h symbols ahead, radix sort (group, rank
//           ahead), split the groups; h doubles.
// Sharding  k2_keyrange_plan / k2_sort_keyrange: the same on one key range of the suffixes
//           (multi-GPU, one range per rank).
#include "kernels.h"
#include "radix.cuh"
#include "scan.cuh"
#include "textview.cuh"
#include "msd.cuh"
#include <time.h>
#include <algorithm>
#include <vector>
#include <stdlib.h>

namespace b3m {

// Input order of round 0: in linear mode the nshort suffixes that run past the window end come
// first, shortest first, so that the stable sort leaves them in front of equal padded keys.
// aux byte of a record: the code preceding the suffix (what K3 needs) in the top `bits` bits and
// the nx = 8/bits - 1 symbols that follow the key below it (they extend the key without a gather)
__global__ void __launch_bounds__(256)
k_make_keys(TextView v, unsigned bits, unsigned k0, uint64_t nshort, uint32_t * __restrict__ key, uint32_t * __restrict__ idx,
            uint8_t * __restrict__ aux) {
	uint64_t const t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= v.W) return;
	uint64_t const i = (t < nshort) ? (v.W - 1 - t) : (t - nshort);
	unsigned const nx = 8u / bits - 1u;
	uint64_t const ks = tv_symbols(v, i, k0 + nx, bits);
	key[t] = (uint32_t)(ks >> (nx * bits));
	idx[t] = (uint32_t)i;
	aux[t] = (uint8_t)((tv_pred(v, i) << (nx * bits)) | (uint32_t)(ks & ((1u << (nx * bits)) - 1u)));
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

__global__ void __launch_bounds__(256)
k_rank_scatter(const uint32_t * __restrict__ sa, uint64_t W, uint32_t * __restrict__ rank) {
	uint64_t const k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k < W) rank[sa[k]] = (uint32_t)k;
}

// ------------------------------------------------------------------------------------------
// fused outputs of one suffix in its final place (K3 + sampling)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fo_pred(TextView const & v, FusedOut const & fo, uint32_t i) {
	if (i == 0) return fo.has_term ? 0u : tv_code(v, v.ntext - 1);
	return tv_code_before(v, i);
}

__device__ __forceinline__ void fo_emit(FusedOut const & fo, uint32_t i, uint64_t k, uint32_t c) {
	// ranks and positions are below 2^32 (the engine rejects longer texts)
	uint32_t const r = (uint32_t)(k + fo.shift);
	fo.bwt[r] = (uint8_t)c;
	if (i == 0) { if (fo.has_term) fo.special[0] = r; fo.special[1] = r; }
	if (fo.prelog >= 32 ? i == 0 : (i & ((1u << fo.prelog) - 1u)) == 0) fo.prerank[fo.prelog >= 32 ? 0 : (i >> fo.prelog)] = r;
	if (fo.isa_s && (fo.isalog >= 32 ? i == 0 : (i & ((1u << fo.isalog) - 1u)) == 0)) fo.isa_s[fo.isalog >= 32 ? 0 : (i >> fo.isalog)] = r;
	if (fo.sa_s && (fo.salog >= 32 ? r == 0 : (r & ((1u << fo.salog) - 1u)) == 0)) fo.sa_s[fo.salog >= 32 ? 0 : (r >> fo.salog)] = i;
}

// K3 on a final suffix array of the whole text (the path taken after prefix doubling)
__global__ void __launch_bounds__(256)
k_extract_sample(TextView v, const uint32_t * __restrict__ sa, uint64_t m, FusedOut fo) {
	uint64_t const k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	uint32_t const i = sa[k];
	fo_emit(fo, i, k, fo_pred(v, fo, i));
}

// ------------------------------------------------------------------------------------------
// Resolve: sorts every run of equal round-0 keys that is at most RS_EXT long inside one CTA.
// A CTA owns the runs that START in its tile; it looks at RS_EXT records on either side so that
// every suffix can find the start and the end of its run: the head flags of the region are kept
// as ballot words, and a 64-bit window around a record gives both ends with one clz / ffs.
// Order inside a run: (1) the symbols carried in the aux byte; only records that tie on those
// (about 1 in 100 on random DNA) read (2) their next 64/bits symbols from the text; records
// that reach the sentinel of a linear window inside that range compare by (3) remaining length
// (shorter = smaller), which is also what keeps equal zero-padded keys apart.  Longer runs and
// records equal on all three stay in their current order and are counted; hflag[k] = 1 where
// the record in final place k starts a new group.
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_EXT = 32;
constexpr int RS_TROWS = 64;                      // tile rows of 32 records
constexpr int RS_TILE = RS_TROWS * 32;
constexpr int RS_ROWS = RS_TROWS + 2;             // + one row on either side
constexpr int RS_REG = RS_ROWS * 32;
constexpr int RS_CSLOTS = 256;

// second key of the suffix at window index i: its symbols from k0+nx on, and its length if it
// reaches the sentinel of a linear window inside that range
__device__ __forceinline__ void rs_second_key(TextView const & v, unsigned bits, unsigned skip, int lin, uint32_t i, unsigned long long & k2, uint32_t & rem) {
	unsigned const k2syms = 64u / bits;
	k2 = tv_symbols(v, (uint64_t)i + skip, k2syms, bits);
	uint64_t const left = v.W - i;
	rem = (uint32_t)((lin && left < skip + k2syms) ? left : skip + k2syms);
}

// ORDER: write the resolved order (suffix array + head flags); a fused whole-text sort leaves it out
// and only comes back for it when something stayed unresolved.
template <bool FUSED, bool ORDER>
__global__ void __launch_bounds__(RS_THREADS)
k_resolve(TextView v, unsigned bits, unsigned k0, int lin, const uint32_t * __restrict__ key, const uint32_t * __restrict__ idx,
          const uint8_t * __restrict__ aux, uint64_t nrec, uint32_t tile0, uint32_t * __restrict__ sa_out, uint8_t * __restrict__ hflag, FusedOut fo, unsigned long long * __restrict__ counters) {
	__shared__ uint32_t s_idx[RS_REG];
	__shared__ uint32_t s_hb[RS_ROWS + 2];        // head flags of row q in s_hb[q + 1]
	__shared__ uint8_t s_aux[RS_REG];
	__shared__ uint32_t s_cnt[3];
	// all indices fit 32 bits (windows of 2^32-16 suffixes at most); places before the first record
	// wrap around to huge values and read as invalid
	uint32_t const NR = (uint32_t)nrec;                       // records to resolve (all W suffixes, or the ones of one key range)
	bool const allshort = v.W < (uint64_t)k0;
	uint32_t const shortlim = allshort ? 0u : (uint32_t)(v.W - k0); // suffix i reaches the sentinel inside its first key iff i > shortlim
	uint32_t const tile = blockIdx.x + tile0;
	uint32_t const kbase = tile * (uint32_t)RS_TILE - (uint32_t)RS_EXT; // region index x <-> global place kbase + x
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned const nx = 8u / bits - 1u;          // symbols carried in the aux byte behind the key
	unsigned const xbits = nx * bits;
	unsigned const xmask = (1u << xbits) - 1u;
	if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
	if (threadIdx.x == 0) { s_hb[0] = 0xffffffffu; s_hb[RS_ROWS + 1] = 0xffffffffu; }

	#pragma unroll 1
	for (unsigned q = w; q < RS_ROWS; q += RS_WARPS) {
		unsigned const x = q * 32 + lane;
		uint32_t const k = kbase + x;
		bool const valid = k < NR;
		uint32_t mk = 0, mi = 0, ma = 0;
		if (valid) { mk = key[k]; mi = idx[k]; ma = aux[k]; }
		s_idx[x] = mi;
		s_aux[x] = (uint8_t)ma;
		// the record before this one: the lane below, or one extra load for lane 0
		uint32_t pk = __shfl_up_sync(0xffffffffu, mk, 1), pi = __shfl_up_sync(0xffffffffu, mi, 1);
		if (lane == 0 && k - 1u < NR) { pk = key[k - 1]; pi = idx[k - 1]; }
		bool head = !valid || k == 0 || mk != pk;
		if (lin) head = head || allshort || mi > shortlim || pi > shortlim;
		uint32_t const hb = __ballot_sync(0xffffffffu, head);
		if (lane == 0) s_hb[q + 1] = hb;
	}
	__syncthreads();

	uint32_t ntied = 0, nunres = 0, ngather = 0;
	#pragma unroll 1
	for (unsigned q = 1 + w; q < RS_ROWS; q += RS_WARPS) {
		int const x = (int)(q * 32 + lane);
		if (kbase + (uint32_t)x >= NR) continue;
		// run start: highest head bit at or below x inside the 64 records that end with this row
		uint32_t const h0 = s_hb[q], h1 = s_hb[q + 1], h2 = s_hb[q + 2];
		uint32_t const b1 = h1 & (0xffffffffu >> (31 - lane));
		int const y = b1 ? (int)(q * 32) + 31 - __clz((int)b1) : (h0 ? (int)(q * 32) - 1 - __clz((int)h0) : -1);
		// run end: lowest head bit above x inside the 64 records that start with this row
		uint32_t const a1 = lane == 31 ? 0u : (h1 >> (lane + 1));
		int const z = a1 ? x + __ffs((int)a1) : (h2 ? (int)(q * 32) + 31 + __ffs((int)h2) : RS_REG + RS_EXT);
		bool const big = y < 0 || x - y >= RS_EXT || z - y > RS_EXT;
		uint32_t const ax = s_aux[x];
		int f = x;
		uint32_t hf = 1;
		if (big) {
			if (q > RS_TROWS) continue;                        // the next tile passes it through
			hf = (h1 >> lane) & 1u;
			++nunres;
		} else {
			if (y < RS_EXT || y >= RS_EXT + RS_TILE) continue; // run of a neighbouring tile
			if (z - y > 1) {
				uint32_t const mx = ax & xmask;
				uint32_t less = 0, eq = 0;
				#pragma unroll 1
				for (int y2 = y; y2 < z; ++y2) {
					uint32_t const ox = s_aux[y2] & xmask;
					less += ox < mx ? 1u : 0u;
					eq += ox == mx ? 1u : 0u;
				}
				++ntied;
				f = y + (int)less;
				if (eq > 1) {
					// the carried symbols do not separate this record from the rest of its run: compare the
					// second keys, read from the text (about 1 record in 100 on random DNA)
					unsigned long long mk2; uint32_t mr;
					rs_second_key(v, bits, k0 + nx, lin, s_idx[x], mk2, mr);
					uint32_t eqb = 0, eqa = 0;
					#pragma unroll 1
					for (int y2 = y; y2 < z; ++y2) {
						if (y2 == x || (s_aux[y2] & xmask) != mx) continue;
						unsigned long long ok2; uint32_t orr;
						rs_second_key(v, bits, k0 + nx, lin, s_idx[y2], ok2, orr);
						bool const same = ok2 == mk2 && orr == mr;
						f += (ok2 < mk2 || (ok2 == mk2 && orr < mr) || (same && y2 < x)) ? 1 : 0;
						eqb += (same && y2 < x) ? 1u : 0u;
						eqa += same ? 1u : 0u;
					}
					hf = eqb == 0 ? 1u : 0u;
					if (eqa) ++nunres;
					++ngather;
				}
			}
		}
		uint32_t const i = s_idx[x];
		uint32_t const kf = kbase + (uint32_t)f;
		if (ORDER) { sa_out[kf] = i; hflag[kf] = (uint8_t)hf; }
		if (FUSED) fo_emit(fo, i, (uint64_t)kf, ax >> xbits);
	}
	// per-CTA totals, spread over RS_CSLOTS counter sets (one hot address would serialise in L2)
	ntied = __reduce_add_sync(0xffffffffu, ntied);
	nunres = __reduce_add_sync(0xffffffffu, nunres);
	ngather = __reduce_add_sync(0xffffffffu, ngather);
	if (lane == 0) {
		if (nunres) atomicAdd(&s_cnt[0], nunres);
		if (ntied) atomicAdd(&s_cnt[1], ntied);
		if (ngather) atomicAdd(&s_cnt[2], ngather);
	}
	__syncthreads();
	if (threadIdx.x < 3 && s_cnt[threadIdx.x])
		atomicAdd(&counters[(tile % RS_CSLOTS) * 4 + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// first place whose key is not below `key0` (the records are sorted by key)
__global__ void k_lower_bound(const uint32_t * __restrict__ key, uint32_t n, uint32_t key0, uint32_t * __restrict__ out) {
	uint32_t lo = 0, hi = n;
	while (lo < hi) { uint32_t const mid = lo + ((hi - lo) >> 1); if (key[mid] < key0) lo = mid + 1; else hi = mid; }
	*out = lo;
}

static double wall_ms() {
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#define RS_LAUNCH(label, bytes_, F, O, grid, ...) B3M_LAUNCH_T(st, label, bytes_, (k_resolve<F, O>), grid, RS_THREADS, 0, __VA_ARGS__)
static bool trace_on() { static int t = -1; if (t < 0) t = getenv("B3M_TRACE") ? 1 : 0; return t == 1; }
#define TRACE(msg) do { if (trace_on()) { cudaStreamSynchronize(st.s); double t_ = wall_ms(); fprintf(stderr, "[T] %-28s %9.3f ms\n", msg, t_ - t_last); t_last = wall_ms(); } } while (0)

static uint32_t fetch_u32(Stream & st, const uint32_t * d) {
	uint32_t h = 0;
	B3M_CUDA(cudaMemcpyAsync(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	return h;
}

// ------------------------------------------------------------------------------------------
// MSD path (msd.cuh): round 0 of a whole-text sort of a 2-bit alphabet, on all suffixes or on the
// key range [d_lo, d_hi) of level-1 bins (suffix-range sharding).
// ------------------------------------------------------------------------------------------
struct MsdGeom { unsigned b1 = 0, b2 = 0; };

// bits of the two global levels: sub-buckets of 3200..6400 suffixes on average (a finish CTA holds MSD_CAP = 8192
// records including the padding of its runs)
// nparts > 1 (position-sharded build): the level-1 runs of a tile cross NVLink as they are stored, and 64-byte runs use
// a third of the link; one bit less doubles them (measured on 2 GPUs at 3.1 Gbp: scatter 21.1 -> 16.2 ms, finish -- twice
// the runs to gather -- 15.8 -> 18.9 ms).  8 bits do not work: a finish CTA holds its sub-bucket plus two padding
// records per tile of the parent bucket, and a bucket of 1/256 of a 3.1 Gbp text has 1478 tiles (MSD_CAP)
static bool msd_geometry(Stream const & st, DevText const & T, uint64_t W, MsdGeom & g, uint32_t nparts = 1) {
	if (st.sortpath == B3M_SORT_LSD) return false;
	if (T.keybits != 2 || !T.packed || W < 64 || W >= 0xFFFFFF00ull) return false;
	if (st.sortpath != B3M_SORT_MSD && W < (1u << 16)) return false;
	unsigned tb = 4;
	while (tb < 21 && (W >> tb) > 6400) ++tb;
	g.b1 = 2 * ((tb + 3) / 4);
	if (g.b1 > 10) g.b1 = 10;
	if (nparts > 1 && g.b1 > 9 && tb - 9 <= 11) g.b1 = 9;
	g.b2 = tb - g.b1;
	if (g.b2 < 1) g.b2 = 1;
	if (g.b2 > 11) g.b2 = 11;
	return true;
}

// sizes of the level-1 buckets of the whole text (the plan of a sharded build)
static void msd_hist(Stream & st, TextView const & v, unsigned b1, std::vector<unsigned long long> & h) {
	unsigned const nb = 1u << b1;
	DevBuf<unsigned long long> gh(st, nb);
	B3M_CUDA(cudaMemsetAsync(gh.get(), 0, gh.bytes(), st.s));
	uint64_t const want = div_up(div_up(v.W, MSD_ITEMS), 256);
	unsigned const grid = (unsigned)(want < (uint64_t)st.sms * 8 ? (want ? want : 1) : (uint64_t)st.sms * 8);
	B3M_LAUNCH_T(st, "msd_hist", v.W / 4, k_msd_hist, grid, 256, 0, v, b1, gh.get());
	h.resize(nb);
	B3M_CUDA(cudaMemcpyAsync(h.data(), gh.get(), nb * 8, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
}


static void msd_configure() {
	static std::atomic<uint64_t> seen{0};
	if (!first_on_device(seen)) return;
	B3M_CUDA(cudaFuncSetAttribute(k_msd_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MSD_TILE * 8));
	B3M_CUDA(cudaFuncSetAttribute(k_msd_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MSD_TILE * 8));
	B3M_CUDA(cudaFuncSetAttribute(k_msd_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (MSD_TILE + 2) * 8));
	B3M_CUDA(cudaFuncSetAttribute(k_msd_finish<true, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, MSD_FIN_SMEM));
	B3M_CUDA(cudaFuncSetAttribute(k_msd_finish<false, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, MSD_FIN_SMEM));
}

static uint32_t low_mask(unsigned log) { return log >= 32 ? 0xffffffffu : (1u << log) - 1u; }

// ---- the three phases of round 0 -------------------------------------------------------------
// (1) counts of the tiles [t_lo, t_hi) of the text for the bins [d_lo, d_lo + nkeep), scanned down the columns;
//     the bin totals of these tiles are left in d_tot (device)
struct MsdCounts {
	uint32_t t_lo = 0, t_hi = 0;
	DevBuf<uint32_t> toff; // [t_hi - t_lo][nkeep] records of bin b in the earlier tiles of the range
};
static void msd_count_phase(Stream & st, TextView const & v, MsdGeom const & g, uint32_t d_lo, uint32_t nkeep, uint32_t t_lo, uint32_t t_hi,
                            MsdCounts & C, unsigned long long * d_tot, SortStats & S) {
	msd_configure();
	uint32_t const nt = t_hi - t_lo, nch = (uint32_t)div_up(nt, MSD_COLCHUNK);
	C.t_lo = t_lo; C.t_hi = t_hi;
	if (nt == 0) { B3M_CUDA(cudaMemsetAsync(d_tot, 0, nkeep * 8, st.s)); return; }
	C.toff.alloc(st, (uint64_t)nt * nkeep);
	DevBuf<uint16_t> tcount(st, (uint64_t)nt * nkeep);
	DevBuf<uint32_t> partial(st, (uint64_t)nch * nkeep);
	uint64_t const npos = std::min<uint64_t>((uint64_t)nt * MSD_TILE, v.W - (uint64_t)t_lo * MSD_TILE);
	B3M_LAUNCH_T(st, "msd_count", npos / 4 + 2ull * nt * nkeep, k_msd_count, nt, MSD_THREADS, 0, v, g.b1, d_lo, nkeep, t_lo, tcount.get());
	B3M_LAUNCH_T(st, "msd_colscan", 2ull * nt * nkeep, k_msd_colsum, nch, 256, 0, (const uint16_t *)tcount.get(), nt, nkeep, partial.get());
	B3M_LAUNCH(st, k_msd_colscan, (unsigned)div_up(nkeep, 256), 256, 0, partial.get(), nch, nkeep, d_tot);
	B3M_LAUNCH_T(st, "msd_colscan", 6ull * nt * nkeep, k_msd_colapply, nch, 256, 0, (const uint16_t *)tcount.get(), nt, nkeep, (const uint32_t *)partial.get(), C.toff.get());
	S.other_bytes += npos / 4 + 10ull * nt * nkeep;
}

// (2) level 1 on the counted tiles: the records of kept bin b go to out[p][destbase[b] ...], p the part that owns b
static void msd_scatter_phase(Stream & st, TextView const & v, MsdGeom const & g, uint32_t d_lo, uint32_t nkeep, MsdCounts & C,
                              std::vector<uint32_t> const & destbase, unsigned nparts, const uint32_t * bnd, unsigned long long * const * out,
                              uint64_t nrec, SortStats & S) {
	uint32_t const nt = C.t_hi - C.t_lo;
	if (nt == 0) return;
	B3M_REQUIRE(nparts >= 1 && nparts <= (unsigned)MSD_MAXPARTS, "too many destinations");
	DevBuf<uint32_t> dbase(st, nkeep);
	B3M_CUDA(cudaMemcpyAsync(dbase.get(), destbase.data(), nkeep * 4, cudaMemcpyHostToDevice, st.s));
	MsdP1 A;
	A.v = v; A.b1 = g.b1; A.d_lo = d_lo; A.nkeep = nkeep; A.t_lo = C.t_lo; A.nt = nt; A.base = dbase.get(); A.toff = C.toff.get(); A.nparts = nparts;
	for (unsigned p = 0; p <= nparts; ++p) A.bnd[p] = bnd[p];
	for (unsigned p = 0; p < nparts; ++p) A.out[p] = out[p];
	uint64_t const npos = std::min<uint64_t>((uint64_t)nt * MSD_TILE, v.W - (uint64_t)C.t_lo * MSD_TILE);
	// several destinations: the runs cross NVLink, two tiles per CTA double their length
	if (nparts > 1) B3M_LAUNCH_T(st, "msd_scatter", npos / 4 + 4ull * nt * nkeep + 8 * nrec, k_msd_scatter<2>, (nt + 1) / 2, 2 * MSD_THREADS, 2 * MSD_TILE * 8, A);
	else B3M_LAUNCH_T(st, "msd_scatter", npos / 4 + 4ull * nt * nkeep + 8 * nrec, k_msd_scatter<1>, nt, MSD_THREADS, MSD_TILE * 8, A);
	S.radix_passes++; S.radix_bytes += npos / 4 + 4ull * nt * nkeep + 8 * nrec;
	C.toff.release();
}

static bool msd_finish_phase(Stream & st, TextView const & v, int lin, MsdGeom const & g, uint32_t d_lo, uint32_t nkeep, const unsigned long long * tot,
                             unsigned long long * recs_ptr, bool whole, FusedOut const & fo, StreamOut * so, DevBuf<uint32_t> * sa_buf,
                             DevBuf<uint8_t> * hflag, SortStats & S, uint64_t & unresolved, uint64_t & hstart);

// Round 0 on the suffixes whose first b1 bits lie in [d_lo, d_hi), all on this device.  Returns false when the
// path does not apply (a level-1 bin of 2^30 suffixes or more).  On return `unresolved` counts the suffixes still
// tied; if that is not zero and sa_buf is given, sa_buf / hflag hold the order reached and the head flags of its
// groups, which share at least `hstart` symbols.
static bool msd_round0(Stream & st, TextView const & v, int lin, MsdGeom const & g, uint32_t d_lo, uint32_t d_hi,
                       FusedOut const & fo, StreamOut * so, DevBuf<uint32_t> * sa_buf, DevBuf<uint8_t> * hflag, SortStats & S,
                       uint64_t & unresolved, uint64_t & hstart) {
	unsigned const nkeep = d_hi - d_lo;
	unresolved = 0;
	S.rounds = 1;
	double t_last = wall_ms();
	MsdCounts C;
	std::vector<unsigned long long> tot(nkeep);
	{
		DevBuf<unsigned long long> dtot(st, nkeep);
		msd_count_phase(st, v, g, d_lo, nkeep, 0u, (uint32_t)div_up(v.W, MSD_TILE), C, dtot.get(), S);
		B3M_CUDA(cudaMemcpyAsync(tot.data(), dtot.get(), nkeep * 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	TRACE("msd count");
	std::vector<uint32_t> destbase(nkeep);
	uint64_t m = 0;
	for (unsigned b = 0; b < nkeep; ++b) {
		if (tot[b] >= (1ull << 30)) return false;
		destbase[b] = (uint32_t)m;
		m += tot[b];
	}
	S.active_sum += m;
	if (m == 0) return true;
	DevBuf<unsigned long long> recs(st, m + 2);
	uint32_t const bnd[2] = {0u, nkeep};
	unsigned long long * const out[1] = {recs.get()};
	msd_scatter_phase(st, v, g, d_lo, nkeep, C, destbase, 1, bnd, out, m, S);
	TRACE("msd level 1");
	return msd_finish_phase(st, v, lin, g, d_lo, nkeep, tot.data(), recs.get(), d_lo == 0 && nkeep == (1u << g.b1), fo, so, sa_buf, hflag, S, unresolved, hstart);
}

// (3) level 2, sub-bucket sizes and the finish on the `nkeep` bins whose records lie bin after bin in recs_ptr
// (tot[b] of them in bin b); `whole`: these are all suffixes of the text (results may be streamed to the host)
static bool msd_finish_phase(Stream & st, TextView const & v, int lin, MsdGeom const & g, uint32_t d_lo, uint32_t nkeep, const unsigned long long * tot,
                             unsigned long long * recs_ptr, bool whole, FusedOut const & fo, StreamOut * so, DevBuf<uint32_t> * sa_buf,
                             DevBuf<uint8_t> * hflag, SortStats & S, uint64_t & unresolved, uint64_t & hstart) {
	unsigned const nb2 = 1u << g.b2;
	uint64_t const W = v.W;
	msd_configure();
	double t_last = wall_ms();
	unresolved = 0;
	std::vector<uint32_t> hb(2 * (nkeep + 1)); // first record | first level-2 tile of every kept bin
	uint64_t m = 0, nt2 = 0;
	for (unsigned b = 0; b < nkeep; ++b) {
		unsigned long long const c = tot[b];
		if (c >= (1ull << 30)) return false;
		hb[b] = (uint32_t)m; hb[nkeep + 1 + b] = (uint32_t)nt2;
		m += c; nt2 += div_up(c, MSD_TILE);
	}
	hb[nkeep] = (uint32_t)m; hb[2 * nkeep + 1] = (uint32_t)nt2;
	if (m == 0) return true;
	B3M_REQUIRE(m < 0xFFFFFF00ull, "too many records for one device");
	DevBuf<uint32_t> dplan(st, hb.size());
	B3M_CUDA(cudaMemcpyAsync(dplan.get(), hb.data(), hb.size() * 4, cudaMemcpyHostToDevice, st.s));
	const uint32_t * d_base = dplan.get(), * d_tpre = dplan.get() + nkeep + 1;
	struct { unsigned long long * p; unsigned long long * get() const { return p; } } recs{recs_ptr};
	DevBuf<uint16_t> table(st, nt2 * (nb2 + 1));
	{
		MsdP2 A{g.b2, nkeep, d_base, d_tpre, recs.get(), table.get()};
		B3M_LAUNCH_T(st, "msd_local", 16 * m + nt2 * (nb2 + 1) * 2ull, k_msd_local, (unsigned)nt2, MSD_THREADS, (MSD_TILE + 2) * 8, A);
		S.radix_passes++; S.radix_bytes += 16 * m + nt2 * (nb2 + 1) * 2ull;
	}
	TRACE("msd level 2");
	uint64_t const nsb = (uint64_t)nkeep * nb2;
	DevBuf<uint32_t> sub(st, nsb + 1), scal(st, 4);
	B3M_CUDA(cudaMemsetAsync(scal.get(), 0, 16, st.s));
	B3M_CUDA(cudaMemsetAsync(sub.get() + nsb, 0, 4, st.s));
	B3M_LAUNCH_T(st, "msd_subtotals", nt2 * (nb2 + 1) * 2ull, k_msd_subtotals, nkeep, 256, 0, g.b2, d_tpre, (const uint16_t *)table.get(), sub.get(), scal.get());
	scan_exclusive_inplace<OpSum>(st, sub.get(), nsb + 1);
	S.other_bytes += nt2 * (nb2 + 1) * 2ull + nsb * 12;
	// lanes per run of the gather: the 16-byte pieces of an average run
	unsigned const pieces = (((unsigned)MSD_TILE >> g.b2) >> 1) + 1;
	unsigned glog = 1;
	while (glog < 5 && (1u << glog) < pieces) ++glog;
	DevBuf<unsigned long long> counters(st, 4 * MSD_CSLOTS);
	B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * MSD_CSLOTS, st.s));
	uint32_t const imask = low_mask(std::min<unsigned>(fo.prelog, fo.isa_s ? fo.isalog : 32u));
	uint32_t const rmask = fo.sa_s ? low_mask(fo.salog) : 0xffffffffu;
	// linear windows: the sub-buckets that hold one of the last suffixes of the window
	DevBuf<uint8_t> shortflag;
	if (lin) {
		shortflag.alloc(st, nsb);
		B3M_CUDA(cudaMemsetAsync(shortflag.get(), 0, nsb, st.s));
		B3M_LAUNCH(st, k_msd_shortflags, 1, 32, 0, v, g.b1, g.b2, d_lo, nkeep, shortflag.get());
	}
	MsdFin F{v, lin, g.b1, g.b2, glog, nkeep, recs.get(), d_base, d_tpre, table.get(), sub.get(), 0u, nullptr, nullptr, fo, imask, rmask, counters.get(), shortflag.get()};
	// 512 threads per finish CTA, 16 records through each thread's registers (1024 x 8 measured 1.5 ms slower at cfg3, profiles/r2c_*)
	auto launch_fin = [&](bool order, const char * label, uint64_t bytes, unsigned grid) {
		if (order) B3M_LAUNCH_T(st, label, bytes, (k_msd_finish<false, true, 512>), grid, 512, MSD_FIN_SMEM, F);
		else B3M_LAUNCH_T(st, label, bytes, (k_msd_finish<true, false, 512>), grid, 512, MSD_FIN_SMEM, F);
	};
	uint64_t const fbytes_per = 8 + 1; // record in, BWT code out (+ samples)
	auto read_counters = [&](unsigned long long * hc) {
		std::vector<unsigned long long> hcs(4 * MSD_CSLOTS);
		B3M_CUDA(cudaMemcpyAsync(hcs.data(), counters.get(), 32 * MSD_CSLOTS, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		for (int c = 0; c < 4; ++c) hc[c] = 0;
		for (int q = 0; q < MSD_CSLOTS; ++q) { for (int c = 0; c < 3; ++c) hc[c] += hcs[4 * q + c]; hc[3] |= hcs[4 * q + 3]; }
	};
	// early delivery (StreamOut): rows below a finished range of sub-buckets are final
	// (a key range of a sharded build: only with a local copy of the samples, and through the kernel's staged path)
	bool const stream_sa = so && so->host_sa && fo.sa_s && st.copy && nsb >= 64 && (whole || (fo.sa_s2 && fo.salog >= 5 && fo.salog < 32));
	const unsigned long long * const sa_src = fo.sa_s2 ? fo.sa_s2 : fo.sa_s;
	bool stream_bwa = so && so->host_bwa && so->d_bwa && fo.has_term && st.copy && nsb >= 64 && whole;
	uint64_t primary = 0;
	if (stream_bwa) {
		// the row of the suffix at position 0 first: its sub-bucket is known from the first text word
		uint64_t w0 = 0;
		B3M_CUDA(cudaMemcpyAsync(&w0, v.packed, 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		F.sb0 = (uint32_t)(w0 >> (64u - g.b1 - g.b2));
		B3M_CUDA(cudaMemsetAsync(fo.special, 0xff, 8, st.s));
		launch_fin(false, "msd_finish", 0, 1);
		uint32_t const row0 = fetch_u32(st, fo.special + 1);
		unsigned long long hc0[4];
		read_counters(hc0);
		if (row0 == 0xffffffffu || hc0[0]) stream_bwa = false; // tied beyond what the finish compares: no early primary
		else primary = row0;
		B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * MSD_CSLOTS, st.s)); // that sub-bucket is counted again below
	}
	unsigned const nchunks = (stream_sa || stream_bwa) ? 16u : 1u;
	std::vector<uint32_t> rows(nchunks + 1, 0);
	if (nchunks > 1) {
		for (unsigned c = 1; c < nchunks; ++c)
			B3M_CUDA(cudaMemcpyAsync(&rows[c], sub.get() + nsb * c / nchunks, 4, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	uint64_t const seq_len = W, nwords_bwa = (W + 15) >> 4;
	uint64_t words_done = 0;
	for (unsigned c = 0; c < nchunks; ++c) {
		uint32_t const s_lo = (uint32_t)(nsb * c / nchunks), s_hi = (uint32_t)(nsb * (c + 1) / nchunks);
		F.sb0 = s_lo;
		uint64_t const cm = nchunks > 1 ? (c + 1 == nchunks ? m : rows[c + 1]) - rows[c] : m;
		launch_fin(false, "msd_finish", cm * fbytes_per + cm / 4, s_hi - s_lo);
		if (stream_sa || stream_bwa) {
			uint64_t const rows_lo = c ? rows[c] + fo.shift : (whole ? 0 : fo.shift), rows_hi = (c + 1 == nchunks) ? m + fo.shift : rows[c + 1] + fo.shift;
			cudaEvent_t ev;
			B3M_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
			B3M_CUDA(cudaEventRecord(ev, st.s));
			B3M_CUDA(cudaStreamWaitEvent(st.copy, ev, 0));
			B3M_CUDA(cudaEventDestroy(ev));
			if (stream_sa) {
				uint64_t const k_lo = div_up(rows_lo, 1ull << fo.salog), k_hi = std::min<uint64_t>(div_up(rows_hi, 1ull << fo.salog), so->nsa);
				if (k_hi > k_lo) B3M_CUDA(cudaMemcpyAsync(so->host_sa + k_lo, sa_src + k_lo, (k_hi - k_lo) * 8, cudaMemcpyDeviceToHost, st.copy));
			}
			if (stream_bwa) {
				// word w reads the rows 16w .. 16w+16 (one further behind the primary)
				uint64_t const w_hi = (c + 1 == nchunks) ? nwords_bwa : (rows_hi >= 17 ? std::min<uint64_t>((rows_hi - 17) / 16 + 1, nwords_bwa) : 0);
				if (w_hi > words_done) {
					k9_pack_bwa_range(st.copy, fo.bwt, seq_len, primary, so->d_bwa, words_done, w_hi);
					++st.launches;
					B3M_CUDA(cudaMemcpyAsync(so->host_bwa + words_done, so->d_bwa + words_done, (w_hi - words_done) * 4, cudaMemcpyDeviceToHost, st.copy));
					words_done = w_hi;
				}
			}
		}
	}
	unsigned long long hc[4];
	read_counters(hc);
	TRACE("msd finish");
	S.other_bytes += m * fbytes_per + m / 4 + 32ull * hc[2];
	S.tied0 = hc[1]; S.unresolved0 = hc[0];
	unresolved = hc[0];
	if (stream_sa) so->delivered = hc[0] == 0;
	if (stream_bwa) so->bwa_delivered = hc[0] == 0;
	if (hc[0] && sa_buf) {
		// something stays tied: write the order reached and its head flags for the doubling rounds
		sa_buf->alloc(st, m);
		hflag->alloc(st, m);
		B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * MSD_CSLOTS, st.s));
		F.sb0 = 0; F.sa_out = sa_buf->get(); F.hflag = hflag->get(); F.fo = FusedOut();
		launch_fin(true, "msd_finish<order>", m * 13ull, (unsigned)nsb);
		read_counters(hc);
		S.other_bytes += m * 13ull + 32ull * hc[2];
		// what every group left shares: a sub-bucket too large for a CTA b1+b2 bits, a crowded local digit at least
		// MSD_LBITS_MIN more, records equal in all they carry and in the 32 symbols behind that
		unsigned const skip = (g.b1 + 30u) / 2;
		hstart = (hc[3] & 2) ? (g.b1 + g.b2) / 2 : ((hc[3] & 1) ? (g.b1 + g.b2 + MSD_LBITS_MIN) / 2 : skip + 32u);
		if (hstart < 1) hstart = 1;
		TRACE("msd finish<order>");
	}
	return true;
}

// ------------------------------------------------------------------------------------------
// Suffix-range sharding (multi-GPU): the suffixes are split by the top 12 bits of their first key
// into `nparts` contiguous key ranges of about equal size; one engine sorts one range and emits
// its slice of the BWT / anchors / samples with the global rank offset of the range.  The text is
// replicated, so every rank derives the same plan without communication, and no gap array or
// merge is needed.  Only texts whose ties all resolve in k_resolve take this path (the caller
// falls back to the block merge tree otherwise).
// ------------------------------------------------------------------------------------------
constexpr int KR_BINS = 4096;
constexpr int KR_TILE = 2048; // text positions per CTA of the filter kernels

__global__ void __launch_bounds__(256)
k_keyrange_hist(RadixTextSrc S, unsigned long long * __restrict__ ghist /* [KR_BINS] */) {
	__shared__ uint32_t sh[KR_BINS];
	for (int i = threadIdx.x; i < KR_BINS; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t const ntiles = div_up(S.v.W, KR_TILE);
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		uint64_t const t0 = tile * KR_TILE + (uint64_t)w * 256;
		uint32_t key[8], aux[8];
		radix_text_load<8>(S, t0, lane, key, aux);
		#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (t0 + j * 32 + lane < S.v.W) atomicAdd(&sh[key[j] >> 20], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < KR_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&ghist[i], (unsigned long long)sh[i]);
}

// The records whose key bin lies in [blo, bhi), in input order, in two passes whose cost follows
// the work: (1) one bit per record "in range" + counts per tile, read off the packed text as 4-mers
// (a few instructions per position); (2) after an exclusive scan of the counts, every warp spreads
// the set bits of its 1024 records evenly over its lanes (prefix sums + find-nth-set) and only
// then builds and writes the records, so the expensive part is proportional to the range's size.
// (The lanes first list the places of their own set bits in shared memory, at the offsets given
// by a prefix sum of their popcounts.)
constexpr int KR_FTILE = 8192; // records per CTA of both passes: 256 flag words

__global__ void __launch_bounds__(256)
k_keyrange_flags(RadixTextSrc S, uint32_t binshift, uint32_t blo, uint32_t bhi, uint32_t * __restrict__ flags, uint32_t * __restrict__ tilecount) {
	__shared__ uint32_t wsum[8];
	uint64_t const t0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 32;
	uint32_t m = 0;
	if (t0 < S.v.W) {
		uint64_t const i0 = t0 - S.nshort;
		uint64_t p0 = S.v.wstart + i0;
		if (S.v.text_wraps && p0 >= S.v.ntext) p0 -= S.v.ntext;
		bool const fast = binshift == 24 && S.bits == 2 && S.v.packed && t0 >= S.nshort && i0 + 67 <= S.v.W && p0 + 67 <= S.v.ntext;
		if (fast) {
			uint64_t const a = pk_window(S.v.packed, p0), b = pk_window(S.v.packed, p0 + 32);
			uint32_t const span = bhi - blo;
			#pragma unroll
			for (int s = 0; s < 32; ++s) {
				uint32_t const kmer = (s <= 28 ? (uint32_t)(a >> (56 - 2 * s)) : (uint32_t)((a << (2 * s - 56)) | (b >> (120 - 2 * s)))) & 255u;
				m |= (kmer - blo < span ? 1u : 0u) << s;
			}
		} else {
			for (int s = 0; s < 32; ++s) {
				uint64_t const t = t0 + s;
				if (t >= S.v.W) break;
				uint32_t key, idx, aux;
				radix_text_record(S, t, key, idx, aux);
				uint32_t const bin = key >> binshift;
				m |= ((bin >= blo && bin < bhi) ? 1u : 0u) << s;
			}
		}
		flags[t0 >> 5] = m;
	}
	uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m));
	if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t tot = 0; for (int i = 0; i < 8; ++i) tot += wsum[i]; tilecount[blockIdx.x] = tot; }
}

__global__ void __launch_bounds__(256)
k_keyrange_gather(RadixTextSrc S, const uint32_t * __restrict__ flags, const uint32_t * __restrict__ tileoff, uint32_t ntiles,
                  uint32_t * __restrict__ okey, uint32_t * __restrict__ oidx, uint8_t * __restrict__ oaux,
                  unsigned long long * __restrict__ ghist /* [4][256]: digit histograms of the keys written, for the radix passes */) {
	__shared__ uint32_t wtot[8];
	__shared__ uint16_t s_list[8][1024];
	__shared__ uint32_t s_hist[RADIX_MAXDIG][RADIX_BINS];
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < RADIX_MAXDIG * RADIX_BINS; i += blockDim.x) (&s_hist[0][0])[i] = 0;
	uint64_t const nwords = div_up(S.v.W, 32);
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		__syncthreads(); // s_hist cleared / wtot of the previous tile consumed
		uint64_t const tbase = (uint64_t)tile * KR_FTILE + (uint64_t)w * 1024; // a warp owns 32 flag words
		uint64_t const wi = (tbase >> 5) + lane;
		uint32_t const m = wi < nwords ? flags[wi] : 0u;
		uint32_t incl = (uint32_t)__popc(m);
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t const x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += x; }
		uint32_t const excl = incl - (uint32_t)__popc(m);
		uint32_t const total = __shfl_sync(0xffffffffu, incl, 31);
		if (lane == 0) wtot[w] = total;
		__syncthreads();
		uint32_t off = tileoff[tile];
		for (unsigned i = 0; i < w; ++i) off += wtot[i];
		// every lane lists the places of its own set bits behind those of the lanes below it ...
		uint16_t * const list = s_list[w];
		{
			uint32_t mm = m, j = excl;
			while (mm) { list[j++] = (uint16_t)(32u * lane + (uint32_t)__ffs((int)mm) - 1u); mm &= mm - 1u; }
		}
		__syncwarp();
		// ... and the warp takes the listed records 32 at a time
		for (uint32_t k = lane; k < total; k += 32) {
			uint64_t const t = tbase + list[k];
			uint32_t key, idx, aux;
			radix_text_record_any(S, t, key, idx, aux);
			okey[off + k] = key; oidx[off + k] = idx; oaux[off + k] = (uint8_t)aux;
			atomicAdd(&s_hist[0][key & 255u], 1u); atomicAdd(&s_hist[1][(key >> 8) & 255u], 1u);
			atomicAdd(&s_hist[2][(key >> 16) & 255u], 1u); atomicAdd(&s_hist[3][key >> 24], 1u);
		}
		__syncwarp();
	}
	__syncthreads();
	for (int i = threadIdx.x; i < RADIX_MAXDIG * RADIX_BINS; i += blockDim.x)
		if ((&s_hist[0][0])[i]) atomicAdd(&ghist[i], (unsigned long long)(&s_hist[0][0])[i]);
}

void k2_keyrange_plan(Stream & st, DevText const & T, int circular, uint32_t nparts, KeyRangePlan & plan) {
	B3M_REQUIRE(nparts >= 1 && nparts <= KR_BINS, "bad number of key ranges");
	unsigned const bits = T.keybits, k0 = 32 / bits;
	uint64_t const W = T.ntext;
	uint64_t const nshort = circular ? 0 : (W < (uint64_t)(k0 - 1) ? W : (uint64_t)(k0 - 1));
	TextView v{T.codes, bits == 2 ? T.packed : nullptr, T.ntext, 0, W, circular, 0, T.has_term};
	RadixTextSrc S{v, nshort, bits, k0};
	std::vector<unsigned long long> h;
	uint32_t nbins;
	MsdGeom mg;
	plan.msd_b1 = plan.msd_b2 = 0; plan.hist.clear();
	if (msd_geometry(st, T, W, mg)) {
		// the level-1 buckets of the MSD path: the histogram the sort needs anyway
		msd_hist(st, v, mg.b1, h);
		nbins = 1u << mg.b1; plan.binshift = 32 - mg.b1;
		plan.msd_b1 = mg.b1; plan.msd_b2 = mg.b2; plan.hist = h;
	} else if (bits == 2 && v.packed) {
		// the leading 4 symbols of every suffix: the top digit of the first key, whose histogram the sort needs anyway
		nbins = RADIX_BINS; plan.binshift = 24;
		DevBuf<unsigned long long> gh(st, RADIX_MAXDIG * RADIX_BINS);
		B3M_CUDA(cudaMemsetAsync(gh.get(), 0, gh.bytes(), st.s));
		uint64_t const want = div_up(div_up(W, 32), 256 * 4);
		unsigned const grid = (unsigned)(want < (uint64_t)st.sms * 8 ? (want ? want : 1) : (uint64_t)st.sms * 8);
		B3M_LAUNCH_T(st, "hist_4mers", W / 4, k_hist_4mers, grid, 256, 0, v, gh.get());
		B3M_LAUNCH(st, k_hist_4mers_fix, 1, 256, 0, v, gh.get());
		h.resize(nbins);
		B3M_CUDA(cudaMemcpyAsync(h.data(), gh.get() + 3 * RADIX_BINS, nbins * 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	} else {
		nbins = KR_BINS; plan.binshift = 20;
		DevBuf<unsigned long long> gh(st, KR_BINS);
		B3M_CUDA(cudaMemsetAsync(gh.get(), 0, KR_BINS * 8, st.s));
		uint64_t const want = div_up(W, KR_TILE);
		unsigned const grid = (unsigned)(want < (uint64_t)st.sms * 8 ? (want ? want : 1) : (uint64_t)st.sms * 8);
		B3M_LAUNCH_T(st, "keyrange_hist", W / 4, k_keyrange_hist, grid, 256, 0, S, gh.get());
		h.resize(nbins);
		B3M_CUDA(cudaMemcpyAsync(h.data(), gh.get(), KR_BINS * 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	plan.nparts = nparts;
	plan.bin_lo.assign(nparts + 1, nbins);
	plan.base.assign(nparts + 1, W);
	plan.bin_lo[0] = 0; plan.base[0] = 0;
	uint64_t acc = 0;
	uint32_t p = 1;
	for (uint32_t b = 0; b < nbins && p < nparts; ++b) {
		// part p starts at the first bin boundary at or past p/nparts of the suffixes
		while (p < nparts && acc >= (W * p) / nparts) { plan.bin_lo[p] = b; plan.base[p] = acc; ++p; }
		acc += h[b];
	}
}

uint64_t k2_sort_keyrange(Stream & st, DevText const & T, int circular, KeyRangePlan const & plan, uint32_t part, FusedOut const & fo0, SortStats * stats) {
	B3M_REQUIRE(part < plan.nparts, "bad key range index");
	unsigned const bits = T.keybits, k0 = 32 / bits;
	uint64_t const W = T.ntext;
	uint64_t const nshort = circular ? 0 : (W < (uint64_t)(k0 - 1) ? W : (uint64_t)(k0 - 1));
	TextView v{T.codes, bits == 2 ? T.packed : nullptr, T.ntext, 0, W, circular, 0, T.has_term};
	RadixTextSrc S{v, nshort, bits, k0};
	uint32_t const blo = plan.bin_lo[part], bhi = plan.bin_lo[part + 1];
	uint64_t const m = plan.base[part + 1] - plan.base[part];
	SortStats St;
	St.rounds = 1;
	if (m == 0) return 0;
	if (plan.msd_b1) {
		MsdGeom mg; mg.b1 = plan.msd_b1; mg.b2 = plan.msd_b2;
		FusedOut fo = fo0;
		fo.shift = fo0.shift + plan.base[part];
		uint64_t unresolved = 0, hstart = 0;
		if (msd_round0(st, v, !circular, mg, blo, bhi, fo, nullptr, nullptr, nullptr, St, unresolved, hstart)) {
			if (stats) {
				stats->rounds = stats->rounds > St.rounds ? stats->rounds : St.rounds;
				stats->radix_passes += St.radix_passes; stats->radix_bytes += St.radix_bytes;
				stats->active_sum += St.active_sum; stats->other_bytes += St.other_bytes;
				stats->tied0 += St.tied0; stats->unresolved0 += St.unresolved0;
			}
			return unresolved;
		}
		return m; // a level-1 bucket too large for the path: the caller takes the merge tree
	}
	// stable compaction of the range's records out of the text
	uint32_t const ntiles = (uint32_t)div_up(W, KR_FTILE);
	DevBuf<uint32_t> tcount(st, ntiles), fl(st, (size_t)ntiles * (KR_FTILE / 32));
	B3M_LAUNCH_T(st, "keyrange_flags", W / 4 + W / 8, k_keyrange_flags, ntiles, 256, 0, S, plan.binshift, blo, bhi, fl.get(), tcount.get());
	scan_exclusive_inplace<OpSum>(st, tcount.get(), ntiles);
	DevBuf<uint32_t> key0(st, m), key1(st, m), idx0(st, m), idx1(st, m);
	DevBuf<uint8_t> aux0(st, m), aux1(st, m);
	DevBuf<unsigned long long> khist(st, RADIX_MAXDIG * RADIX_BINS);
	B3M_CUDA(cudaMemsetAsync(khist.get(), 0, khist.bytes(), st.s));
	unsigned const ggrid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)st.sms * 8);
	B3M_LAUNCH_T(st, "keyrange_gather", W / 8 + 9 * m, k_keyrange_gather, ggrid, 256, 0, S, (const uint32_t *)fl.get(), (const uint32_t *)tcount.get(), ntiles,
	             key0.get(), idx0.get(), aux0.get(), khist.get());
	St.other_bytes += W / 4 + W / 4 + 9 * m;
	RadixRec<2> cur{{key0.get(), idx0.get()}, aux0.get()}, alt{{key1.get(), idx1.get()}, aux1.get()};
	RadixStats rs;
	radix_sort_bits<2>(st, cur, alt, 0, m, 0, 32, &rs, khist.get());
	St.radix_passes += rs.passes; St.radix_bytes += rs.bytes; St.active_sum += m;
	DevBuf<unsigned long long> counters(st, 4 * RS_CSLOTS);
	DevBuf<uint8_t> hflag(st, m);
	B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * RS_CSLOTS, st.s));
	FusedOut fo = fo0;
	fo.shift = fo0.shift + plan.base[part];
	uint64_t const rbytes = m * 10ull;
	RS_LAUNCH("resolve_extract", rbytes, true, false, (unsigned)div_up(m, RS_TILE), v, bits, k0, !circular, (const uint32_t *)cur.a[0],
	             (const uint32_t *)cur.a[1], (const uint8_t *)cur.aux, m, 0u, alt.a[1], hflag.get(), fo, counters.get());
	std::vector<unsigned long long> hcs(4 * RS_CSLOTS);
	B3M_CUDA(cudaMemcpyAsync(hcs.data(), counters.get(), 32 * RS_CSLOTS, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	unsigned long long hc[4] = {0, 0, 0, 0};
	for (int q = 0; q < RS_CSLOTS; ++q) for (int c = 0; c < 3; ++c) hc[c] += hcs[4 * q + c];
	St.tied0 = hc[1]; St.unresolved0 = hc[0];
	St.other_bytes += rbytes + 32ull * hc[2];
	if (stats) {
		stats->rounds = stats->rounds > St.rounds ? stats->rounds : St.rounds;
		stats->radix_passes += St.radix_passes; stats->radix_bytes += St.radix_bytes;
		stats->active_sum += St.active_sum; stats->other_bytes += St.other_bytes;
		stats->tied0 += St.tied0; stats->unresolved0 += St.unresolved0;
	}
	return hc[0];
}

// ------------------------------------------------------------------------------------------
// Position sharding of level 1 (XShard, kernels.h)
// ------------------------------------------------------------------------------------------
bool k2_xshard_count(Stream & st, DevText const & T, int circular, uint32_t part, uint32_t nparts, XShard & X, unsigned long long * d_totals, uint32_t * nbins) {
	B3M_REQUIRE(nparts >= 1 && nparts <= (uint32_t)MSD_MAXPARTS && part < nparts, "bad part index");
	MsdGeom g;
	*nbins = 0;
	if (!msd_geometry(st, T, T.ntext, g, nparts)) return false;
	TextView v{T.codes, T.packed, T.ntext, 0, T.ntext, circular, 0, T.has_term};
	uint32_t const nt1 = (uint32_t)div_up(T.ntext, MSD_TILE);
	X = XShard();
	X.b1 = g.b1; X.b2 = g.b2; X.part = part; X.nparts = nparts;
	X.t_lo = (uint32_t)((uint64_t)nt1 * part / nparts); X.t_hi = (uint32_t)((uint64_t)nt1 * (part + 1) / nparts);
	MsdCounts C;
	SortStats S;
	msd_count_phase(st, v, g, 0u, 1u << g.b1, X.t_lo, X.t_hi, C, d_totals, S);
	X.toff = std::move(C.toff);
	*nbins = 1u << g.b1;
	return true;
}

void k2_xshard_scatter(Stream & st, DevText const & T, int circular, XShard & X, const unsigned long long * h_alltot, unsigned long long * const * recs,
                       const uint64_t * cap, SortStats * stats) {
	B3M_REQUIRE(X.nparts, "xshard_count was not called");
	unsigned const nb1 = 1u << X.b1;
	uint32_t const P = X.nparts;
	MsdGeom g; g.b1 = X.b1; g.b2 = X.b2;
	TextView v{T.codes, T.packed, T.ntext, 0, T.ntext, circular, 0, T.has_term};
	// global bin sizes; the bins are cut into P ranges of about W / P suffixes (the same on every part)
	X.total.assign(nb1, 0);
	for (uint32_t p = 0; p < P; ++p) for (unsigned b = 0; b < nb1; ++b) X.total[b] += h_alltot[(uint64_t)p * nb1 + b];
	uint64_t const W = T.ntext;
	X.bnd.assign(P + 1, nb1);
	X.bnd[0] = 0;
	std::vector<uint64_t> first(P + 1, W);
	first[0] = 0;
	{
		uint64_t acc = 0;
		uint32_t p = 1;
		for (unsigned b = 0; b < nb1 && p < P; ++b) {
			while (p < P && acc >= (W * p) / P) { X.bnd[p] = b; first[p] = acc; ++p; }
			acc += X.total[b];
		}
	}
	for (unsigned b = 0; b < nb1; ++b) B3M_REQUIRE(X.total[b] < (1ull << 30), "a level-1 bucket of 2^30 suffixes or more: use another strategy");
	for (uint32_t p = 0; p < P; ++p) B3M_REQUIRE(first[p + 1] - first[p] + 2 <= cap[p] && first[p + 1] - first[p] < 0xFFFFFF00ull, "a key range exceeds its record array: use another strategy");
	X.rank_base = first[X.part]; X.records = first[X.part + 1] - first[X.part];
	// where this part's records of bin b go: start of the bin in its owner's array + the records of the lower parts
	std::vector<uint32_t> destbase(nb1);
	uint64_t nrec = 0;
	for (uint32_t p = 0; p < P; ++p) {
		uint64_t acc = 0;
		for (unsigned b = X.bnd[p]; b < X.bnd[p + 1]; ++b) {
			uint64_t below = 0;
			for (uint32_t q = 0; q < X.part; ++q) below += h_alltot[(uint64_t)q * nb1 + b];
			destbase[b] = (uint32_t)(acc + below);
			acc += X.total[b];
			nrec += h_alltot[(uint64_t)X.part * nb1 + b];
		}
	}
	MsdCounts C;
	C.t_lo = X.t_lo; C.t_hi = X.t_hi; C.toff = std::move(X.toff);
	SortStats S;
	msd_scatter_phase(st, v, g, 0u, nb1, C, destbase, P, X.bnd.data(), recs, nrec, S);
	if (stats) { stats->radix_passes += S.radix_passes; stats->radix_bytes += S.radix_bytes; stats->other_bytes += S.other_bytes; }
}

uint64_t k2_xshard_finish(Stream & st, DevText const & T, int circular, XShard & X, unsigned long long * recs_own, FusedOut const & fo0, SortStats * stats,
                          StreamOut * so) {
	B3M_REQUIRE(X.nparts && !X.bnd.empty(), "xshard_scatter was not called");
	MsdGeom g; g.b1 = X.b1; g.b2 = X.b2;
	TextView v{T.codes, T.packed, T.ntext, 0, T.ntext, circular, 0, T.has_term};
	uint32_t const d_lo = X.bnd[X.part], nkeep = X.bnd[X.part + 1] - d_lo;
	FusedOut fo = fo0;
	fo.shift = fo0.shift + X.rank_base;
	SortStats S;
	S.rounds = 1; S.active_sum = X.records;
	uint64_t unresolved = 0, hstart = 0;
	if (nkeep && X.records) {
		bool const ok = msd_finish_phase(st, v, !circular, g, d_lo, nkeep, X.total.data() + d_lo, recs_own, false, fo, so, nullptr, nullptr, S, unresolved, hstart);
		B3M_REQUIRE(ok, "internal: xshard finish does not apply");
	}
	if (stats) {
		stats->rounds = stats->rounds > S.rounds ? stats->rounds : S.rounds;
		stats->radix_passes += S.radix_passes; stats->radix_bytes += S.radix_bytes;
		stats->active_sum += S.active_sum; stats->other_bytes += S.other_bytes;
		stats->tied0 += S.tied0; stats->unresolved0 += S.unresolved0;
	}
	return unresolved;
}

// ------------------------------------------------------------------------------------------
// Prefix doubling, one round on the groups that fit a CTA (k_dbl_tile + k_dbl_compact).  The active list holds
// the suffixes of all groups of two or more, group after group: (g, i) = (place of the group's first member
// in the suffix array, suffix).  A CTA owns the groups that START in its tile of the list and have at
// most DT_GMAX members (it sees DT_GMAX elements on either side of the tile, so it knows).  It reads the
// rank of the suffix h symbols ahead of every member (the only scattered access), and every member counts
// the members of its group with a smaller key: that is its new place, and the new group id of all members
// with an equal key.  Most members of a group of near-identical copies carry the key of the group's first
// member: those get their place from three group totals (members below that key, members off that key in
// front of them -- a bitmap and its prefix counts) without looking at any other member; the others
// ("deviants") are queued and counted against the whole group by consecutive threads afterwards, so that no
// warp walks a group for one lane.  Outputs, at the member's new place in the list: new group id, suffix,
// flags (1: its new group still has two or more members, 2: member of a larger group -- passed through for
// the radix round, 4: the group id changed); and, per tile of list places, how many entries carry flag 1
// (tcount; a CTA's groups reach into the next tile at most).  The suffix array is updated in place (a group
// owns its range); the rank array is NOT written here: other CTAs read it in the same round, and a mixture
// of old and new ranks inside one comparison would order two suffixes of one old group wrongly.
// k_dbl_compact, after a scan of tcount, writes the changed ranks and the next list.
// ------------------------------------------------------------------------------------------
constexpr int DT_THREADS = 512;
constexpr int DT_TILE = 2048;
constexpr int DT_GMAX = 256;
constexpr int DT_REG = DT_TILE + 2 * DT_GMAX;
constexpr int DT_PER = (DT_REG - DT_GMAX + DT_THREADS - 1) / DT_THREADS;
constexpr int DT_WORDS = DT_REG / 32;
static_assert(DT_REG % 32 == 0 && DT_GMAX % 32 == 0 && DT_TILE % DT_THREADS == 0, "tile geometry");

// highest head at or below x, not below lo (-1: none)
__device__ __forceinline__ int dbl_prev_head(const uint32_t * hb, int x, int lo) {
	int q = x >> 5;
	uint32_t w = hb[q] & (0xffffffffu >> (31 - (x & 31)));
	for (;;) {
		if (w) { int const p = q * 32 + 31 - __clz((int)w); return p >= lo ? p : -1; }
		if (q * 32 <= lo) return -1;
		w = hb[--q];
	}
}
// lowest head above x, not above hi (-1: none)
__device__ __forceinline__ int dbl_next_head(const uint32_t * hb, int x, int hi) {
	int const x1 = x + 1;
	int q = x1 >> 5;
	uint32_t w = hb[q] & (0xffffffffu << (x1 & 31));
	for (;;) {
		if (w) { int const p = q * 32 + __ffs((int)w) - 1; return p <= hi ? p : -1; }
		if ((q + 1) * 32 > hi) return -1;
		w = hb[++q];
	}
}

__global__ void __launch_bounds__(DT_THREADS)
k_dbl_tile(const uint32_t * __restrict__ cg, const uint32_t * __restrict__ ci, uint32_t na, const uint32_t * __restrict__ rank,
           uint32_t * __restrict__ sa, uint64_t h, uint64_t W, int circular, uint32_t * __restrict__ og, uint32_t * __restrict__ oi,
           uint8_t * __restrict__ of, uint32_t * __restrict__ nbig, uint32_t * __restrict__ tcount) {
	__shared__ uint32_t s_g[DT_REG], s_i[DT_REG], s_k[DT_REG];
	__shared__ uint32_t s_nlt[DT_REG];      // at a group's first member: members with a key below that member's
	__shared__ uint16_t s_q[DT_REG];        // queue of deviants
	__shared__ uint32_t s_hb[DT_WORDS + 1]; // heads
	__shared__ uint32_t s_db[DT_WORDS + 1]; // deviants
	__shared__ uint32_t s_dpre[DT_WORDS + 1];
	__shared__ uint32_t s_big, s_nq, s_cnt[3];
	uint32_t const t0 = blockIdx.x * (uint32_t)DT_TILE;
	uint32_t const abase = t0 - (uint32_t)DT_GMAX; // region index x <-> list index abase + x; places before the list wrap to huge values
	unsigned const lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) { s_big = 0; s_nq = 0; s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; }
	for (int x = threadIdx.x; x < DT_REG; x += DT_THREADS) {
		uint32_t const a = abase + (uint32_t)x;
		bool const valid = a < na;
		s_g[x] = valid ? cg[a] : 0xffffffffu;
		s_i[x] = (valid && x >= DT_GMAX) ? ci[a] : 0u;
		s_nlt[x] = 0;
	}
	if (threadIdx.x <= DT_WORDS) s_db[threadIdx.x] = 0;
	__syncthreads();
	for (int q = w; q < DT_WORDS; q += DT_THREADS / 32) {
		int const x = q * 32 + (int)lane;
		uint32_t const a = abase + (uint32_t)x;
		bool const valid = a < na;
		bool const head = !valid || a == 0 || x == 0 || s_g[x] != s_g[x - 1];
		uint32_t const hb = __ballot_sync(0xffffffffu, head);
		if (lane == 0) s_hb[q] = hb;
	}
	if (threadIdx.x == 0) s_hb[DT_WORDS] = 0xffffffffu;
	__syncthreads();
	// group of every element of the tile and of the DT_GMAX elements behind it; the key of the members of my groups
	uint32_t span[DT_PER]; // (first member) | (members << 16); 0: not mine
	uint32_t nb = 0;
	#pragma unroll
	for (int j = 0; j < DT_PER; ++j) {
		span[j] = 0;
		int const x = DT_GMAX + j * DT_THREADS + (int)threadIdx.x;
		if (x >= DT_REG) continue;
		uint32_t const a = abase + (uint32_t)x;
		if (a >= na) continue;
		int const sx = dbl_prev_head(s_hb, x, x - DT_GMAX + 1);
		int const ex = sx < 0 ? -1 : dbl_next_head(s_hb, x, sx + DT_GMAX);
		bool const owned = x < DT_GMAX + DT_TILE;
		if (sx < 0 || ex < 0) {
			// a group of more than DT_GMAX members: left to the radix round
			if (owned) { og[a] = s_g[x]; oi[a] = s_i[x]; of[a] = 2; ++nb; }
			continue;
		}
		if (sx < DT_GMAX || sx >= DT_GMAX + DT_TILE) continue; // a group of the neighbouring tile
		span[j] = (uint32_t)sx | ((uint32_t)(ex - sx) << 16);
		uint64_t jj = (uint64_t)s_i[x] + h;
		uint32_t k;
		if (circular) { if (jj >= W) jj %= W; k = rank[jj]; }
		else k = (jj < W) ? rank[jj] + 1u : 0u;
		s_k[x] = k;
	}
	if (nb) atomicAdd(&s_big, nb);
	__syncthreads();
	// deviants: members whose key is not the one of their group's first member
	#pragma unroll
	for (int j = 0; j < DT_PER; ++j) {
		int const x = DT_GMAX + j * DT_THREADS + (int)threadIdx.x; // a warp covers the 32 places of one bitmap word
		bool dev = false;
		if (span[j]) {
			int const sx = (int)(span[j] & 0xffffu);
			uint32_t const k = s_k[x], k0 = s_k[sx];
			dev = k != k0;
			if (k < k0) atomicAdd(&s_nlt[sx], 1u);
		}
		uint32_t const b = __ballot_sync(0xffffffffu, dev);
		if (x < DT_REG) {
			uint32_t qb = 0;
			if (lane == 0) { s_db[x >> 5] = b; if (b) qb = atomicAdd(&s_nq, (uint32_t)__popc(b)); }
			qb = __shfl_sync(0xffffffffu, qb, 0);
			if (dev) s_q[qb + __popc(b & lanemask_lt())] = (uint16_t)x;
		}
	}
	__syncthreads();
	if (w == 0) { // prefix counts of the deviant bitmap, word by word
		uint32_t run = 0;
		for (int q0 = 0; q0 < DT_WORDS + 1; q0 += 32) {
			int const q = q0 + (int)lane;
			uint32_t const c = q < DT_WORDS ? (uint32_t)__popc(s_db[q]) : 0u;
			uint32_t incl = c;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
			if (q <= DT_WORDS) s_dpre[q] = run + incl - c;
			run += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
	__syncthreads();
	uint32_t n_own = 0, n_next = 0; // entries still tied, by the tile of the list they land in
	uint32_t n_chg = 0;             // entries whose rank changes
	auto emit = [&](int x, int sx, uint32_t less, uint32_t eqb, uint32_t eq) {
		uint32_t const g = s_g[x], i = s_i[x];
		uint32_t const a2 = abase + (uint32_t)sx + less + eqb;
		sa[g + less + eqb] = i;
		og[a2] = g + less;
		oi[a2] = i;
		of[a2] = (uint8_t)((eq > 1 ? 1u : 0u) | (less ? 4u : 0u));
		if (eq > 1) { if (a2 - t0 < (uint32_t)DT_TILE) ++n_own; else ++n_next; }
		if (less) ++n_chg;
	};
	auto devs_before = [&](int x) -> uint32_t { return s_dpre[x >> 5] + (uint32_t)__popc(s_db[x >> 5] & ((1u << (x & 31)) - 1u)); };
	#pragma unroll
	for (int j = 0; j < DT_PER; ++j) {
		if (!span[j]) continue;
		int const x = DT_GMAX + j * DT_THREADS + (int)threadIdx.x;
		int const sx = (int)(span[j] & 0xffffu), n = (int)(span[j] >> 16);
		if (s_k[x] != s_k[sx]) continue; // queued
		uint32_t const d0 = devs_before(sx);
		emit(x, sx, s_nlt[sx], (uint32_t)(x - sx) - (devs_before(x) - d0), (uint32_t)n - (devs_before(sx + n) - d0));
	}
	uint32_t const nq = s_nq;
	for (uint32_t q = threadIdx.x; q < nq; q += DT_THREADS) {
		int const x = (int)s_q[q];
		int const sx = dbl_prev_head(s_hb, x, x - DT_GMAX + 1);
		int const ex = dbl_next_head(s_hb, x, sx + DT_GMAX);
		uint32_t const k = s_k[x];
		uint32_t less = 0, eqb = 0, eq = 0;
		#pragma unroll 4
		for (int y = sx; y < ex; ++y) {
			uint32_t const o = s_k[y];
			less += o < k ? 1u : 0u;
			eq += o == k ? 1u : 0u;
			eqb += (o == k && y < x) ? 1u : 0u;
		}
		emit(x, sx, less, eqb, eq);
	}
	n_own = __reduce_add_sync(0xffffffffu, n_own);
	n_next = __reduce_add_sync(0xffffffffu, n_next);
	n_chg = __reduce_add_sync(0xffffffffu, n_chg);
	if (lane == 0) { if (n_own) atomicAdd(&s_cnt[0], n_own); if (n_next) atomicAdd(&s_cnt[1], n_next); if (n_chg) atomicAdd(&s_cnt[2], n_chg); }
	__syncthreads();
	if (threadIdx.x == 0) {
		if (s_big) atomicAdd(nbig, s_big);
		if (s_cnt[2]) atomicAdd(nbig + 1, s_cnt[2]); // nbig[1]: ranks that change in this round
		if (s_cnt[0]) atomicAdd(&tcount[blockIdx.x], s_cnt[0]);
		if (s_cnt[1]) atomicAdd(&tcount[blockIdx.x + 1], s_cnt[1]);
	}
}

// The pass behind k_dbl_tile: list places in order, DT_TILE per CTA.  Writes the ranks that changed in this round
// (now that every CTA has read the old ones) and moves the entries still tied to the front: tbase = exclusive scan
// of k_dbl_tile's per-tile counts, the place inside the tile from ballots (warp w handles places w*32.. of every
// DT_THREADS-wide row, all accesses coalesced).
// DIRECT = false: the changed ranks are not written here; oi[a] is replaced by the suffix whose rank changed (or
// 0xffffffff), so that (oi, og) are the pairs rank_write_partitioned takes.
template <bool DIRECT>
__global__ void __launch_bounds__(DT_THREADS)
k_dbl_compact(const uint32_t * __restrict__ og, uint32_t * __restrict__ oi, const uint8_t * __restrict__ of, uint32_t na,
              const uint32_t * __restrict__ tbase, uint32_t * __restrict__ rank, uint32_t * __restrict__ cg, uint32_t * __restrict__ ci) {
	constexpr int ROWS = DT_TILE / DT_THREADS, WARPS = DT_THREADS / 32;
	__shared__ uint32_t s_pre[ROWS * WARPS];
	uint32_t const t0 = blockIdx.x * (uint32_t)DT_TILE;
	unsigned const lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t f[ROWS], bal[ROWS];
	#pragma unroll
	for (int j = 0; j < ROWS; ++j) {
		uint32_t const a = t0 + (uint32_t)(j * DT_THREADS) + threadIdx.x;
		f[j] = a < na ? (uint32_t)of[a] : 0u;
		bal[j] = __ballot_sync(0xffffffffu, f[j] & 1u);
		if (lane == 0) s_pre[j * WARPS + w] = (uint32_t)__popc(bal[j]);
	}
	__syncthreads();
	if (w == 0) {
		uint32_t run = 0;
		#pragma unroll
		for (int q0 = 0; q0 < ROWS * WARPS; q0 += 32) {
			uint32_t const c = s_pre[q0 + lane];
			uint32_t incl = c;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
			s_pre[q0 + lane] = run + incl - c;
			run += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
	__syncthreads();
	uint32_t const base = tbase[blockIdx.x];
	#pragma unroll
	for (int j = 0; j < ROWS; ++j) {
		uint32_t const a = t0 + (uint32_t)(j * DT_THREADS) + threadIdx.x;
		if (!(f[j] & 5u)) { if (!DIRECT && a < na) oi[a] = 0xffffffffu; continue; }
		uint32_t const g = og[a], i = oi[a];
		if (DIRECT) { if (f[j] & 4u) rank[i] = g; }
		else if (!(f[j] & 4u)) oi[a] = 0xffffffffu;
		if (f[j] & 1u) {
			uint32_t const dst = base + s_pre[j * WARPS + w] + (uint32_t)__popc(bal[j] & lanemask_lt());
			cg[dst] = g; ci[dst] = i;
		}
	}
}

// rank[pi[k]] = pg[k] for the pairs of a list (pi = 0xffffffff: no pair).  A rank array beyond the L2 size takes one
// 4-byte store at a random place as a read-modify-write of a 32-byte sector (27 G stores/s measured on cfg4); with the
// pairs partitioned by the top 8 bits of their target (one radix pass, rank_write_partitioned) the stores in flight fall
// into 1/256 of the array, merge in L2 and leave as whole sectors.
__global__ void __launch_bounds__(256)
k_rank_write(const uint32_t * __restrict__ pi, const uint32_t * __restrict__ pg, uint64_t n, uint32_t * __restrict__ rank) {
	uint64_t const i0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 4;
	if (i0 + 4 <= n) {
		uint4 const a = ld_stream_u4(reinterpret_cast<const uint4 *>(pi + i0)), g = ld_stream_u4(reinterpret_cast<const uint4 *>(pg + i0));
		if (a.x != 0xffffffffu) rank[a.x] = g.x;
		if (a.y != 0xffffffffu) rank[a.y] = g.y;
		if (a.z != 0xffffffffu) rank[a.z] = g.z;
		if (a.w != 0xffffffffu) rank[a.w] = g.w;
	} else
		for (uint64_t i = i0; i < n; ++i) if (pi[i] != 0xffffffffu) rank[pi[i]] = pg[i];
}

// pairs (pi, pg) in list order -> rank; t0 / t1 = scratch of n words each (16-byte aligned, like pi / pg); pi / pg are kept
static void rank_write_partitioned(Stream & st, uint32_t * pi, uint32_t * pg, uint64_t n, uint32_t * t0, uint32_t * t1, uint64_t W, uint32_t * rank,
                                   const char * label, SortStats & S) {
	int const bw = (int)ceil_log2_u64(W + 1);
	const uint32_t * qi = pi, * qg = pg;
	if (bw > 8) {
		RadixRec<2> cur{{pi, pg}}, alt{{t0, t1}};
		RadixStats rs;
		radix_sort_bits<2>(st, cur, alt, 0, n, bw - 8, bw, &rs); // one pass: the result lies in t0 / t1 (or in place when every target shares the digit)
		S.radix_passes += rs.passes; S.radix_bytes += rs.bytes;
		qi = cur.a[0]; qg = cur.a[1];
	}
	B3M_LAUNCH_T(st, label, n * 12ull, k_rank_write, (unsigned)div_up(div_up(n, 4), 256), 256, 0, qi, qg, n, rank);
	S.other_bytes += n * 12ull;
}

void k2_suffix_sort(Stream & st, DevText const & T, uint64_t wstart, uint64_t W, int circular, int text_wraps,
                    DevBuf<uint32_t> & sa_buf, uint32_t * rank, SortStats * stats, const FusedOut * fo, StreamOut * so) {
	if (W == 0) return;
	B3M_REQUIRE(W < 0xFFFFFF00ull, "window too large for 32-bit suffix indices");
	B3M_REQUIRE(!fo || (wstart == 0 && W == T.ntext), "internal: fused outputs need the whole text in one window");
	unsigned const bits = T.keybits;
	unsigned const k0 = 32 / bits;
	uint64_t const nshort = circular ? 0 : (W < (uint64_t)(k0 - 1) ? W : (uint64_t)(k0 - 1));
	TextView v{T.codes, bits == 2 ? T.packed : nullptr, T.ntext, wstart, W, circular, text_wraps, T.has_term};
	int const lin = !circular;
	SortStats S;
	double t_last = wall_ms();

	DevBuf<uint32_t> scalar(st, 4);
	uint32_t * d_total = scalar.get();
	DevBuf<unsigned long long> counters(st, 4 * RS_CSLOTS);
	DevBuf<uint8_t> hflag;
	uint64_t unresolved = 0;
	uint64_t hstart = k0; // every group left by round 0 shares at least this many symbols
	bool round0_done = false;
	MsdGeom mg;
	if (fo && msd_geometry(st, T, W, mg)) {
		// ---------------- round 0, MSD path (msd.cuh) ----------------
		round0_done = msd_round0(st, v, lin, mg, 0u, 1u << mg.b1, *fo, so, &sa_buf, &hflag, S, unresolved, hstart);
	}
	if (!round0_done) {
		// ---------------- round 0, LSD path ----------------
		hflag.alloc(st, W);
		DevBuf<uint32_t> key0(st, W), key1(st, W), idx0(st, W), idx1(st, W);
		DevBuf<uint8_t> aux0(st, W), aux1(st, W);
		RadixRec<2> cur{{key0.get(), idx0.get()}, aux0.get()}, alt{{key1.get(), idx1.get()}, aux1.get()};
		unsigned const grid = (unsigned)div_up(W, 256);
		TRACE("r0 alloc");
		RadixStats rs;
		if (bits == 2 && v.packed) {
			// the first pass reads its records straight from the packed text
			radix_sort_suffix_keys(st, v, nshort, k0, cur, alt, &rs);
		} else {
			B3M_LAUNCH_T(st, "make_keys", W * 10ull, k_make_keys, grid, 256, 0, v, bits, k0, nshort, cur.a[0], cur.a[1], cur.aux);
			S.other_bytes += W * (1ull + 9ull);
			TRACE("r0 make_keys");
			radix_sort_bits<2>(st, cur, alt, 0, W, 0, 32, &rs);
		}
		TRACE("r0 radix");
		S.radix_passes += rs.passes; S.radix_bytes += rs.bytes; S.active_sum += W; S.rounds = 1;
		// ---------------- resolve (+ fused extraction) ----------------
		B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * RS_CSLOTS, st.s));
		unsigned const rgrid = (unsigned)div_up(W, RS_TILE);
		// key + index + aux in, suffix array + head flag (+ BWT code) out; second-key gathers are added below
		uint64_t const rbytes = W * (9ull + 4ull + 1ull);
		auto read_counters = [&](unsigned long long * hc) {
			std::vector<unsigned long long> hcs(4 * RS_CSLOTS);
			B3M_CUDA(cudaMemcpyAsync(hcs.data(), counters.get(), 32 * RS_CSLOTS, cudaMemcpyDeviceToHost, st.s));
			B3M_CUDA(cudaStreamSynchronize(st.s));
			for (int c = 0; c < 4; ++c) hc[c] = 0;
			for (int q = 0; q < RS_CSLOTS; ++q) for (int c = 0; c < 3; ++c) hc[c] += hcs[4 * q + c];
		};
		unsigned long long hc[4];
		if (fo) {
			// fused: emit BWT / anchors / samples and skip the order; it is only needed if something stays unresolved
			bool const stream_sa = so && so->host_sa && fo->sa_s && st.copy && rgrid >= 64;
			// BWA words: find the row of the suffix at position 0 first (its tile is resolved ahead of the rest:
			// resolving a tile twice writes the same values), then every chunk of rows can be packed and sent
			bool stream_bwa = so && so->host_bwa && so->d_bwa && fo->has_term && st.copy && rgrid >= 64 && bits == 2 && W >= 64;
			uint64_t primary = 0;
			if (stream_bwa) {
				uint32_t key0 = 0;
				{ // the first key of the suffix at position 0, from the packed text (W >= 64: no padding)
					uint64_t w0 = 0;
					B3M_CUDA(cudaMemcpyAsync(&w0, v.packed, 8, cudaMemcpyDeviceToHost, st.s));
					B3M_CUDA(cudaStreamSynchronize(st.s));
					key0 = (uint32_t)(w0 >> 32);
				}
				B3M_LAUNCH(st, k_lower_bound, 1, 1, 0, (const uint32_t *)cur.a[0], (uint32_t)W, key0, d_total);
				uint32_t const lb = fetch_u32(st, d_total);
				unsigned const t0 = lb / RS_TILE, tn = std::min<unsigned>(2u, rgrid - t0);
				B3M_CUDA(cudaMemsetAsync(fo->special, 0xff, 8, st.s));
				RS_LAUNCH("resolve_extract", (uint64_t)tn * RS_TILE * 10ull, true, false, tn, v, bits, k0, lin,
				             (const uint32_t *)cur.a[0], (const uint32_t *)cur.a[1], (const uint8_t *)cur.aux, W, t0, alt.a[1], hflag.get(), *fo, counters.get());
				uint32_t const row0 = fetch_u32(st, fo->special);
				if (row0 == 0xffffffffu) stream_bwa = false; // in a run too long for the CTA: no early primary
				else primary = row0;
				B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * RS_CSLOTS, st.s)); // those tiles are counted again below
			}
			uint64_t const seq_len = W, nwords_bwa = (W + 15) >> 4; // terminated text: n - 1 = W bases
			uint64_t words_done = 0;
			unsigned const nchunks = (stream_sa || stream_bwa) ? 16u : 1u; // the last chunk's copies are the exposed tail
			for (unsigned c = 0; c < nchunks; ++c) {
				unsigned const t_lo = (unsigned)((uint64_t)rgrid * c / nchunks), t_hi = (unsigned)((uint64_t)rgrid * (c + 1) / nchunks);
				RS_LAUNCH("resolve_extract", (uint64_t)(t_hi - t_lo) * RS_TILE * 10ull, true, false, t_hi - t_lo, v, bits, k0, lin,
				             (const uint32_t *)cur.a[0], (const uint32_t *)cur.a[1], (const uint8_t *)cur.aux, W, t_lo, alt.a[1], hflag.get(), *fo, counters.get());
				if (stream_sa || stream_bwa) {
					// rows below t_hi * RS_TILE + shift are final once these tiles are done: their SA samples and BWA words go to the host now
					uint64_t const rows_lo = c ? (uint64_t)t_lo * RS_TILE + fo->shift : 0, rows_hi = (c + 1 == nchunks) ? W + fo->shift : (uint64_t)t_hi * RS_TILE + fo->shift;
					cudaEvent_t ev;
					B3M_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
					B3M_CUDA(cudaEventRecord(ev, st.s));
					B3M_CUDA(cudaStreamWaitEvent(st.copy, ev, 0));
					B3M_CUDA(cudaEventDestroy(ev));
					if (stream_sa) {
						uint64_t const k_lo = div_up(rows_lo, 1ull << fo->salog), k_hi = std::min<uint64_t>(div_up(rows_hi, 1ull << fo->salog), so->nsa);
						if (k_hi > k_lo) B3M_CUDA(cudaMemcpyAsync(so->host_sa + k_lo, fo->sa_s + k_lo, (k_hi - k_lo) * 8, cudaMemcpyDeviceToHost, st.copy));
					}
					if (stream_bwa) {
						// word w reads the rows 16w .. 16w+16 (one further behind the primary)
						uint64_t const w_hi = (c + 1 == nchunks) ? nwords_bwa : (rows_hi >= 17 ? std::min<uint64_t>((rows_hi - 17) / 16 + 1, nwords_bwa) : 0);
						if (w_hi > words_done) {
							k9_pack_bwa_range(st.copy, fo->bwt, seq_len, primary, so->d_bwa, words_done, w_hi);
							++st.launches;
							B3M_CUDA(cudaMemcpyAsync(so->host_bwa + words_done, so->d_bwa + words_done, (w_hi - words_done) * 4, cudaMemcpyDeviceToHost, st.copy));
							words_done = w_hi;
						}
					}
				}
			}
			read_counters(hc);
			S.other_bytes += W * 10ull + 32ull * hc[2];
			if (stream_sa) so->delivered = hc[0] == 0; // otherwise the samples are rewritten after the doubling rounds
			if (stream_bwa) so->bwa_delivered = hc[0] == 0;
			if (hc[0]) {
				B3M_CUDA(cudaMemsetAsync(counters.get(), 0, 32 * RS_CSLOTS, st.s));
				RS_LAUNCH("resolve", rbytes, false, true, rgrid, v, bits, k0, lin, (const uint32_t *)cur.a[0],
				             (const uint32_t *)cur.a[1], (const uint8_t *)cur.aux, W, 0u, alt.a[1], hflag.get(), FusedOut(), counters.get());
				read_counters(hc);
				S.other_bytes += rbytes + 32ull * hc[2];
			}
		} else {
			RS_LAUNCH("resolve", rbytes, false, true, rgrid, v, bits, k0, lin, (const uint32_t *)cur.a[0],
			             (const uint32_t *)cur.a[1], (const uint8_t *)cur.aux, W, 0u, alt.a[1], hflag.get(), FusedOut(), counters.get());
			read_counters(hc);
			S.other_bytes += rbytes + 32ull * hc[2];
		}
		unresolved = hc[0];
		S.tied0 = hc[1]; S.unresolved0 = hc[0];
		sa_buf = (alt.a[1] == idx1.get()) ? std::move(idx1) : std::move(idx0);
		TRACE("r0 resolve");
	}
	uint32_t * const sa = sa_buf.get();

	if (unresolved == 0) {
		if (rank) {
			B3M_LAUNCH_T(st, "rank_scatter", W * 36ull, k_rank_scatter, (unsigned)div_up(W, 256), 256, 0, (const uint32_t *)sa, W, rank);
			S.other_bytes += W * 36ull;
		}
	} else {
		// ---------------- prefix doubling on the suffixes that are still tied ----------------
		DevBuf<uint32_t> own_rank;
		if (!rank) { own_rank.alloc(st, W); rank = own_rank.get(); }
		uint64_t na = 0;
		DevBuf<uint32_t> pool[6];
		{
			DevBuf<uint32_t> grpb(st, W);
			uint32_t * grp = grpb.get();
			const uint8_t * hf = hflag.get();
			uint64_t const Wm = W;
			// head flags -> group head position (max-scan); rank of every suffix = head of its group
			// rank array beyond the L2 size: see k_rank_write (sortpath=msd forces it on small texts, for the tests)
			bool const part_ranks = W * 4ull > (64ull << 20) ? st.sortpath != B3M_SORT_LSD : st.sortpath == B3M_SORT_MSD;
			if (!part_ranks) {
				scan_apply<OpMax>(st, W,
					[=] __device__(uint64_t k) -> uint32_t { return (k && hf[k]) ? (uint32_t)k : 0u; },
					[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) {
						uint32_t const head = excl > v0 ? excl : v0;
						grp[k] = head;
						rank[sa[k]] = head;
					}, "heads_rank_scatter", W * 42ull);
				S.other_bytes += W * (2ull + 4ull + 4ull + 32ull);
			} else {
				scan_apply<OpMax>(st, W,
					[=] __device__(uint64_t k) -> uint32_t { return (k && hf[k]) ? (uint32_t)k : 0u; },
					[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) { grp[k] = excl > v0 ? excl : v0; }, "heads", W * 6ull);
				S.other_bytes += W * 6ull;
				DevBuf<uint32_t> t0(st, W + 4), t1(st, W + 4);
				rank_write_partitioned(st, sa, grp, W, t0.get(), t1.get(), W, rank, "heads_rank_write", S);
				B3M_CUDA(cudaStreamSynchronize(st.s)); // t0 / t1 are released here
			}
			TRACE("heads+rank scatter");
			auto active = [=] __device__(uint64_t k) -> uint32_t {
				bool const hk = grp[k] == (uint32_t)k;
				bool const hn = (k + 1 == Wm) || (grp[k + 1] == (uint32_t)(k + 1));
				return (hk && hn) ? 0u : 1u;
			};
			B3M_CUDA(cudaMemsetAsync(d_total, 0, 4, st.s));
			scan_apply<OpSum>(st, W, active,
				[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) { if (k + 1 == Wm) *d_total = excl + v0; });
			na = fetch_u32(st, d_total);
			S.other_bytes += W * 8ull;
			if (na) {
				for (int b = 0; b < 6; ++b) pool[b].alloc(st, na);
				uint32_t * agrp = pool[0].get();
				uint32_t * aidx = pool[1].get();
				scan_apply<OpSum>(st, W, active,
					[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) {
						if (v0) { agrp[excl] = grp[k]; aidx[excl] = sa[k]; }
					});
				S.other_bytes += W * 8ull + na * 8ull;
			}
			TRACE("compact");
		}
		hflag.release();
		uint32_t * bufs[6];
		for (int b = 0; b < 6; ++b) bufs[b] = pool[b].get();
		// roles: bufs[0]=grp, bufs[1]=idx, bufs[2]=key2, bufs[3..5]=ping-pong partners
		uint64_t h = hstart;
		int const bw = (int)ceil_log2_u64(W + 2);
		// one round of the radix path on the list b[0] = group, b[1] = suffix (n entries; b[2] = key if key_ready): gather the
		// rank ahead, sort by (group, rank ahead), split the groups; returns the entries still active, left in b[0], b[1]
		auto radix_round = [&](uint32_t ** b, uint64_t n, bool key_ready) -> uint64_t {
			unsigned const grid = (unsigned)div_up(n, 256);
			if (!key_ready) {
				B3M_LAUNCH_T(st, "gather_ahead", n * 40ull, k_gather_ahead, grid, 256, 0, (const uint32_t *)b[1], n, (const uint32_t *)rank, h, W, circular, b[2]);
				S.other_bytes += n * (4ull + 32ull + 4ull);
			}
			TRACE("rN gather");
			RadixRec<3> cur{{b[0], b[2], b[1]}}, alt{{b[3], b[4], b[5]}};
			RadixStats rs;
			radix_sort_bits<3>(st, cur, alt, 1, n, 0, bw, &rs); // rank ahead (minor key)
			radix_sort_bits<3>(st, cur, alt, 0, n, 0, bw, &rs); // group (major key)
			S.radix_passes += rs.passes; S.radix_bytes += rs.bytes;
			TRACE("rN radix");
			const uint32_t * sg = cur.a[0];
			const uint32_t * sk = cur.a[1];
			const uint32_t * si = cur.a[2];
			uint32_t * ngrp = alt.a[0];
			uint64_t const nam = n;
			scan_apply<OpMaxMax>(st, n,
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
				}, "split_scatter", n * 92ull);
			S.other_bytes += n * (2 * 12ull + 4ull + 32ull + 32ull);
			auto active = [=] __device__(uint64_t a) -> uint32_t {
				bool const hk = (a == 0) || (ngrp[a] != ngrp[a - 1]);
				bool const hn = (a + 1 == nam) || (ngrp[a + 1] != ngrp[a]);
				return (hk && hn) ? 0u : 1u;
			};
			uint32_t * ogrp = alt.a[1];
			uint32_t * oidx = alt.a[2];
			scan_apply<OpSum>(st, n, active,
				[=] __device__(uint64_t a, uint32_t excl, uint32_t v0) {
					if (v0) { ogrp[excl] = ngrp[a]; oidx[excl] = si[a]; }
					if (a + 1 == nam) *d_total = excl + v0;
				});
			S.other_bytes += n * (2 * 8ull + 8ull);
			uint64_t const nn = fetch_u32(st, d_total);
			TRACE("rN split+compact");
			// next round: grp = alt[1], idx = alt[2]; everything else is free
			uint32_t * nbuf[6] = {alt.a[1], alt.a[2], cur.a[0], cur.a[1], cur.a[2], alt.a[0]};
			for (int q = 0; q < 6; ++q) b[q] = nbuf[q];
			return nn;
		};
		bool const tile_rounds = st.sortpath != B3M_SORT_LSD; // B3M_SORT_LSD keeps the plain radix rounds (tests compare the two)
		while (na) {
			if (circular && h >= W) break; // non-primitive text: ties stay in current order (unpinned)
			S.active_sum += na; S.rounds++;
			if (tile_rounds) {
				// groups of at most DT_GMAX members are sorted inside a CTA; the members of larger ones are passed through
				uint32_t * cg = bufs[0], * ci = bufs[1], * og = bufs[2], * oi = bufs[3];
				uint8_t * of = reinterpret_cast<uint8_t *>(bufs[4]);
				uint32_t const ntiles = (uint32_t)div_up(na, DT_TILE);
				DevBuf<uint32_t> tcount(st, (uint64_t)ntiles + 1);
				B3M_CUDA(cudaMemsetAsync(tcount.get(), 0, tcount.bytes(), st.s));
				B3M_CUDA(cudaMemsetAsync(d_total + 1, 0, 8, st.s));
				B3M_LAUNCH_T(st, "dbl_tile", na * 59ull, k_dbl_tile, ntiles, DT_THREADS, 0, (const uint32_t *)cg, (const uint32_t *)ci, (uint32_t)na,
				             (const uint32_t *)rank, sa, h, W, circular, og, oi, of, d_total + 1, tcount.get());
				S.other_bytes += na * (8ull + 32ull + 4ull + 9ull);
				uint64_t const nbig = fetch_u32(st, d_total + 1), nchg = fetch_u32(st, d_total + 2);
				TRACE("rN tile");
				if (nbig <= na / 4) {
					DevBuf<uint32_t> bpool[6];
					uint32_t * bb[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
					if (nbig) {
						// the larger groups: their members, in list order, and the ranks ahead of them (before any rank of this round is written)
						for (int q = 0; q < 6; ++q) { bpool[q].alloc(st, nbig); bb[q] = bpool[q].get(); }
						uint32_t * bg = bb[0], * bi = bb[1];
						scan_apply<OpSum>(st, na,
							[=] __device__(uint64_t k) -> uint32_t { return (of[k] >> 1) & 1u; },
							[=] __device__(uint64_t k, uint32_t excl, uint32_t v0) { if (v0) { bg[excl] = og[k]; bi[excl] = oi[k]; } });
						B3M_LAUNCH_T(st, "gather_ahead", nbig * 40ull, k_gather_ahead, (unsigned)div_up(nbig, 256), 256, 0, (const uint32_t *)bi, nbig, (const uint32_t *)rank, h, W, circular, bb[2]);
						S.other_bytes += na * 9ull + nbig * 48ull;
					}
					// the changed ranks are written now; members of groups still tied make the next list
					scan_exclusive_inplace<OpSum>(st, tcount.get(), (uint64_t)ntiles + 1);
					// scattered rank stores cost a 64-byte read-modify-write each once the rank array outgrows the L2 (27 G stores/s);
					// partitioning the (suffix, rank) pairs of the round streams 53 B per list entry: taken when a quarter of the
					// list or more changes its rank (cfg4: compaction + rank stores 697 -> 507 ms over the 11 rounds)
					bool const part = (W * 4ull > (64ull << 20) || st.sortpath == B3M_SORT_MSD) && nchg * 4ull > na;
					if (!part) {
						B3M_LAUNCH_T(st, "dbl_compact", na * 49ull, k_dbl_compact<true>, ntiles, DT_THREADS, 0, (const uint32_t *)og, oi, (const uint8_t *)of, (uint32_t)na,
						             (const uint32_t *)tcount.get(), rank, cg, ci);
						S.other_bytes += na * (9ull + 8ull) + nchg * 32ull;
					} else {
						B3M_LAUNCH_T(st, "dbl_compact", na * 21ull, k_dbl_compact<false>, ntiles, DT_THREADS, 0, (const uint32_t *)og, oi, (const uint8_t *)of, (uint32_t)na,
						             (const uint32_t *)tcount.get(), rank, cg, ci);
						S.other_bytes += na * (9ull + 8ull + 4ull);
						rank_write_partitioned(st, oi, og, na, bufs[4], bufs[5], W, rank, "dbl_rank_write", S); // `of` (bufs[4]) is dead by now
					}
					B3M_CUDA(cudaMemcpyAsync(d_total, tcount.get() + ntiles, 4, cudaMemcpyDeviceToDevice, st.s));
					uint64_t nn = fetch_u32(st, d_total);
					TRACE("rN compact");
					if (nbig) {
						uint64_t const nn2 = radix_round(bb, nbig, true);
						if (nn2) {
							B3M_CUDA(cudaMemcpyAsync(cg + nn, bb[0], nn2 * 4, cudaMemcpyDeviceToDevice, st.s));
							B3M_CUDA(cudaMemcpyAsync(ci + nn, bb[1], nn2 * 4, cudaMemcpyDeviceToDevice, st.s));
							B3M_CUDA(cudaStreamSynchronize(st.s)); // bpool is released below
						}
						nn += nn2;
					}
					na = nn;
					h *= 2;
					continue;
				}
				// mostly large groups: the radix round redoes the whole list (the tile kernel wrote nothing that round reads)
			}
			na = radix_round(bufs, na, false);
			h *= 2;
		}
		if (fo) {
			B3M_LAUNCH_T(st, "extract_sample", W * 37ull, k_extract_sample, (unsigned)div_up(W, 256), 256, 0, v, (const uint32_t *)sa, W, *fo);
			S.other_bytes += W * 37ull;
		}
	}
	if (stats) { // accumulated over the leaves of a multi-block build
		stats->rounds = stats->rounds > S.rounds ? stats->rounds : S.rounds;
		stats->radix_passes += S.radix_passes; stats->radix_bytes += S.radix_bytes;
		stats->active_sum += S.active_sum; stats->other_bytes += S.other_bytes;
		stats->tied0 += S.tied0; stats->unresolved0 += S.unresolved0;
	}
}

} // namespace b3m
