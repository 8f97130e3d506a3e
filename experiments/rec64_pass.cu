// EXPERIMENT (standalone, not part of libb3m.so, not yet run on a GPU): does a radix pass get faster when a
// (key, index) record is ONE 64-bit word in global memory instead of two 32-bit arrays?
//
// the pass is bound by instruction issue; the static SASS mix
// (profiles/r03_onesweep_sass_mix.txt) says a 64-bit record saves one LDG, one STG, one LDS/STS pair, one barrier
// and the address arithmetic of one output array per record, and doubles the length of the scattered runs.
// The price: the keys are read twice (once alone for the ranking -- 32-bit loads at stride 8 -- and once inside
// the LDG.64 of the whole record right before the scatter; the second read should hit L1/L2).
//
// This program times, on the same random records and the same digit, (a) the product's kernel
// k_radix_onesweep<2, aux> on structure-of-arrays records and (b) k_onesweep_rec64 below, and checks that
// (b) produces exactly the permutation of (a).   make experiments && bin/exp_rec64_pass [log2 records, default 28]
#include "../bwtb3m_b200/csrc/radix.cuh"
#include <stdio.h>
#include <stdlib.h>

using namespace b3m;

template <bool AUX>
__global__ void __launch_bounds__(RADIX_THREADS, RADIX_CTAS_PER_SM)
k_onesweep_rec64(const unsigned long long * __restrict__ in, unsigned long long * __restrict__ out, const uint8_t * __restrict__ aux_in,
                 uint8_t * __restrict__ aux_out, int shift, uint32_t mask, const uint32_t * __restrict__ base /* [256] */,
                 unsigned long long * __restrict__ status /* [ntiles][256] */, uint32_t * __restrict__ ticket) {
	__shared__ uint16_t wcnt[RADIX_WARPS][RADIX_BINS];
	__shared__ uint32_t gbase[RADIX_BINS];
	__shared__ uint32_t wsum[RADIX_BINS / 32];
	extern __shared__ __align__(16) uint8_t radix_dyn[];
	unsigned long long * const srec = reinterpret_cast<unsigned long long *>(radix_dyn); // the tile's records in sorted order
	uint8_t * const saux = radix_dyn + (size_t)RADIX_TILE * 8;
	__shared__ uint32_t s_tile;
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < RADIX_WARPS * RADIX_BINS / 2; i += RADIX_THREADS) reinterpret_cast<uint32_t *>(&wcnt[0][0])[i] = 0;
	__syncthreads();
	uint32_t const tile = s_tile;
	uint64_t const tbase = (uint64_t)tile * RADIX_TILE;
	uint64_t const chunk = tbase + (uint64_t)w * (32 * RADIX_ITEMS);

	// keys alone for the ranking: the high word of every record (little-endian: word 2i+1)
	const uint32_t * const in32 = reinterpret_cast<const uint32_t *>(in);
	uint32_t k[RADIX_ITEMS];
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) k[j] = in32[2 * (chunk + j * 32 + lane) + 1];
	uint16_t slot[RADIX_ITEMS];
	uint16_t * mycnt = wcnt[w];
	unsigned const lt = lanemask_lt();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint32_t const d = (k[j] >> shift) & mask;
		unsigned const peers = warp_peers8(d);
		uint32_t const before = mycnt[d];
		__syncwarp();
		if ((peers & lt) == 0) mycnt[d] = (uint16_t)(before + __popc(peers));
		__syncwarp();
		slot[j] = (uint16_t)(before + __popc(peers & lt));
	}
	__syncthreads();
	bool const binthread = threadIdx.x < RADIX_BINS;
	uint32_t bs = 0, bincl = 0;
	volatile unsigned long long * stw = status + (uint64_t)tile * RADIX_BINS + (threadIdx.x & (RADIX_BINS - 1));
	if (binthread) {
		uint32_t const d = threadIdx.x;
		#pragma unroll
		for (int ww = 0; ww < RADIX_WARPS; ++ww) { uint32_t const t = wcnt[ww][d]; wcnt[ww][d] = (uint16_t)bs; bs += t; }
		*stw = (tile == 0 ? RADIX_FLAG_INC : RADIX_FLAG_AGG) | bs;
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
			*stw = RADIX_FLAG_INC | (unsigned long long)(excl + bs);
		}
		gbase[d] = base[d] + excl - dstart;
	}
	__syncthreads();
	// final slots (the keys die here), then the whole records: one 64-bit load, one 64-bit scatter
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) slot[j] = (uint16_t)(slot[j] + wcnt[w][(k[j] >> shift) & mask]);
	{
		unsigned long long r[RADIX_ITEMS];
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) r[j] = in[chunk + j * 32 + lane];
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) srec[slot[j]] = r[j];
	}
	__syncthreads();
	#pragma unroll
	for (int j = 0; j < RADIX_ITEMS; ++j) {
		uint32_t const s = j * RADIX_THREADS + threadIdx.x;
		unsigned long long const r = srec[s];
		uint32_t const o = gbase[((uint32_t)(r >> 32) >> shift) & mask] + s;
		out[o] = r;
		k[j] = o;
	}
	if (AUX) {
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) saux[slot[j]] = aux_in[chunk + j * 32 + lane];
		__syncthreads();
		#pragma unroll
		for (int j = 0; j < RADIX_ITEMS; ++j) aux_out[k[j]] = saux[j * RADIX_THREADS + threadIdx.x];
	}
}

__global__ void k_init(uint64_t n, uint32_t * key, uint32_t * idx, uint8_t * aux, unsigned long long * rec) {
	uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t z = i + 0x9E3779B97F4A7C15ull; // splitmix64
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	key[i] = (uint32_t)z; idx[i] = (uint32_t)i; aux[i] = (uint8_t)(z >> 40);
	rec[i] = ((unsigned long long)(uint32_t)z << 32) | (uint32_t)i;
}

__global__ void k_compare(uint64_t n, const uint32_t * key, const uint32_t * idx, const uint8_t * aux, const unsigned long long * rec, const uint8_t * aux64,
                          unsigned long long * bad) {
	uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (rec[i] != (((unsigned long long)key[i] << 32) | idx[i]) || aux[i] != aux64[i]) atomicAdd(bad, 1ull);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char ** argv) {
	int const lg = argc > 1 ? atoi(argv[1]) : 28;
	uint64_t const n = 1ull << lg; // whole tiles only
	uint32_t const ntiles = (uint32_t)(n / RADIX_TILE);
	int const shift = 8, reps = 5;
	uint32_t *key, *idx, *key2, *idx2, *base, *skip, *ticket;
	uint8_t *aux, *aux2, *aux3;
	unsigned long long *rec, *rec2, *status, *ghist, *bad;
	CK(cudaMalloc(&key, 4 * n)); CK(cudaMalloc(&idx, 4 * n)); CK(cudaMalloc(&key2, 4 * n)); CK(cudaMalloc(&idx2, 4 * n));
	CK(cudaMalloc(&aux, n)); CK(cudaMalloc(&aux2, n)); CK(cudaMalloc(&aux3, n));
	CK(cudaMalloc(&rec, 8 * n)); CK(cudaMalloc(&rec2, 8 * n));
	CK(cudaMalloc(&status, (size_t)ntiles * RADIX_BINS * 8)); CK(cudaMalloc(&ghist, RADIX_MAXDIG * RADIX_BINS * 8));
	CK(cudaMalloc(&base, (RADIX_MAXDIG * RADIX_BINS + 16) * 4)); CK(cudaMalloc(&bad, 8));
	skip = base + RADIX_MAXDIG * RADIX_BINS; ticket = skip + 4;
	k_init<<<(unsigned)((n + 255) / 256), 256>>>(n, key, idx, aux, rec);
	CK(cudaMemset(ghist, 0, RADIX_MAXDIG * RADIX_BINS * 8)); CK(cudaMemset(skip, 0, 16 * 4));
	k_radix_hist<<<148 * 8, 256>>>(key, n, shift, 1, 255u, ghist);
	k_radix_hist_scan<<<1, 256>>>(ghist, 1, n, base, skip);
	CK(cudaDeviceSynchronize());
	size_t const smem_soa = radix_smem_bytes<2, true>(), smem_rec = (size_t)RADIX_TILE * 9;
	CK(cudaFuncSetAttribute(k_radix_onesweep<2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_soa));
	CK(cudaFuncSetAttribute(k_onesweep_rec64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rec));
	RadixPassArgs<2> A;
	A.in[0] = key; A.in[1] = idx; A.out[0] = key2; A.out[1] = idx2; A.aux_in = aux; A.aux_out = aux2;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float ms_soa = 0, ms_rec = 0;
	for (int variant = 0; variant < 2; ++variant) {
		for (int r = 0; r < reps + 1; ++r) { // the first repetition is a warm-up
			CK(cudaMemset(status, 0, (size_t)ntiles * RADIX_BINS * 8)); CK(cudaMemset(ticket, 0, 4));
			CK(cudaEventRecord(e0));
			if (variant == 0) k_radix_onesweep<2, true, false, true><<<ntiles, RADIX_THREADS, smem_soa>>>(A, RadixTextSrc(), n, shift, 255u, base, status, ticket);
			else k_onesweep_rec64<true><<<ntiles, RADIX_THREADS, smem_rec>>>(rec, rec2, aux, aux3, shift, 255u, base, status, ticket);
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			CK(cudaGetLastError());
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
			if (r) (variant ? ms_rec : ms_soa) += ms / reps;
		}
	}
	CK(cudaMemset(bad, 0, 8));
	k_compare<<<(unsigned)((n + 255) / 256), 256>>>(n, key2, idx2, aux2, rec2, aux3, bad);
	unsigned long long hbad = 0;
	CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
	double const gb = 18.0 * (double)n / 1e9;
	printf("records 2^%d, digit at bit %d\n  SoA (key, idx, aux) : %.3f ms  %.0f GB/s\n  64-bit record + aux : %.3f ms  %.0f GB/s\n  outputs %s (%llu differing records)\n",
	       lg, shift, ms_soa, gb / (ms_soa * 1e-3), ms_rec, gb / (ms_rec * 1e-3), hbad ? "DIFFER" : "identical", hbad);
	return hbad ? 2 : 0;
}
