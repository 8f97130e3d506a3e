// Device-wide prefix scans (reduce-then-scan, three launches per level, no inter-CTA spinning).
// Inputs are produced by a functor of the element index and outputs are consumed by a functor
// (index, exclusive prefix, own value), so flagging, ranking and compaction fuse into the scan.
#pragma once
#include "common.cuh"

namespace b3m {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct OpSum {
	typedef uint32_t T;
	__host__ __device__ static T identity() { return 0u; }
	__host__ __device__ static T apply(T a, T b) { return a + b; }
};
struct OpMax {
	typedef uint32_t T;
	__host__ __device__ static T identity() { return 0u; }
	__host__ __device__ static T apply(T a, T b) { return a > b ? a : b; }
};
// pair scans: .x and .y scanned independently
struct OpMaxMax {
	typedef uint2 T;
	__host__ __device__ static T identity() { return make_uint2(0u, 0u); }
	__host__ __device__ static T apply(T a, T b) { return make_uint2(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y); }
};
struct OpMaxSum {
	typedef uint2 T;
	__host__ __device__ static T identity() { return make_uint2(0u, 0u); }
	__host__ __device__ static T apply(T a, T b) { return make_uint2(a.x > b.x ? a.x : b.x, a.y + b.y); }
};
struct OpSumSum {
	typedef uint2 T;
	__host__ __device__ static T identity() { return make_uint2(0u, 0u); }
	__host__ __device__ static T apply(T a, T b) { return make_uint2(a.x + b.x, a.y + b.y); }
};

struct OpSum4 {
	typedef uint4 T;
	__host__ __device__ static T identity() { return make_uint4(0u, 0u, 0u, 0u); }
	__host__ __device__ static T apply(T a, T b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
};

__device__ __forceinline__ uint4 shfl_up_t(uint4 v, int d) {
	return make_uint4(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d),
	                  __shfl_up_sync(0xffffffffu, v.z, d), __shfl_up_sync(0xffffffffu, v.w, d));
}
__device__ __forceinline__ uint32_t shfl_up_t(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint2 shfl_up_t(uint2 v, int d) {
	return make_uint2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

// inclusive scan of one value per thread across the CTA; returns inclusive value, total in *total
template <typename Op>
__device__ __forceinline__ typename Op::T block_scan_inclusive(typename Op::T v, typename Op::T * total) {
	typedef typename Op::T T;
	__shared__ T warpsum[SCAN_THREADS / 32];
	__shared__ T tot;
	unsigned const lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		T const o = shfl_up_t(v, d);
		if (lane >= (unsigned)d) v = Op::apply(o, v);
	}
	if (lane == 31) warpsum[w] = v;
	__syncthreads();
	if (w == 0) {
		T s = lane < SCAN_THREADS / 32 ? warpsum[lane] : Op::identity();
		#pragma unroll
		for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
			T const o = shfl_up_t(s, d);
			if (lane >= (unsigned)d) s = Op::apply(o, s);
		}
		if (lane < SCAN_THREADS / 32) warpsum[lane] = s;
		if (lane == SCAN_THREADS / 32 - 1) tot = s;
	}
	__syncthreads();
	if (w > 0) v = Op::apply(warpsum[w - 1], v);
	*total = tot;
	__syncthreads();
	return v;
}

template <typename Op, typename In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(uint64_t n, In in, typename Op::T * partial) {
	typedef typename Op::T T;
	uint64_t const base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	T acc = Op::identity();
	#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		uint64_t const i = base + k;
		if (i < n) acc = Op::apply(acc, in(i));
	}
	T total;
	block_scan_inclusive<Op>(acc, &total);
	if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

template <typename Op, typename In, typename Out>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(uint64_t n, In in, Out out, const typename Op::T * prefix) {
	typedef typename Op::T T;
	uint64_t const base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	T v[SCAN_ITEMS];
	T acc = Op::identity();
	#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		uint64_t const i = base + k;
		v[k] = (i < n) ? in(i) : Op::identity();
		acc = Op::apply(acc, v[k]);
	}
	T total;
	T const incl = block_scan_inclusive<Op>(acc, &total);
	// exclusive prefix of this thread = inclusive of previous thread; recover by a shuffle + smem hop
	__shared__ T carry[SCAN_THREADS];
	carry[threadIdx.x] = incl;
	__syncthreads();
	T run = prefix ? prefix[blockIdx.x] : Op::identity();
	if (threadIdx.x > 0) run = Op::apply(run, carry[threadIdx.x - 1]);
	#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		uint64_t const i = base + k;
		if (i < n) out(i, run, v[k]);
		run = Op::apply(run, v[k]);
	}
}

template <typename Op, typename In, typename Out>
void scan_apply(Stream & st, uint64_t n, In in, Out out, const char * label = nullptr, uint64_t bytes = 0);

// exclusive scan of an array in place
template <typename Op>
void scan_exclusive_inplace(Stream & st, typename Op::T * a, uint64_t n) {
	typedef typename Op::T T;
	scan_apply<Op>(st, n,
		[=] __device__(uint64_t i) -> T { return a[i]; },
		[=] __device__(uint64_t i, T excl, T) { a[i] = excl; });
}

template <typename Op, typename In, typename Out>
void scan_apply(Stream & st, uint64_t n, In in, Out out, const char * label, uint64_t bytes) {
	typedef typename Op::T T;
	if (n == 0) return;
	uint64_t const ntiles = div_up(n, SCAN_TILE);
	if (ntiles == 1) {
		B3M_LAUNCH(st, (k_scan_tile<Op, In, Out>), 1, SCAN_THREADS, 0, n, in, out, (const T *)nullptr);
		return;
	}
	DevBuf<T> partial(st, ntiles);
	B3M_LAUNCH(st, (k_scan_reduce<Op, In>), (unsigned)ntiles, SCAN_THREADS, 0, n, in, partial.get());
	scan_exclusive_inplace<Op>(st, partial.get(), ntiles);
	if (label) B3M_LAUNCH_T(st, label, bytes, (k_scan_tile<Op, In, Out>), (unsigned)ntiles, SCAN_THREADS, 0, n, in, out, (const T *)partial.get());
	else B3M_LAUNCH(st, (k_scan_tile<Op, In, Out>), (unsigned)ntiles, SCAN_THREADS, 0, n, in, out, (const T *)partial.get());
}

} // namespace b3m
