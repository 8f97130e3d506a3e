// Shared helpers for the b3m CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <vector>

namespace b3m {

struct Error : std::runtime_error {
	explicit Error(std::string const & s) : std::runtime_error(s) {}
};

#define B3M_CUDA(x)                                                                               \
	do {                                                                                          \
		cudaError_t e_ = (x);                                                                     \
		if (e_ != cudaSuccess)                                                                    \
			throw ::b3m::Error(std::string(#x) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
			                   ":" + std::to_string(__LINE__) + ")");                             \
	} while (0)

#define B3M_REQUIRE(cond, msg)                                                                    \
	do {                                                                                          \
		if (!(cond)) throw ::b3m::Error(std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
	} while (0)

// Per-engine launch context: the stream every kernel goes to, and counters the bench reports.
struct Stream {
	cudaStream_t s = nullptr;
	uint64_t launches = 0;      // kernels launched by this library (bench: "gpu_launches")
	int sms = 148;              // multiprocessor count of the device
};

#define B3M_LAUNCH(st, kernel, grid, block, smem, ...)                                            \
	do {                                                                                          \
		kernel<<<(grid), (block), (smem), (st).s>>>(__VA_ARGS__);                                 \
		++(st).launches;                                                                          \
		B3M_CUDA(cudaGetLastError());                                                             \
	} while (0)

// Stream-ordered device buffer (cudaMallocAsync from the device's default pool).
template <typename T>
struct DevBuf {
	T * p = nullptr;
	size_t n = 0;
	cudaStream_t s = nullptr;
	DevBuf() {}
	DevBuf(Stream & st, size_t count) { alloc(st, count); }
	DevBuf(DevBuf const &) = delete;
	DevBuf & operator=(DevBuf const &) = delete;
	DevBuf(DevBuf && o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
	DevBuf & operator=(DevBuf && o) noexcept {
		if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
		return *this;
	}
	~DevBuf() { release(); }
	void alloc(Stream & st, size_t count) {
		release();
		s = st.s; n = count;
		if (count) B3M_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), s));
	}
	void release() {
		if (p) { cudaFreeAsync(p, s); p = nullptr; n = 0; }
	}
	T * get() const { return p; }
	size_t bytes() const { return n * sizeof(T); }
};

static inline unsigned ceil_log2_u64(uint64_t v) {
	unsigned b = 0;
	while (b < 64 && (1ull << b) < v) ++b;
	return b;
}

__host__ __device__ static inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// 128-bit streaming loads/stores (read-only path, no L1 allocation) for single-use data.
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 * p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
	unsigned m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

} // namespace b3m
