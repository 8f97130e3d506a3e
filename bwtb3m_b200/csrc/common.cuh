// Shared helpers for the b3m CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include "../../include/b3m.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <stdexcept>
#include <vector>
#include <map>
#include <iterator>
#include <atomic>

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

// Device memory arena: a few large cudaMalloc slabs carved by a host-side first-fit free list.
// Every user of the arena enqueues on the engine's single stream, so a block may be handed out
// again as soon as it is freed (stream order protects it).  Measured reason: cudaMallocAsync
// stalled the host for 100-600 ms at unpredictable points of a 1 Gbp build.
struct Arena {
	struct Slab { char * base; size_t size; std::map<size_t, size_t> free_; };
	std::vector<Slab> slabs;
	std::map<void *, std::pair<int, size_t>> live; // ptr -> (slab, size)
	size_t in_use = 0, peak = 0, capacity = 0;
	static size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }
	~Arena() { release_all(); }
	void release_all() {
		for (auto & s : slabs) cudaFree(s.base);
		slabs.clear(); live.clear(); in_use = 0; capacity = 0;
	}
	void add_slab(size_t bytes) {
		Slab s; s.size = round_up(bytes);
		cudaError_t const e = cudaMalloc((void **)&s.base, s.size);
		if (e != cudaSuccess) { cudaGetLastError(); throw Error(std::string("out of device memory allocating ") + std::to_string(s.size >> 20) + " MiB: " + cudaGetErrorString(e)); }
		s.free_[0] = s.size;
		capacity += s.size;
		slabs.push_back(std::move(s));
	}
	// make sure one slab can hold `bytes` more; only reshapes the arena when nothing is live
	void reserve(size_t bytes) {
		bytes = round_up(bytes);
		for (auto & s : slabs) for (auto & f : s.free_) if (f.second >= bytes) return;
		if (live.empty()) release_all();
		add_slab(bytes);
	}
	// Slab `idx` (0: the text, 1: the working set of a build) holds at least `bytes`: it is created, or -- when
	// nothing in it is live -- replaced by a larger one.  Keeps the two apart, so that neither `mem=` / `numblocks=`
	// nor the choice of sorter is decided before the text has been seen (the working set is sized in build()).
	void ensure_slab(size_t idx, size_t bytes) {
		bytes = round_up(bytes);
		while (slabs.size() <= idx) {
			// placeholder slabs of the minimum size keep the index meaning stable
			add_slab(slabs.size() == idx ? bytes : 512);
		}
		Slab & s = slabs[idx];
		if (s.size >= bytes) return;
		for (auto & l : live) if (l.second.first == (int)idx) return; // in use: alloc() grows the arena if it must
		cudaFree(s.base);
		capacity -= s.size;
		s.size = bytes;
		s.free_.clear();
		cudaError_t const e = cudaMalloc((void **)&s.base, s.size);
		if (e != cudaSuccess) {
			cudaGetLastError();
			s.base = nullptr; s.size = 0;
			throw Error(std::string("out of device memory allocating ") + std::to_string(bytes >> 20) + " MiB: " + cudaGetErrorString(e));
		}
		s.free_[0] = s.size;
		capacity += s.size;
	}
	void * alloc(size_t bytes) {
		if (!bytes) return nullptr;
		bytes = round_up(bytes);
		for (int pass = 0; pass < 2; ++pass) {
			for (size_t si = 0; si < slabs.size(); ++si) {
				auto & fl = slabs[si].free_;
				for (auto it = fl.begin(); it != fl.end(); ++it) {
					if (it->second >= bytes) {
						size_t const off = it->first, sz = it->second;
						fl.erase(it);
						if (sz > bytes) fl[off + bytes] = sz - bytes;
						void * p = slabs[si].base + off;
						live[p] = std::make_pair((int)si, bytes);
						in_use += bytes; if (in_use > peak) peak = in_use;
						return p;
					}
				}
			}
			size_t const grow = bytes > ((size_t)1 << 30) ? bytes : ((size_t)1 << 30);
			add_slab(grow);
		}
		throw Error("arena allocation failed");
	}
	void free(void * p) {
		if (!p) return;
		auto it = live.find(p);
		if (it == live.end()) return;
		int const si = it->second.first; size_t sz = it->second.second;
		size_t off = (size_t)((char *)p - slabs[si].base);
		live.erase(it);
		in_use -= sz;
		auto & fl = slabs[si].free_;
		auto nx = fl.lower_bound(off);
		if (nx != fl.end() && off + sz == nx->first) { sz += nx->second; nx = fl.erase(nx); }
		if (nx != fl.begin()) { auto pv = std::prev(nx); if (pv->first + pv->second == off) { off = pv->first; sz += pv->second; fl.erase(pv); } }
		fl[off] = sz;
	}
};

// Opt-in per-kernel timing (CUDA events on the launching stream) for the roofline report.
struct KernelTimes {
	struct Rec { std::string name; uint64_t bytes; cudaEvent_t a, b; };
	bool on = false;
	std::vector<Rec> recs;
	std::vector<cudaEvent_t> spare;
	cudaEvent_t get() {
		cudaEvent_t e;
		if (!spare.empty()) { e = spare.back(); spare.pop_back(); return e; }
		cudaEventCreate(&e);
		return e;
	}
	void clear() { for (auto & r : recs) { spare.push_back(r.a); spare.push_back(r.b); } recs.clear(); }
	~KernelTimes() { clear(); for (auto e : spare) cudaEventDestroy(e); }
};

// Per-engine launch context: the stream every kernel goes to, the arena, and counters the bench reports.
struct Stream {
	cudaStream_t s = nullptr;
	cudaStream_t copy = nullptr;  // second stream: results that are final early travel to the host while the build runs
	uint64_t launches = 0;      // kernels launched by this library (bench: "gpu_launches")
	int sms = 148;              // multiprocessor count of the device
	Arena * arena = nullptr;
	KernelTimes kt;
	int sortpath = 0;           // B3M_SORT_*: which round-0 sort a whole-text build of a 2-bit alphabet takes
};

// launch + (when profiling is switched on) bracket with events; `bytes` = algorithmic HBM bytes
#define B3M_LAUNCH_T(st, label, bytes_, kernel, grid, block, smem, ...)                           \
	do {                                                                                          \
		if ((st).kt.on) {                                                                         \
			::b3m::KernelTimes::Rec r_{label, (uint64_t)(bytes_), (st).kt.get(), (st).kt.get()};   \
			cudaEventRecord(r_.a, (st).s);                                                        \
			kernel<<<(grid), (block), (smem), (st).s>>>(__VA_ARGS__);                             \
			cudaEventRecord(r_.b, (st).s);                                                        \
			(st).kt.recs.push_back(r_);                                                           \
		} else {                                                                                  \
			kernel<<<(grid), (block), (smem), (st).s>>>(__VA_ARGS__);                             \
		}                                                                                         \
		++(st).launches;                                                                          \
		B3M_CUDA(cudaGetLastError());                                                             \
	} while (0)

#define B3M_LAUNCH(st, kernel, grid, block, smem, ...)                                            \
	do {                                                                                          \
		kernel<<<(grid), (block), (smem), (st).s>>>(__VA_ARGS__);                                 \
		++(st).launches;                                                                          \
		B3M_CUDA(cudaGetLastError());                                                             \
	} while (0)

// Arena-backed device buffer.
template <typename T>
struct DevBuf {
	T * p = nullptr;
	size_t n = 0;
	Arena * a = nullptr;
	DevBuf() {}
	DevBuf(Stream & st, size_t count) { alloc(st, count); }
	DevBuf(DevBuf const &) = delete;
	DevBuf & operator=(DevBuf const &) = delete;
	DevBuf(DevBuf && o) noexcept : p(o.p), n(o.n), a(o.a) { o.p = nullptr; o.n = 0; }
	DevBuf & operator=(DevBuf && o) noexcept {
		if (this != &o) { release(); p = o.p; n = o.n; a = o.a; o.p = nullptr; o.n = 0; }
		return *this;
	}
	~DevBuf() { release(); }
	void alloc(Stream & st, size_t count) {
		release();
		a = st.arena; n = count;
		if (count) p = (T *)a->alloc(count * sizeof(T));
	}
	void release() {
		if (p) { a->free(p); p = nullptr; n = 0; }
	}
	T * get() const { return p; }
	size_t bytes() const { return n * sizeof(T); }
};

// true the first time it is called with this flag word on the current device (kernel attributes are per device,
// and the multi-GPU host runs one thread per device inside one process)
static inline bool first_on_device(std::atomic<uint64_t> & seen) {
	int dev = 0;
	B3M_CUDA(cudaGetDevice(&dev));
	uint64_t const bit = 1ull << (dev & 63);
	return !(seen.fetch_or(bit) & bit);
}

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
