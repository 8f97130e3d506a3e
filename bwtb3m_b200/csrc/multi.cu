// Multi-GPU build behind the C ABI (include/b3m.h, b3m_multi_*; `ngpus=` of bwtb3m): N engines in ONE
// process, one host thread per GPU, peer access between all of them.  The reference scales through its one
// call (numthreads workers inside BwtMergeSort::computeBwt, /root/reference/src/bwtb3m.cpp:48-50,62-63);
// here the same call spreads over the GPUs of a box:
//   load    every GPU uploads 1/N of the input file over ITS PCIe link, the pieces are exchanged with peer
//           copies over NVLink, every GPU decodes the text (K1);
//   build   position-sharded MSD sort (b3m_engine_xshard_*, sufsort.cu): count own tiles -> the counts meet in
//           host memory (threads share it: no collective) -> records cross NVLink as the stores of the scatter
//           kernel -> every GPU sorts its key range and stores BWT rows / anchors / samples into GPU 0's result
//           buffers (peer stores).  Texts that path does not take (more than four codes): key ranges of the
//           LSD sorter (b3m_engine_shard_build) with the same direct stores.  Texts with repeats neither sorts
//           completely: GPU 0 builds alone (the general path).
// No NCCL: inside one process the phases are ordered by joining the host threads.
#include "engine.h"
#include <string.h>
#include <time.h>
#include <algorithm>
#include <memory>
#include <mutex>
#include <thread>

namespace b3m {

namespace {
double mono_ms() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
} // namespace

struct MultiEngine {
	std::vector<int> dev;
	std::vector<std::unique_ptr<Engine>> e;
	// staged input file, one full copy per GPU
	std::vector<uint8_t *> stage;
	uint64_t stage_cap = 0;
	// record arrays of the position-sharded sort, one per GPU
	std::vector<unsigned long long *> recs;
	uint64_t rec_cap = 0;
	std::vector<unsigned long long *> d_tot; // [2048] bin counts of a GPU's tiles
	// result buffers on GPU 0, written by every GPU
	void * r_bwt = nullptr, * r_pre = nullptr, * r_sa = nullptr, * r_isa = nullptr, * r_special = nullptr;
	uint64_t rb_bwt = 0, rb_pre = 0, rb_sa = 0, rb_isa = 0;
	// what the last build did
	std::string strategy;
	double ms_load = 0, ms_build = 0;

	unsigned N() const { return (unsigned)dev.size(); }

	// f(p) on one host thread per GPU; the first exception is rethrown on the caller's thread
	template <typename F>
	void parallel(F f) {
		std::vector<std::thread> th;
		std::mutex mu;
		std::string first;
		for (unsigned p = 0; p < N(); ++p)
			th.emplace_back([&, p]() {
				try {
					B3M_CUDA(cudaSetDevice(dev[p]));
					f(p);
				} catch (std::exception const & ex) {
					std::lock_guard<std::mutex> g(mu);
					if (first.empty()) first = std::string("GPU ") + std::to_string(dev[p]) + ": " + ex.what();
				} catch (...) {
					std::lock_guard<std::mutex> g(mu);
					if (first.empty()) first = "unknown error";
				}
			});
		for (auto & t : th) t.join();
		if (!first.empty()) throw Error(first);
	}

	MultiEngine(int ngpus, const int * devices) {
		int ndev = 0;
		cudaError_t const ce = cudaGetDeviceCount(&ndev);
		if (ce != cudaSuccess || ndev <= 0)
			throw Error(std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); this library has no CPU fallback");
		B3M_REQUIRE(ngpus >= 1 && ngpus <= 16, "ngpus must be 1..16");
		for (int p = 0; p < ngpus; ++p) {
			int const d = devices ? devices[p] : p;
			if (d < 0 || d >= ndev) throw Error("ngpus=" + std::to_string(ngpus) + ": device " + std::to_string(d) + " does not exist (" + std::to_string(ndev) + " visible)");
			for (int q : dev) B3M_REQUIRE(q != d, "the same device is listed twice");
			dev.push_back(d);
		}
		for (unsigned p = 0; p < N(); ++p) {
			B3M_CUDA(cudaSetDevice(dev[p]));
			for (unsigned q = 0; q < N(); ++q) {
				if (q == p) continue;
				int can = 0;
				B3M_CUDA(cudaDeviceCanAccessPeer(&can, dev[p], dev[q]));
				if (!can) throw Error("devices " + std::to_string(dev[p]) + " and " + std::to_string(dev[q]) + " are not peers: the multi-GPU build needs NVLink / peer access");
				cudaError_t const pe = cudaDeviceEnablePeerAccess(dev[q], 0);
				if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) B3M_CUDA(pe);
				cudaGetLastError();
			}
		}
		e.resize(N());
		stage.assign(N(), nullptr);
		recs.assign(N(), nullptr);
		d_tot.assign(N(), nullptr);
		parallel([&](unsigned p) {
			e[p].reset(new Engine(dev[p], nullptr));
			B3M_CUDA(cudaMalloc((void **)&d_tot[p], 2048 * 8));
		});
	}

	void free_on(unsigned p, void * ptr) {
		if (!ptr) return;
		cudaSetDevice(dev[p]);
		cudaFree(ptr);
	}

	~MultiEngine() {
		for (unsigned p = 0; p < N(); ++p) {
			if (e[p]) { cudaSetDevice(dev[p]); cudaDeviceSynchronize(); }
		}
		e.clear(); // engines first: engine 0 may refer to the result buffers
		for (unsigned p = 0; p < N(); ++p) { free_on(p, stage[p]); free_on(p, recs[p]); free_on(p, d_tot[p]); }
		free_on(0, r_bwt); free_on(0, r_pre); free_on(0, r_sa); free_on(0, r_isa); free_on(0, r_special);
	}

	// K1 on every GPU; the file crosses PCIe once, in N pieces over N links
	void load_host(const void * input, uint64_t nbytes, int itype) {
		B3M_REQUIRE(input && nbytes, "empty input");
		double const t0 = mono_ms();
		e[0]->reset_results(); // may refer to result buffers that are resized below
		if (nbytes + 64 > stage_cap) {
			parallel([&](unsigned p) {
				if (stage[p]) { B3M_CUDA(cudaFree(stage[p])); stage[p] = nullptr; }
				B3M_CUDA(cudaMalloc((void **)&stage[p], nbytes + 64));
			});
			stage_cap = nbytes + 64;
		}
		uint64_t const piece = (div_up(nbytes, N()) + 255) & ~255ull;
		auto lo = [&](unsigned q) { return std::min<uint64_t>(nbytes, piece * q); };
		parallel([&](unsigned p) {
			cudaStream_t const s = e[p]->st.s;
			if (lo(p + 1) > lo(p))
				B3M_CUDA(cudaMemcpyAsync(stage[p] + lo(p), (const uint8_t *)input + lo(p), lo(p + 1) - lo(p), cudaMemcpyHostToDevice, s));
			B3M_CUDA(cudaStreamSynchronize(s));
		});
		parallel([&](unsigned p) {
			cudaStream_t const s = e[p]->st.s;
			for (unsigned k = 1; k < N(); ++k) {
				unsigned const q = (p + k) % N(); // every GPU starts with another peer
				if (lo(q + 1) > lo(q))
					B3M_CUDA(cudaMemcpyPeerAsync(stage[p] + lo(q), dev[p], stage[q] + lo(q), dev[q], lo(q + 1) - lo(q), s));
			}
			e[p]->load(stage[p], nbytes, itype, true);
		});
		ms_load = mono_ms() - t0;
	}

	void ensure_results(uint64_t n, uint64_t npre, uint64_t nsa, uint64_t nisa) {
		B3M_CUDA(cudaSetDevice(dev[0]));
		auto fit = [&](void *& ptr, uint64_t & have, uint64_t want) {
			if (want <= have && ptr) return;
			if (ptr) { B3M_CUDA(cudaFree(ptr)); ptr = nullptr; have = 0; }
			B3M_CUDA(cudaMalloc(&ptr, want ? want : 16));
			have = want;
		};
		fit(r_bwt, rb_bwt, n + 16);
		fit(r_pre, rb_pre, 4 * npre + 16);
		fit(r_sa, rb_sa, 8 * nsa + 16);
		fit(r_isa, rb_isa, 8 * nisa + 16);
		if (!r_special) B3M_CUDA(cudaMalloc(&r_special, 64));
	}

	void build(b3m_build_params const & p0) {
		B3M_REQUIRE(e[0]->loaded, "no input loaded");
		double const t0 = mono_ms();
		b3m_build_params p = p0;
		uint64_t const n = e[0]->T.n;
		if (!p.preisarate) p.preisarate = p.bwtonly ? 64 : e[0]->choose_preisarate_pub(n);
		if (N() == 1 || p.numblocks > 1 || p.sampling == B3M_SAMPLING_WALK) {
			// one GPU, or a request only the block path serves (forced blocks, the LF walk)
			e[0]->build(p);
			strategy = "single";
			ms_build = mono_ms() - t0;
			return;
		}
		p.numblocks = 1;
		uint64_t const npre = div_up(n, p.preisarate);
		uint64_t const nsa = p.bwtonly ? 0 : div_up(n, p.sasamplingrate), nisa = p.bwtonly ? 0 : div_up(n, p.isasamplingrate);
		e[0]->reset_results();
		ensure_results(n, npre, nsa, nisa);
		void * const o_sa = p.bwtonly ? nullptr : r_sa, * const o_isa = p.bwtonly ? nullptr : r_isa;
		std::vector<uint64_t> unres(N(), 0);
		bool done = false;
		// ---- position-sharded MSD sort ----
		{
			std::vector<uint32_t> nbins(N(), 0);
			std::vector<uint64_t> alltot((size_t)N() * 2048, 0);
			parallel([&](unsigned q) {
				e[q]->xs_count(q, N(), p, d_tot[q], &nbins[q]);
				if (nbins[q]) {
					B3M_CUDA(cudaMemcpyAsync(alltot.data() + (size_t)q * 2048, d_tot[q], nbins[q] * 8, cudaMemcpyDeviceToHost, e[q]->st.s));
					B3M_CUDA(cudaStreamSynchronize(e[q]->st.s));
				}
			});
			uint32_t const nb = nbins[0];
			if (nb) {
				// [N][nb] as xshard_scatter reads it
				std::vector<uint64_t> tot((size_t)N() * nb);
				for (unsigned q = 0; q < N(); ++q) memcpy(tot.data() + (size_t)q * nb, alltot.data() + (size_t)q * 2048, nb * 8);
				uint64_t const cap = n / N() + n / (4 * N()) + (1u << 22);
				if (cap > rec_cap) {
					parallel([&](unsigned q) {
						if (recs[q]) { B3M_CUDA(cudaFree(recs[q])); recs[q] = nullptr; }
						B3M_CUDA(cudaMalloc((void **)&recs[q], cap * 8));
					});
					rec_cap = cap;
				}
				std::vector<uint64_t> caps(N(), rec_cap);
				std::vector<void *> rp(recs.begin(), recs.end());
				bool fits = true;
				try {
					parallel([&](unsigned q) {
						e[q]->xs_scatter(tot.data(), rp.data(), caps.data());
						B3M_CUDA(cudaStreamSynchronize(e[q]->st.s));
					});
				} catch (Error const &) { fits = false; } // a key range larger than its array: decided alike on every GPU, before any store
				if (fits) {
					parallel([&](unsigned q) {
						e[q]->xs_finish(recs[q], r_bwt, r_pre, o_sa, o_isa, r_special, &unres[q]);
						B3M_CUDA(cudaStreamSynchronize(e[q]->st.s));
					});
					uint64_t u = 0;
					for (auto x : unres) u += x;
					if (u == 0) { done = true; strategy = "xshard"; }
					else strategy = "repeats";
				}
			}
		}
		// ---- key ranges of the LSD sorter (alphabets the MSD sorter does not take) ----
		if (!done && strategy != "repeats") {
			parallel([&](unsigned q) {
				e[q]->kr_build_part(q, N(), p, r_bwt, r_pre, o_sa, o_isa, r_special, &unres[q]);
				B3M_CUDA(cudaStreamSynchronize(e[q]->st.s));
			});
			uint64_t u = 0;
			for (auto x : unres) u += x;
			if (u == 0) { done = true; strategy = "shard"; }
		}
		if (done) {
			B3M_CUDA(cudaSetDevice(dev[0]));
			e[0]->kr_finish(r_bwt, r_pre, o_sa, o_isa, r_special, N(), true);
		} else {
			// repeats longer than the sort keys: the general path (prefix doubling) on one GPU
			B3M_CUDA(cudaSetDevice(dev[0]));
			e[0]->build(p);
			strategy = "single (text with long repeats)";
		}
		ms_build = mono_ms() - t0;
	}
};

} // namespace b3m

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
struct b3m_multi {
	b3m::MultiEngine * m = nullptr;
	std::vector<b3m_engine> handles; // non-owning engine handles for the b3m_engine_* calls
	std::string err;
};

static void set_err(char * err, size_t errlen, const char * msg) {
	if (err && errlen) { strncpy(err, msg, errlen - 1); err[errlen - 1] = 0; }
}

#define B3M_MGUARD(h, ...)                                                   \
	if (!(h)) return 1;                                                      \
	try { __VA_ARGS__; (h)->err.clear(); return 0; }                         \
	catch (std::exception const & ex) { (h)->err = ex.what(); return 2; }    \
	catch (...) { (h)->err = "unknown error"; return 3; }

extern "C" {

int b3m_multi_create(int ngpus, const int * devices, b3m_multi ** out, char * err, size_t errlen) {
	if (!out) return 1;
	*out = nullptr;
	try {
		b3m_multi * h = new b3m_multi();
		try { h->m = new b3m::MultiEngine(ngpus, devices); } catch (...) { delete h; throw; }
		h->handles.resize(h->m->N());
		for (unsigned p = 0; p < h->m->N(); ++p) h->handles[p].e = h->m->e[p].get();
		*out = h;
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

void b3m_multi_destroy(b3m_multi * h) {
	if (!h) return;
	delete h->m;
	delete h;
}

const char * b3m_multi_last_error(const b3m_multi * h) { return h ? h->err.c_str() : "null handle"; }

int b3m_multi_load_host(b3m_multi * h, const void * input, uint64_t nbytes, int inputtype) {
	B3M_MGUARD(h, h->m->load_host(input, nbytes, inputtype));
}

int b3m_multi_build(b3m_multi * h, const b3m_build_params * p) {
	B3M_MGUARD(h, { if (!p) throw b3m::Error("null params"); h->m->build(*p); });
}

b3m_engine * b3m_multi_engine(b3m_multi * h, int i) {
	if (!h || i < 0 || (unsigned)i >= h->m->N()) return nullptr;
	return &h->handles[i];
}

int b3m_multi_stats(b3m_multi * h, char * strategy, size_t len, double * ms_load, double * ms_build) {
	B3M_MGUARD(h, {
		if (strategy && len) { strncpy(strategy, h->m->strategy.c_str(), len - 1); strategy[len - 1] = 0; }
		if (ms_load) *ms_load = h->m->ms_load;
		if (ms_build) *ms_build = h->m->ms_build;
	});
}

} // extern "C"
