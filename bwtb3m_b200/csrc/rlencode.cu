// K8: device-side run-length + Huffman encoding of the final BWT into the .bwt container of
// formats.h (replaces libmaus2::huffman::RLEncoder on the output side of
// BwtMergeSort::computeBwt, /root/reference/src/bwtb3m.cpp:63; reader contract
// /root/reference/src/bwtb3mdecoderl.cpp:27-34).  Runs are found by a flag scan, the two code
// tables are built on the host from 2 x 256 counters, code lengths are scanned per block and
// every thread packs its 16 runs into whole 32-bit words, so that only the compressed stream
// crosses PCIe (DNA: ~0.3 B/symbol instead of 1).
#include "engine.h"
#include "formats.h"
#include "scan.cuh"
#include <string.h>
#include <algorithm>

namespace b3m {

constexpr int RLE_THREADS = 256;
constexpr int RLE_ITEMS = RL_RUNS_PER_BLOCK / RLE_THREADS; // 16 runs per thread
static_assert(RLE_ITEMS * RLE_THREADS == (int)RL_RUNS_PER_BLOCK, "block geometry");

__global__ void __launch_bounds__(256)
k_rl_hist(const uint8_t * __restrict__ s, const uint32_t * __restrict__ start, uint64_t nruns, uint64_t n,
          unsigned long long * __restrict__ hsym, unsigned long long * __restrict__ hlen) {
	__shared__ uint32_t sh[2][256];
	sh[0][threadIdx.x] = 0; sh[1][threadIdx.x] = 0;
	__syncthreads();
	for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nruns; k += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t const b = start[k];
		uint64_t const e = k + 1 < nruns ? (uint64_t)start[k + 1] : n;
		uint64_t const l = e - b;
		atomicAdd(&sh[0][s[b]], 1u);
		atomicAdd(&sh[1][l < RL_LENBINS ? (uint32_t)l : 0u], 1u);
	}
	__syncthreads();
	if (sh[0][threadIdx.x]) atomicAdd(&hsym[threadIdx.x], (unsigned long long)sh[0][threadIdx.x]);
	if (sh[1][threadIdx.x]) atomicAdd(&hlen[threadIdx.x], (unsigned long long)sh[1][threadIdx.x]);
}

__global__ void __launch_bounds__(256)
k_rl_gather(const uint8_t * __restrict__ s, const uint32_t * __restrict__ start, uint64_t nruns, uint64_t n,
            uint8_t * __restrict__ osym, unsigned long long * __restrict__ olen) {
	uint64_t const k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nruns) return;
	uint64_t const b = start[k], e = k + 1 < nruns ? (uint64_t)start[k + 1] : n;
	osym[k] = s[b]; olen[k] = e - b;
}

struct RlTabs { uint32_t sym[256]; uint32_t len[256]; }; // (code length << 24) | code

__device__ __forceinline__ uint32_t run_bits(RlTabs const & tb, uint32_t sym, uint32_t l) {
	uint32_t b = tb.sym[sym] >> 24;
	if (l < RL_LENBINS) b += tb.len[l] >> 24;
	else b += (tb.len[0] >> 24) + 6u + (32u - __clz(l));
	return b;
}

// pass 1: 64-bit words per block
__global__ void __launch_bounds__(RLE_THREADS)
k_rl_blockbits(const uint8_t * __restrict__ s, const uint32_t * __restrict__ start, uint64_t nruns, uint64_t n, const RlTabs * __restrict__ tabs,
               uint32_t * __restrict__ blockwords, unsigned long long * __restrict__ blocksym) {
	__shared__ RlTabs tb;
	for (int i = threadIdx.x; i < 512; i += RLE_THREADS) (&tb.sym[0])[i] = (&tabs->sym[0])[i];
	__syncthreads();
	uint64_t const k0 = (uint64_t)blockIdx.x * RL_RUNS_PER_BLOCK + (uint64_t)threadIdx.x * RLE_ITEMS;
	uint32_t bits = 0;
	#pragma unroll 4
	for (int j = 0; j < RLE_ITEMS; ++j) {
		uint64_t const k = k0 + j;
		if (k < nruns) {
			uint64_t const b = start[k];
			uint64_t const e = k + 1 < nruns ? (uint64_t)start[k + 1] : n;
			bits += run_bits(tb, s[b], (uint32_t)(e - b));
		}
	}
	uint32_t total;
	block_scan_inclusive<OpSum>(bits, &total);
	if (threadIdx.x == 0) {
		blockwords[blockIdx.x] = (total + 63u) >> 6;
		blocksym[blockIdx.x] = start[(uint64_t)blockIdx.x * RL_RUNS_PER_BLOCK];
	}
}

// MSB-first bit packer over big-endian 32-bit words: complete interior words are stored, the
// first and last (shared with the neighbouring threads) are OR-ed in
struct WordPacker {
	uint32_t * out; uint64_t w; uint64_t acc; uint32_t fill; bool shared_first;
	__device__ WordPacker(uint32_t * o, uint64_t bitpos) : out(o), w(bitpos >> 5), acc(0), fill((uint32_t)(bitpos & 31)), shared_first((bitpos & 31) != 0) {}
	__device__ __forceinline__ void put(uint32_t v, uint32_t nb) { // nb <= 32, fill < 32
		if (!nb) return;
		acc |= (uint64_t)v << (64 - fill - nb);
		fill += nb;
		if (fill >= 32) {
			uint32_t const word = __byte_perm((uint32_t)(acc >> 32), 0, 0x0123);
			if (shared_first) { atomicOr(&out[w], word); shared_first = false; } else out[w] = word;
			acc <<= 32; fill -= 32; ++w;
		}
	}
	__device__ __forceinline__ void flush() {
		if (fill) atomicOr(&out[w], __byte_perm((uint32_t)(acc >> 32), 0, 0x0123));
	}
};

// pass 2: emit
__global__ void __launch_bounds__(RLE_THREADS)
k_rl_emit(const uint8_t * __restrict__ s, const uint32_t * __restrict__ start, uint64_t nruns, uint64_t n, const RlTabs * __restrict__ tabs,
          const unsigned long long * __restrict__ blockoff /* 64-bit words */, uint32_t * __restrict__ out) {
	__shared__ RlTabs tb;
	for (int i = threadIdx.x; i < 512; i += RLE_THREADS) (&tb.sym[0])[i] = (&tabs->sym[0])[i];
	__syncthreads();
	uint64_t const k0 = (uint64_t)blockIdx.x * RL_RUNS_PER_BLOCK + (uint64_t)threadIdx.x * RLE_ITEMS;
	uint32_t sym[RLE_ITEMS], len[RLE_ITEMS];
	uint32_t bits = 0;
	#pragma unroll
	for (int j = 0; j < RLE_ITEMS; ++j) {
		uint64_t const k = k0 + j;
		len[j] = 0; sym[j] = 0;
		if (k < nruns) {
			uint64_t const b = start[k];
			uint64_t const e = k + 1 < nruns ? (uint64_t)start[k + 1] : n;
			sym[j] = s[b]; len[j] = (uint32_t)(e - b);
			bits += run_bits(tb, sym[j], len[j]);
		}
	}
	uint32_t total;
	uint32_t const incl = block_scan_inclusive<OpSum>(bits, &total);
	if (!bits) return;
	WordPacker wp(out, (uint64_t)blockoff[blockIdx.x] * 64ull + (incl - bits));
	#pragma unroll
	for (int j = 0; j < RLE_ITEMS; ++j) {
		if (!len[j]) continue;
		uint32_t const ts = tb.sym[sym[j]];
		wp.put(ts & 0xffffffu, ts >> 24);
		if (len[j] < RL_LENBINS) { uint32_t const tl = tb.len[len[j]]; wp.put(tl & 0xffffffu, tl >> 24); }
		else {
			uint32_t const tl = tb.len[0];
			uint32_t const nb = 32u - __clz(len[j]);
			wp.put(tl & 0xffffffu, tl >> 24);
			wp.put(nb - 1, 6);
			wp.put(len[j], nb);
		}
	}
	wp.flush();
}

// BWT in the reference's symbol space on the device (pacterm: 0 at the terminator row)
void Engine::symbols_device(DevBuf<uint8_t> & out) {
	B3M_REQUIRE(have_results, "no results");
	B3M_REQUIRE(!ssa_only, "the engine holds sampled SA/ISA only (computed from an existing BWT)");
	out.alloc(st, T.n + 16);
	DevBuf<uint8_t> dlut(st, 256);
	memset(pinned, 0, 256);
	for (uint32_t c = 0; c < T.sigma; ++c) pinned[c] = code2sym[c];
	B3M_CUDA(cudaMemcpyAsync(dlut.get(), pinned, 256, cudaMemcpyHostToDevice, st.s));
	k1_map_bytes(st, bwt.get(), T.n, dlut.get(), out.get());
	if (T.has_term) B3M_CUDA(cudaMemsetAsync(out.get() + root_exc_pos, 0, 1, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s)); // pinned is reused by the caller
}

// ------------------------------------------------------------------------------------------
// K9: BWA's packed BWT straight from the device BWT (engine half of MausFmToBwaConversion::rewrite,
// /root/reference/src/bwtb3mtobwa.cpp:29; layout: bwa's bwt_dump_bwt): 16 symbols per
// uint32, symbol k at bits (15-(k&15))*2, the terminator row (primary) removed, so rows behind it
// move up by one.  pacterm codes 0..3 are BWA's A,C,G,T.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_bwa(const uint8_t * __restrict__ bwt, uint64_t seq_len, uint64_t primary, uint32_t * __restrict__ words, uint64_t w_lo, uint64_t nwords) {
	uint64_t const w = w_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords) return;
	uint64_t const k0 = w << 4;
	uint32_t acc = 0;
	if (k0 + 16 <= seq_len && (k0 + 16 <= primary || k0 >= primary)) {
		// 16 rows on one side of the primary: 16 consecutive bytes (unaligned when shifted)
		const uint8_t * src = bwt + k0 + (k0 >= primary ? 1 : 0);
		#pragma unroll
		for (int j = 0; j < 16; ++j) acc = (acc << 2) | (uint32_t)(src[j] & 3u);
	} else {
		for (uint64_t k = k0; k < k0 + 16; ++k) {
			uint32_t c = 0;
			if (k < seq_len) c = bwt[k < primary ? k : k + 1] & 3u;
			acc = (acc << 2) | c;
		}
	}
	words[w] = acc;
}

// ------------------------------------------------------------------------------------------
// 2-bit transport packing of BWT rows (codes < 4) for the multi-GPU slice exchange: row lo+4q+j of
// the slice sits in bits [2j, 2j+1] of byte q.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_rows(const uint8_t * __restrict__ rows, uint64_t nrows, uint8_t * __restrict__ packed) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= div_up(nrows, 4)) return;
	uint32_t b = 0;
	#pragma unroll
	for (int j = 0; j < 4; ++j) if (4 * q + j < nrows) b |= (uint32_t)(rows[4 * q + j] & 3u) << (2 * j);
	packed[q] = (uint8_t)b;
}
__global__ void __launch_bounds__(256)
k_unpack_rows(const uint8_t * __restrict__ packed, uint64_t nrows, uint8_t * __restrict__ rows) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= div_up(nrows, 4)) return;
	uint32_t const b = packed[q];
	#pragma unroll
	for (int j = 0; j < 4; ++j) if (4 * q + j < nrows) rows[4 * q + j] = (uint8_t)((b >> (2 * j)) & 3u);
}

void Engine::pack_rows(const void * d_rows, uint64_t nrows, void * d_packed, bool unpack) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded && T.sigma <= 4, "2-bit row packing needs an alphabet of at most four codes");
	if (!nrows) return;
	unsigned const grid = (unsigned)div_up(div_up(nrows, 4), 256);
	if (unpack) B3M_LAUNCH(st, k_unpack_rows, grid, 256, 0, (const uint8_t *)d_packed, nrows, (uint8_t *)const_cast<void *>(d_rows));
	else B3M_LAUNCH(st, k_pack_rows, grid, 256, 0, (const uint8_t *)d_rows, nrows, (uint8_t *)d_packed);
}

// words [w_lo, w_hi) on any stream (used while the sort is still emitting later rows)
void k9_pack_bwa_range(cudaStream_t s, const uint8_t * bwt, uint64_t seq_len, uint64_t primary, uint32_t * words, uint64_t w_lo, uint64_t w_hi) {
	if (w_hi <= w_lo) return;
	k_pack_bwa<<<(unsigned)div_up(w_hi - w_lo, 256), 256, 0, s>>>(bwt, seq_len, primary, words, w_lo, w_hi);
	B3M_CUDA(cudaGetLastError());
}

void Engine::fetch_bwa(uint32_t * h_words, uint64_t cap, uint64_t * primary, uint64_t * L2, uint64_t * seq_len_out) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(have_results && !ssa_only, "no BWT to export");
	B3M_REQUIRE(T.has_term && T.sigma <= 4, "the BWA export needs a pacterm build (bases A,C,G,T and one terminator)");
	uint64_t const seq_len = T.n - 1;
	uint64_t const nwords = (seq_len + 15) >> 4;
	if (seq_len_out) *seq_len_out = seq_len;
	if (primary) *primary = root_exc_pos;
	if (L2) { L2[0] = 0; for (int c = 0; c < 4; ++c) L2[c + 1] = L2[c] + codehist[c]; }
	if (!h_words || h_words == bwa_on_host) return; // already delivered during the build (b3m_build_params.host_bwa)
	B3M_REQUIRE(cap >= nwords, "BWA word buffer too small");
	DevBuf<uint32_t> words(st, nwords);
	B3M_LAUNCH_T(st, "pack_bwa", T.n + 4 * nwords, k_pack_bwa, (unsigned)div_up(nwords, 256), 256, 0, (const uint8_t *)bwt.get(), seq_len, (uint64_t)root_exc_pos, words.get(), (uint64_t)0, nwords);
	B3M_CUDA(cudaMemcpyAsync(h_words, words.get(), 4 * nwords, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
}

void Engine::pack_bwa_device(uint32_t * d_words, uint64_t w_lo, uint64_t w_hi) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(have_results && !ssa_only, "no BWT to export");
	B3M_REQUIRE(T.has_term && T.sigma <= 4, "the BWA export needs a pacterm build (bases A,C,G,T and one terminator)");
	uint64_t const seq_len = T.n - 1;
	uint64_t const nwords = (seq_len + 15) >> 4;
	B3M_REQUIRE(d_words && w_lo <= w_hi && w_hi <= nwords, "bad word range");
	if (w_hi > w_lo) k9_pack_bwa_range(st.s, bwt.get(), seq_len, (uint64_t)root_exc_pos, d_words, w_lo, w_hi);
	++st.launches;
}

// runs of the BWT: start positions (device) and their number
uint64_t Engine::rl_runs(const uint8_t * s, DevBuf<uint32_t> & start) {
	uint64_t const n = T.n;
	start.alloc(st, n + 1); // worst case: every symbol its own run
	uint32_t * sp = start.get();
	uint32_t * total = d_special.get() + 4;
	scan_apply<OpSum>(st, n,
		[=] __device__(uint64_t i) -> uint32_t { return (i == 0 || s[i] != s[i - 1]) ? 1u : 0u; },
		[=] __device__(uint64_t i, uint32_t excl, uint32_t v) { if (v) sp[excl] = (uint32_t)i; if (i + 1 == n) *total = excl + v; });
	return fetch_special(4);
}

void Engine::write_bwt(const char * fn) {
	B3M_CUDA(cudaSetDevice(device));
	DevBuf<uint8_t> syms;
	symbols_device(syms);
	DevBuf<uint32_t> start;
	uint64_t const n = T.n;
	uint64_t const nruns = rl_runs(syms.get(), start);
	RlHeader h;
	h.n = n; h.nruns = nruns; h.runs_per_block = RL_RUNS_PER_BLOCK; h.nblocks = div_up(nruns, RL_RUNS_PER_BLOCK);
	// code tables from the two histograms
	DevBuf<unsigned long long> dh(st, 512);
	B3M_CUDA(cudaMemsetAsync(dh.get(), 0, 512 * 8, st.s));
	unsigned const hgrid = (unsigned)std::min<uint64_t>(std::max<uint64_t>(div_up(nruns, 256 * 16), 1), (uint64_t)st.sms * 8);
	B3M_LAUNCH_T(st, "rl_hist", 5 * nruns, k_rl_hist, hgrid, 256, 0, (const uint8_t *)syms.get(), (const uint32_t *)start.get(), nruns, n, dh.get(), dh.get() + 256);
	std::vector<uint64_t> hh(512);
	B3M_CUDA(cudaMemcpyAsync(hh.data(), dh.get(), 512 * 8, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	h.sym = huff_build(hh.data(), 256);
	h.len = huff_build(hh.data() + 256, RL_LENBINS);
	RlTabs tabs;
	for (int i = 0; i < 256; ++i) { tabs.sym[i] = ((uint32_t)h.sym.len[i] << 24) | h.sym.code[i]; tabs.len[i] = ((uint32_t)h.len.len[i] << 24) | h.len.code[i]; }
	DevBuf<RlTabs> dtabs(st, 1);
	B3M_CUDA(cudaMemcpyAsync(dtabs.get(), &tabs, sizeof(tabs), cudaMemcpyHostToDevice, st.s));
	// block sizes -> offsets (host scan: nblocks is nruns/4096)
	DevBuf<uint32_t> bw(st, h.nblocks);
	DevBuf<unsigned long long> bsym(st, h.nblocks), boff(st, h.nblocks);
	B3M_LAUNCH_T(st, "rl_blockbits", 5 * nruns, k_rl_blockbits, (unsigned)h.nblocks, RLE_THREADS, 0, (const uint8_t *)syms.get(), (const uint32_t *)start.get(), nruns, n,
	           (const RlTabs *)dtabs.get(), bw.get(), bsym.get());
	std::vector<uint32_t> hbw(h.nblocks);
	std::vector<uint64_t> woff(h.nblocks), soff(h.nblocks);
	B3M_CUDA(cudaMemcpyAsync(hbw.data(), bw.get(), 4 * h.nblocks, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaMemcpyAsync(soff.data(), bsym.get(), 8 * h.nblocks, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	uint64_t nwords = 0;
	for (uint64_t b = 0; b < h.nblocks; ++b) { woff[b] = nwords; nwords += hbw[b]; }
	B3M_CUDA(cudaMemcpyAsync(boff.get(), woff.data(), 8 * h.nblocks, cudaMemcpyHostToDevice, st.s));
	DevBuf<uint32_t> payload(st, 2 * nwords + 2);
	B3M_CUDA(cudaMemsetAsync(payload.get(), 0, 8 * nwords + 8, st.s));
	B3M_LAUNCH_T(st, "rl_emit", n + 8 * nwords, k_rl_emit, (unsigned)h.nblocks, RLE_THREADS, 0, (const uint8_t *)syms.get(), (const uint32_t *)start.get(), nruns, n,
	             (const RlTabs *)dtabs.get(), (const unsigned long long *)boff.get(), payload.get());
	B3M_CUDA(cudaStreamSynchronize(st.s));
	rl_bytes = 8 * nwords; rl_nruns = nruns;
	// stream the payload to the file through two pinned staging buffers
	RlContainerWriter wr(fn, h);
	size_t const chunk = (size_t)32 << 20;
	uint8_t * stage[2] = {nullptr, nullptr};
	cudaEvent_t done[2];
	try {
		for (int i = 0; i < 2; ++i) { B3M_CUDA(cudaMallocHost((void **)&stage[i], chunk)); B3M_CUDA(cudaEventCreate(&done[i])); }
		uint64_t const total = 8 * nwords;
		uint64_t issued = 0, written = 0;
		int slot = 0;
		size_t len[2] = {0, 0};
		// prime
		for (int i = 0; i < 2 && issued < total; ++i) {
			len[i] = (size_t)std::min<uint64_t>(chunk, total - issued);
			B3M_CUDA(cudaMemcpyAsync(stage[i], (const uint8_t *)payload.get() + issued, len[i], cudaMemcpyDeviceToHost, st.s));
			B3M_CUDA(cudaEventRecord(done[i], st.s));
			issued += len[i];
		}
		while (written < total) {
			B3M_CUDA(cudaEventSynchronize(done[slot]));
			wr.payload(stage[slot], len[slot]);
			written += len[slot];
			if (issued < total) {
				len[slot] = (size_t)std::min<uint64_t>(chunk, total - issued);
				B3M_CUDA(cudaMemcpyAsync(stage[slot], (const uint8_t *)payload.get() + issued, len[slot], cudaMemcpyDeviceToHost, st.s));
				B3M_CUDA(cudaEventRecord(done[slot], st.s));
				issued += len[slot];
			}
			slot ^= 1;
		}
	} catch (...) {
		for (int i = 0; i < 2; ++i) if (stage[i]) { cudaFreeHost(stage[i]); cudaEventDestroy(done[i]); }
		throw;
	}
	for (int i = 0; i < 2; ++i) { cudaFreeHost(stage[i]); cudaEventDestroy(done[i]); }
	wr.finish(woff.data(), soff.data());
}

// run stream for a reference-side binding that feeds libmaus2's own RLEncoder
void Engine::fetch_runs(uint8_t * h_sym, uint64_t * h_len, uint64_t cap, uint64_t * nruns_out) {
	B3M_CUDA(cudaSetDevice(device));
	DevBuf<uint8_t> syms;
	symbols_device(syms);
	DevBuf<uint32_t> start;
	uint64_t const n = T.n;
	uint64_t const nruns = rl_runs(syms.get(), start);
	if (nruns_out) *nruns_out = nruns;
	if (!h_sym && !h_len) return;
	B3M_REQUIRE(cap >= nruns, "run buffers too small");
	DevBuf<uint8_t> rs(st, nruns);
	DevBuf<unsigned long long> rl(st, nruns);
	B3M_LAUNCH(st, k_rl_gather, (unsigned)div_up(nruns, 256), 256, 0, (const uint8_t *)syms.get(), (const uint32_t *)start.get(), nruns, n, rs.get(), rl.get());
	if (h_sym) B3M_CUDA(cudaMemcpyAsync(h_sym, rs.get(), nruns, cudaMemcpyDeviceToHost, st.s));
	if (h_len) B3M_CUDA(cudaMemcpyAsync(h_len, rl.get(), 8 * nruns, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
}

} // namespace b3m
