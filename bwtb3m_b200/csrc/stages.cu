// K1 (input decoding + histogram), K3 (BWT extraction), K4 (rank dictionary build),
// K7 (sampled SA/ISA LF walk).  See kernels.h for the reference code each one replaces.
#include "kernels.h"
#include "rankdict.cuh"
#include "scan.cuh"

namespace b3m {

// ------------------------------------------------------------------------------------------
// K1: replaces the libmaus2 input wrappers {Byte,Pac,PacTerm,Compact}InputTypes
// (/root/reference/src/checkbwt.cpp:254-270) and the symbol histogram pass.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hist_bytes(const uint8_t * __restrict__ in, uint64_t n, unsigned long long * __restrict__ hist) {
	__shared__ uint32_t sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	uint64_t const nvec = n / 16;
	const uint4 * v = reinterpret_cast<const uint4 *>(in);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 const q = ld_stream_u4(v + i);
		uint32_t const w[4] = {q.x, q.y, q.z, q.w};
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			atomicAdd(&sh[w[k] & 255u], 1u);
			atomicAdd(&sh[(w[k] >> 8) & 255u], 1u);
			atomicAdd(&sh[(w[k] >> 16) & 255u], 1u);
			atomicAdd(&sh[w[k] >> 24], 1u);
		}
	}
	if (blockIdx.x == 0)
		for (uint64_t i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sh[in[i]], 1u);
	__syncthreads();
	if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

void k1_hist_bytes(Stream & st, const uint8_t * d_in, uint64_t nbytes, uint64_t * d_hist256) {
	B3M_CUDA(cudaMemsetAsync(d_hist256, 0, 256 * sizeof(uint64_t), st.s));
	if (!nbytes) return;
	// one CTA handles at most 2^32/16 vectors between flushes of its 32-bit shared counters
	uint64_t want = div_up(nbytes / 16 + 1, 256 * 64);
	unsigned grid = (unsigned)(want < (uint64_t)st.sms * 8 ? (want ? want : 1) : (uint64_t)st.sms * 8);
	B3M_LAUNCH(st, k_hist_bytes, grid, 256, 0, d_in, nbytes, (unsigned long long *)d_hist256);
}

__global__ void __launch_bounds__(256) k_map_bytes(const uint8_t * __restrict__ in, uint64_t n, const uint8_t * __restrict__ lut, uint8_t * __restrict__ out) {
	__shared__ uint8_t sl[256];
	sl[threadIdx.x] = lut[threadIdx.x];
	__syncthreads();
	uint64_t const nvec = n / 16;
	const uint4 * v = reinterpret_cast<const uint4 *>(in);
	uint4 * o = reinterpret_cast<uint4 *>(out);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 const q = ld_stream_u4(v + i);
		uint32_t w[4] = {q.x, q.y, q.z, q.w};
		#pragma unroll
		for (int k = 0; k < 4; ++k)
			w[k] = (uint32_t)sl[w[k] & 255u] | ((uint32_t)sl[(w[k] >> 8) & 255u] << 8) |
			       ((uint32_t)sl[(w[k] >> 16) & 255u] << 16) | ((uint32_t)sl[w[k] >> 24] << 24);
		o[i] = make_uint4(w[0], w[1], w[2], w[3]);
	}
	if (blockIdx.x == 0)
		for (uint64_t i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) out[i] = sl[in[i]];
}

void k1_map_bytes(Stream & st, const uint8_t * d_in, uint64_t n, const uint8_t * d_lut256, uint8_t * d_out) {
	if (!n) return;
	uint64_t want = div_up(n / 16 + 1, 256 * 8);
	unsigned grid = (unsigned)(want < (uint64_t)st.sms * 16 ? (want ? want : 1) : (uint64_t)st.sms * 16);
	B3M_LAUNCH(st, k_map_bytes, grid, 256, 0, d_in, n, d_lut256, d_out);
}

// pac: 2 bit/symbol, MSB first inside each byte (BWA fa2pac layout).
// pac -> the packed text directly: 8 pac bytes, read big-endian, ARE one word of the packed text (32 symbols, first
// symbol in the top bits); the histogram of the codes comes from popcounts of the word.  The byte codes are not
// written: the MSD sorter, the key-range sorter and the resolve kernels read the packed text only, and the paths that
// want one byte per symbol (block merge tree, checkbwt) expand it on demand (k1_unpack_packed).
__global__ void __launch_bounds__(256) k_pac_to_packed(const uint8_t * __restrict__ pac, uint64_t l, uint64_t * __restrict__ out, uint64_t nwords,
                                                       int aligned, unsigned long long * __restrict__ hist, uint8_t * __restrict__ lastcode) {
	__shared__ uint32_t sh[4];
	if (threadIdx.x < 4) sh[threadIdx.x] = 0;
	__syncthreads();
	uint64_t const nbytes = (l + 3) >> 2;
	uint32_t c1 = 0, c2 = 0, c3 = 0, cv = 0;
	for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t const first = w * 32;
		uint32_t const valid = first >= l ? 0u : (l - first < 32 ? (uint32_t)(l - first) : 32u);
		uint64_t x = 0;
		if (valid) {
			if (aligned && 8 * w + 8 <= nbytes) {
				uint64_t const v = __ldg(reinterpret_cast<const uint64_t *>(pac) + w);
				x = ((uint64_t)__byte_perm((uint32_t)v, 0u, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0u, 0x0123);
			} else {
				for (uint32_t b = 0; b < 8; ++b) x = (x << 8) | (8 * w + b < nbytes ? (uint64_t)pac[8 * w + b] : 0ull);
			}
			if (valid < 32) x &= ~0ull << (64u - 2u * valid);
			uint64_t const lo = x & 0x5555555555555555ull, hi = (x >> 1) & 0x5555555555555555ull;
			c3 += (uint32_t)__popcll(lo & hi); c1 += (uint32_t)__popcll(lo & ~hi); c2 += (uint32_t)__popcll(hi & ~lo);
			cv += valid;
			if (l - 1 - first < 32) *lastcode = (uint8_t)((x >> (62u - 2u * (uint32_t)(l - 1 - first))) & 3u);
		}
		out[w] = x;
	}
	uint32_t const c0 = cv - c1 - c2 - c3;
	if (c0) atomicAdd(&sh[0], c0);
	if (c1) atomicAdd(&sh[1], c1);
	if (c2) atomicAdd(&sh[2], c2);
	if (c3) atomicAdd(&sh[3], c3);
	__syncthreads();
	if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

void k1_pac_to_packed(Stream & st, const uint8_t * d_pac, uint64_t l, uint64_t * d_out, uint64_t * d_hist256, uint8_t * d_lastcode) {
	B3M_CUDA(cudaMemsetAsync(d_hist256, 0, 256 * sizeof(uint64_t), st.s));
	if (!l) return;
	uint64_t const nwords = l / 32 + 3;
	uint64_t const want = div_up(nwords, 256 * 4);
	unsigned const grid = (unsigned)(want < (uint64_t)st.sms * 16 ? (want ? want : 1) : (uint64_t)st.sms * 16);
	B3M_LAUNCH(st, k_pac_to_packed, grid, 256, 0, d_pac, l, d_out, nwords, ((uintptr_t)d_pac & 7u) == 0 ? 1 : 0, (unsigned long long *)d_hist256, d_lastcode);
}

// one byte per symbol out of the packed text: a thread expands one word into 32 codes (two 128-bit stores); the 16
// bytes behind the last code are cleared (readers of 16-byte blocks may touch them)
__global__ void __launch_bounds__(256) k_unpack_packed(const uint64_t * __restrict__ packed, uint64_t n, uint8_t * __restrict__ out) {
	uint64_t const w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t const first = w * 32;
	if (first >= n + 16) return;
	uint64_t const x = first < n ? __ldg(packed + w) : 0ull;
	if (first + 32 <= n) {
		uint32_t o[8];
		#pragma unroll
		for (int q = 0; q < 8; ++q) {
			uint32_t const y = (uint32_t)(x >> (56 - 8 * q)) & 255u; // four symbols, first in the top bits
			o[q] = ((y >> 6) & 3u) | (((y >> 4) & 3u) << 8) | (((y >> 2) & 3u) << 16) | ((y & 3u) << 24);
		}
		uint4 * dst = reinterpret_cast<uint4 *>(out + first);
		dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
		dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
	} else {
		for (uint32_t j = 0; j < 32 && first + j < n + 16; ++j) out[first + j] = first + j < n ? (uint8_t)((x >> (62u - 2u * j)) & 3u) : (uint8_t)0;
	}
}

void k1_unpack_packed(Stream & st, const uint64_t * d_packed, uint64_t n, uint8_t * d_out) {
	uint64_t const nwords = (n + 16 + 31) / 32;
	B3M_LAUNCH(st, k_unpack_packed, (unsigned)div_up(nwords, 256), 256, 0, d_packed, n, d_out);
}

// compactstream payload: b-bit symbols, MSB first, in 64-bit words [layout unpinned, SURVEY 8c].  The words lie
// either as a big-endian byte stream (flip = 0) or as native little-endian uint64 (flip = 7: byte k of the bit
// stream is byte k ^ 7 of the file, what libmaus2's native Serialize<uint64_t> writes); one thread per symbol (b <= 8).
__global__ void __launch_bounds__(256) k_unpack_compact(const uint8_t * __restrict__ d, uint64_t n, unsigned b, unsigned flip, uint8_t * __restrict__ out,
                                                        unsigned long long * __restrict__ hist) {
	__shared__ uint32_t sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t const bit = i * b;
		uint64_t const byte = bit >> 3;
		uint32_t const two = ((uint32_t)d[byte ^ flip] << 8) | (uint32_t)d[(byte + 1) ^ flip]; // eight pad bytes are guaranteed by the loader
		uint32_t const v = (two >> (16 - (bit & 7) - b)) & ((1u << b) - 1u);
		out[i] = (uint8_t)v;
		atomicAdd(&sh[v], 1u);
	}
	__syncthreads();
	if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

void k1_unpack_compact(Stream & st, const uint8_t * d_words, uint64_t n, unsigned b, bool le_words, uint8_t * d_out, uint64_t * d_hist256) {
	B3M_CUDA(cudaMemsetAsync(d_hist256, 0, 256 * sizeof(uint64_t), st.s));
	if (!n) return;
	uint64_t want = div_up(n, 256 * 16);
	unsigned grid = (unsigned)(want < (uint64_t)st.sms * 16 ? (want ? want : 1) : (uint64_t)st.sms * 16);
	B3M_LAUNCH(st, k_unpack_compact, grid, 256, 0, d_words, n, b, le_words ? 7u : 0u, d_out, (unsigned long long *)d_hist256);
}

// 2-bit packed copy of the codes (textview.cuh): 32 symbols per uint64, first symbol in the top bits
__global__ void __launch_bounds__(256) k_pack2(const uint8_t * __restrict__ codes, uint64_t n, uint64_t * __restrict__ out, uint64_t nwords) {
	uint64_t const w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords) return;
	uint64_t const first = w * 32;
	uint64_t acc = 0;
	if (first + 32 <= n) {
		const uint4 * src = reinterpret_cast<const uint4 *>(codes + first);
		#pragma unroll
		for (int q = 0; q < 2; ++q) {
			uint4 const x4 = ld_stream_u4(src + q);
			uint32_t const x[4] = {x4.x, x4.y, x4.z, x4.w};
			#pragma unroll
			for (int k = 0; k < 4; ++k) {
				uint32_t const y = x[k] & 0x03030303u;
				uint32_t const z = ((y & 3u) << 6) | ((y >> 4) & 0x30u) | ((y >> 14) & 0x0cu) | ((y >> 24) & 3u);
				acc = (acc << 8) | z;
			}
		}
	} else {
		for (uint32_t j = 0; j < 32; ++j) acc = (acc << 2) | ((first + j < n) ? (uint64_t)(codes[first + j] & 3u) : 0ull);
	}
	out[w] = acc;
}

void k1_pack2(Stream & st, const uint8_t * d_codes, uint64_t n, uint64_t * d_out) {
	uint64_t const nwords = n / 32 + 3;
	B3M_LAUNCH(st, k_pack2, (unsigned)div_up(nwords, 256), 256, 0, d_codes, n, d_out, nwords);
}

// K3 (BWT extraction) lives with the sort: k_resolve / k_extract_sample in sufsort.cu, k_leaf_emit in blocks.cu.

// ------------------------------------------------------------------------------------------
// K4: dictionary build
// ------------------------------------------------------------------------------------------
size_t dict_bytes(int flavour, uint64_t n, uint32_t sigma) {
	if (flavour == 2) return (size_t)(n / D2_SYMS + 1) * 64;
	return (size_t)(n / D8_SYMS + 1) * (4u * d8_spad(sigma) + D8_SYMS);
}

// one thread packs one 64-byte line and leaves the line's own symbol counts in the counter slot
__global__ void __launch_bounds__(128) k_dict2_pack(const uint8_t * __restrict__ bwt, uint64_t n, uint4 * __restrict__ lines, uint64_t nlines) {
	uint64_t const line = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (line >= nlines) return;
	uint64_t const first = line * D2_SYMS;
	uint64_t w[6] = {0, 0, 0, 0, 0, 0};
	uint32_t cnt[4] = {0, 0, 0, 0};
	uint32_t const have = first >= n ? 0u : ((n - first) < D2_SYMS ? (uint32_t)(n - first) : D2_SYMS);
	if (have == D2_SYMS) {
		const uint4 * src = reinterpret_cast<const uint4 *>(bwt + first); // 192*line is 16-byte aligned
		#pragma unroll
		for (int q = 0; q < 12; ++q) {
			uint4 const v = __ldg(src + q);
			uint32_t const x[4] = {v.x, v.y, v.z, v.w};
			uint64_t acc = 0;
			#pragma unroll
			for (int k = 0; k < 4; ++k) {
				uint32_t const y = x[k] & 0x03030303u;
				// gather the four 2-bit codes of this word into 8 contiguous bits
				uint32_t const z = (y | (y >> 6) | (y >> 12) | (y >> 18)) & 0xffu;
				acc |= (uint64_t)z << (8 * k);
			}
			w[q >> 1] |= acc << (32 * (q & 1));
		}
	} else {
		for (uint32_t j = 0; j < have; ++j) w[j >> 5] |= (uint64_t)(bwt[first + j] & 3u) << (2 * (j & 31));
	}
	#pragma unroll
	for (int k = 0; k < 6; ++k) {
		uint32_t const valid = have > 32u * k ? ((have - 32u * k) < 32u ? have - 32u * k : 32u) : 0u;
		uint64_t const pm = valid >= 32u ? ~0ull : ((1ull << (2 * valid)) - 1ull);
		#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) cnt[c] += popc_code2(w[k], c, pm);
	}
	uint4 * lp = lines + line * 4;
	lp[0] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
	lp[1] = make_uint4((uint32_t)w[0], (uint32_t)(w[0] >> 32), (uint32_t)w[1], (uint32_t)(w[1] >> 32));
	lp[2] = make_uint4((uint32_t)w[2], (uint32_t)(w[2] >> 32), (uint32_t)w[3], (uint32_t)(w[3] >> 32));
	lp[3] = make_uint4((uint32_t)w[4], (uint32_t)(w[4] >> 32), (uint32_t)w[5], (uint32_t)(w[5] >> 32));
}

// flavour 8: one warp per 128-symbol block: local counts by warp match, symbols copied
__global__ void __launch_bounds__(256) k_dict8_pack(const uint8_t * __restrict__ bwt, uint64_t n, uint8_t * __restrict__ base,
                                                    uint32_t stride, uint32_t spad, uint64_t nblocks) {
	__shared__ uint32_t cnt[8][256];
	unsigned const w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t const blk = (uint64_t)blockIdx.x * 8 + w;
	if (blk >= nblocks) return; // whole warp leaves together
	for (uint32_t c = lane; c < 256; c += 32) cnt[w][c] = 0;
	__syncwarp();
	uint64_t const first = blk * D8_SYMS + 4ull * lane;
	uint32_t word = 0;
	uint32_t nb = 0;
	if (first < n) {
		nb = (n - first) < 4 ? (uint32_t)(n - first) : 4u;
		if (nb == 4) word = *reinterpret_cast<const uint32_t *>(bwt + first);
		else for (uint32_t k = 0; k < nb; ++k) word |= (uint32_t)bwt[first + k] << (8 * k);
	}
	#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		bool const valid = k < nb;
		uint32_t const c = (word >> (8 * k)) & 255u;
		unsigned const peers = __match_any_sync(0xffffffffu, valid ? c : 0xffffffffu);
		if (valid && (peers & lanemask_lt()) == 0) cnt[w][c] += __popc(peers);
		__syncwarp();
	}
	uint8_t * bp = base + blk * stride;
	for (uint32_t c = lane; c < spad; c += 32) reinterpret_cast<uint32_t *>(bp)[c] = cnt[w][c];
	reinterpret_cast<uint32_t *>(bp + 4u * spad)[lane] = word;
}

constexpr uint32_t D8_CHUNK = 256; // blocks per chunk in the column scans
__global__ void __launch_bounds__(256) k_dict8_chunksum(const uint8_t * __restrict__ base, uint32_t stride, uint32_t spad, uint64_t nblocks,
                                                        uint32_t * __restrict__ chunktot) {
	uint32_t const c = threadIdx.x;
	if (c >= spad) return;
	uint64_t const b0 = (uint64_t)blockIdx.x * D8_CHUNK;
	uint64_t const b1 = b0 + D8_CHUNK < nblocks ? b0 + D8_CHUNK : nblocks;
	uint32_t s = 0;
	for (uint64_t b = b0; b < b1; ++b) s += reinterpret_cast<const uint32_t *>(base + b * stride)[c];
	chunktot[(uint64_t)blockIdx.x * spad + c] = s;
}
__global__ void __launch_bounds__(256) k_dict8_chunkscan(uint32_t * __restrict__ chunktot, uint32_t spad, uint64_t nchunks) {
	uint32_t const c = threadIdx.x;
	if (c >= spad) return;
	uint32_t run = 0;
	for (uint64_t k = 0; k < nchunks; ++k) { uint32_t const t = chunktot[k * spad + c]; chunktot[k * spad + c] = run; run += t; }
}
__global__ void __launch_bounds__(256) k_dict8_apply(uint8_t * __restrict__ base, uint32_t stride, uint32_t spad, uint64_t nblocks,
                                                     const uint32_t * __restrict__ chunktot) {
	uint32_t const c = threadIdx.x;
	if (c >= spad) return;
	uint64_t const b0 = (uint64_t)blockIdx.x * D8_CHUNK;
	uint64_t const b1 = b0 + D8_CHUNK < nblocks ? b0 + D8_CHUNK : nblocks;
	uint32_t run = chunktot[(uint64_t)blockIdx.x * spad + c];
	for (uint64_t b = b0; b < b1; ++b) {
		uint32_t * p = reinterpret_cast<uint32_t *>(base + b * stride) + c;
		uint32_t const t = *p; *p = run; run += t;
	}
}

void k4_build_dict(Stream & st, const uint8_t * bwt, uint64_t n, int flavour, uint32_t sigma, void * lines) {
	if (flavour == 2) {
		uint64_t const nlines = n / D2_SYMS + 1;
		uint4 * L = reinterpret_cast<uint4 *>(lines);
		B3M_LAUNCH(st, k_dict2_pack, (unsigned)div_up(nlines, 128), 128, 0, bwt, n, L, nlines);
		scan_apply<OpSum4>(st, nlines,
			[=] __device__(uint64_t i) -> uint4 { return L[i * 4]; },
			[=] __device__(uint64_t i, uint4 excl, uint4) { L[i * 4] = excl; });
		return;
	}
	uint32_t const spad = d8_spad(sigma);
	uint32_t const stride = 4u * spad + D8_SYMS;
	uint64_t const nblocks = n / D8_SYMS + 1;
	uint8_t * base = reinterpret_cast<uint8_t *>(lines);
	B3M_LAUNCH(st, k_dict8_pack, (unsigned)div_up(nblocks, 8), 256, 0, bwt, n, base, stride, spad, nblocks);
	uint64_t const nchunks = div_up(nblocks, D8_CHUNK);
	DevBuf<uint32_t> chunktot(st, nchunks * spad);
	B3M_LAUNCH(st, k_dict8_chunksum, (unsigned)nchunks, 256, 0, (const uint8_t *)base, stride, spad, nblocks, chunktot.get());
	B3M_LAUNCH(st, k_dict8_chunkscan, 1, 256, 0, chunktot.get(), spad, nchunks);
	B3M_LAUNCH(st, k_dict8_apply, (unsigned)nchunks, 256, 0, base, stride, spad, nblocks, (const uint32_t *)chunktot.get());
}

// ------------------------------------------------------------------------------------------
// K7: sampled SA / ISA by LF walk from the anchors; one chain per thread
// (in-repo twin of the reference loop: /root/reference/src/hwtPreIsaToIsa.cpp:114-161;
//  `curpos=(curpos+n-1)%n; currank=LF(currank)`, `ISA[curpos/rate]=currank`).
// SA is sampled by rank, ISA by position.
// ------------------------------------------------------------------------------------------
struct CTable { uint32_t c[257]; };

__device__ __forceinline__ uint32_t lf_step(DictView const & D, CTable const & C, uint32_t exc_lf, uint32_t r) {
	if (r == D.exc_pos) return exc_lf;
	if (D.flavour == 2) { uint32_t s; return dict_lf2(D, C.c, r, &s); }
	uint32_t const s = dict_symbol(D, r);
	return C.c[s] + dict_rank(D, s, r);
}

__global__ void __launch_bounds__(256)
k_walk(DictView D, CTable C, uint32_t exc_lf, const uint32_t * __restrict__ anchor_rank, uint64_t nanchors, uint64_t arate, uint64_t n,
       uint32_t samask, uint32_t sashift, uint32_t isamask, uint32_t isashift,
       unsigned long long * __restrict__ sa_out, unsigned long long * __restrict__ isa_out, uint64_t q_lo, uint64_t q_hi) {
	uint64_t const q = q_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= q_hi) return;
	uint32_t r = anchor_rank[q];
	uint64_t p = q * arate;
	uint64_t steps = q ? arate : n - (nanchors - 1) * arate;
	while (steps--) {
		if ((p & isamask) == 0) isa_out[p >> isashift] = r;
		if ((r & samask) == 0) sa_out[r >> sashift] = p;
		p = p ? p - 1 : n - 1;
		r = lf_step(D, C, exc_lf, r);
	}
}

// anchors at arbitrary positions (bwtcomputessa: whatever the .preisa file holds)
__global__ void __launch_bounds__(256)
k_walk_anchors(DictView D, CTable C, uint32_t exc_lf, const uint32_t * __restrict__ anchor_rank, const unsigned long long * __restrict__ anchor_pos,
               const unsigned long long * __restrict__ anchor_steps, uint64_t nanchors, uint64_t n,
               uint32_t samask, uint32_t sashift, uint32_t isamask, uint32_t isashift,
               unsigned long long * __restrict__ sa_out, unsigned long long * __restrict__ isa_out) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nanchors) return;
	uint32_t r = anchor_rank[q];
	uint64_t p = anchor_pos[q];
	uint64_t steps = anchor_steps[q];
	while (steps--) {
		if ((p & isamask) == 0) isa_out[p >> isashift] = r;
		if ((r & samask) == 0) sa_out[r >> sashift] = p;
		p = p ? p - 1 : n - 1;
		r = lf_step(D, C, exc_lf, r);
	}
}

__global__ void __launch_bounds__(256)
k_lfbench(DictView D, CTable C, uint32_t exc_lf, const uint32_t * __restrict__ start, uint64_t nchains, uint64_t steps, uint32_t * __restrict__ out) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nchains) return;
	uint32_t r = start[q];
	for (uint64_t s = 0; s < steps; ++s) r = lf_step(D, C, exc_lf, r);
	out[q] = r;
}

// checkbwt (/root/reference/src/checkbwt.cpp:176-243): from every anchor walk LF back to the
// previous anchor and compare the symbol the BWT holds at the current rank with the text symbol
// in front of the current position ("text circularly reversed at the anchor").  Codes are the
// text's dense codes; exc_pos is the row of the terminator symbol of a terminated text.
__global__ void __launch_bounds__(256)
k_check_walk(DictView D, CTable C, uint32_t exc_lf, const uint8_t * __restrict__ codes, uint64_t ntext, int has_term,
             const uint32_t * __restrict__ anchor_rank, const unsigned long long * __restrict__ anchor_pos,
             const unsigned long long * __restrict__ anchor_steps, uint64_t nanchors, uint64_t n,
             unsigned long long * __restrict__ result /* [0] mismatches, [1] rank of one of them */) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nanchors) return;
	uint32_t r = anchor_rank[q];
	uint64_t p = anchor_pos[q];
	uint64_t steps = anchor_steps[q];
	uint32_t bad = 0, badrank = 0;
	while (steps--) {
		p = p ? p - 1 : n - 1;
		int const want = (has_term && p == ntext) ? -1 : (int)codes[p];
		int const have = (r == D.exc_pos) ? -1 : (int)dict_symbol(D, r);
		if (want != have) { if (!bad) badrank = r; ++bad; }
		r = lf_step(D, C, exc_lf, r);
	}
	if (bad) { atomicAdd(&result[0], (unsigned long long)bad); result[1] = badrank; }
}

static DictView make_view(DevDict const & D) {
	DictView v;
	v.base = reinterpret_cast<const uint8_t *>(D.lines);
	v.flavour = (uint32_t)D.flavour;
	v.spad = d8_spad(D.sigma);
	v.stride = D.flavour == 2 ? 64u : 4u * v.spad + D8_SYMS;
	v.exc_pos = D.exc_pos;
	v.exc_code = D.exc_code;
	return v;
}

static unsigned ilog2_exact(uint64_t v) {
	B3M_REQUIRE(v && !(v & (v - 1)), "sampling rates must be powers of two"); // hwtPreIsaToIsa.cpp:90-97
	unsigned s = 0;
	while ((1ull << s) < v) ++s;
	return s;
}

void k7_walk(Stream & st, DevDict const & D, const uint32_t * anchor_rank, uint64_t nanchors, uint64_t arate,
             uint64_t n, uint64_t sarate, uint64_t isarate, uint64_t * sa_out, uint64_t * isa_out, WalkStats * ws, uint64_t q_lo, uint64_t q_hi) {
	if (q_hi > nanchors) q_hi = nanchors;
	if (q_lo >= q_hi) return;
	CTable C;
	for (int i = 0; i < 257; ++i) C.c[i] = D.C[i];
	// anchor q walks arate steps back to anchor q-1; anchor 0 walks around the end of the text
	uint64_t steps = (q_hi - q_lo) * arate;
	if (q_lo == 0) steps = steps - arate + (n - (nanchors - 1) * arate);
	B3M_LAUNCH_T(st, "lf_walk", steps * 64ull, k_walk, (unsigned)div_up(q_hi - q_lo, 256), 256, 0, make_view(D), C, D.exc_lf, anchor_rank, nanchors, arate, n,
	           (uint32_t)(sarate - 1), ilog2_exact(sarate), (uint32_t)(isarate - 1), ilog2_exact(isarate),
	           (unsigned long long *)sa_out, (unsigned long long *)isa_out, q_lo, q_hi);
	if (ws) { ws->steps += steps; ws->chains += q_hi - q_lo; }
}

void k7_walk_anchors(Stream & st, DevDict const & D, const uint32_t * anchor_rank, const uint64_t * anchor_pos, const uint64_t * anchor_steps,
                     uint64_t nanchors, uint64_t n, uint64_t sarate, uint64_t isarate, uint64_t * sa_out, uint64_t * isa_out, WalkStats * ws) {
	if (!nanchors) return;
	CTable C;
	for (int i = 0; i < 257; ++i) C.c[i] = D.C[i];
	B3M_LAUNCH_T(st, "lf_walk", n * 64ull, k_walk_anchors, (unsigned)div_up(nanchors, 256), 256, 0, make_view(D), C, D.exc_lf, anchor_rank,
	           (const unsigned long long *)anchor_pos, (const unsigned long long *)anchor_steps, nanchors, n,
	           (uint32_t)(sarate - 1), ilog2_exact(sarate), (uint32_t)(isarate - 1), ilog2_exact(isarate),
	           (unsigned long long *)sa_out, (unsigned long long *)isa_out);
	if (ws) { ws->steps += n; ws->chains += nanchors; }
}

void k7_check_walk(Stream & st, DevDict const & D, const uint8_t * codes, uint64_t ntext, int has_term, const uint32_t * anchor_rank,
                   const uint64_t * anchor_pos, const uint64_t * anchor_steps, uint64_t nanchors, uint64_t n, uint64_t * d_result) {
	if (!nanchors) return;
	CTable C;
	for (int i = 0; i < 257; ++i) C.c[i] = D.C[i];
	B3M_LAUNCH_T(st, "check_walk", n * 65ull, k_check_walk, (unsigned)div_up(nanchors, 256), 256, 0, make_view(D), C, D.exc_lf, codes, ntext, has_term, anchor_rank,
	             (const unsigned long long *)anchor_pos, (const unsigned long long *)anchor_steps, nanchors, n, (unsigned long long *)d_result);
}

void k7_lfbench(Stream & st, DevDict const & D, const uint32_t * start_rank, uint64_t nchains, uint64_t steps, uint32_t * out_rank) {
	if (!nchains) return;
	CTable C;
	for (int i = 0; i < 257; ++i) C.c[i] = D.C[i];
	B3M_LAUNCH(st, k_lfbench, (unsigned)div_up(nchains, 256), 256, 0, make_view(D), C, D.exc_lf, start_rank, nchains, steps, out_rank);
}

} // namespace b3m
