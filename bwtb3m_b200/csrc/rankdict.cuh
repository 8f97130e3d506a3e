// Occurrence/rank dictionaries queried by popc (K4).  They replace the Huffman-shaped wavelet
// tree rank of the reference (ImpCompactHuffmanWaveletLF, /root/reference/src/checkbwt.cpp:165-166,
// LF(r) = C[L[r]] + rank_{L[r]}(L,r), /root/reference/src/lcpbit.cpp:3362-3365).
//
// Flavour 2 (sigma <= 4): one 64-byte line per 192 symbols: uint32 cum[4] | 6 x uint64 holding
//   32 symbols each, symbol j of a word at bits [2j, 2j+1].  One LF step touches one line.
// Flavour 8 (sigma <= 256): one block per 128 symbols: uint32 cum[spad] | 128 symbol bytes,
//   spad = sigma rounded up to 8 (so counters start on 32-byte sectors).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace b3m {

constexpr uint32_t D2_SYMS = 192;
constexpr uint32_t D8_SYMS = 128;

struct DictView {
	const uint8_t * base;
	uint32_t flavour;     // 2 or 8
	uint32_t stride;      // bytes per line/block
	uint32_t spad;        // flavour 8: counters per block
	uint32_t exc_pos;     // position whose stored symbol is not a real occurrence (0xffffffff: none)
	uint32_t exc_code;    // code stored at exc_pos
};

__host__ __device__ inline uint32_t d8_spad(uint32_t sigma) { return (sigma + 7u) & ~7u; }

__device__ __forceinline__ uint32_t popc_code2(uint64_t w, uint32_t c, uint64_t posmask) {
	uint64_t const y = w ^ (0x5555555555555555ull * (uint64_t)c);
	uint64_t const m = ~(y | (y >> 1)) & 0x5555555555555555ull;
	return (uint32_t)__popcll(m & posmask);
}

// returns rank_c(L, r) = number of occurrences of code c in L[0..r), exception position excluded
__device__ __forceinline__ uint32_t dict_rank(DictView const & D, uint32_t c, uint32_t r) {
	uint32_t cnt;
	if (D.flavour == 2) {
		uint32_t const line = r / D2_SYMS, off = r - line * D2_SYMS;
		const uint4 * lp = reinterpret_cast<const uint4 *>(D.base + (uint64_t)line * 64u);
		uint4 const q0 = __ldg(lp), q1 = __ldg(lp + 1), q2 = __ldg(lp + 2), q3 = __ldg(lp + 3);
		cnt = c == 0 ? q0.x : (c == 1 ? q0.y : (c == 2 ? q0.z : q0.w));
		uint64_t const w[6] = {
			((uint64_t)q1.y << 32) | q1.x, ((uint64_t)q1.w << 32) | q1.z,
			((uint64_t)q2.y << 32) | q2.x, ((uint64_t)q2.w << 32) | q2.z,
			((uint64_t)q3.y << 32) | q3.x, ((uint64_t)q3.w << 32) | q3.z };
		uint32_t const fw = off >> 5, fb = off & 31;
		#pragma unroll
		for (uint32_t k = 0; k < 6; ++k) {
			uint64_t const pm = k < fw ? ~0ull : (k == fw ? ((1ull << (2 * fb)) - 1ull) : 0ull);
			cnt += popc_code2(w[k], c, pm);
		}
	} else {
		uint32_t const blk = r / D8_SYMS, off = r - blk * D8_SYMS;
		const uint8_t * bp = D.base + (uint64_t)blk * D.stride;
		cnt = __ldg(reinterpret_cast<const uint32_t *>(bp) + c);
		const uint4 * sp = reinterpret_cast<const uint4 *>(bp + 4u * D.spad);
		uint32_t const cc = c * 0x01010101u;
		uint32_t const nq = (off + 15u) >> 4;
		for (uint32_t q = 0; q < nq; ++q) {
			uint4 const v = __ldg(sp + q);
			uint32_t const wv[4] = {v.x, v.y, v.z, v.w};
			#pragma unroll
			for (uint32_t k = 0; k < 4; ++k) {
				uint32_t const first = q * 16u + k * 4u;
				if (first < off) {
					uint32_t const nb = off - first; // valid bytes in this word (>= 1)
					uint32_t const bm = nb >= 4u ? 0xffffffffu : ((1u << (8u * nb)) - 1u);
					cnt += (uint32_t)__popc(__vcmpeq4(wv[k], cc) & bm) >> 3;
				}
			}
		}
	}
	if (D.exc_pos < r && D.exc_code == c) --cnt;
	return cnt;
}

// symbol code stored at position r (the caller handles r == exc_pos)
__device__ __forceinline__ uint32_t dict_symbol(DictView const & D, uint32_t r) {
	if (D.flavour == 2) {
		uint32_t const line = r / D2_SYMS, off = r - line * D2_SYMS;
		const uint32_t * wp = reinterpret_cast<const uint32_t *>(D.base + (uint64_t)line * 64u + 16u);
		return (__ldg(wp + (off >> 4)) >> (2u * (off & 15u))) & 3u;
	}
	uint32_t const blk = r / D8_SYMS, off = r - blk * D8_SYMS;
	return __ldg(D.base + (uint64_t)blk * D.stride + 4u * D.spad + off);
}

// one fused LF step for flavour 2: symbol at r and its rank from a single line fetch
__device__ __forceinline__ uint32_t dict_lf2(DictView const & D, const uint32_t * __restrict__ C, uint32_t r, uint32_t * sym_out) {
	uint32_t const line = r / D2_SYMS, off = r - line * D2_SYMS;
	const uint4 * lp = reinterpret_cast<const uint4 *>(D.base + (uint64_t)line * 64u);
	uint4 const q0 = __ldg(lp), q1 = __ldg(lp + 1), q2 = __ldg(lp + 2), q3 = __ldg(lp + 3);
	uint64_t const w[6] = {
		((uint64_t)q1.y << 32) | q1.x, ((uint64_t)q1.w << 32) | q1.z,
		((uint64_t)q2.y << 32) | q2.x, ((uint64_t)q2.w << 32) | q2.z,
		((uint64_t)q3.y << 32) | q3.x, ((uint64_t)q3.w << 32) | q3.z };
	uint32_t const fw = off >> 5, fb = off & 31;
	uint64_t wsel = w[0];
	#pragma unroll
	for (uint32_t k = 1; k < 6; ++k) wsel = (k == fw) ? w[k] : wsel;
	uint32_t const c = (uint32_t)(wsel >> (2 * fb)) & 3u;
	uint32_t cnt = c == 0 ? q0.x : (c == 1 ? q0.y : (c == 2 ? q0.z : q0.w));
	#pragma unroll
	for (uint32_t k = 0; k < 6; ++k) {
		uint64_t const pm = k < fw ? ~0ull : (k == fw ? ((1ull << (2 * fb)) - 1ull) : 0ull);
		cnt += popc_code2(w[k], c, pm);
	}
	if (D.exc_pos < r && D.exc_code == c) --cnt;
	*sym_out = c;
	return C[c] + cnt;
}

} // namespace b3m
