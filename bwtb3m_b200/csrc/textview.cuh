// Text access shared by the sort kernels (K2) and the fused extraction (K3).
//
// A window is the range of W suffix start positions wstart+i, 0 <= i < W, that one call sorts:
//   circular   : the window is the whole terminator-free text, indices wrap modulo W
//   linear     : the end of the window is a sentinel smaller than every symbol (pacterm's
//                terminator, or the end of a leaf's look-ahead window);
//                text positions wrap modulo ntext when text_wraps != 0
// Symbols past a linear window read as code 0; the order of suffixes that run into the sentinel
// is fixed by their remaining length (shorter = smaller), see k_resolve in sufsort.cu.
//
// For alphabets of at most four codes the text is also held packed, 2 bit per symbol, 32 symbols
// per uint64 word, first symbol in the most significant bits, so that an unsigned comparison of
// words is the lexicographic comparison of the symbols.  The packed array is padded with two
// zero words.
#pragma once
#include "common.cuh"

namespace b3m {

struct TextView {
	const uint8_t * codes;
	const uint64_t * packed; // nullptr unless keybits == 2
	uint64_t ntext;
	uint64_t wstart;
	uint64_t W;
	int circular;
	int text_wraps;
	int has_term;            // the text ends in an implicit unique terminator (pacterm)
};

// code at text position p: out of the packed text when there is one (the byte codes of a pac input exist only on demand)
__device__ __forceinline__ uint32_t tv_code(TextView const & v, uint64_t p) {
	if (v.packed) return (uint32_t)(__ldg(v.packed + (p >> 5)) >> (62u - 2u * (unsigned)(p & 31u))) & 3u;
	return v.codes[p];
}

// symbol at window index i (any i >= 0)
__device__ __forceinline__ uint32_t tv_symbol(TextView const & v, uint64_t i) {
	if (v.circular) {
		if (i >= v.W) { i -= v.W; if (i >= v.W) i %= v.W; }
		return tv_code(v, i);
	}
	if (i >= v.W) return 0u;
	uint64_t p = v.wstart + i;
	if (v.text_wraps && p >= v.ntext) p %= v.ntext;
	return tv_code(v, p);
}

// 32 symbols starting at text position p, no wrap: needs p + 32 <= 32 * (words of the packed array)
__device__ __forceinline__ uint64_t pk_window(const uint64_t * __restrict__ P, uint64_t p) {
	uint64_t const w = p >> 5;
	unsigned const s = (unsigned)(p & 31u) << 1;
	uint64_t const a = __ldg(P + w), b = __ldg(P + w + 1);
	return s ? ((a << s) | (b >> (64u - s))) : a;
}

// text position of window index i when [i, i+len) touches neither the window end nor the text
// end (so that packed words can be read directly); ~0 otherwise
__device__ __forceinline__ uint64_t tv_interior(TextView const & v, uint64_t i, uint32_t len) {
	if (i + len > v.W) return ~0ull;
	uint64_t p = v.wstart + i;
	if (v.text_wraps && p >= v.ntext) p -= v.ntext;
	return (p + len <= v.ntext) ? p : ~0ull;
}

// the `count` (<= 64/bits) symbols from window index i on, `bits` bits each, first symbol in the
// most significant position of the result's low count*bits bits
__device__ __forceinline__ uint64_t tv_symbols(TextView const & v, uint64_t i, uint32_t count, uint32_t bits) {
	if (bits == 2 && v.packed) {
		uint64_t const p = tv_interior(v, i, 32);
		if (p != ~0ull) return pk_window(v.packed, p) >> (64u - 2u * count);
	}
	if (bits == 8 && count <= 8) {
		// byte codes: one unaligned 8-byte window out of two aligned words, first symbol most significant
		uint64_t const p = tv_interior(v, i, 16);
		if (p != ~0ull) {
			const uint64_t * wp = reinterpret_cast<const uint64_t *>(v.codes) + (p >> 3); // the code array is 16-byte aligned and padded
			unsigned const sh = (unsigned)(p & 7u) << 3;
			uint64_t const a = __ldg(wp), b = __ldg(wp + 1);
			uint64_t const x = sh ? ((a >> sh) | (b << (64u - sh))) : a;
			uint64_t const be = ((uint64_t)__byte_perm((uint32_t)x, 0u, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0u, 0x0123);
			return be >> (64u - 8u * count);
		}
	}
	uint64_t k = 0;
	for (uint32_t s = 0; s < count; ++s) k = (k << bits) | tv_symbol(v, i + s);
	return k;
}

// Code preceding the suffix at window index i, as it appears in the BWT: 0 when the suffix starts
// at text position 0 of a terminated text (the row of the terminator, fixed up by the caller).
__device__ __forceinline__ uint32_t tv_pred(TextView const & v, uint64_t i) {
	uint64_t p = v.wstart + i;
	if (p >= v.ntext) p %= v.ntext;
	if (p == 0) return v.has_term ? 0u : tv_code(v, v.ntext - 1);
	return tv_code(v, p - 1);
}

// code preceding text position p (p >= 1)
__device__ __forceinline__ uint32_t tv_code_before(TextView const & v, uint64_t p) { return tv_code(v, p - 1); }

} // namespace b3m
