/*
 * b3m_oracle.c -- CPU restatement of the gt1/bwtb3m hot path (BWT by balanced block merging).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and only as the checker / the timed CPU arm.  The product path (bwtb3m_b200/csrc)
 * never links or calls it and has no CPU fallback.
 *
 * PARITY STATUS.  The reference's arithmetic for this path lives in libmaus2 (>= 2.0.349,
 * /root/reference/configure.ac:39), which is absent from /root/reference and from this image, so
 * the reference cannot be compiled or run here.  This file restates the published algorithm and
 * anchors every function on the reference's own call sites, definitions and verifier:
 *   - BWT / ISA definition ............ /root/reference/src/lcpbit.cpp:3658-3669,3688-3712
 *   - LF-walk verifier (checkbwt) ..... /root/reference/src/checkbwt.cpp:126-243
 *   - ISA-from-preisa walk ............ /root/reference/src/hwtPreIsaToIsa.cpp:114-161
 *   - LF identity ..................... /root/reference/src/lcpbit.cpp:3362-3365
 *   - .sa/.isa/.preisa layouts ........ /root/reference/src/sasubsample.cpp:34-58,
 *                                       /root/reference/src/hwtPreIsaToIsa.cpp:55-77
 *   - LF-steps/s instrument ........... /root/reference/src/bwttestdecodespeed.cpp:67-97
 * It is pinned against the known-answer vectors of SURVEY.md section 4 (derived from the
 * reference's definition) in tests/test_oracle.py.  The byte layout of libmaus2's run-length
 * Huffman .bwt container and of the compactstream container is "parity unpinned" (no golden
 * bytes exist anywhere in the reference).
 *
 * Conventions: symbols are uint8 in the reference's symbol space (pacterm: 0 = terminator,
 * 1..4 = A,C,G,T; pac: 0..3; bytestream: raw bytes).  n < 2^32.  The BWT is circular:
 * BWT[r] = T[(SA[r]+n-1) % n], rotations ordered lexicographically (T primitive).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

typedef uint32_t idx_t;

static double now_sec(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------
 * A3: input decoders.  pac: 2 bit/symbol MSB first, length = (fsize-2)*4 + lastbyte (BWA
 * fa2pac layout).  pacterm: pac symbols +1, then one 0 terminator.
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_pac_numsyms(const uint8_t *file, uint64_t fsize)
{
	if (fsize < 2) return 0;
	return (fsize - 2) * 4 + file[fsize - 1];
}

/* out must hold l (+1 if term) symbols; returns n */
uint64_t orc_decode_pac(const uint8_t *file, uint64_t fsize, int term, uint8_t *out)
{
	uint64_t const l = orc_pac_numsyms(file, fsize);
	for (uint64_t i = 0; i < l; ++i) {
		uint8_t const s = (file[i >> 2] >> ((~i & 3) << 1)) & 3;
		out[i] = term ? (uint8_t)(s + 1) : s;
	}
	if (term) { out[l] = 0; return l + 1; }
	return l;
}

/* inverse, used by generators/tests: syms in 0..3 */
uint64_t orc_encode_pac(const uint8_t *syms, uint64_t l, uint8_t *file)
{
	uint64_t const nb = (l >> 2) + ((l & 3) ? 1 : 0);
	memset(file, 0, nb + 2);
	for (uint64_t i = 0; i < l; ++i)
		file[i >> 2] |= (uint8_t)(syms[i] << ((~i & 3) << 1));
	uint64_t p = nb;
	if ((l & 3) == 0) file[p++] = 0;
	file[p++] = (uint8_t)(l & 3);
	return p;
}

/* compactstream [parity unpinned, SURVEY 8c]: the serialised CompactArray the reference's writers leave
 * behind [REF digitsToCompact.cpp:35,122; arraytocompact.cpp:78,100]: four uint64 (b, n, word count, word
 * count again as the array header), then 64-bit words holding b-bit symbols MSB first.  The byte order of
 * the numbers and words is not pinned by anything in the reference: libmaus2's Serialize<uint64_t> (the
 * facility the .sa/.isa files use, [REF sasubsample.cpp:34-58]) is native little-endian; a big-endian
 * reading is accepted too.  b is 1..64 in exactly one of the two readings. */
static uint64_t be64(const uint8_t *p)
{
	uint64_t v = 0;
	for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
	return v;
}
static uint64_t le64(const uint8_t *p)
{
	uint64_t v = 0;
	for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
	return v;
}
/* returns 0 = big-endian layout, 1 = little-endian words, -1 = not a compact file */
int orc_compact_header(const uint8_t *file, uint64_t fsize, uint64_t *b, uint64_t *n)
{
	if (fsize < 32) return -1;
	uint64_t const bb = be64(file), bl = le64(file);
	if (bb >= 1 && bb <= 64) { *b = bb; *n = be64(file + 8); return 0; }
	if (bl >= 1 && bl <= 64) { *b = bl; *n = le64(file + 8); return 1; }
	return -1;
}
uint64_t orc_decode_compact(const uint8_t *file, uint64_t fsize, uint8_t *out)
{
	uint64_t b, n;
	int const le = orc_compact_header(file, fsize, &b, &n);
	if (le < 0 || b > 8) return 0;
	const uint8_t *d = file + 32;
	for (uint64_t i = 0; i < n; ++i) {
		uint64_t v = 0;
		for (uint64_t k = 0; k < b; ++k) {
			uint64_t const bit = i * b + k; /* bit 0 = MSB of word 0 */
			uint64_t const word = bit >> 6, inword = bit & 63;
			uint64_t const w = le ? le64(d + 8 * word) : be64(d + 8 * word);
			v = (v << 1) | ((w >> (63 - inword)) & 1);
		}
		out[i] = (uint8_t)v;
	}
	return n;
}

/* ------------------------------------------------------------------------------------------
 * Independent small-n definition: naive rotation sort (build plan 1a).
 * ---------------------------------------------------------------------------------------- */
typedef struct { const uint8_t *t; uint64_t n; } rotctx_t;

static int rotcmp(const void *pa, const void *pb, void *vctx)
{
	rotctx_t const *c = (rotctx_t const *)vctx;
	idx_t const a = *(idx_t const *)pa, b = *(idx_t const *)pb;
	uint64_t ia = a, ib = b;
	for (uint64_t k = 0; k < c->n; ++k) {
		uint8_t const x = c->t[ia], y = c->t[ib];
		if (x != y) return x < y ? -1 : 1;
		if (++ia == c->n) ia = 0;
		if (++ib == c->n) ib = 0;
	}
	return a < b ? -1 : (a > b ? 1 : 0); /* non-primitive text: tie by position (unpinned) */
}

void orc_naive_rotation_sort(const uint8_t *t, uint64_t n, idx_t *sa)
{
	rotctx_t c = { t, n };
	for (uint64_t i = 0; i < n; ++i) sa[i] = (idx_t)i;
	qsort_r(sa, n, sizeof(idx_t), rotcmp, &c);
}

/* BWT[i] = s[(SA[i]+n-1)%n]; ISA[SA[i]] = i   [REF lcpbit.cpp:3668-3669,3688-3690] */
void orc_bwt_from_sa(const uint8_t *t, uint64_t n, const idx_t *sa, uint8_t *bwt, idx_t *isa)
{
	for (uint64_t i = 0; i < n; ++i) {
		bwt[i] = t[(sa[i] + n - 1) % n];
		if (isa) isa[sa[i]] = (idx_t)i;
	}
}

/* ------------------------------------------------------------------------------------------
 * Suffix sorter used by the block sort: radix sort on a packed K-symbol key, then
 * Larsson-Sadakane style prefix doubling on the groups that are still tied.
 * mode 0 (linear): window w[0..W), end of window is a sentinel smaller than every symbol
 *                  (SURVEY Appendix A.1).
 * mode 1 (circular): w is the whole text, indices wrap modulo W.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t key; idx_t idx; } kv_t;

static void radix_kv(kv_t *a, kv_t *tmp, uint64_t n, unsigned keybits)
{
	for (unsigned shift = 0; shift < keybits; shift += 8) {
		uint64_t cnt[257];
		memset(cnt, 0, sizeof(cnt));
		for (uint64_t i = 0; i < n; ++i) cnt[((a[i].key >> shift) & 255) + 1]++;
		int single = 0;
		for (int d = 0; d < 256; ++d) if (cnt[d + 1] == n) single = 1;
		if (single) continue;
		for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
		for (uint64_t i = 0; i < n; ++i) tmp[cnt[(a[i].key >> shift) & 255]++] = a[i];
		memcpy(a, tmp, n * sizeof(kv_t));
	}
}

typedef struct { idx_t key; idx_t idx; } k2_t;
static int k2cmp(const void *a, const void *b)
{
	k2_t const *x = (k2_t const *)a, *y = (k2_t const *)b;
	if (x->key != y->key) return x->key < y->key ? -1 : 1;
	return 0;
}

/* returns 0 on success; sa[W], rank[W] (rank = final ISA of the window) */
static int window_suffix_sort(const uint8_t *w, uint64_t W, unsigned sigma, int circular, idx_t *sa, idx_t *rank)
{
	if (W == 0) return 0;
	unsigned const vals = circular ? sigma : sigma + 1; /* +1: sentinel code 0 */
	unsigned bits = 1;
	while ((1u << bits) < vals) ++bits;
	unsigned K = 64 / bits;
	if (K > W) K = (unsigned)W;
	if (K == 0) K = 1;

	kv_t *kv = (kv_t *)malloc(W * sizeof(kv_t));
	kv_t *tmp = (kv_t *)malloc(W * sizeof(kv_t));
	if (!kv || !tmp) { free(kv); free(tmp); return -1; }
	uint64_t const mask = (bits * K == 64) ? ~0ull : ((1ull << (bits * K)) - 1);
	/* rolling key, filled right to left */
	{
		uint64_t key = 0;
		/* key for position W-1+K .. built by scanning from the right */
		/* first compute key(W-1) explicitly, then slide left */
		for (unsigned k = 0; k < K; ++k) {
			uint64_t const p = (W - 1) + k;
			uint64_t c;
			if (circular) c = w[p % W]; else c = (p < W) ? (uint64_t)w[p] + 1 : 0;
			key = (key << bits) | c;
		}
		kv[W - 1].key = key; kv[W - 1].idx = (idx_t)(W - 1);
		for (uint64_t i = W - 1; i-- > 0;) {
			uint64_t const c = circular ? (uint64_t)w[i] : (uint64_t)w[i] + 1;
			key = (key >> bits) | (c << (bits * (K - 1)));
			key &= mask;
			kv[i].key = key; kv[i].idx = (idx_t)i;
		}
	}
	radix_kv(kv, tmp, W, bits * K);
	free(tmp);

	/* groups: rank = index of group head; collect unsorted groups */
	uint64_t ngroups = 0, gcap = 1024;
	uint64_t *gs = (uint64_t *)malloc(gcap * 2 * sizeof(uint64_t));
	{
		uint64_t i = 0;
		while (i < W) {
			uint64_t j = i + 1;
			while (j < W && kv[j].key == kv[i].key) ++j;
			for (uint64_t k = i; k < j; ++k) { sa[k] = kv[k].idx; rank[kv[k].idx] = (idx_t)i; }
			if (j - i > 1) {
				if (ngroups == gcap) { gcap *= 2; gs = (uint64_t *)realloc(gs, gcap * 2 * sizeof(uint64_t)); }
				gs[2 * ngroups] = i; gs[2 * ngroups + 1] = j - i; ++ngroups;
			}
			i = j;
		}
	}
	free(kv);

	uint64_t h = K;
	k2_t *buf = NULL; uint64_t bufcap = 0;
	while (ngroups) {
		if (circular && h >= W) break; /* non-primitive text: leave ties in current order */
		uint64_t nn = 0, ncap = 1024;
		uint64_t *ns = (uint64_t *)malloc(ncap * 2 * sizeof(uint64_t));
		/* pass 1: sort every group by the rank h ahead, using ranks from before this round */
		uint64_t total = 0;
		for (uint64_t g = 0; g < ngroups; ++g) total += gs[2 * g + 1];
		if (total > bufcap) { free(buf); bufcap = total; buf = (k2_t *)malloc(bufcap * sizeof(k2_t)); }
		uint64_t o = 0;
		for (uint64_t g = 0; g < ngroups; ++g) {
			uint64_t const s = gs[2 * g], len = gs[2 * g + 1];
			for (uint64_t k = 0; k < len; ++k) {
				uint64_t const p = (uint64_t)sa[s + k] + h;
				idx_t key;
				if (circular) key = rank[p % W]; else key = (p < W) ? rank[p] + 1 : 0;
				buf[o + k].key = key; buf[o + k].idx = sa[s + k];
			}
			o += len;
		}
		/* pass 2: sort and re-rank */
		o = 0;
		for (uint64_t g = 0; g < ngroups; ++g) {
			uint64_t const s = gs[2 * g], len = gs[2 * g + 1];
			k2_t *b = buf + o;
			qsort(b, len, sizeof(k2_t), k2cmp);
			uint64_t i = 0;
			while (i < len) {
				uint64_t j = i + 1;
				while (j < len && b[j].key == b[i].key) ++j;
				for (uint64_t k = i; k < j; ++k) { sa[s + k] = b[k].idx; rank[b[k].idx] = (idx_t)(s + i); }
				if (j - i > 1) {
					if (nn == ncap) { ncap *= 2; ns = (uint64_t *)realloc(ns, ncap * 2 * sizeof(uint64_t)); }
					ns[2 * nn] = s + i; ns[2 * nn + 1] = j - i; ++nn;
				}
				i = j;
			}
			o += len;
		}
		free(gs); gs = ns; ngroups = nn;
		h *= 2;
	}
	free(gs); free(buf);
	return 0;
}

/* whole-text circular suffix (rotation) sort */
int orc_sa_circular(const uint8_t *t, uint64_t n, idx_t *sa)
{
	unsigned sigma = 0;
	for (uint64_t i = 0; i < n; ++i) if (t[i] + 1u > sigma) sigma = t[i] + 1u;
	idx_t *rank = (idx_t *)malloc((n ? n : 1) * sizeof(idx_t));
	int const r = window_suffix_sort(t, n, sigma, 1, sa, rank);
	free(rank);
	return r;
}

/* ------------------------------------------------------------------------------------------
 * Rank dictionary over a symbol array (stands in for libmaus2's Huffman-shaped wavelet tree,
 * A6).  Sampled cumulative counts every RB symbols + scan.  Symbols are uint16 so that block
 * BWTs can carry bwtterm = maxsym+1 (A5).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	const uint16_t *L;
	uint64_t n;
	unsigned sigma;   /* symbols 0..sigma-1 are counted (bwtterm excluded: it is == sigma) */
	unsigned rb;      /* block size */
	uint32_t *cnt;    /* [(n/rb)+1][sigma] */
	uint64_t *C;      /* [sigma+1] exclusive prefix sums of symbol counts of L */
} rankdict_t;

static void rd_build(rankdict_t *d, const uint16_t *L, uint64_t n, unsigned sigma)
{
	d->L = L; d->n = n; d->sigma = sigma;
	d->rb = sigma <= 8 ? 64 : 1024;
	uint64_t const nb = n / d->rb + 1;
	d->cnt = (uint32_t *)malloc(nb * sigma * sizeof(uint32_t));
	d->C = (uint64_t *)calloc(sigma + 2, sizeof(uint64_t));
	uint32_t *run = (uint32_t *)calloc(sigma + 1, sizeof(uint32_t));
	for (uint64_t i = 0; i < n; ++i) {
		if (i % d->rb == 0) memcpy(d->cnt + (i / d->rb) * sigma, run, sigma * sizeof(uint32_t));
		if (L[i] < sigma) run[L[i]]++;
	}
	if (n % d->rb == 0) memcpy(d->cnt + (n / d->rb) * sigma, run, sigma * sizeof(uint32_t));
	uint64_t acc = 0;
	for (unsigned c = 0; c < sigma; ++c) { d->C[c] = acc; acc += run[c]; }
	d->C[sigma] = acc;
	free(run);
}
static void rd_free(rankdict_t *d) { free(d->cnt); free(d->C); }
static inline uint64_t rd_rank(rankdict_t const *d, unsigned c, uint64_t r)
{
	uint64_t const b = r / d->rb;
	uint64_t v = d->cnt[b * d->sigma + c];
	for (uint64_t i = b * d->rb; i < r; ++i) v += (d->L[i] == c);
	return v;
}

/* ------------------------------------------------------------------------------------------
 * checkbwt restatement [REF checkbwt.cpp:126-243]: pick one anchor per thread slot from the
 * (rank,pos) pairs (smallest pos >= i*ceil(n/numthreads)), then for adjacent anchors (PL,PH)
 * LF-walk from PH.rank and require every extendedLF symbol to equal the circularly reversed
 * text read from PH.pos.  Returns 1 if all symbols match (gok), 0 otherwise, <0 on error.
 * Deviation, stated: segment length is (PH.pos-PL.pos) mod n, = n when only one anchor
 * exists, so that every text position is always covered exactly once (the reference relies
 * on position 0 being among the samples for its wrap-around segment).
 * ---------------------------------------------------------------------------------------- */
int orc_checkbwt(const uint8_t *text, uint64_t n, const uint8_t *bwt, const uint64_t *preisa_pairs,
                 uint64_t npairs, unsigned numthreads, uint64_t *checked)
{
	if (!n || !npairs || !numthreads) return -1;
	uint64_t const o = (n + numthreads - 1) / numthreads;
	uint64_t *ap = (uint64_t *)malloc(numthreads * sizeof(uint64_t));
	uint64_t *ar = (uint64_t *)malloc(numthreads * sizeof(uint64_t));
	for (unsigned i = 0; i < numthreads; ++i) ap[i] = ar[i] = UINT64_MAX;
	for (uint64_t k = 0; k < npairs; ++k) {
		uint64_t const r = preisa_pairs[2 * k], p = preisa_pairs[2 * k + 1];
		if (p >= n || r >= n) { free(ap); free(ar); return -2; }
		uint64_t slot = p / o;
		if (slot >= numthreads) slot = numthreads - 1;
		if (p < ap[slot]) { ap[slot] = p; ar[slot] = r; }
	}
	unsigned na = 0;
	for (unsigned i = 0; i < numthreads; ++i) if (ap[i] != UINT64_MAX) { ap[na] = ap[i]; ar[na] = ar[i]; ++na; }

	unsigned sigma = 0;
	uint16_t *L = (uint16_t *)malloc(n * sizeof(uint16_t));
	for (uint64_t i = 0; i < n; ++i) { L[i] = bwt[i]; if (bwt[i] + 1u > sigma) sigma = bwt[i] + 1u; }
	rankdict_t d; rd_build(&d, L, n, sigma);
	int gok = 1;
	uint64_t gc = 0;
	#if defined(_OPENMP)
	#pragma omp parallel for schedule(dynamic,1) num_threads(numthreads)
	#endif
	for (unsigned t = 0; t < na; ++t) {
		unsigned const t1 = (t + 1) % na;
		uint64_t const plp = ap[t], php = ap[t1];
		uint64_t b = (php + n - plp) % n;
		if (b == 0) b = n;
		uint64_t r = ar[t1];
		uint64_t tp = php; /* circular reverse wrapper starting at PH.pos */
		int tok = 1;
		for (uint64_t i = 0; i < b; ++i) {
			unsigned const sym = L[r];                       /* extendedLF(r).first */
			uint64_t const nr = d.C[sym] + rd_rank(&d, sym, r); /* extendedLF(r).second */
			tp = tp ? tp - 1 : n - 1;
			if (sym != text[tp]) { tok = 0; }
			r = nr;
		}
		if (r != ar[t]) tok = 0; /* arriving at PL.pos must give PL.rank */
		#if defined(_OPENMP)
		#pragma omp critical
		#endif
		{ gc += b; if (!tok) gok = 0; }
	}
	if (checked) *checked = gc;
	rd_free(&d); free(L); free(ap); free(ar);
	return gok;
}

/* ------------------------------------------------------------------------------------------
 * oracle_b3m: block partition -> per-block suffix sort with look-ahead (A4,A5) -> gap arrays
 * by backward search (A7) -> gap-driven merge (A8) -> final BWT + (rank,pos) preisa pairs (A9).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	uint64_t a0, a1;      /* text range [a0,a1) */
	uint16_t *L;          /* block/merged BWT, TERM (=sigma) at the rank of suffix a0 */
	uint8_t *gt;          /* gt[i-a0] = rot(i) > rot(a0) */
	uint64_t *srank;      /* rank of sampled positions p (p % rate == 0, a0 <= p < a1), in position order */
	uint64_t nsamp;
	uint64_t isa_a0;      /* local rank of suffix a0 */
} node_t;

typedef struct {
	const uint8_t *t; uint64_t n; unsigned sigma;
	uint64_t rate;           /* preisa sampling rate */
	uint64_t largelcpthres;
	uint64_t nblocks, bs;
	idx_t **leafsa;          /* per leaf: suffix array of the block, global positions */
	uint64_t *leaflen;
	unsigned nthreads;
	uint64_t lfsteps;        /* statistics: gap LF steps */
	uint64_t maxlcpnext;
} b3m_t;

static inline uint64_t first_sample_at_or_after(uint64_t p, uint64_t rate) { return (p + rate - 1) / rate; }

/* bounded then exact LCP of rot(i) and rot(e) [A4] */
static uint64_t lcp_rot(const uint8_t *t, uint64_t n, uint64_t i, uint64_t e, uint64_t cap)
{
	uint64_t k = 0;
	while (k < cap && t[(i + k) % n] == t[(e + k) % n]) ++k;
	return k;
}

static int cmp_rot(const uint8_t *t, uint64_t n, uint64_t a, uint64_t b)
{
	for (uint64_t k = 0; k < n; ++k) {
		uint8_t const x = t[a], y = t[b];
		if (x != y) return x < y ? -1 : 1;
		if (++a == n) a = 0;
		if (++b == n) b = 0;
	}
	return 0;
}

/* number of suffixes of leaf b smaller than rot(z) (z-rank, A5(iv)): binary search in SA_b */
static uint64_t zrank_leaf(b3m_t const *B, uint64_t b, uint64_t z)
{
	idx_t const *sa = B->leafsa[b];
	uint64_t lo = 0, hi = B->leaflen[b];
	while (lo < hi) {
		uint64_t const mid = (lo + hi) / 2;
		if (cmp_rot(B->t, B->n, sa[mid], z) < 0) lo = mid + 1; else hi = mid;
	}
	return lo;
}

static int leaf_build(b3m_t *B, uint64_t b, node_t *N)
{
	const uint8_t *t = B->t; uint64_t const n = B->n;
	uint64_t const s = b * B->bs, e = (s + B->bs < n) ? s + B->bs : n;
	uint64_t const m = e - s;
	uint64_t const en = e % n;
	/* lcpnext: bounded by largelcpthres first, exact only for the capped ones [A4] */
	uint64_t lcpnext = 0;
	if (B->nblocks > 1) {
		for (uint64_t i = s; i < e; ++i) {
			if (i == en) continue;
			uint64_t l = lcp_rot(t, n, i, en, B->largelcpthres);
			if (l >= B->largelcpthres) l = lcp_rot(t, n, i, en, n); /* large-LCP escape */
			if (l > lcpnext) lcpnext = l;
		}
	}
	#if defined(_OPENMP)
	#pragma omp critical
	#endif
	{ if (lcpnext > B->maxlcpnext) B->maxlcpnext = lcpnext; }
	int const whole = (B->nblocks == 1);
	uint64_t const W = whole ? n : m + lcpnext + 1;
	uint8_t *w = (uint8_t *)malloc(W);
	for (uint64_t i = 0; i < W; ++i) w[i] = t[(s + i) % n];
	idx_t *sa = (idx_t *)malloc(W * sizeof(idx_t));
	idx_t *rank = (idx_t *)malloc(W * sizeof(idx_t));
	if (window_suffix_sort(w, W, B->sigma, whole, sa, rank)) return -1;
	free(w);
	/* keep only the block's own suffixes, in order */
	idx_t *bsa = (idx_t *)malloc(m * sizeof(idx_t));
	uint64_t o = 0;
	for (uint64_t k = 0; k < W; ++k) if (sa[k] < m) bsa[o++] = (idx_t)(s + sa[k]);
	free(sa); free(rank);
	if (o != m) return -2;
	B->leafsa[b] = bsa; B->leaflen[b] = m;

	N->a0 = s; N->a1 = e;
	N->L = (uint16_t *)malloc(m * sizeof(uint16_t));
	N->gt = (uint8_t *)malloc(m);
	idx_t *bisa = (idx_t *)malloc(m * sizeof(idx_t));
	for (uint64_t k = 0; k < m; ++k) {
		uint64_t const p = bsa[k];
		bisa[p - s] = (idx_t)k;
		N->L[k] = (p == s) ? (uint16_t)B->sigma : t[(p + n - 1) % n]; /* bwtterm at block start */
	}
	N->isa_a0 = bisa[0];
	for (uint64_t i = 0; i < m; ++i) N->gt[i] = bisa[i] > bisa[0];
	uint64_t const f = first_sample_at_or_after(s, B->rate), l = first_sample_at_or_after(e, B->rate);
	N->nsamp = l - f;
	N->srank = (uint64_t *)malloc((N->nsamp ? N->nsamp : 1) * sizeof(uint64_t));
	for (uint64_t q = f; q < l; ++q) N->srank[q - f] = bisa[q * B->rate - s];
	free(bisa);
	return 0;
}

static void node_free(node_t *N) { free(N->L); free(N->gt); free(N->srank); }

/* merge A (left) and R (right) into M  [A7, A8, Appendix A.2-A.4] */
static int node_merge(b3m_t *B, node_t *A, node_t *R, uint64_t blo, uint64_t bmid, node_t *M)
{
	const uint8_t *t = B->t; uint64_t const n = B->n; unsigned const sigma = B->sigma;
	uint64_t const na = A->a1 - A->a0, nr = R->a1 - R->a0;
	uint64_t const a1 = R->a0, r1 = R->a1;
	rankdict_t d; rd_build(&d, A->L, na, sigma);
	/* C_A counted over A's text: counts of L_A (TERM excluded) plus the last symbol of A */
	uint64_t *CA = (uint64_t *)calloc(sigma + 1, sizeof(uint64_t));
	{
		uint64_t *h = (uint64_t *)calloc(sigma + 1, sizeof(uint64_t));
		for (uint64_t i = A->a0; i < A->a1; ++i) h[t[i]]++;
		uint64_t acc = 0;
		for (unsigned c = 0; c < sigma; ++c) { CA[c] = acc; acc += h[c]; }
		free(h);
	}
	uint8_t const lastA = t[a1 - 1];
	uint32_t *G = (uint32_t *)calloc(na + 1, sizeof(uint32_t));
	uint64_t *rsamp = (uint64_t *)malloc((R->nsamp ? R->nsamp : 1) * sizeof(uint64_t)); /* r(j) at sampled j */
	uint8_t *gtnew = (uint8_t *)malloc(nr);
	uint64_t const fR = first_sample_at_or_after(a1, B->rate);

	/* chains: split R into pieces, start rank of each piece from z-ranks of A's leaves */
	uint64_t nch = B->nthreads * 4;
	if (nch > nr) nch = nr;
	if (nch == 0) nch = 1;
	uint64_t const chl = (nr + nch - 1) / nch;
	nch = (nr + chl - 1) / chl;
	#if defined(_OPENMP)
	#pragma omp parallel for schedule(dynamic,1) num_threads(B->nthreads)
	#endif
	for (uint64_t c = 0; c < nch; ++c) {
		uint64_t const zlo = a1 + c * chl;
		uint64_t const zhi = (zlo + chl < r1) ? zlo + chl : r1;
		/* r(zhi) = #{a in A : rot(a) < rot(zhi)} : sum of leaf z-ranks */
		uint64_t r = 0;
		for (uint64_t b = blo; b < bmid; ++b) r += zrank_leaf(B, b, zhi % n);
		for (uint64_t j = zhi; j > zlo; --j) {
			/* gt_R[j] for j == r1 is rot(r1) > rot(a1): decided directly */
			int gtj;
			if (j < r1) gtj = R->gt[j - a1];
			else gtj = cmp_rot(t, n, r1 % n, a1) > 0;
			uint8_t const c0 = t[j - 1];
			r = CA[c0] + rd_rank(&d, c0, r) + ((c0 == lastA && gtj) ? 1 : 0);
			__atomic_fetch_add(&G[r], 1, __ATOMIC_RELAXED);
			gtnew[j - 1 - a1] = r > A->isa_a0; /* A.3 */
			if ((j - 1) % B->rate == 0) rsamp[(j - 1) / B->rate - fR] = r;
		}
	}
	__atomic_fetch_add(&B->lfsteps, nr, __ATOMIC_RELAXED);

	/* merge by G [A.4] */
	M->a0 = A->a0; M->a1 = r1;
	M->L = (uint16_t *)malloc((na + nr) * sizeof(uint16_t));
	uint64_t *S = (uint64_t *)malloc((na + 1) * sizeof(uint64_t)); /* inclusive prefix sums of G */
	{
		uint64_t acc = 0;
		for (uint64_t k = 0; k <= na; ++k) { acc += G[k]; S[k] = acc; }
		if (acc != nr) return -3;
		#if defined(_OPENMP)
		#pragma omp parallel for schedule(static) num_threads(B->nthreads)
		#endif
		for (uint64_t k = 0; k <= na; ++k) {
			uint64_t q = S[k] - G[k];     /* first R symbol of gap k */
			uint64_t o = k + q;           /* its place in the merged sequence */
			for (uint32_t g = 0; g < G[k]; ++g) {
				uint16_t s = R->L[q++];
				if (s == sigma) s = t[a1 - 1]; /* stale bwtterm of R -> true seam symbol */
				M->L[o++] = s;
			}
			if (k < na) M->L[o] = A->L[k];
		}
	}
	M->isa_a0 = A->isa_a0 + S[A->isa_a0];
	M->gt = (uint8_t *)malloc(na + nr);
	memcpy(M->gt, A->gt, na);
	memcpy(M->gt + na, gtnew, nr);
	M->nsamp = A->nsamp + R->nsamp;
	M->srank = (uint64_t *)malloc((M->nsamp ? M->nsamp : 1) * sizeof(uint64_t));
	for (uint64_t q = 0; q < A->nsamp; ++q) M->srank[q] = A->srank[q] + S[A->srank[q]];
	for (uint64_t q = 0; q < R->nsamp; ++q) M->srank[A->nsamp + q] = R->srank[q] + rsamp[q];
	free(S); free(G); free(rsamp); free(gtnew); free(CA); rd_free(&d);
	return 0;
}

static int build_rec(b3m_t *B, node_t *leaves, uint64_t lo, uint64_t hi, node_t *out)
{
	if (hi - lo == 1) { *out = leaves[lo]; memset(&leaves[lo], 0, sizeof(node_t)); return 0; }
	uint64_t const mid = (lo + hi) / 2;
	node_t A, R;
	int rc;
	if ((rc = build_rec(B, leaves, lo, mid, &A))) return rc;
	if ((rc = build_rec(B, leaves, mid, hi, &R))) return rc;
	rc = node_merge(B, &A, &R, lo, mid, out);
	node_free(&A); node_free(&R);
	return rc;
}

/* default block size of the reference [RECALL, SURVEY 3.1 step 3 / Appendix B]:
 * tblock = min(max(0.95*mem/(5*threads),1), ceil(fs/threads)); numblocks = ceil(fs/tblock) */
uint64_t orc_default_numblocks(uint64_t fs, uint64_t mem, uint64_t threads)
{
	if (!fs) return 1;
	if (!threads) threads = 1;
	uint64_t tb = (uint64_t)(0.95 * (double)mem / (5.0 * (double)threads));
	if (tb < 1) tb = 1;
	uint64_t const per = (fs + threads - 1) / threads;
	if (per < tb) tb = per;
	return (fs + tb - 1) / tb;
}

/* text: n symbols (reference symbol space).  Outputs: bwt[n]; preisa pairs (rank,pos) for every
 * pos % rate == 0 written as 2*ceil(n/rate) uint64.  stats (nullable): [0]=gap LF steps,
 * [1]=max lcpnext, [2]=seconds leaves, [3]=seconds merges (double bit patterns avoided: ms). */
int orc_b3m(const uint8_t *text, uint64_t n, uint64_t nblocks, uint64_t rate, uint64_t largelcpthres,
            unsigned nthreads, uint8_t *bwt, uint64_t *preisa_pairs, uint64_t *stats)
{
	if (!n || !rate) return -1;
	if (n >= 0xFFFFFFF0ull) return -4;
	if (nblocks < 1) nblocks = 1;
	if (nblocks > n) nblocks = n;
	if (!nthreads) nthreads = 1;
	b3m_t B; memset(&B, 0, sizeof(B));
	B.t = text; B.n = n; B.rate = rate; B.largelcpthres = largelcpthres ? largelcpthres : 16384;
	B.nthreads = nthreads;
	for (uint64_t i = 0; i < n; ++i) if (text[i] + 1u > B.sigma) B.sigma = text[i] + 1u;
	B.bs = (n + nblocks - 1) / nblocks;
	B.nblocks = (n + B.bs - 1) / B.bs;
	B.leafsa = (idx_t **)calloc(B.nblocks, sizeof(idx_t *));
	B.leaflen = (uint64_t *)calloc(B.nblocks, sizeof(uint64_t));
	node_t *leaves = (node_t *)calloc(B.nblocks, sizeof(node_t));
	int rc = 0;
	double const t0 = now_sec();
	#if defined(_OPENMP)
	#pragma omp parallel for schedule(dynamic,1) num_threads(nthreads)
	#endif
	for (uint64_t b = 0; b < B.nblocks; ++b) {
		int const r = leaf_build(&B, b, &leaves[b]);
		if (r) rc = r;
	}
	double const t1 = now_sec();
	node_t root; memset(&root, 0, sizeof(root));
	if (!rc) rc = build_rec(&B, leaves, 0, B.nblocks, &root);
	double const t2 = now_sec();
	if (!rc) {
		for (uint64_t k = 0; k < n; ++k) {
			uint16_t s = root.L[k];
			if (s == B.sigma) s = text[n - 1]; /* root: remaining bwtterm -> T[n-1] (A.4) */
			bwt[k] = (uint8_t)s;
		}
		uint64_t const ns = (n + rate - 1) / rate;
		if (root.nsamp != ns) rc = -5;
		else for (uint64_t q = 0; q < ns; ++q) { preisa_pairs[2 * q] = root.srank[q]; preisa_pairs[2 * q + 1] = q * rate; }
	}
	if (stats) {
		stats[0] = B.lfsteps; stats[1] = B.maxlcpnext;
		stats[2] = (uint64_t)((t1 - t0) * 1e6); stats[3] = (uint64_t)((t2 - t1) * 1e6);
	}
	node_free(&root);
	for (uint64_t b = 0; b < B.nblocks; ++b) { free(B.leafsa[b]); node_free(&leaves[b]); }
	free(B.leafsa); free(B.leaflen); free(leaves);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * A10: sampled SA / ISA by LF walk from the (rank,pos) anchors
 * [REF hwtPreIsaToIsa.cpp:79,114-161; SURVEY Appendix A.5].  sa_out has ceil(n/sarate)
 * entries (SA sampled by rank), isa_out has ceil(n/isarate) entries (ISA sampled by position).
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t p, r; } anchor_t;
static int anchor_cmp(const void *a, const void *b)
{
	anchor_t const *x = (anchor_t const *)a, *y = (anchor_t const *)b;
	if (x->p != y->p) return x->p < y->p ? -1 : 1;
	return 0;
}

int orc_ssa(const uint8_t *bwt, uint64_t n, const uint64_t *preisa_pairs, uint64_t npairs,
            uint64_t sarate, uint64_t isarate, uint64_t *sa_out, uint64_t *isa_out, unsigned nthreads)
{
	if (!n || !npairs) return -1;
	if ((sarate & (sarate - 1)) || (isarate & (isarate - 1)) || !sarate || !isarate) return -2; /* powers of two */
	if (!nthreads) nthreads = 1;
	anchor_t *A = (anchor_t *)malloc(npairs * sizeof(anchor_t));
	for (uint64_t k = 0; k < npairs; ++k) { A[k].r = preisa_pairs[2 * k]; A[k].p = preisa_pairs[2 * k + 1]; }
	qsort(A, npairs, sizeof(anchor_t), anchor_cmp);
	unsigned sigma = 0;
	uint16_t *L = (uint16_t *)malloc(n * sizeof(uint16_t));
	for (uint64_t i = 0; i < n; ++i) { L[i] = bwt[i]; if (bwt[i] + 1u > sigma) sigma = bwt[i] + 1u; }
	rankdict_t d; rd_build(&d, L, n, sigma);
	uint64_t const nsa = (n + sarate - 1) / sarate, nisa = (n + isarate - 1) / isarate;
	for (uint64_t i = 0; i < nsa; ++i) sa_out[i] = UINT64_MAX;
	for (uint64_t i = 0; i < nisa; ++i) isa_out[i] = UINT64_MAX;
	#if defined(_OPENMP)
	#pragma omp parallel for schedule(dynamic,16) num_threads(nthreads)
	#endif
	for (uint64_t k = 0; k < npairs; ++k) {
		uint64_t const prev = (k + npairs - 1) % npairs;
		uint64_t todo = (A[k].p + n - A[prev].p) % n;
		if (todo == 0) todo = (npairs == 1) ? n : 0;
		uint64_t p = A[k].p, r = A[k].r;
		while (todo--) {
			if ((p & (isarate - 1)) == 0) isa_out[p / isarate] = r;
			if ((r & (sarate - 1)) == 0) sa_out[r / sarate] = p;
			p = p ? p - 1 : n - 1;
			unsigned const sym = L[r];
			r = d.C[sym] + rd_rank(&d, sym, r); /* LF(r) [REF lcpbit.cpp:3362-3365] */
		}
	}
	int rc = 0;
	for (uint64_t i = 0; i < nsa; ++i) if (sa_out[i] == UINT64_MAX) rc = -3; /* [REF hwtPreIsaToIsa.cpp:166-167] */
	for (uint64_t i = 0; i < nisa; ++i) if (isa_out[i] == UINT64_MAX) rc = -3;
	rd_free(&d); free(L); free(A);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * LF-steps/s instrument [REF bwttestdecodespeed.cpp:67-97]: `par` interleaved dependent LF
 * chains started at evenly spaced sampled-ISA ranks, after a 128 MiB cache flush; returns
 * steps per second (par*tsteps/t).  maxsteps bounds tsteps (the reference uses 128 Mi).
 * ---------------------------------------------------------------------------------------- */
double orc_lf_speed(const uint8_t *bwt, uint64_t n, const uint64_t *isa_samples, uint64_t nisa,
                    unsigned tpar, uint64_t maxsteps, uint64_t *checksum)
{
	unsigned sigma = 0;
	uint16_t *L = (uint16_t *)malloc(n * sizeof(uint16_t));
	for (uint64_t i = 0; i < n; ++i) { L[i] = bwt[i]; if (bwt[i] + 1u > sigma) sigma = bwt[i] + 1u; }
	rankdict_t d; rd_build(&d, L, n, sigma);
	{
		uint64_t const fl = 128ull * 1024 * 1024;
		uint8_t *a = (uint8_t *)malloc(fl), *b = (uint8_t *)malloc(fl);
		uint64_t x = 88172645463325252ull;
		for (uint64_t i = 0; i < fl; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; a[i] = (uint8_t)x; }
		memcpy(b, a, fl);
		volatile uint8_t sink = b[fl / 2]; (void)sink;
		free(a); free(b);
	}
	uint64_t const step = (nisa + tpar - 1) / tpar;
	uint64_t const par = (nisa + step - 1) / step;
	uint64_t tsteps = (n + par - 1) / par;
	if (tsteps > maxsteps) tsteps = maxsteps;
	uint64_t R[64];
	for (uint64_t i = 0; i < par && i < 64; ++i) R[i] = isa_samples[i * step];
	double const t0 = now_sec();
	for (uint64_t i = 0; i < tsteps; ++i)
		for (uint64_t j = 0; j < par; ++j) {
			unsigned const sym = L[R[j]];
			R[j] = d.C[sym] + rd_rank(&d, sym, R[j]);
		}
	double const t = now_sec() - t0;
	uint64_t cs = 0;
	for (uint64_t j = 0; j < par; ++j) cs ^= R[j];
	if (checksum) *checksum = cs;
	rd_free(&d); free(L);
	return (double)(par * tsteps) / t;
}

/* ------------------------------------------------------------------------------------------
 * BWA export [REF bwtb3mtobwa.cpp:23-30; formats: SURVEY 8f-1, public BWA bwt_dump_bwt /
 * bwt_dump_sa].  Input: pacterm BWT (symbols 0..4, exactly one 0), sampled SA by rank with
 * rate sarate, primary = ISA[0].  Output buffers sized by the caller:
 *   bwa_bwt: 8 + 32 + 4*ceil(seq_len/16) bytes;  bwa_sa: 8 + 32 + 8 + 8 + 8*(n_sa-1) bytes.
 * ---------------------------------------------------------------------------------------- */
int orc_to_bwa(const uint8_t *bwt, uint64_t n, const uint64_t *sa_samples, uint64_t sarate,
               uint8_t *bwa_bwt, uint64_t *bwa_bwt_len, uint8_t *bwa_sa, uint64_t *bwa_sa_len)
{
	if (n < 2) return -1;
	uint64_t const seq_len = n - 1;
	uint64_t primary = UINT64_MAX, cnt[5] = {0, 0, 0, 0, 0};
	for (uint64_t i = 0; i < n; ++i) {
		if (bwt[i] > 4) return -2;
		cnt[bwt[i]]++;
		if (bwt[i] == 0) primary = i;
	}
	if (cnt[0] != 1) return -3;
	uint64_t L2[5]; L2[0] = 0;
	for (int c = 0; c < 4; ++c) L2[c + 1] = L2[c] + cnt[c + 1];
	uint64_t const nw = (seq_len + 15) >> 4;
	uint8_t *o = bwa_bwt;
	memcpy(o, &primary, 8); o += 8;
	memcpy(o, L2 + 1, 32); o += 32;
	uint32_t *wds = (uint32_t *)calloc(nw ? nw : 1, 4);
	uint64_t k = 0;
	for (uint64_t i = 0; i < n; ++i) {
		if (i == primary) continue;
		wds[k >> 4] |= (uint32_t)(bwt[i] - 1) << ((15 - (k & 15)) << 1);
		++k;
	}
	memcpy(o, wds, nw * 4); o += nw * 4;
	free(wds);
	*bwa_bwt_len = (uint64_t)(o - bwa_bwt);
	uint64_t const n_sa = (seq_len + sarate) / sarate;
	o = bwa_sa;
	memcpy(o, &primary, 8); o += 8;
	memcpy(o, L2 + 1, 32); o += 32;
	memcpy(o, &sarate, 8); o += 8;
	memcpy(o, &seq_len, 8); o += 8;
	for (uint64_t q = 1; q < n_sa; ++q) { memcpy(o, &sa_samples[q], 8); o += 8; }
	*bwa_sa_len = (uint64_t)(o - bwa_sa);
	return 0;
}

unsigned orc_max_threads(void)
{
	#if defined(_OPENMP)
	return (unsigned)omp_get_max_threads();
	#else
	return 1;
	#endif
}
