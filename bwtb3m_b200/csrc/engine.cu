// Engine: owns the device text and results, runs K1..K7 on one stream, exports the C ABI of
// include/b3m.h.  Host orchestration only; all arithmetic is in the kernels.
#include "engine.h"
#include "scan.cuh"
#include <string.h>
#include <algorithm>
#include <map>
#include <tuple>

namespace b3m {

// ------------------------------------------------------------------------------------------

Engine::Engine(int dev, void * stream) : device(dev) {
	B3M_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	B3M_CUDA(cudaGetDeviceProperties(&prop, device));
	st.sms = prop.multiProcessorCount;
	if (stream) { st.s = (cudaStream_t)stream; own_stream = false; }
	else { B3M_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking)); own_stream = true; }
	st.arena = &arena;
	B3M_CUDA(cudaStreamCreateWithFlags(&st.copy, cudaStreamNonBlocking));
	B3M_CUDA(cudaMallocHost((void **)&pinned, 4096));
}

Engine::~Engine() {
	cudaSetDevice(device);
	cudaStreamSynchronize(st.s);
	raw.release(); codes.release(); packed.release(); bwt.release(); prerank.release(); sa.release(); isa.release(); dict.release();
	d_hist.release(); d_special.release(); xs.toff.release();
	delete xs_pt; xs_pt = nullptr;
	cudaStreamSynchronize(st.s);
	arena.release_all();
	if (pinned) cudaFreeHost(pinned);
	if (st.copy) cudaStreamDestroy(st.copy);
	if (own_stream) cudaStreamDestroy(st.s);
}

void Engine::reset_results() {
	ssa_only = false;
	bwt.release(); prerank.release(); sa.release(); isa.release(); dict.release();
	D = DevDict();
	sa_on_host = nullptr; bwa_on_host = nullptr; bwa_words.release();
	have_results = false;
}

// K1 ---------------------------------------------------------------------------------------
void Engine::load(const void * input, uint64_t nbytes, int itype, bool on_device) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(itype >= 0 && itype <= 3, "unknown input type");
	reset_results();
	codes.release(); packed.release(); raw.release(); d_hist.release(); d_special.release(); lastcode.release();
	kr_plan = KeyRangePlan();
	xs = XShard();
	inputtype = itype;
	// slab 0 of the arena: the staged file, the byte codes, the packed text.  The working set of a build is
	// sized in build(), once the block plan and the sorter are known (HBM layout)
	{
		uint64_t const nsym_est = (itype == B3M_INPUT_PAC || itype == B3M_INPUT_PACTERM) ? nbytes * 4 : nbytes;
		arena.ensure_slab(0, (size_t)(nbytes + nsym_est + nsym_est / 4 + (8u << 20)));
	}
	PhaseTimer pt(st);
	pt.mark();
	const uint8_t * d_in = nullptr;
	auto stage = [&](uint64_t off, uint64_t len, uint64_t pad) {
		raw.alloc(st, len + pad + 16);
		if (pad) B3M_CUDA(cudaMemsetAsync(raw.get() + len, 0, pad + 16, st.s));
		B3M_CUDA(cudaMemcpyAsync(raw.get(), (const uint8_t *)input + off, len,
		                         on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st.s));
		d_in = raw.get();
	};
	auto peek = [&](uint64_t off, uint64_t len, uint8_t * dst) {
		if (on_device) {
			B3M_CUDA(cudaMemcpyAsync(pinned, (const uint8_t *)input + off, len, cudaMemcpyDeviceToHost, st.s));
			B3M_CUDA(cudaStreamSynchronize(st.s));
			memcpy(dst, pinned, len);
		} else memcpy(dst, (const uint8_t *)input + off, len);
	};
	d_hist.alloc(st, 256);
	memset(hist, 0, sizeof(hist));
	T = DevText();
	uint64_t hcodes[256];
	if (itype == B3M_INPUT_PAC || itype == B3M_INPUT_PACTERM) {
		B3M_REQUIRE(nbytes >= 2, "pac file too short");
		uint8_t last;
		peek(nbytes - 1, 1, &last);
		B3M_REQUIRE(last < 4, "pac file: bad count byte");
		uint64_t const l = (nbytes - 2) * 4 + last; // BWA fa2pac layout
		B3M_REQUIRE(l > 0, "empty input");
		if (on_device) d_in = (const uint8_t *)input; else stage(0, nbytes, 0);
		// straight to the packed text; the byte codes are made when a path asks for them (ensure_codes)
		packed.alloc(st, l / 32 + 3);
		lastcode.alloc(st, 16);
		k1_pac_to_packed(st, d_in, l, packed.get(), d_hist.get(), lastcode.get());
		B3M_CUDA(cudaMemcpyAsync(pinned, d_hist.get(), 256 * 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		memcpy(hcodes, pinned, sizeof(hcodes));
		T.ntext = l; T.sigma = 4; T.keybits = 2;
		if (itype == B3M_INPUT_PACTERM) {
			T.has_term = 1; T.n = l + 1;
			hist[0] = 1;
			for (int c = 0; c < 4; ++c) { hist[c + 1] = hcodes[c]; code2sym[c] = (uint8_t)(c + 1); }
		} else {
			T.has_term = 0; T.n = l;
			for (int c = 0; c < 4; ++c) { hist[c] = hcodes[c]; code2sym[c] = (uint8_t)c; }
		}
		for (int c = 0; c < 4; ++c) codehist[c] = hcodes[c];
		decode_bytes = nbytes + l / 4;
	} else {
		uint64_t nsym;
		if (itype == B3M_INPUT_COMPACTSTREAM) {
			// [layout unpinned, SURVEY 8c] the serialised CompactArray the reference's writers leave behind
			// (/root/reference/src/digitsToCompact.cpp:35,122): 4 uint64 (bits, n, words, words), then the words.
			// Both byte orders are accepted: native little-endian numbers and words (what libmaus2's
			// Serialize<uint64_t> writes) or big-endian numbers with a big-endian bit stream; bits per
			// symbol is 1..8 in exactly one of the two readings.
			B3M_REQUIRE(nbytes >= 32, "compact file too short");
			uint8_t hdr[32];
			peek(0, 32, hdr);
			// the four numbers in both readings; a reading is plausible when bits is 1..8 and both word counts equal ceil(n*bits/64)
			uint64_t be[4], le[4];
			for (int f = 0; f < 4; ++f) {
				be[f] = le[f] = 0;
				for (int i = 0; i < 8; ++i) { be[f] = (be[f] << 8) | hdr[8 * f + i]; le[f] = (le[f] << 8) | hdr[8 * f + 7 - i]; }
			}
			auto plausible = [](const uint64_t * h) {
				if (h[0] < 1 || h[0] > 8 || h[1] > (~0ull >> 4)) return false;
				uint64_t const w = (h[1] * h[0] + 63) / 64;
				return h[2] == w && h[3] == w;
			};
			bool const ok_be = plausible(be), ok_le = plausible(le);
			// both or neither plausible (e.g. word counts written differently): fall back to "bits in 1..8", big-endian first
			bool const le_words = (ok_be != ok_le) ? ok_le : !(be[0] >= 1 && be[0] <= 8);
			uint64_t const b = le_words ? le[0] : be[0], n = le_words ? le[1] : be[1];
			B3M_REQUIRE(b >= 1 && b <= 8, "compact file: unsupported bits per symbol");
			B3M_REQUIRE(n <= (nbytes - 32) * 8 / b && (!le_words || (n * b + 63) / 64 * 8 <= nbytes - 32), "compact file: truncated");
			B3M_REQUIRE(n > 0, "empty input");
			stage(32, nbytes - 32, 8);
			codes.alloc(st, n + 16);
			k1_unpack_compact(st, d_in, n, (unsigned)b, le_words, codes.get(), d_hist.get());
			nsym = n;
		} else {
			B3M_REQUIRE(nbytes > 0, "empty input");
			if (on_device) d_in = (const uint8_t *)input; else stage(0, nbytes, 0);
			codes.alloc(st, nbytes + 16);
			k1_hist_bytes(st, d_in, nbytes, d_hist.get());
			nsym = nbytes;
		}
		B3M_CUDA(cudaMemcpyAsync(pinned, d_hist.get(), 256 * 8, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		memcpy(hist, pinned, sizeof(hist));
		// dense, order preserving alphabet
		uint8_t lut[256];
		memset(lut, 0, sizeof(lut));
		uint32_t sigma = 0;
		for (int s = 0; s < 256; ++s) if (hist[s]) { lut[s] = (uint8_t)sigma; code2sym[sigma] = (uint8_t)s; codehist[sigma] = hist[s]; ++sigma; }
		memcpy(pinned, lut, 256);
		DevBuf<uint8_t> dlut(st, 256);
		B3M_CUDA(cudaMemcpyAsync(dlut.get(), pinned, 256, cudaMemcpyHostToDevice, st.s));
		const uint8_t * src = (itype == B3M_INPUT_COMPACTSTREAM) ? (const uint8_t *)codes.get() : d_in;
		k1_map_bytes(st, src, nsym, dlut.get(), codes.get());
		B3M_CUDA(cudaStreamSynchronize(st.s)); // pinned staging buffer is reused below
		T.ntext = nsym; T.n = nsym; T.sigma = sigma; T.has_term = 0;
		T.keybits = sigma <= 4 ? 2 : (sigma <= 16 ? 4 : 8);
		decode_bytes = 3 * nsym;
	}
	B3M_REQUIRE(T.n < 0xFFFFFF00ull, "inputs of 2^32 - 256 symbols or more are not supported yet");
	T.codes = codes.get(); // nullptr for pac / pacterm until ensure_codes()
	if (packed.get()) T.packed = packed.get();
	else {
		lastcode.alloc(st, 16);
		B3M_CUDA(cudaMemcpyAsync(lastcode.get(), T.codes + T.ntext - 1, 1, cudaMemcpyDeviceToDevice, st.s));
		if (T.keybits == 2) {
			packed.alloc(st, T.ntext / 32 + 3);
			k1_pack2(st, T.codes, T.ntext, packed.get());
			T.packed = packed.get();
			decode_bytes += T.ntext + T.ntext / 4;
		}
	}
	raw.release();
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_decode = pt.ms(0, 1);
	loaded = true;
}

void Engine::ensure_codes() {
	if (T.codes) return;
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded && T.packed, "internal: no packed text to expand");
	codes.alloc(st, T.ntext + 16);
	k1_unpack_packed(st, T.packed, T.ntext, codes.get());
	T.codes = codes.get();
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pairs(const uint32_t * __restrict__ prerank, uint64_t ns, uint64_t rate, unsigned long long * __restrict__ pairs) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= ns) return;
	pairs[2 * q] = prerank[q];
	pairs[2 * q + 1] = q * rate;
}

__global__ void __launch_bounds__(256)
k_fill_u64(unsigned long long * p, uint64_t n, unsigned long long v) {
	uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

static uint64_t choose_preisarate(uint64_t n, int sms) {
	// enough independent LF chains to fill every SM (2048 threads each) a few times over,
	// but chains of at least 64 steps
	uint64_t const want = (uint64_t)sms * 2048 * 2;
	uint64_t r = 64;
	while (r < (1u << 18) && n / r > want) r <<= 1;
	return r;
}

uint64_t Engine::choose_preisarate_pub(uint64_t n) const { return choose_preisarate(n, st.sms); }

// Device bytes the working set of a build needs in the common case (rarer needs -- the prefix-doubling rounds
// of a repetitive text -- grow the arena on demand): `window` suffixes are sorted at a time.  MSD sorter
// (msd.cuh): 8 B records + 0.75 B tables; LSD sorter and leaves: two (key, index, aux) record sets, order, flags.
uint64_t Engine::work_bytes(uint64_t nblocks, uint64_t window, bool with_results) const {
	bool const msd = nblocks == 1 && T.keybits == 2 && T.packed && st.sortpath != B3M_SORT_LSD && T.ntext >= 64 &&
	                 (st.sortpath == B3M_SORT_MSD || T.ntext >= (1u << 16));
	uint64_t need = (msd ? 10 : 30) * window + (64ull << 20);
	if (nblocks > 1) need += 8 * T.n;          // node BWTs, gap array, gt bits, dictionary of the top merges
	if (with_results) need += 2 * T.n + T.n / 2; // BWT, sampled SA / ISA, anchors, BWA words
	return need;
}

void Engine::make_dict(uint32_t exc_pos, uint32_t exc_code, uint32_t exc_lf) {
	int const flavour = T.sigma <= 4 ? 2 : 8;
	size_t const bytes = dict_bytes(flavour, T.n, T.sigma);
	dict.alloc(st, bytes);
	k4_build_dict(st, bwt.get(), T.n, flavour, T.sigma, dict.get());
	D = DevDict();
	D.flavour = flavour; D.n = T.n; D.lines = dict.get(); D.sigma = T.sigma;
	uint64_t acc = T.has_term ? 1 : 0;
	for (uint32_t c = 0; c < 257; ++c) { D.C[c] = (uint32_t)acc; if (c < T.sigma) acc += codehist[c]; }
	D.exc_pos = exc_pos; D.exc_code = exc_code; D.exc_lf = exc_lf;
	dict_bytes_moved = T.n + bytes;
}

void Engine::build(b3m_build_params const & p) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no input loaded");
	B3M_REQUIRE(p.numblocks >= 1, "numblocks must be >= 1");
	auto pow2 = [](uint64_t v) { return v && !(v & (v - 1)); };
	B3M_REQUIRE(pow2(p.sasamplingrate) && pow2(p.isasamplingrate), "sampling rates must be powers of two");
	B3M_REQUIRE(p.sortpath >= B3M_SORT_AUTO && p.sortpath <= B3M_SORT_MSD, "unknown sortpath");
	reset_results();
	params = p;
	st.sortpath = p.sortpath;
	prerate = p.preisarate ? p.preisarate : (p.bwtonly ? 64 : choose_preisarate(T.n, st.sms));
	B3M_REQUIRE(pow2(prerate), "preisarate must be a power of two");
	npre = div_up(T.n, prerate);
	sortstats = SortStats(); walkstats = WalkStats();
	gap_lf_steps = gap_chains = merge_bytes = extract_bytes = 0; max_lcpnext = large_lcp_blocks = 0;
	ms_sort = ms_extract = ms_dict = ms_gap = ms_merge = ms_walk = 0;
	numblocks = std::min<uint64_t>(p.numblocks, T.n);
	arena.ensure_slab(1, (size_t)work_bytes(numblocks, div_up(T.n, numblocks), true));

	PhaseTimer pt(st);
	pt.mark(); // 0
	d_special.alloc(st, 8);
	B3M_CUDA(cudaMemsetAsync(d_special.get(), 0xff, 16, st.s));
	bwt.alloc(st, T.n + 16);
	prerank.alloc(st, npre);
	uint32_t exc_pos = 0xffffffffu;
	// With the whole text in one block the sorted order itself yields the sampled SA/ISA; the LF
	// walk (K7) is what the reference must do because it never holds a full suffix array.
	bool const direct = numblocks == 1 && !p.bwtonly && p.sampling != B3M_SAMPLING_WALK;
	nsa = nisa = 0;
	if (!p.bwtonly) {
		nsa = div_up(T.n, p.sasamplingrate);
		nisa = div_up(T.n, p.isasamplingrate);
		sa.alloc(st, nsa); isa.alloc(st, nisa);
		if (!direct) {
			B3M_LAUNCH(st, k_fill_u64, (unsigned)div_up(nsa, 256), 256, 0, (unsigned long long *)sa.get(), nsa, ~0ull);
			B3M_LAUNCH(st, k_fill_u64, (unsigned)div_up(nisa, 256), 256, 0, (unsigned long long *)isa.get(), nisa, ~0ull);
		}
	}
	if (numblocks == 1) {
		uint64_t const W = T.ntext;
		uint64_t const shift = T.has_term ? 1 : 0;
		FusedOut fo;
		fo.bwt = bwt.get(); fo.shift = shift; fo.has_term = T.has_term; fo.special = d_special.get();
		fo.prerank = prerank.get(); fo.prelog = ceil_log2_u64(prerate);
		if (direct) {
			fo.sa_s = (unsigned long long *)sa.get(); fo.salog = ceil_log2_u64(p.sasamplingrate);
			fo.isa_s = (unsigned long long *)isa.get(); fo.isalog = ceil_log2_u64(p.isasamplingrate);
		}
		StreamOut so;
		if (direct && p.host_sa) { so.host_sa = (unsigned long long *)p.host_sa; so.nsa = nsa; }
		if (T.has_term) // rank 0 is the terminator suffix (text position ntext); its predecessor is the last base
			B3M_CUDA(cudaMemcpyAsync(bwt.get(), lastcode.get(), 1, cudaMemcpyDeviceToDevice, st.s));
		if (p.host_bwa && T.has_term && T.sigma <= 4) {
			bwa_words.alloc(st, (T.ntext + 15) / 16 + 1);
			so.host_bwa = p.host_bwa; so.d_bwa = bwa_words.get();
		}
		{
			DevBuf<uint32_t> dsa;
			k2_suffix_sort(st, T, 0, W, T.has_term ? 0 : 1, 0, dsa, nullptr, &sortstats, &fo, &so);
		}
		if (so.host_sa) {
			B3M_CUDA(cudaStreamSynchronize(st.copy));
			if (!so.delivered) B3M_CUDA(cudaMemcpyAsync(so.host_sa, sa.get(), 8 * nsa, cudaMemcpyDeviceToHost, st.s)); // rewritten after doubling
			B3M_CUDA(cudaStreamSynchronize(st.s));
			if (T.has_term) so.host_sa[0] = T.ntext; // the terminator suffix, set on the device below
			sa_on_host = (uint64_t *)so.host_sa;
		}
		if (so.host_bwa) {
			B3M_CUDA(cudaStreamSynchronize(st.copy));
			if (so.bwa_delivered) bwa_on_host = so.host_bwa; // otherwise b3m_engine_fetch_bwa packs and copies as usual
			bwa_words.release();
		}
		if (T.has_term) {
			uint64_t const zero = 0, pos = T.ntext;
			if ((T.ntext & (prerate - 1)) == 0) B3M_CUDA(cudaMemsetAsync(prerank.get() + T.ntext / prerate, 0, 4, st.s));
			if (direct) {
				B3M_CUDA(cudaMemcpyAsync(sa.get(), &pos, 8, cudaMemcpyHostToDevice, st.s));
				if ((T.ntext & (p.isasamplingrate - 1)) == 0) B3M_CUDA(cudaMemcpyAsync(isa.get() + T.ntext / p.isasamplingrate, &zero, 8, cudaMemcpyHostToDevice, st.s));
			}
			B3M_CUDA(cudaMemcpyAsync(pinned, d_special.get(), 16, cudaMemcpyDeviceToHost, st.s));
			B3M_CUDA(cudaStreamSynchronize(st.s));
			exc_pos = ((uint32_t *)pinned)[0];
			B3M_REQUIRE(exc_pos != 0xffffffffu, "internal: terminator row not found");
		}
		extract_bytes = W * 37 + npre * 4 + (direct ? 8 * (nsa + nisa) : 0);
		pt.mark(); // 1
		pt.mark(); // 2
	} else {
		build_blocks(pt, &exc_pos);
		pt.mark(); pt.mark(); // 1, 2: the phases are timed inside build_blocks
	}
	root_exc_pos = exc_pos;
	// the rank dictionary (K4) serves the LF walk; with direct sampling it is built on first use (lf_bench)
	dict.release(); D = DevDict();
	if (!p.bwtonly && !direct) make_dict(exc_pos, 0, 0);
	pt.mark(); // 3
	if (!p.bwtonly && !direct)
		k7_walk(st, D, prerank.get(), npre, prerate, T.n, p.sasamplingrate, p.isasamplingrate, sa.get(), isa.get(), &walkstats, 0, npre);
	pt.mark(); // 4
	B3M_CUDA(cudaStreamSynchronize(st.s));
	if (numblocks == 1) { ms_sort = pt.ms(0, 1); ms_extract = pt.ms(1, 2); }
	ms_dict = pt.ms(2, 3);
	ms_walk = pt.ms(3, 4);
	ms_total = pt.ms(0, 4);
	have_results = true;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU, suffix-range sharding: every rank holds the whole text and sorts ONE key range of
// the suffixes; slices of BWT / anchors / samples land in caller-owned, zero-initialised buffers
// at their global places, so that the ranks' buffers combine by a sum (bwtb3m_b200/multigpu.py).
// ------------------------------------------------------------------------------------------
void Engine::kr_build_part(uint32_t part, uint32_t nparts, b3m_build_params const & p, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa,
                           void * d_special, uint64_t * unresolved) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no input loaded");
	B3M_REQUIRE(d_bwt && d_prerank && d_special && unresolved, "null argument");
	B3M_REQUIRE(part < nparts, "bad part index");
	auto pow2 = [](uint64_t v) { return v && !(v & (v - 1)); };
	B3M_REQUIRE(pow2(p.sasamplingrate) && pow2(p.isasamplingrate), "sampling rates must be powers of two");
	B3M_REQUIRE(p.bwtonly || (d_sa && d_isa), "null sample buffers");
	B3M_REQUIRE(p.sortpath >= B3M_SORT_AUTO && p.sortpath <= B3M_SORT_MSD, "unknown sortpath");
	reset_results();
	params = p;
	if (st.sortpath != p.sortpath) kr_plan = KeyRangePlan(); // the plan depends on the sorter
	st.sortpath = p.sortpath;
	prerate = p.preisarate ? p.preisarate : (p.bwtonly ? 64 : choose_preisarate(T.n, st.sms));
	B3M_REQUIRE(pow2(prerate), "preisarate must be a power of two");
	npre = div_up(T.n, prerate);
	nsa = p.bwtonly ? 0 : div_up(T.n, p.sasamplingrate);
	nisa = p.bwtonly ? 0 : div_up(T.n, p.isasamplingrate);
	numblocks = nparts;
	sortstats = SortStats(); walkstats = WalkStats();
	gap_lf_steps = gap_chains = merge_bytes = extract_bytes = 0; max_lcpnext = large_lcp_blocks = 0;
	ms_sort = ms_extract = ms_dict = ms_gap = ms_merge = ms_walk = ms_total = 0;
	arena.ensure_slab(1, (size_t)work_bytes(1, div_up(T.n, nparts) + T.n / 16, false));
	PhaseTimer pt(st);
	pt.mark();
	int const circular = T.has_term ? 0 : 1;
	if (kr_plan.nparts != nparts) k2_keyrange_plan(st, T, circular, nparts, kr_plan);
	FusedOut fo;
	fo.bwt = (uint8_t *)d_bwt; fo.shift = T.has_term ? 1 : 0; fo.has_term = T.has_term; fo.special = (uint32_t *)d_special;
	fo.prerank = (uint32_t *)d_prerank; fo.prelog = ceil_log2_u64(prerate);
	if (!p.bwtonly) {
		fo.sa_s = (unsigned long long *)d_sa; fo.salog = ceil_log2_u64(p.sasamplingrate);
		fo.isa_s = (unsigned long long *)d_isa; fo.isalog = ceil_log2_u64(p.isasamplingrate);
	}
	*unresolved = k2_sort_keyrange(st, T, circular, kr_plan, part, fo, &sortstats);
	if (part == 0 && T.has_term) {
		// rank 0 is the terminator suffix (text position ntext): its predecessor is the last base; its
		// anchor / ISA entries are rank 0 (written explicitly: the buffers may be another GPU's, not zeroed)
		B3M_CUDA(cudaMemcpyAsync(d_bwt, lastcode.get(), 1, cudaMemcpyDeviceToDevice, st.s));
		uint64_t const pos = T.ntext;
		if ((T.ntext & (prerate - 1)) == 0) B3M_CUDA(cudaMemsetAsync((uint32_t *)d_prerank + T.ntext / prerate, 0, 4, st.s));
		if (!p.bwtonly) {
			B3M_CUDA(cudaMemcpyAsync(d_sa, &pos, 8, cudaMemcpyHostToDevice, st.s));
			if ((T.ntext & (p.isasamplingrate - 1)) == 0) B3M_CUDA(cudaMemsetAsync((uint64_t *)d_isa + T.ntext / p.isasamplingrate, 0, 8, st.s));
		}
	}
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_sort = pt.ms(0, 1); ms_total = ms_sort;
	extract_bytes = (kr_plan.base[part + 1] - kr_plan.base[part]) * 6;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU, position sharding of level 1 (XShard): count -> [caller exchanges the counts] -> scatter (peer
// stores) -> [caller: barrier] -> finish.  Outputs as in kr_build_part.
// ------------------------------------------------------------------------------------------
void Engine::prepare_shard_params(b3m_build_params const & p, uint32_t nparts) {
	auto pow2 = [](uint64_t v) { return v && !(v & (v - 1)); };
	B3M_REQUIRE(pow2(p.sasamplingrate) && pow2(p.isasamplingrate), "sampling rates must be powers of two");
	B3M_REQUIRE(p.sortpath >= B3M_SORT_AUTO && p.sortpath <= B3M_SORT_MSD, "unknown sortpath");
	reset_results();
	params = p;
	if (st.sortpath != p.sortpath) kr_plan = KeyRangePlan();
	st.sortpath = p.sortpath;
	prerate = p.preisarate ? p.preisarate : (p.bwtonly ? 64 : choose_preisarate(T.n, st.sms));
	B3M_REQUIRE(pow2(prerate), "preisarate must be a power of two");
	npre = div_up(T.n, prerate);
	nsa = p.bwtonly ? 0 : div_up(T.n, p.sasamplingrate);
	nisa = p.bwtonly ? 0 : div_up(T.n, p.isasamplingrate);
	numblocks = nparts;
	sortstats = SortStats(); walkstats = WalkStats();
	gap_lf_steps = gap_chains = merge_bytes = extract_bytes = 0; max_lcpnext = large_lcp_blocks = 0;
	ms_sort = ms_extract = ms_dict = ms_gap = ms_merge = ms_walk = ms_total = 0;
}

void Engine::xs_count(uint32_t part, uint32_t nparts, b3m_build_params const & p, void * d_totals, uint32_t * nbins) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no input loaded");
	B3M_REQUIRE(d_totals && nbins, "null argument");
	prepare_shard_params(p, nparts);
	arena.ensure_slab(1, (size_t)work_bytes(1, div_up(T.n, nparts) + T.n / 16, false));
	delete xs_pt;
	xs_pt = new PhaseTimer(st);
	xs_pt->mark();
	k2_xshard_count(st, T, T.has_term ? 0 : 1, part, nparts, xs, (unsigned long long *)d_totals, nbins);
}

void Engine::xs_scatter(const uint64_t * h_alltot, void * const * d_recs, const uint64_t * caps) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded && xs.nparts && xs_pt, "xshard_count was not called");
	B3M_REQUIRE(h_alltot && d_recs && caps, "null argument");
	k2_xshard_scatter(st, T, T.has_term ? 0 : 1, xs, (const unsigned long long *)h_alltot, (unsigned long long * const *)d_recs, caps, &sortstats);
}

void Engine::xs_stream_sa(void * d_sa_local, uint64_t * host_sa) {
	xs_sa_local = (unsigned long long *)d_sa_local;
	xs_sa_host = (unsigned long long *)host_sa;
}

void Engine::xs_finish(void * d_recs_own, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa, void * d_special, uint64_t * unresolved) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded && xs.nparts && xs_pt, "xshard_count was not called");
	B3M_REQUIRE(d_recs_own && d_bwt && d_prerank && d_special && unresolved, "null argument");
	b3m_build_params const & p = params;
	B3M_REQUIRE(p.bwtonly || (d_sa && d_isa), "null sample buffers");
	FusedOut fo;
	fo.bwt = (uint8_t *)d_bwt; fo.shift = T.has_term ? 1 : 0; fo.has_term = T.has_term; fo.special = (uint32_t *)d_special;
	fo.prerank = (uint32_t *)d_prerank; fo.prelog = ceil_log2_u64(prerate);
	if (!p.bwtonly) {
		fo.sa_s = (unsigned long long *)d_sa; fo.salog = ceil_log2_u64(p.sasamplingrate);
		fo.isa_s = (unsigned long long *)d_isa; fo.isalog = ceil_log2_u64(p.isasamplingrate);
	}
	StreamOut so;
	bool const stream = !p.bwtonly && xs_sa_local && xs_sa_host;
	if (stream) { fo.sa_s2 = xs_sa_local; so.host_sa = xs_sa_host; so.nsa = nsa; }
	xs_sa_local = nullptr; xs_sa_host = nullptr; // one build
	*unresolved = k2_xshard_finish(st, T, T.has_term ? 0 : 1, xs, (unsigned long long *)d_recs_own, fo, &sortstats, stream ? &so : nullptr);
	if (stream) B3M_CUDA(cudaStreamSynchronize(st.copy)); // this rank's samples are in the host buffer (all but sample 0 of a terminated text)
	xs_sa_delivered = stream && so.delivered;
	if (xs.part == 0 && T.has_term) {
		// rank 0 is the terminator suffix (text position ntext): written explicitly, the buffers are not zeroed
		B3M_CUDA(cudaMemcpyAsync(d_bwt, lastcode.get(), 1, cudaMemcpyDeviceToDevice, st.s));
		uint64_t const pos = T.ntext;
		if ((T.ntext & (prerate - 1)) == 0) B3M_CUDA(cudaMemsetAsync((uint32_t *)d_prerank + T.ntext / prerate, 0, 4, st.s));
		if (!p.bwtonly) {
			B3M_CUDA(cudaMemcpyAsync(d_sa, &pos, 8, cudaMemcpyHostToDevice, st.s));
			if ((T.ntext & (p.isasamplingrate - 1)) == 0) B3M_CUDA(cudaMemsetAsync((uint64_t *)d_isa + T.ntext / p.isasamplingrate, 0, 8, st.s));
		}
	}
	// kr_rows / shard_adopt describe the result by the same plan structure
	kr_plan = KeyRangePlan();
	kr_plan.nparts = xs.nparts;
	kr_plan.bin_lo.assign(xs.bnd.begin(), xs.bnd.end());
	kr_plan.base.assign(xs.nparts + 1, T.ntext);
	{
		uint64_t acc = 0;
		for (uint32_t q = 0, b = 0; q <= xs.nparts; ++q) {
			for (; b < (q < xs.nparts ? xs.bnd[q] : (uint32_t)xs.total.size()); ++b) acc += xs.total[b];
			kr_plan.base[q] = acc;
		}
	}
	xs_pt->mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_sort = xs_pt->ms(0, 1); ms_total = ms_sort;
	extract_bytes = xs.records * 6;
}

// first BWT row of every key range (rows of range p: [first[p], first[p+1])); range 0 also owns the terminator row
void Engine::kr_rows(uint32_t nparts, uint64_t * first) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no input loaded");
	B3M_REQUIRE(first, "null argument");
	if (kr_plan.nparts != nparts) k2_keyrange_plan(st, T, T.has_term ? 0 : 1, nparts, kr_plan);
	uint64_t const shift = T.has_term ? 1 : 0;
	for (uint32_t p = 0; p <= nparts; ++p) first[p] = kr_plan.base[p] + shift;
	first[0] = 0; first[nparts] = T.n;
}

// adopt: the engine refers to the caller's buffers instead of copying them (they must stay valid until the
// next load / build; the arena ignores pointers it does not own when the results are released)
void Engine::kr_finish(const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa, const void * d_special, uint32_t nparts, bool adopt) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded && npre, "kr_build_part was not called");
	PhaseTimer pt(st);
	pt.mark();
	if (adopt) {
		bwt.release(); prerank.release(); sa.release(); isa.release();
		bwt.p = (uint8_t *)const_cast<void *>(d_bwt); bwt.n = T.n + 16; bwt.a = &arena;
		prerank.p = (uint32_t *)const_cast<void *>(d_prerank); prerank.n = npre; prerank.a = &arena;
		if (nsa) { sa.p = (uint64_t *)const_cast<void *>(d_sa); sa.n = nsa; sa.a = &arena; }
		if (nisa) { isa.p = (uint64_t *)const_cast<void *>(d_isa); isa.n = nisa; isa.a = &arena; }
	} else {
	bwt.alloc(st, T.n + 16);
	B3M_CUDA(cudaMemcpyAsync(bwt.get(), d_bwt, T.n, cudaMemcpyDeviceToDevice, st.s));
	prerank.alloc(st, npre);
	B3M_CUDA(cudaMemcpyAsync(prerank.get(), d_prerank, 4 * npre, cudaMemcpyDeviceToDevice, st.s));
	if (nsa) { sa.alloc(st, nsa); B3M_CUDA(cudaMemcpyAsync(sa.get(), d_sa, 8 * nsa, cudaMemcpyDeviceToDevice, st.s)); }
	if (nisa) { isa.alloc(st, nisa); B3M_CUDA(cudaMemcpyAsync(isa.get(), d_isa, 8 * nisa, cudaMemcpyDeviceToDevice, st.s)); }
	}
	if (!this->d_special.get()) this->d_special.alloc(st, 8); // scratch words of the output stages (K8 counts its runs there)
	uint32_t exc_pos = 0xffffffffu;
	if (T.has_term) {
		B3M_CUDA(cudaMemcpyAsync(pinned, d_special, 16, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
		exc_pos = ((uint32_t *)pinned)[0];
	}
	root_exc_pos = exc_pos;
	numblocks = nparts;
	dict.release(); D = DevDict();
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_dict = pt.ms(0, 1);
	have_results = true;
}

void Engine::fetch(uint8_t * h_bwt, uint64_t * h_pairs, uint64_t * h_sa, uint64_t * h_isa) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(have_results, "no results");
	B3M_REQUIRE(!(ssa_only && (h_bwt || h_pairs)), "the engine holds sampled SA/ISA only (computed from an existing BWT)");
	if (h_bwt) {
		DevBuf<uint8_t> out;
		symbols_device(out);
		B3M_CUDA(cudaMemcpyAsync(h_bwt, out.get(), T.n, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	if (h_pairs) {
		DevBuf<uint64_t> pairs(st, 2 * npre);
		B3M_LAUNCH(st, k_pairs, (unsigned)div_up(npre, 256), 256, 0, (const uint32_t *)prerank.get(), npre, prerate,
		           (unsigned long long *)pairs.get());
		B3M_CUDA(cudaMemcpyAsync(h_pairs, pairs.get(), 16 * npre, cudaMemcpyDeviceToHost, st.s));
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	if (h_sa && nsa && h_sa != sa_on_host) B3M_CUDA(cudaMemcpyAsync(h_sa, sa.get(), 8 * nsa, cudaMemcpyDeviceToHost, st.s));
	if (h_isa && nisa) B3M_CUDA(cudaMemcpyAsync(h_isa, isa.get(), 8 * nisa, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
}

void Engine::info(b3m_info * o) {
	memset(o, 0, sizeof(*o));
	o->n = T.n; o->sigma = T.sigma + (T.has_term ? 1 : 0); o->numblocks = numblocks;
	o->preisarate = prerate; o->npreisa = npre;
	o->sasamplingrate = params.sasamplingrate; o->nsa = nsa;
	o->isasamplingrate = params.isasamplingrate; o->nisa = nisa;
	memcpy(o->hist, hist, sizeof(hist));
	o->sort_rounds = sortstats.rounds; o->radix_passes = sortstats.radix_passes; o->radix_bytes = sortstats.radix_bytes;
	o->sort_active_sum = sortstats.active_sum; o->sort_other_bytes = sortstats.other_bytes;
	o->gap_lf_steps = gap_lf_steps; o->walk_lf_steps = walkstats.steps; o->walk_chains = walkstats.chains; o->gap_chains = gap_chains;
	o->merge_bytes = merge_bytes; o->extract_bytes = extract_bytes; o->dict_bytes = dict_bytes_moved; o->decode_bytes = decode_bytes;
	o->launches = st.launches; o->max_lcpnext = max_lcpnext;
	o->sort_tied0 = sortstats.tied0; o->sort_unresolved0 = sortstats.unresolved0;
	o->arena_capacity = arena.capacity; o->arena_peak = arena.peak;
	o->ms_decode = ms_decode; o->ms_sort = ms_sort; o->ms_extract = ms_extract; o->ms_dict = ms_dict;
	o->ms_gap = ms_gap; o->ms_merge = ms_merge; o->ms_walk = ms_walk; o->ms_total = ms_total;
}

// ------------------------------------------------------------------------------------------
// bwtcomputessa path: sampled SA/ISA from an existing BWT and (rank,pos) anchors
// (replaces BwtComputeSSA::computeSSA, /root/reference/src/bwtcomputessa.cpp:51; the walk is
//  /root/reference/src/hwtPreIsaToIsa.cpp:79,114-161: anchors sorted by position, each walks
//  back to its predecessor)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_find_byte(const uint8_t * __restrict__ s, uint64_t n, uint32_t v, uint32_t * __restrict__ out) {
	uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && s[i] == v) *out = (uint32_t)i;
}

// K4 on a BWT that arrives in the reference's symbol space (a .bwt file): dense codes + rank dictionary
void Engine::install_bwt_symbols(const uint8_t * h_bwt, uint64_t n, uint64_t extra_bytes) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(n > 0 && n < 0xFFFFFFF0ull, "BWT length out of range");
	reset_results();
	codes.release(); packed.release(); raw.release(); d_hist.release(); d_special.release();
	loaded = false;
	arena.reserve((size_t)(n * 4 + extra_bytes + (64u << 20)));
	raw.alloc(st, n + 16);
	B3M_CUDA(cudaMemcpyAsync(raw.get(), h_bwt, n, cudaMemcpyHostToDevice, st.s));
	d_hist.alloc(st, 256);
	d_special.alloc(st, 8);
	k1_hist_bytes(st, raw.get(), n, d_hist.get());
	B3M_CUDA(cudaMemcpyAsync(pinned, d_hist.get(), 256 * 8, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	memcpy(hist, pinned, sizeof(hist));
	// dense codes; a symbol that occurs once (pacterm's terminator) is kept out of a 4-symbol
	// alphabet so that the 2-bit dictionary can be used
	int distinct = 0, unique_sym = -1;
	for (int c = 0; c < 256; ++c) if (hist[c]) { ++distinct; if (hist[c] == 1 && unique_sym < 0) unique_sym = c; }
	bool const use_exc = distinct == 5 && unique_sym >= 0;
	uint8_t lut[256];
	memset(lut, 0, sizeof(lut));
	uint32_t sigma = 0;
	uint64_t acc = 0, csym[257];
	for (int c = 0; c < 256; ++c) { csym[c] = acc; acc += hist[c]; }
	D = DevDict();
	for (int c = 0; c < 256; ++c) if (hist[c] && !(use_exc && c == unique_sym)) { lut[c] = (uint8_t)sigma; code2sym[sigma] = (uint8_t)c; D.C[sigma] = (uint32_t)csym[c]; ++sigma; }
	for (uint32_t c = sigma; c < 257; ++c) D.C[c] = (uint32_t)n;
	uint32_t exc_pos = 0xffffffffu;
	if (use_exc) {
		B3M_CUDA(cudaMemsetAsync(d_special.get(), 0xff, 4, st.s));
		B3M_LAUNCH(st, k_find_byte, (unsigned)div_up(n, 256), 256, 0, (const uint8_t *)raw.get(), n, (uint32_t)unique_sym, d_special.get());
		exc_pos = fetch_special(0);
		B3M_REQUIRE(exc_pos != 0xffffffffu, "internal: unique symbol not found");
	}
	T = DevText();
	T.n = n; T.ntext = n; T.sigma = sigma; T.has_term = 0; T.keybits = 8;
	bwt.alloc(st, n + 16);
	{
		memcpy(pinned, lut, 256);
		DevBuf<uint8_t> dlut(st, 256);
		B3M_CUDA(cudaMemcpyAsync(dlut.get(), pinned, 256, cudaMemcpyHostToDevice, st.s));
		k1_map_bytes(st, raw.get(), n, dlut.get(), bwt.get());
		B3M_CUDA(cudaStreamSynchronize(st.s));
	}
	raw.release();
	int const flavour = sigma <= 4 ? 2 : 8;
	size_t const bytes = dict_bytes(flavour, n, sigma);
	dict.alloc(st, bytes);
	k4_build_dict(st, bwt.get(), n, flavour, sigma, dict.get());
	D.flavour = flavour; D.n = n; D.lines = dict.get(); D.sigma = sigma;
	D.exc_pos = exc_pos; D.exc_code = 0; D.exc_lf = use_exc ? (uint32_t)csym[unique_sym] : 0;
	dict_bytes_moved = n + bytes;
	root_exc_pos = 0xffffffffu;
	npre = 0; prerate = 0; numblocks = 0;
	ssa_only = true;
}

void Engine::ssa_from_bwt(const uint8_t * h_bwt, uint64_t n, const uint64_t * h_pairs, uint64_t npairs, uint64_t sarate, uint64_t isarate) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(npairs > 0, "no (rank,pos) anchors: the .preisa file is empty");
	auto pow2 = [](uint64_t v) { return v && !(v & (v - 1)); };
	B3M_REQUIRE(pow2(sarate) && pow2(isarate), "sampling rates must be powers of two");
	PhaseTimer pt(st);
	pt.mark();
	install_bwt_symbols(h_bwt, n, 16 * npairs);
	pt.mark();
	// anchors sorted by position; each walks back to its predecessor
	std::vector<std::pair<uint64_t, uint64_t>> A(npairs);
	for (uint64_t k = 0; k < npairs; ++k) {
		A[k] = std::make_pair(h_pairs[2 * k + 1], h_pairs[2 * k]);
		B3M_REQUIRE(A[k].first < n && A[k].second < n, "anchor out of range in .preisa");
	}
	std::sort(A.begin(), A.end());
	std::vector<uint32_t> ar(npairs);
	std::vector<uint64_t> ap(npairs), as(npairs);
	for (uint64_t k = 0; k < npairs; ++k) {
		uint64_t const prev = A[(k + npairs - 1) % npairs].first;
		uint64_t todo = (A[k].first + n - prev) % n;
		if (todo == 0) todo = npairs == 1 ? n : 0;
		ar[k] = (uint32_t)A[k].second; ap[k] = A[k].first; as[k] = todo;
	}
	DevBuf<uint32_t> dar(st, npairs);
	DevBuf<uint64_t> dap(st, npairs), das(st, npairs);
	B3M_CUDA(cudaMemcpyAsync(dar.get(), ar.data(), 4 * npairs, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaMemcpyAsync(dap.get(), ap.data(), 8 * npairs, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaMemcpyAsync(das.get(), as.data(), 8 * npairs, cudaMemcpyHostToDevice, st.s));
	params = b3m_build_params();
	params.sasamplingrate = sarate; params.isasamplingrate = isarate;
	nsa = div_up(n, sarate); nisa = div_up(n, isarate);
	sa.alloc(st, nsa); isa.alloc(st, nisa);
	B3M_LAUNCH(st, k_fill_u64, (unsigned)div_up(nsa, 256), 256, 0, (unsigned long long *)sa.get(), nsa, ~0ull);
	B3M_LAUNCH(st, k_fill_u64, (unsigned)div_up(nisa, 256), 256, 0, (unsigned long long *)isa.get(), nisa, ~0ull);
	walkstats = WalkStats();
	k7_walk_anchors(st, D, dar.get(), dap.get(), das.get(), npairs, n, sarate, isarate, sa.get(), isa.get(), &walkstats);
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_dict = pt.ms(0, 1); ms_walk = pt.ms(1, 2); ms_total = pt.ms(0, 2);
	ms_decode = ms_sort = ms_extract = ms_gap = ms_merge = 0;
	npre = 0; prerate = 0; numblocks = 0;
	root_exc_pos = 0xffffffffu;
	ssa_only = true;
	have_results = true;
}

// ------------------------------------------------------------------------------------------
// checkbwt (/root/reference/src/checkbwt.cpp:26-246) on the device: the text has been loaded
// (K1); the BWT comes in the reference's symbol space, the anchors as (rank,pos) pairs.  Every
// text position is compared exactly once; the symbol counts must agree as well.
// ------------------------------------------------------------------------------------------
uint64_t Engine::check_bwt(const uint8_t * h_bwt, uint64_t n, const uint64_t * h_pairs, uint64_t npairs, uint64_t * badrank) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(loaded, "no text loaded");
	B3M_REQUIRE(npairs > 0, "no (rank,pos) anchors: the .preisa file is empty");
	if (badrank) *badrank = ~0ull;
	if (n != T.n) return n > T.n ? n : T.n; // lengths differ: nothing matches
	reset_results();
	// symbols -> the text's codes; the terminator of a terminated text is the exception row
	int lut[256];
	for (int c = 0; c < 256; ++c) lut[c] = -1;
	for (uint32_t c = 0; c < T.sigma; ++c) lut[code2sym[c]] = (int)c;
	std::vector<uint8_t> hc(n);
	uint64_t cnt[256];
	memset(cnt, 0, sizeof(cnt));
	uint64_t exc = ~0ull, nterm = 0, unknown = 0;
	for (uint64_t i = 0; i < n; ++i) {
		uint8_t const sym = h_bwt[i];
		if (T.has_term && sym == 0) { exc = i; ++nterm; hc[i] = 0; }
		else if (lut[sym] < 0) { ++unknown; hc[i] = 0; }
		else { hc[i] = (uint8_t)lut[sym]; cnt[lut[sym]]++; }
	}
	uint64_t histdiff = unknown + (T.has_term ? (nterm > 1 ? nterm - 1 : 1 - nterm) : 0);
	for (uint32_t c = 0; c < T.sigma; ++c) histdiff += cnt[c] > codehist[c] ? cnt[c] - codehist[c] : codehist[c] - cnt[c];
	if (histdiff) return histdiff; // a walk over a BWT with other symbol counts proves nothing more
	bwt.alloc(st, n + 16);
	B3M_CUDA(cudaMemcpyAsync(bwt.get(), hc.data(), n, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	make_dict(T.has_term ? (uint32_t)exc : 0xffffffffu, 0, 0);
	std::vector<std::pair<uint64_t, uint64_t>> A(npairs);
	for (uint64_t k = 0; k < npairs; ++k) {
		A[k] = std::make_pair(h_pairs[2 * k + 1], h_pairs[2 * k]);
		B3M_REQUIRE(A[k].first < n && A[k].second < n, "anchor out of range in .preisa");
	}
	std::sort(A.begin(), A.end());
	std::vector<uint32_t> ar(npairs);
	std::vector<uint64_t> ap(npairs), as(npairs);
	for (uint64_t k = 0; k < npairs; ++k) {
		uint64_t const prev = A[(k + npairs - 1) % npairs].first;
		uint64_t todo = (A[k].first + n - prev) % n;
		if (todo == 0) todo = npairs == 1 ? n : 0;
		ar[k] = (uint32_t)A[k].second; ap[k] = A[k].first; as[k] = todo;
	}
	DevBuf<uint32_t> dar(st, npairs);
	DevBuf<uint64_t> dap(st, npairs), das(st, npairs), dres(st, 2);
	B3M_CUDA(cudaMemcpyAsync(dar.get(), ar.data(), 4 * npairs, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaMemcpyAsync(dap.get(), ap.data(), 8 * npairs, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaMemcpyAsync(das.get(), as.data(), 8 * npairs, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaMemsetAsync(dres.get(), 0, 16, st.s));
	PhaseTimer pt(st);
	pt.mark();
	ensure_codes();
	k7_check_walk(st, D, T.codes, T.ntext, T.has_term, dar.get(), dap.get(), das.get(), npairs, n, dres.get());
	pt.mark();
	uint64_t res[2];
	B3M_CUDA(cudaMemcpyAsync(res, dres.get(), 16, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	ms_walk = pt.ms(0, 1);
	walkstats = WalkStats(); walkstats.steps = n; walkstats.chains = npairs;
	if (badrank && res[0]) *badrank = res[1];
	bwt.release(); dict.release(); D = DevDict();
	return res[0];
}

__global__ void __launch_bounds__(256)
k_pick_starts(const uint32_t * __restrict__ prerank, uint64_t npre, uint64_t nchains, uint32_t * __restrict__ start) {
	uint64_t const q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nchains) return;
	start[q] = prerank[(q * npre) / nchains];
}

// LF-steps/s of the dictionary in place, chains started at the given ranks (bwttestdecodespeed: evenly spaced .isa samples)
void Engine::lf_speed(const uint64_t * h_start, uint64_t nstart, uint64_t nchains, uint64_t steps, float * ms) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(D.lines, "no dictionary");
	B3M_REQUIRE(nstart >= 1 && nchains >= 1, "need start ranks and at least one chain");
	std::vector<uint32_t> h(nchains);
	for (uint64_t c = 0; c < nchains; ++c) h[c] = (uint32_t)h_start[(c * nstart) / nchains];
	DevBuf<uint32_t> start(st, nchains), out(st, nchains);
	B3M_CUDA(cudaMemcpyAsync(start.get(), h.data(), 4 * nchains, cudaMemcpyHostToDevice, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	PhaseTimer pt(st);
	pt.mark();
	k7_lfbench(st, D, start.get(), nchains, steps, out.get());
	pt.mark();
	B3M_CUDA(cudaStreamSynchronize(st.s));
	*ms = pt.ms(0, 1);
}

void Engine::lf_bench(uint64_t nchains, uint64_t steps, float * ms, uint64_t * checksum) {
	B3M_CUDA(cudaSetDevice(device));
	B3M_REQUIRE(have_results, "no results");
	B3M_REQUIRE(nchains >= 1, "nchains must be >= 1");
	B3M_REQUIRE(!ssa_only || D.lines, "no dictionary");
	if (!D.lines) make_dict(root_exc_pos, 0, 0);
	DevBuf<uint32_t> start(st, nchains), out(st, nchains);
	B3M_LAUNCH(st, k_pick_starts, (unsigned)div_up(nchains, 256), 256, 0, (const uint32_t *)prerank.get(), npre, nchains, start.get());
	PhaseTimer pt(st);
	pt.mark();
	k7_lfbench(st, D, start.get(), nchains, steps, out.get());
	pt.mark();
	std::vector<uint32_t> h(nchains);
	B3M_CUDA(cudaMemcpyAsync(h.data(), out.get(), 4 * nchains, cudaMemcpyDeviceToHost, st.s));
	B3M_CUDA(cudaStreamSynchronize(st.s));
	*ms = pt.ms(0, 1);
	uint64_t cs = 0;
	for (uint64_t i = 0; i < nchains; ++i) cs = cs * 1000003ull + h[i];
	if (checksum) *checksum = cs;
}

} // namespace b3m

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using b3m::Engine;


static void set_err(char * err, size_t errlen, const char * msg) {
	if (err && errlen) { strncpy(err, msg, errlen - 1); err[errlen - 1] = 0; }
}

#define B3M_GUARD(h, ...)                                                    \
	if (!(h)) return 1;                                                      \
	try { __VA_ARGS__; (h)->err.clear(); return 0; }                                \
	catch (std::exception const & ex) { (h)->err = ex.what(); return 2; }    \
	catch (...) { (h)->err = "unknown error"; return 3; }

extern "C" {

const char * b3m_version(void) { return "b3m-b200 0.1 (sm_100a)"; }

int b3m_parse_inputtype(const char * name) {
	if (!name) return -1;
	if (!strcmp(name, "bytestream")) return B3M_INPUT_BYTESTREAM;
	if (!strcmp(name, "compactstream")) return B3M_INPUT_COMPACTSTREAM;
	if (!strcmp(name, "pac")) return B3M_INPUT_PAC;
	if (!strcmp(name, "pacterm")) return B3M_INPUT_PACTERM;
	return -1;
}

int b3m_engine_create(int device, void * cuda_stream, b3m_engine ** out, char * err, size_t errlen) {
	if (!out) return 1;
	*out = nullptr;
	try {
		int ndev = 0;
		cudaError_t const ce = cudaGetDeviceCount(&ndev);
		if (ce != cudaSuccess || ndev <= 0)
			throw b3m::Error(std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); this library has no CPU fallback");
		if (device < 0 || device >= ndev) throw b3m::Error("bad device ordinal");
		b3m_engine * h = new b3m_engine();
		h->e = new Engine(device, cuda_stream);
		*out = h;
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

void b3m_engine_destroy(b3m_engine * h) {
	if (!h) return;
	delete h->e;
	delete h;
}

const char * b3m_engine_last_error(const b3m_engine * h) { return h ? h->err.c_str() : "null engine"; }

int b3m_engine_load_host(b3m_engine * h, const void * input, uint64_t nbytes, int inputtype) {
	B3M_GUARD(h, h->e->load(input, nbytes, inputtype, false));
}
int b3m_engine_load_device(b3m_engine * h, const void * d_input, uint64_t nbytes, int inputtype) {
	B3M_GUARD(h, h->e->load(d_input, nbytes, inputtype, true));
}
int b3m_engine_build(b3m_engine * h, const b3m_build_params * p) {
	B3M_GUARD(h, { if (!p) throw b3m::Error("null params"); h->e->build(*p); });
}
int b3m_engine_info(b3m_engine * h, b3m_info * info) {
	B3M_GUARD(h, { if (!info) throw b3m::Error("null info"); h->e->info(info); });
}
int b3m_engine_fetch(b3m_engine * h, uint8_t * bwt, uint64_t * preisa_pairs, uint64_t * sa, uint64_t * isa) {
	B3M_GUARD(h, h->e->fetch(bwt, preisa_pairs, sa, isa));
}
int b3m_engine_device_results(b3m_engine * h, const void ** d_bwt_codes, const void ** d_preisa_rank,
                              const void ** d_sa, const void ** d_isa) {
	B3M_GUARD(h, {
		if (!h->e->have_results) throw b3m::Error("no results");
		if (d_bwt_codes) *d_bwt_codes = h->e->bwt.get();
		if (d_preisa_rank) *d_preisa_rank = h->e->prerank.get();
		if (d_sa) *d_sa = h->e->sa.get();
		if (d_isa) *d_isa = h->e->isa.get();
	});
}
int b3m_engine_write_bwt(b3m_engine * h, const char * bwtfn) {
	B3M_GUARD(h, { if (!bwtfn) throw b3m::Error("null file name"); h->e->write_bwt(bwtfn); });
}
int b3m_engine_fetch_runs(b3m_engine * h, uint8_t * syms, uint64_t * lens, uint64_t cap, uint64_t * nruns) {
	B3M_GUARD(h, h->e->fetch_runs(syms, lens, cap, nruns));
}
int b3m_engine_shard_build(b3m_engine * h, uint32_t part, uint32_t nparts, const b3m_build_params * p, void * d_bwt, void * d_prerank,
                           void * d_sa, void * d_isa, void * d_special, uint64_t * unresolved) {
	B3M_GUARD(h, { if (!p) throw b3m::Error("null params"); h->e->kr_build_part(part, nparts, *p, d_bwt, d_prerank, d_sa, d_isa, d_special, unresolved); });
}
int b3m_engine_pack_rows(b3m_engine * h, const void * d_rows, uint64_t nrows, void * d_packed) {
	B3M_GUARD(h, h->e->pack_rows(d_rows, nrows, d_packed, false));
}
int b3m_engine_unpack_rows(b3m_engine * h, const void * d_packed, uint64_t nrows, void * d_rows) {
	B3M_GUARD(h, h->e->pack_rows(d_rows, nrows, const_cast<void *>(d_packed), true));
}
int b3m_engine_shard_rows(b3m_engine * h, uint32_t nparts, uint64_t * first_row) {
	B3M_GUARD(h, h->e->kr_rows(nparts, first_row));
}
int b3m_engine_shard_finish(b3m_engine * h, const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa, const void * d_special,
                            uint32_t nparts) {
	B3M_GUARD(h, h->e->kr_finish(d_bwt, d_prerank, d_sa, d_isa, d_special, nparts, false));
}
int b3m_engine_xshard_count(b3m_engine * h, uint32_t part, uint32_t nparts, const b3m_build_params * p, void * d_totals, uint32_t * nbins) {
	B3M_GUARD(h, { if (!p) throw b3m::Error("null params"); h->e->xs_count(part, nparts, *p, d_totals, nbins); });
}
int b3m_engine_xshard_scatter(b3m_engine * h, const uint64_t * all_totals, void * const * d_recs, const uint64_t * caps) {
	B3M_GUARD(h, h->e->xs_scatter(all_totals, d_recs, caps));
}
int b3m_engine_xshard_stream_sa(b3m_engine * h, void * d_sa_local, uint64_t * host_sa) {
	B3M_GUARD(h, h->e->xs_stream_sa(d_sa_local, host_sa));
}
int b3m_engine_xshard_sa_delivered(b3m_engine * h, int * delivered) {
	B3M_GUARD(h, { if (!delivered) throw b3m::Error("null argument"); *delivered = h->e->xs_sa_delivered ? 1 : 0; });
}
int b3m_engine_xshard_finish(b3m_engine * h, void * d_recs_own, void * d_bwt, void * d_prerank, void * d_sa, void * d_isa, void * d_special, uint64_t * unresolved) {
	B3M_GUARD(h, h->e->xs_finish(d_recs_own, d_bwt, d_prerank, d_sa, d_isa, d_special, unresolved));
}
int b3m_engine_shard_adopt(b3m_engine * h, const void * d_bwt, const void * d_prerank, const void * d_sa, const void * d_isa, const void * d_special,
                           uint32_t nparts) {
	B3M_GUARD(h, h->e->kr_finish(d_bwt, d_prerank, d_sa, d_isa, d_special, nparts, true));
}
int b3m_engine_fetch_bwa(b3m_engine * h, uint32_t * bwt_words, uint64_t cap_words, uint64_t * primary, uint64_t * L2, uint64_t * seq_len) {
	B3M_GUARD(h, h->e->fetch_bwa(bwt_words, cap_words, primary, L2, seq_len));
}
int b3m_engine_ssa_from_bwt(b3m_engine * h, const uint8_t * bwt, uint64_t n, const uint64_t * pairs, uint64_t npairs, uint64_t sarate, uint64_t isarate) {
	B3M_GUARD(h, { if (!bwt || !pairs) throw b3m::Error("null argument"); h->e->ssa_from_bwt(bwt, n, pairs, npairs, sarate, isarate); });
}
int b3m_engine_lf_bench(b3m_engine * h, uint64_t nchains, uint64_t steps, float * ms, uint64_t * checksum) {
	B3M_GUARD(h, { float t = 0; h->e->lf_bench(nchains, steps, &t, checksum); if (ms) *ms = t; });
}
int b3m_engine_set_profile(b3m_engine * h, int on) {
	B3M_GUARD(h, { h->e->st.kt.clear(); h->e->st.kt.on = on != 0; });
}
int b3m_engine_kernel_times(b3m_engine * h, char * buf, size_t buflen) {
	B3M_GUARD(h, {
		B3M_CUDA(cudaSetDevice(h->e->device));
		B3M_CUDA(cudaStreamSynchronize(h->e->st.s));
		std::map<std::string, std::tuple<uint64_t, double, uint64_t>> agg;
		for (auto & r : h->e->st.kt.recs) {
			float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
			auto & t = agg[r.name];
			std::get<0>(t) += 1; std::get<1>(t) += ms; std::get<2>(t) += r.bytes;
		}
		h->e->st.kt.clear();
		std::string out;
		for (auto & kv : agg) {
			char line[256];
			snprintf(line, sizeof(line), "%s %llu %.6f %llu\n", kv.first.c_str(), (unsigned long long)std::get<0>(kv.second),
			         std::get<1>(kv.second), (unsigned long long)std::get<2>(kv.second));
			out += line;
		}
		if (buf && buflen) { strncpy(buf, out.c_str(), buflen - 1); buf[buflen - 1] = 0; }
	});
}
// ---- device buffers shared between the processes of a multi-GPU build (CUDA IPC) ----
#define B3M_PLAIN(...)                                                                  \
	try { __VA_ARGS__; return 0; }                                                      \
	catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }    \
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }

int b3m_dev_alloc(int device, uint64_t bytes, void ** dptr, char * err, size_t errlen) {
	B3M_PLAIN({
		if (!dptr) throw b3m::Error("null argument");
		B3M_CUDA(cudaSetDevice(device));
		B3M_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
	});
}
int b3m_dev_free(int device, void * dptr, char * err, size_t errlen) {
	B3M_PLAIN({ B3M_CUDA(cudaSetDevice(device)); B3M_CUDA(cudaFree(dptr)); });
}
int b3m_ipc_export(int device, const void * dptr, void * handle64, char * err, size_t errlen) {
	B3M_PLAIN({
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
		if (!dptr || !handle64) throw b3m::Error("null argument");
		B3M_CUDA(cudaSetDevice(device));
		cudaIpcMemHandle_t hd;
		B3M_CUDA(cudaIpcGetMemHandle(&hd, const_cast<void *>(dptr)));
		memcpy(handle64, &hd, 64);
	});
}
int b3m_ipc_open(int device, const void * handle64, void ** dptr, char * err, size_t errlen) {
	B3M_PLAIN({
		if (!dptr || !handle64) throw b3m::Error("null argument");
		B3M_CUDA(cudaSetDevice(device));
		cudaIpcMemHandle_t hd;
		memcpy(&hd, handle64, 64);
		B3M_CUDA(cudaIpcOpenMemHandle(dptr, hd, cudaIpcMemLazyEnablePeerAccess));
	});
}
int b3m_ipc_close(int device, void * dptr, char * err, size_t errlen) {
	B3M_PLAIN({ B3M_CUDA(cudaSetDevice(device)); B3M_CUDA(cudaIpcCloseMemHandle(dptr)); });
}

int b3m_dev_copy(int device, void * dst, const void * src, uint64_t bytes, void * cuda_stream, char * err, size_t errlen) {
	B3M_PLAIN({
		if (bytes && (!dst || !src)) throw b3m::Error("null argument");
		B3M_CUDA(cudaSetDevice(device));
		if (bytes) B3M_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
	});
}
int b3m_engine_pack_bwa(b3m_engine * h, void * d_words, uint64_t w_lo, uint64_t w_hi) {
	B3M_GUARD(h, h->e->pack_bwa_device((uint32_t *)d_words, w_lo, w_hi));
}

int b3m_engine_sync(b3m_engine * h) {
	B3M_GUARD(h, { B3M_CUDA(cudaSetDevice(h->e->device)); B3M_CUDA(cudaStreamSynchronize(h->e->st.s)); });
}

} // extern "C"
