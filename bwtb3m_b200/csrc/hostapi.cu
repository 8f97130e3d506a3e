// File-level C ABI (include/b3m.h): the reference's three library calls on this path,
//   BwtMergeSort::computeBwt              /root/reference/src/bwtb3m.cpp:62-63
//   BwtComputeSSA::computeSSA             /root/reference/src/bwtcomputessa.cpp:39-51
//   MausFmToBwaConversion::rewrite        /root/reference/src/bwtb3mtobwa.cpp:29
// Host orchestration and file formats only; every step of the hot path runs in the CUDA engine.
#include "engine.h"
#include "formats.h"
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <thread>
#include <array>
#include <memory>

using namespace b3m;

namespace {

double now_sec() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void set_err(char * err, size_t errlen, const char * msg) {
	if (err && errlen) { strncpy(err, msg, errlen - 1); err[errlen - 1] = 0; }
}

std::string clip_off(std::string const & s, std::string const & suffix) {
	if (s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0) return s.substr(0, s.size() - suffix.size());
	return s;
}

void copy_name(char * dst, std::string const & s) {
	strncpy(dst, s.c_str(), 1023);
	dst[1023] = 0;
}

// pinned host buffer holding a whole file
struct PinnedFile {
	uint8_t * p = nullptr;
	uint64_t n = 0;
	explicit PinnedFile(std::string const & fn) {
		n = file_size(fn);
		if (cudaMallocHost((void **)&p, n ? n : 1) != cudaSuccess) { cudaGetLastError(); throw Error("cannot allocate pinned host memory for " + fn); }
		FILE * f = fopen(fn.c_str(), "rb");
		if (!f) { cudaFreeHost(p); throw IoError("cannot open " + fn + " for reading"); }
		size_t const got = fread(p, 1, n, f);
		fclose(f);
		if (got != n) { cudaFreeHost(p); throw IoError("short read on " + fn); }
	}
	~PinnedFile() { if (p) cudaFreeHost(p); }
};

void check_device_available() {
	int ndev = 0;
	cudaError_t const ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev <= 0)
		throw Error(std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); this library has no CPU fallback");
}

// bytes of device memory one suffix of a block costs while the block is sorted (text, suffix
// array, ranks, radix ping-pong buffers; DESIGN.md "HBM layout")
constexpr uint64_t BYTES_PER_SUFFIX = 30;

} // namespace

extern "C" {

void b3m_options_init(b3m_options * o) {
	if (!o) return;
	memset(o, 0, sizeof(*o));
	o->inputtype = "bytestream";
	o->sasamplingrate = 32;          // BwtMergeSortOptions::getDefaultSaSamplingRate, README.md:40
	o->isasamplingrate = 262144;     // getDefaultIsaSamplingRate, README.md:41
	o->mem = 0;                      // 0: bounded by the free device memory (the reference's 2 GiB is a host RAM target)
	long const nc = sysconf(_SC_NPROCESSORS_ONLN);
	o->numthreads = nc > 0 ? (uint64_t)nc : 1;
	o->bwtonly = 0;
	o->copyinputtomemory = 0;
	o->largelcpthres = 16384;
	o->verbose = 0;
	o->device = 0;
	o->numblocks = 0;
	o->ngpus = 1;
}

int b3m_compute_bwt(const b3m_options * o, b3m_result * res, char * err, size_t errlen) {
	try {
		if (!o || !res) throw Error("null argument");
		if (!o->fn || !*o->fn) throw Error("no input file name");
		double const t0 = now_sec();
		memset(res, 0, sizeof(*res));
		int const itype = b3m_parse_inputtype(o->inputtype ? o->inputtype : "bytestream");
		if (itype < 0) throw Error(std::string("unknown input type ") + (o->inputtype ? o->inputtype : "(null)") + " (lz4 and utf-8 are not supported by this implementation)");
		std::string const tmpprefix = (o->tmpprefix && *o->tmpprefix) ? o->tmpprefix : "bwtb3m_tmp";
		std::string const bwtfn = (o->outputfilename && *o->outputfilename) ? o->outputfilename : tmpprefix + ".bwt";
		std::string const prefix = clip_off(bwtfn, ".bwt");
		check_device_available();
		if (!file_exists(o->fn)) throw IoError(std::string("input file ") + o->fn + " does not exist");
		bool const verbose = o->verbose != 0;

		PinnedFile in(o->fn);
		if (verbose) fprintf(stderr, "[V] read %llu bytes from %s in %.3f s\n", (unsigned long long)in.n, o->fn, now_sec() - t0);
		// ngpus > 1: the GPUs device .. device+ngpus-1 of this box share the build (multi.cu); engine 0 ends up with the results
		int const ngpus = o->ngpus > 1 ? o->ngpus : 1;
		struct MultiHolder { b3m_multi * m = nullptr; ~MultiHolder() { b3m_multi_destroy(m); } } mh;
		std::unique_ptr<Engine> single;
		Engine * ep = nullptr;
		if (ngpus > 1) {
			std::vector<int> devs(ngpus);
			for (int i = 0; i < ngpus; ++i) devs[i] = o->device + i;
			char merr[512] = {0};
			if (b3m_multi_create(ngpus, devs.data(), &mh.m, merr, sizeof(merr))) throw Error(merr);
			if (b3m_multi_load_host(mh.m, in.p, in.n, itype)) throw Error(b3m_multi_last_error(mh.m));
			ep = b3m_multi_engine(mh.m, 0)->e;
		} else {
			single.reset(new Engine(o->device, nullptr));
			single->load(in.p, in.n, itype, false);
			ep = single.get();
		}
		Engine & e = *ep;
		uint64_t const n = e.T.n;

		// block size: the whole text is one block when its working set fits the memory target;
		// mem=0 means the free device memory (SURVEY 3.1 step 3 for the reference's rule)
		size_t freeb = 0, totalb = 0;
		B3M_CUDA(cudaMemGetInfo(&freeb, &totalb));
		uint64_t const avail = (uint64_t)freeb + e.arena.capacity - e.arena.in_use;
		uint64_t target = o->mem ? std::min<uint64_t>(o->mem, avail) : avail;
		uint64_t numblocks = o->numblocks;
		if (!numblocks) {
			uint64_t const per_block = std::max<uint64_t>(target / BYTES_PER_SUFFIX, 1);
			numblocks = std::max<uint64_t>(1, (n + per_block - 1) / per_block);
		}
		b3m_build_params p;
		memset(&p, 0, sizeof(p));
		p.numblocks = numblocks;
		p.preisarate = 0;
		p.sasamplingrate = o->sasamplingrate ? o->sasamplingrate : 32;
		p.isasamplingrate = o->isasamplingrate ? o->isasamplingrate : 262144;
		p.bwtonly = o->bwtonly;
		p.largelcpthres = o->largelcpthres ? o->largelcpthres : 16384;
		if (verbose) fprintf(stderr, "[V] n=%llu sigma=%u numblocks=%llu (memory target %llu MiB)\n", (unsigned long long)n, e.T.sigma + (e.T.has_term ? 1 : 0),
		                     (unsigned long long)numblocks, (unsigned long long)(target >> 20));
		if (ngpus > 1) {
			if (b3m_multi_build(mh.m, &p)) throw Error(b3m_multi_last_error(mh.m));
			if (verbose) {
				char strat[64]; double msl = 0, msb = 0;
				b3m_multi_stats(mh.m, strat, sizeof(strat), &msl, &msb);
				fprintf(stderr, "[V] %d GPUs: load %.3f ms, build %.3f ms, strategy %s\n", ngpus, msl, msb, strat);
			}
		} else e.build(p);
		b3m_info info;
		e.info(&info);
		if (verbose)
			fprintf(stderr, "[V] device: decode %.3f ms, sort %.3f ms, extract %.3f ms, gap %.3f ms, merge %.3f ms, dict %.3f ms, walk %.3f ms\n",
			        info.ms_decode, info.ms_sort, info.ms_extract, info.ms_gap, info.ms_merge, info.ms_dict, info.ms_walk);

		// outputs: .bwt, .hist, then either .preisa(+.meta) or .sa/.isa
		double const t1 = now_sec();
		e.write_bwt(bwtfn.c_str());
		write_hist(prefix + ".hist", info.hist);
		copy_name(res->textfn, o->fn);
		copy_name(res->bwtfn, bwtfn);
		copy_name(res->histfn, prefix + ".hist");
		if (o->bwtonly) {
			std::vector<uint64_t> pairs(2 * info.npreisa);
			e.fetch(nullptr, pairs.data(), nullptr, nullptr);
			write_preisa(prefix + ".preisa", pairs.data(), info.npreisa, info.preisarate);
			copy_name(res->preisafn, prefix + ".preisa");
			copy_name(res->metafn, prefix + ".preisa.meta");
		} else {
			// the reference removes its .preisa files in this mode (ChangeLog 0.0.49)
			std::vector<uint64_t> sa(info.nsa), isa(info.nisa);
			e.fetch(nullptr, nullptr, sa.data(), isa.data());
			write_sampled(prefix + ".sa", info.sasamplingrate, sa.data(), info.nsa);
			write_sampled(prefix + ".isa", info.isasamplingrate, isa.data(), info.nisa);
			copy_name(res->safn, prefix + ".sa");
			copy_name(res->isafn, prefix + ".isa");
		}
		if (verbose) fprintf(stderr, "[V] wrote %s (%llu runs, %llu payload bytes) and siblings in %.3f s\n", bwtfn.c_str(),
		                     (unsigned long long)e.rl_nruns, (unsigned long long)e.rl_bytes, now_sec() - t1);
		res->n = n;
		res->numblocks = info.numblocks;
		res->seconds_device = 1e-3 * (info.ms_decode + info.ms_total);
		res->seconds_total = now_sec() - t0;
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

int b3m_compute_ssa(const char * bwtfn, uint64_t sasamplingrate, uint64_t isasamplingrate, const char * tmpprefix, int copyinputtomemory,
                    uint64_t numthreads, uint64_t maxsortmem, uint64_t maxtmpfiles, int verbose, const char * ref_isa_fn, const char * ref_sa_fn,
                    int device, char * err, size_t errlen) {
	(void)tmpprefix; (void)copyinputtomemory; (void)maxsortmem; (void)maxtmpfiles; // no external memory is needed: the BWT lives in HBM
	try {
		if (!bwtfn || !*bwtfn) throw Error("no .bwt file name");
		check_device_available();
		double const t0 = now_sec();
		std::string const prefix = clip_off(bwtfn, ".bwt");
		std::string const preisafn = prefix + ".preisa";
		if (!file_exists(bwtfn)) throw IoError(std::string(bwtfn) + " does not exist");
		if (!file_exists(preisafn)) throw IoError(preisafn + " does not exist (run bwtb3m with bwtonly=1 first)");
		if (!numthreads) numthreads = 1;
		std::vector<uint8_t> const L = RlDecoder::decodeAll(std::vector<std::string>(1, bwtfn), numthreads);
		std::vector<uint64_t> const pairs = read_preisa(preisafn);
		if (verbose) fprintf(stderr, "[V] decoded %llu symbols, %llu anchors in %.3f s\n", (unsigned long long)L.size(), (unsigned long long)(pairs.size() / 2), now_sec() - t0);
		Engine e(device, nullptr);
		e.ssa_from_bwt(L.data(), L.size(), pairs.data(), pairs.size() / 2, sasamplingrate ? sasamplingrate : 32, isasamplingrate ? isasamplingrate : 32);
		std::vector<uint64_t> sa(e.nsa), isa(e.nisa);
		e.fetch(nullptr, nullptr, sa.data(), isa.data());
		for (auto v : sa) if (v == ~0ull) throw Error("sampled suffix array is incomplete: the anchors do not cover the text");   // hwtPreIsaToIsa.cpp:166-167
		for (auto v : isa) if (v == ~0ull) throw Error("sampled inverse suffix array is incomplete: the anchors do not cover the text");
		write_sampled(prefix + ".sa", e.params.sasamplingrate, sa.data(), sa.size());
		write_sampled(prefix + ".isa", e.params.isasamplingrate, isa.data(), isa.size());
		// optional comparison with reference files (ref_isa / ref_sa arguments of the reference tool)
		auto compare = [&](const char * fn, std::vector<uint64_t> const & mine, uint64_t rate, const char * what) {
			if (!fn || !*fn) return;
			uint64_t r = 0; std::vector<uint64_t> v;
			read_sampled(fn, &r, &v);
			if (r != rate || v != mine) throw Error(std::string("sampled ") + what + " differs from reference file " + fn);
			if (verbose) fprintf(stderr, "[V] sampled %s equals %s\n", what, fn);
		};
		compare(ref_sa_fn, sa, e.params.sasamplingrate, "suffix array");
		compare(ref_isa_fn, isa, e.params.isasamplingrate, "inverse suffix array");
		if (verbose) fprintf(stderr, "[V] dictionary %.3f ms, walk %.3f ms, total %.3f s\n", e.ms_dict, e.ms_walk, now_sec() - t0);
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// checkbwt <in.bwt> <text>  (/root/reference/src/checkbwt.cpp:26-246): *ok = 1 iff the LF walk from the
// anchors of <prefix>.preisa reproduces the text, every position compared once
int b3m_check_bwt(const char * bwtfn, const char * textfn, const char * inputtype, uint64_t numthreads, int device, int verbose,
                  int * ok, uint64_t * mismatches, char * err, size_t errlen) {
	try {
		if (!bwtfn || !*bwtfn || !textfn || !*textfn) throw Error("checkbwt needs a .bwt file and the text");
		check_device_available();
		int const itype = b3m_parse_inputtype(inputtype && *inputtype ? inputtype : "bytestream");
		if (itype < 0) throw Error("unknown/unsupported input type");
		double const t0 = now_sec();
		std::string const prefix = clip_off(bwtfn, ".bwt");
		std::string const preisafn = prefix + ".preisa";
		if (!file_exists(bwtfn)) throw IoError(std::string(bwtfn) + " does not exist");
		if (!file_exists(textfn)) throw IoError(std::string(textfn) + " does not exist");
		if (!file_exists(preisafn)) throw IoError(preisafn + " does not exist (checkbwt needs the anchors: run bwtb3m with bwtonly=1)");
		if (!numthreads) numthreads = 1;
		std::vector<uint8_t> const L = RlDecoder::decodeAll(std::vector<std::string>(1, bwtfn), numthreads);
		std::vector<uint64_t> const pairs = read_preisa(preisafn);
		Engine e(device, nullptr);
		{
			PinnedFile text(textfn);
			e.load(text.p, text.n, itype, false);
		}
		uint64_t badrank = ~0ull;
		uint64_t const bad = e.check_bwt(L.data(), L.size(), pairs.data(), pairs.size() / 2, &badrank);
		if (verbose) {
			if (bad && badrank != ~0ull) fprintf(stderr, "[E] failure at rank %llu\n", (unsigned long long)badrank);
			fprintf(stderr, "[V] %llu/%llu symbols compared from %llu anchors, %llu mismatches, walk %.3f ms, total %.3f s\n", (unsigned long long)L.size(),
			        (unsigned long long)e.T.n, (unsigned long long)(pairs.size() / 2), (unsigned long long)bad, e.ms_walk, now_sec() - t0);
		}
		if (ok) *ok = bad == 0;
		if (mismatches) *mismatches = bad;
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// bwttestdecodespeed <in.bwt>  (/root/reference/src/bwttestdecodespeed.cpp:27-97): dependent LF chains started at
// evenly spaced samples of <prefix>.isa over a rank dictionary of the BWT; steps_per_s = nchains * steps / seconds
int b3m_lf_speed(const char * bwtfn, uint64_t nchains, uint64_t steps, uint64_t numthreads, int device, double * steps_per_s, double * seconds,
                 char * err, size_t errlen) {
	try {
		if (!bwtfn || !*bwtfn) throw Error("no .bwt file name");
		check_device_available();
		std::string const prefix = clip_off(bwtfn, ".bwt");
		std::string const isafn = prefix + ".isa";
		if (!file_exists(bwtfn)) throw IoError(std::string(bwtfn) + " does not exist");
		if (!file_exists(isafn)) throw IoError(isafn + " does not exist (bwttestdecodespeed starts its chains at the sampled ISA)");
		if (!numthreads) numthreads = 1;
		std::vector<uint8_t> const L = RlDecoder::decodeAll(std::vector<std::string>(1, bwtfn), numthreads);
		uint64_t rate = 0; std::vector<uint64_t> isa;
		read_sampled(isafn, &rate, &isa);
		if (isa.empty()) throw Error("empty sampled ISA");
		if (nchains < 1) throw Error("nchains must be at least 1");
		if (L.empty()) throw Error("empty BWT");
		for (auto v : isa) if (v >= L.size()) throw Error(isafn + " holds a rank outside the BWT (" + std::to_string(v) + " >= " + std::to_string(L.size()) + "): not the sampled ISA of this .bwt");
		Engine e(device, nullptr);
		e.install_bwt_symbols(L.data(), L.size(), 16 * nchains);
		if (!steps) steps = std::min<uint64_t>(div_up(L.size(), nchains), 128ull << 20); // bwttestdecodespeed.cpp:84
		float ms = 0;
		e.lf_speed(isa.data(), isa.size(), nchains, 16, &ms); // warm up
		e.lf_speed(isa.data(), isa.size(), nchains, steps, &ms);
		if (seconds) *seconds = ms * 1e-3;
		if (steps_per_s) *steps_per_s = (double)nchains * (double)steps / (ms * 1e-3);
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// BWA's on-disk formats (public bwt_dump_bwt / bwt_dump_sa):
//   .bwt: primary, L2[1..4], then ceil(seq_len/16) uint32 words, 16 symbols per word, symbol i
//         at bits (15-(i&15))*2, the terminator row removed
//   .sa : primary, L2[1..4], sa_intv, seq_len, then SA[k*sa_intv] for k = 1..floor(seq_len/sa_intv)
int b3m_to_bwa(const char * inbwt, const char * outbwt, const char * outsa, char * err, size_t errlen) {
	try {
		if (!inbwt || !outbwt || !outsa) throw Error("null argument");
		std::string const prefix = clip_off(inbwt, ".bwt");
		std::string const safn = prefix + ".sa";
		if (!file_exists(inbwt)) throw IoError(std::string(inbwt) + " does not exist");
		if (!file_exists(safn)) throw IoError(safn + " does not exist (bwtb3m must run with bwtonly=0, or run bwtcomputessa)");
		unsigned const nthreads = std::max(1u, std::thread::hardware_concurrency());
		std::vector<uint8_t> const L = RlDecoder::decodeAll(std::vector<std::string>(1, inbwt), nthreads);
		uint64_t const n = L.size();
		if (n < 2) throw Error("BWT too short for a BWA index");
		uint64_t sarate = 0; std::vector<uint64_t> sa;
		read_sampled(safn, &sarate, &sa);
		uint64_t const seq_len = n - 1;
		// symbol counts and the terminator row, one slice of the BWT per thread
		uint64_t primary = ~0ull, cnt[5] = {0, 0, 0, 0, 0};
		{
			std::vector<std::array<uint64_t, 8>> part(nthreads); // cnt[0..4], position of a terminator, bad symbol seen
			std::vector<std::thread> th;
			for (unsigned t = 0; t < nthreads; ++t) th.emplace_back([&, t]() {
				std::array<uint64_t, 8> a{}; a[5] = ~0ull;
				uint64_t c[256] = {0};
				uint64_t const i0 = n * t / nthreads, i1 = n * (t + 1) / nthreads;
				for (uint64_t i = i0; i < i1; ++i) c[L[i]]++;
				for (int s = 0; s < 5; ++s) a[s] = c[s];
				for (int s = 5; s < 256; ++s) if (c[s]) a[6] = 1;
				if (c[0]) for (uint64_t i = i0; i < i1; ++i) if (L[i] == 0) a[5] = i; // the last one, as a scan would leave it
				part[t] = a;
			});
			for (auto & x : th) x.join();
			for (auto const & a : part) {
				if (a[6]) throw Error("bwtb3mtobwa needs a pacterm BWT (symbols 0..4)");
				for (int s = 0; s < 5; ++s) cnt[s] += a[s];
				if (a[5] != ~0ull) primary = a[5];
			}
		}
		if (cnt[0] != 1) throw Error("bwtb3mtobwa needs exactly one terminator symbol in the BWT");
		uint64_t L2[5]; L2[0] = 0;
		for (int c = 0; c < 4; ++c) L2[c + 1] = L2[c] + cnt[c + 1];
		uint64_t const nw = (seq_len + 15) >> 4;
		std::vector<uint32_t> wds(nw ? nw : 1, 0);
		// pack in parallel: rows before the primary keep their index, rows after it move up by one
		{
			std::vector<std::thread> th;
			for (unsigned t = 0; t < nthreads; ++t) th.emplace_back([&, t]() {
				uint64_t const w0 = nw * t / nthreads, w1 = nw * (t + 1) / nthreads;
				for (uint64_t w = w0; w < w1; ++w) {
					uint32_t v = 0;
					for (uint64_t k = w << 4; k < std::min<uint64_t>((w + 1) << 4, seq_len); ++k) {
						uint64_t const i = k < primary ? k : k + 1;
						v |= (uint32_t)(L[i] - 1) << ((15 - (k & 15)) << 1);
					}
					wds[w] = v;
				}
			});
			for (auto & x : th) x.join();
		}
		{
			std::vector<uint8_t> o(40 + 4 * nw);
			memcpy(o.data(), &primary, 8); memcpy(o.data() + 8, L2 + 1, 32); memcpy(o.data() + 40, wds.data(), 4 * nw);
			write_file(outbwt, o.data(), o.size());
		}
		{
			uint64_t const n_sa = (seq_len + sarate) / sarate;
			if (sa.size() < n_sa) throw Error("sampled suffix array " + safn + " is too short");
			std::vector<uint8_t> o(56 + 8 * (n_sa - 1));
			memcpy(o.data(), &primary, 8); memcpy(o.data() + 8, L2 + 1, 32); memcpy(o.data() + 40, &sarate, 8); memcpy(o.data() + 48, &seq_len, 8);
			if (n_sa > 1) memcpy(o.data() + 56, sa.data() + 1, 8 * (n_sa - 1));
			write_file(outsa, o.data(), o.size());
		}
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

int b3m_bwt_block_sym_histograms(const char * bwtfn, const char * outfn, int64_t minsym, int64_t maxsym, uint64_t numthreads, uint64_t * nblocks,
                                 char * err, size_t errlen) {
	try {
		if (!bwtfn || !outfn) throw Error("null argument");
		uint64_t const nb = RlDecoder::getBlockSymHistograms(bwtfn, outfn, minsym, maxsym, numthreads);
		if (nblocks) *nblocks = nb;
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}
int b3m_bwt_rank(const char * bwtfn, const char * sparserankfn, int64_t minsym, int64_t maxsym, int64_t sym, uint64_t i, uint64_t * rank,
                 char * err, size_t errlen) {
	try {
		if (!bwtfn || !sparserankfn || !rank) throw Error("null argument");
		*rank = RlDecoder::rankm(bwtfn, sparserankfn, minsym, maxsym, sym, i);
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// reader entry points for bindings that do not want to link C++: decode a .bwt into memory
int b3m_bwt_length(const char * bwtfn, uint64_t * n, char * err, size_t errlen) {
	try {
		if (!bwtfn || !n) throw Error("null argument");
		*n = RlDecoder::getLength(std::string(bwtfn));
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

int b3m_bwt_decode(const char * bwtfn, uint8_t * out, uint64_t cap, uint64_t numthreads, char * err, size_t errlen) {
	try {
		if (!bwtfn || !out) throw Error("null argument");
		std::vector<uint8_t> const L = RlDecoder::decodeAll(std::vector<std::string>(1, bwtfn), numthreads ? numthreads : 1);
		if (L.size() > cap) throw Error("output buffer too small");
		memcpy(out, L.data(), L.size());
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// host encoder of the same container (tools/tests; the product path encodes on the device)
int b3m_bwt_encode_host(const char * bwtfn, const uint8_t * syms, uint64_t n, char * err, size_t errlen) {
	try {
		if (!bwtfn || (!syms && n)) throw Error("null argument");
		rl_encode_host(bwtfn, syms, n);
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

// the compactstream container on the host (converters, tests, bindings)
int b3m_compact_write(const char * fn, unsigned bits, const uint8_t * syms, uint64_t n, char * err, size_t errlen) {
	try {
		if (!fn || (!syms && n)) throw Error("null argument");
		CompactWriter w(fn, bits);
		w.write(syms, n);
		w.flush();
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

int b3m_compact_info(const char * fn, unsigned * bits, uint64_t * n, char * err, size_t errlen) {
	try {
		if (!fn || !bits || !n) throw Error("null argument");
		CompactReader r(fn);
		*bits = r.bits();
		*n = r.size();
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

int b3m_compact_read(const char * fn, uint8_t * out, uint64_t cap, char * err, size_t errlen) {
	try {
		if (!fn || !out) throw Error("null argument");
		CompactReader r(fn);
		if (r.size() > cap) throw Error("output buffer too small");
		uint64_t done = 0;
		size_t got;
		while ((got = r.read(out + done, (size_t)std::min<uint64_t>(r.size() - done, 1u << 20))) != 0) done += got;
		if (done != r.size()) throw Error("compact file shorter than its header says");
		return 0;
	} catch (std::exception const & ex) { set_err(err, errlen, ex.what()); return 2; }
	catch (...) { set_err(err, errlen, "unknown error"); return 3; }
}

} // extern "C"
