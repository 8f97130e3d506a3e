// Host-side file formats; see formats.h for what is pinned by the reference and what is not.
#include "formats.h"
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <queue>
#include <thread>

namespace b3m {

void put_be64(std::vector<uint8_t> & o, uint64_t v) {
	for (int i = 7; i >= 0; --i) o.push_back((uint8_t)(v >> (8 * i)));
}
static void put_be16(std::vector<uint8_t> & o, uint32_t v) { o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }
uint64_t get_be64(const uint8_t * p) {
	uint64_t v = 0;
	for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
	return v;
}
static uint32_t get_be16(const uint8_t * p) { return ((uint32_t)p[0] << 8) | p[1]; }

bool file_exists(std::string const & fn) { struct stat s; return ::stat(fn.c_str(), &s) == 0; }
uint64_t file_size(std::string const & fn) {
	struct stat s;
	if (::stat(fn.c_str(), &s) != 0) throw IoError("cannot stat " + fn);
	return (uint64_t)s.st_size;
}

struct File {
	FILE * f;
	std::string fn;
	File(std::string const & name, const char * mode) : f(fopen(name.c_str(), mode)), fn(name) {
		if (!f) throw IoError(std::string("cannot open ") + name + (mode[0] == 'r' ? " for reading" : " for writing"));
	}
	~File() { if (f) fclose(f); }
	void write(const void * p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw IoError("write failed on " + fn); }
	void read(void * p, size_t n) { if (n && fread(p, 1, n, f) != n) throw IoError("short read on " + fn); }
	void seek(uint64_t off) { if (fseeko(f, (off_t)off, SEEK_SET) != 0) throw IoError("seek failed on " + fn); }
	void close() { if (f) { int const r = fclose(f); f = nullptr; if (r != 0) throw IoError("close failed on " + fn); } }
};

std::vector<uint8_t> read_file(std::string const & fn) {
	uint64_t const sz = file_size(fn);
	std::vector<uint8_t> v(sz);
	File f(fn, "rb");
	f.read(v.data(), sz);
	return v;
}
void write_file(std::string const & fn, const void * data, size_t bytes) {
	File f(fn, "wb");
	f.write(data, bytes);
	f.close();
}

// ---- .hist: NumberMapSerialisation [layout unpinned] --------------------------------------
void write_hist(std::string const & fn, const uint64_t hist[256]) {
	std::vector<uint8_t> o;
	uint64_t cnt = 0;
	for (int s = 0; s < 256; ++s) if (hist[s]) ++cnt;
	put_be64(o, cnt);
	for (int s = 0; s < 256; ++s) if (hist[s]) { put_be64(o, (uint64_t)s); put_be64(o, hist[s]); }
	write_file(fn, o.data(), o.size());
}
std::map<int64_t, uint64_t> read_hist(std::string const & fn) {
	std::vector<uint8_t> const v = read_file(fn);
	if (v.size() < 8) throw IoError("truncated .hist file " + fn);
	uint64_t const cnt = get_be64(v.data());
	if (v.size() != 8 + 16 * cnt) throw IoError("malformed .hist file " + fn);
	std::map<int64_t, uint64_t> m;
	for (uint64_t i = 0; i < cnt; ++i) m[(int64_t)get_be64(v.data() + 8 + 16 * i)] = get_be64(v.data() + 16 + 16 * i);
	return m;
}

// ---- .sa / .isa ----------------------------------------------------------------------------
void write_sampled(std::string const & fn, uint64_t rate, const uint64_t * v, uint64_t count) {
	File f(fn, "wb");
	f.write(&rate, 8); f.write(&count, 8); f.write(v, 8 * count);
	f.close();
}
void read_sampled(std::string const & fn, uint64_t * rate, std::vector<uint64_t> * v) {
	uint64_t const sz = file_size(fn);
	if (sz < 16) throw IoError("truncated sampled array " + fn);
	File f(fn, "rb");
	uint64_t r, c;
	f.read(&r, 8); f.read(&c, 8);
	if (sz != 16 + 8 * c) throw IoError("malformed sampled array " + fn);
	if (rate) *rate = r;
	if (v) { v->resize(c); f.read(v->data(), 8 * c); }
}

// ---- .preisa ---------------------------------------------------------------------------------
void write_preisa(std::string const & fn, const uint64_t * pairs, uint64_t npairs, uint64_t rate) {
	write_file(fn, pairs, 16 * npairs);
	std::vector<uint8_t> m;
	put_be64(m, rate);
	write_file(fn + ".meta", m.data(), m.size());
}
std::vector<uint64_t> read_preisa(std::string const & fn) {
	uint64_t const sz = file_size(fn);
	if (sz % 16) throw IoError("size of " + fn + " is not a multiple of 16"); // hwtPreIsaToIsa.cpp:55-62
	std::vector<uint64_t> v(sz / 8);
	File f(fn, "rb");
	f.read(v.data(), sz);
	return v;
}

// ---- canonical Huffman -----------------------------------------------------------------------
static void huff_lengths(std::vector<uint64_t> const & freq, std::vector<uint8_t> & len) {
	size_t const n = freq.size();
	len.assign(n, 0);
	std::vector<size_t> used;
	for (size_t i = 0; i < n; ++i) if (freq[i]) used.push_back(i);
	if (used.empty()) return;
	if (used.size() == 1) { len[used[0]] = 1; return; }
	// tree as parent links over leaves + internal nodes
	size_t const L = used.size();
	std::vector<uint64_t> w(2 * L);
	std::vector<int> parent(2 * L, -1);
	typedef std::pair<uint64_t, int> QE;
	std::priority_queue<QE, std::vector<QE>, std::greater<QE>> q;
	for (size_t i = 0; i < L; ++i) { w[i] = freq[used[i]]; q.push(QE(w[i], (int)i)); }
	int next = (int)L;
	while (q.size() > 1) {
		QE const a = q.top(); q.pop();
		QE const b = q.top(); q.pop();
		w[next] = a.first + b.first;
		parent[a.second] = parent[b.second] = next;
		q.push(QE(w[next], next));
		++next;
	}
	for (size_t i = 0; i < L; ++i) {
		unsigned d = 0;
		for (int p = parent[i]; p >= 0; p = parent[p]) ++d;
		len[used[i]] = (uint8_t)(d > 255 ? 255 : d);
	}
}

HuffCode huff_build(const uint64_t * freq, size_t nsyms) {
	std::vector<uint64_t> f(freq, freq + nsyms);
	HuffCode h;
	while (true) {
		huff_lengths(f, h.len);
		unsigned mx = 0;
		for (auto l : h.len) mx = std::max<unsigned>(mx, l);
		if (mx <= RL_MAXCODELEN) break;
		// flatten the distribution until the depth bound holds
		for (auto & x : f) if (x) x = (x >> 2) + 1;
	}
	huff_assign(h);
	return h;
}

void huff_assign(HuffCode & h) {
	h.code.assign(h.len.size(), 0);
	uint32_t code = 0;
	for (unsigned l = 1; l <= RL_MAXCODELEN; ++l) {
		for (size_t s = 0; s < h.len.size(); ++s) if (h.len[s] == l) h.code[s] = code++;
		code <<= 1;
	}
}

// canonical decoder: first code and first index per length
struct HuffDecoder {
	static constexpr unsigned LUT_BITS = 12; // codes of up to 12 bits decode with one table look-up
	uint32_t first_code[RL_MAXCODELEN + 2];
	uint32_t first_idx[RL_MAXCODELEN + 2];
	uint32_t count[RL_MAXCODELEN + 2];
	std::vector<uint16_t> sorted; // symbols by (length, symbol)
	std::vector<uint32_t> lut;    // [next LUT_BITS bits] -> (symbol << 8) | code length, 0: a longer code
	explicit HuffDecoder(HuffCode const & h) {
		memset(count, 0, sizeof(count));
		for (auto l : h.len) if (l) count[l]++;
		uint32_t code = 0, idx = 0;
		for (unsigned l = 1; l <= RL_MAXCODELEN; ++l) {
			first_code[l] = code; first_idx[l] = idx;
			code = (code + count[l]) << 1; idx += count[l];
		}
		sorted.resize(idx);
		std::vector<uint32_t> fill(first_idx, first_idx + RL_MAXCODELEN + 2);
		for (size_t s = 0; s < h.len.size(); ++s) if (h.len[s]) sorted[fill[h.len[s]]++] = (uint16_t)s;
		lut.assign((size_t)1 << LUT_BITS, 0);
		for (unsigned l = 1; l <= LUT_BITS; ++l)
			for (uint32_t k = 0; k < count[l]; ++k) {
				uint32_t const c = first_code[l] + k, sym = sorted[first_idx[l] + k];
				uint32_t const lo = c << (LUT_BITS - l), hi = (c + 1) << (LUT_BITS - l);
				for (uint32_t x = lo; x < hi && x < lut.size(); ++x) lut[x] = (sym << 8) | l;
			}
	}
};

// MSB-first bit stream with a 64-bit window
struct BitReader {
	const uint8_t * p; uint64_t nbytes; uint64_t next = 0; // next byte to load
	uint64_t acc = 0; unsigned have = 0;                    // `have` valid bits at the top of acc
	uint64_t used = 0;                                      // bits consumed
	BitReader(const uint8_t * d, uint64_t bytes) : p(d), nbytes(bytes) {}
	inline void refill() { while (have <= 56 && next < nbytes) { acc |= (uint64_t)p[next++] << (56 - have); have += 8; } }
	// the next n <= 32 bits, zero padded behind the end of the block (consuming them there is the error)
	inline uint32_t peek(unsigned n) { if (have < n) refill(); return n ? (uint32_t)(acc >> (64 - n)) : 0u; }
	inline void skip(unsigned n) {
		if (have < n) { refill(); if (have < n) throw IoError("run-length stream: read past the end of a block"); }
		acc <<= n; have -= n; used += n;
	}
	inline uint64_t bits(unsigned n) { // n <= 64
		uint64_t v = 0;
		while (n > 32) { v = (v << 32) | peek(32); skip(32); n -= 32; }
		if (n) { v = (v << n) | peek(n); skip(n); }
		return v;
	}
	inline uint32_t huff(HuffDecoder const & d) {
		uint32_t const e = d.lut[peek(HuffDecoder::LUT_BITS)];
		if (e) { skip(e & 255u); return e >> 8; }
		uint32_t const w = peek(RL_MAXCODELEN);
		for (unsigned l = HuffDecoder::LUT_BITS + 1; l <= RL_MAXCODELEN; ++l) {
			uint32_t const code = w >> (RL_MAXCODELEN - l);
			if (d.count[l] && code >= d.first_code[l] && code - d.first_code[l] < d.count[l]) { skip(l); return d.sorted[d.first_idx[l] + (code - d.first_code[l])]; }
		}
		throw IoError("run-length stream: invalid Huffman code");
	}
};

struct BitWriter {
	std::vector<uint8_t> & o; uint64_t acc = 0; unsigned fill = 0;
	explicit BitWriter(std::vector<uint8_t> & out) : o(out) {}
	void put(uint64_t v, unsigned n) { // n <= 32
		acc = (acc << n) | (v & ((n >= 64) ? ~0ull : ((1ull << n) - 1)));
		fill += n;
		while (fill >= 8) { o.push_back((uint8_t)(acc >> (fill - 8))); fill -= 8; }
	}
	void align64() { if (fill) put(0, 8 - fill); while (o.size() % 8) o.push_back(0); }
};

static void serialise_table(std::vector<uint8_t> & o, HuffCode const & h) {
	uint32_t cnt = 0;
	for (auto l : h.len) if (l) ++cnt;
	put_be16(o, cnt);
	for (size_t s = 0; s < h.len.size(); ++s) if (h.len[s]) { put_be16(o, (uint32_t)s); o.push_back(h.len[s]); }
}

static const char RL_MAGIC[8] = {'B', '3', 'M', 'R', 'L', '0', '1', 0};

struct RlContainerWriter::Impl {
	File f;
	RlHeader h;
	uint64_t hdbytes = 0, paybytes = 0;
	Impl(std::string const & fn, RlHeader const & hh) : f(fn, "wb"), h(hh) {}
};
RlContainerWriter::RlContainerWriter(std::string const & fn, RlHeader const & h) : impl(new Impl(fn, h)) {
	std::vector<uint8_t> hd(RL_MAGIC, RL_MAGIC + 8);
	put_be64(hd, h.n); put_be64(hd, h.nruns); put_be64(hd, h.runs_per_block); put_be64(hd, h.nblocks);
	serialise_table(hd, h.sym);
	serialise_table(hd, h.len);
	while (hd.size() % 8) hd.push_back(0);
	impl->f.write(hd.data(), hd.size());
	impl->hdbytes = hd.size();
}
RlContainerWriter::~RlContainerWriter() {}
void RlContainerWriter::payload(const void * bytes, size_t nbytes) { impl->f.write(bytes, nbytes); impl->paybytes += nbytes; }
void RlContainerWriter::finish(const uint64_t * block_word_off, const uint64_t * block_sym_off) {
	std::vector<uint8_t> idx;
	idx.reserve(16 * impl->h.nblocks + 8);
	for (uint64_t b = 0; b < impl->h.nblocks; ++b) { put_be64(idx, block_word_off[b]); put_be64(idx, block_sym_off[b]); }
	put_be64(idx, impl->hdbytes + impl->paybytes);
	impl->f.write(idx.data(), idx.size());
	impl->f.close();
}

static inline unsigned bitlen64(uint64_t v) { unsigned b = 0; while (v) { ++b; v >>= 1; } return b; }

void rl_encode_host(std::string const & fn, const uint8_t * syms, uint64_t n) {
	std::vector<uint8_t> rsym; std::vector<uint64_t> rlen;
	for (uint64_t i = 0; i < n;) {
		uint64_t j = i + 1;
		while (j < n && syms[j] == syms[i]) ++j;
		rsym.push_back(syms[i]); rlen.push_back(j - i);
		i = j;
	}
	uint64_t fs[256] = {0}, fl[RL_LENBINS] = {0};
	for (size_t k = 0; k < rsym.size(); ++k) { fs[rsym[k]]++; fl[rlen[k] < RL_LENBINS ? rlen[k] : 0]++; }
	RlHeader h;
	h.n = n; h.nruns = rsym.size(); h.nblocks = (h.nruns + h.runs_per_block - 1) / h.runs_per_block;
	h.sym = huff_build(fs, 256); h.len = huff_build(fl, RL_LENBINS);
	std::vector<uint8_t> payload;
	std::vector<uint64_t> woff(h.nblocks), soff(h.nblocks);
	BitWriter bw(payload);
	uint64_t symsbefore = 0;
	for (uint64_t b = 0; b < h.nblocks; ++b) {
		woff[b] = payload.size() / 8; soff[b] = symsbefore;
		uint64_t const k1 = std::min<uint64_t>((b + 1) * h.runs_per_block, h.nruns);
		for (uint64_t k = b * h.runs_per_block; k < k1; ++k) {
			bw.put(h.sym.code[rsym[k]], h.sym.len[rsym[k]]);
			uint64_t const l = rlen[k];
			if (l < RL_LENBINS) bw.put(h.len.code[l], h.len.len[l]);
			else {
				unsigned const nb = bitlen64(l);
				bw.put(h.len.code[0], h.len.len[0]);
				bw.put(nb - 1, 6);
				if (nb > 32) { bw.put(l >> 32, nb - 32); bw.put(l & 0xffffffffull, 32); } else bw.put(l, nb);
			}
			symsbefore += l;
		}
		bw.align64();
	}
	RlContainerWriter wr(fn, h);
	wr.payload(payload.data(), payload.size());
	wr.finish(woff.data(), soff.data());
}

// ---- reader -----------------------------------------------------------------------------------
struct RlFile {
	std::string fn;
	RlHeader h;
	std::unique_ptr<HuffDecoder> dsym, dlen;
	uint64_t payload_off = 0, payload_words = 0;
	std::vector<uint64_t> woff, soff;

	static void read_table(const uint8_t *& p, const uint8_t * end, HuffCode & h, size_t nsyms, std::string const & fn) {
		if (end - p < 2) throw IoError("truncated header in " + fn);
		uint32_t const cnt = get_be16(p); p += 2;
		if ((size_t)(end - p) < 3 * (size_t)cnt) throw IoError("truncated code table in " + fn);
		h.len.assign(nsyms, 0);
		for (uint32_t i = 0; i < cnt; ++i, p += 3) {
			uint32_t const s = get_be16(p);
			if (s >= nsyms || p[2] == 0 || p[2] > RL_MAXCODELEN) throw IoError("bad code table in " + fn);
			h.len[s] = p[2];
		}
		huff_assign(h);
	}

	explicit RlFile(std::string const & name) : fn(name) {
		uint64_t const sz = file_size(fn);
		if (sz < 48) throw IoError(fn + " is not a b3m run-length container (too short)");
		File f(fn, "rb");
		uint8_t fix[40];
		f.read(fix, 40);
		if (memcmp(fix, RL_MAGIC, 8)) throw IoError(fn + " is not a b3m run-length container (bad magic); files written by libmaus2's RLEncoder are not readable by this library and vice versa -- see README.md, file formats");
		h.n = get_be64(fix + 8); h.nruns = get_be64(fix + 16); h.runs_per_block = get_be64(fix + 24); h.nblocks = get_be64(fix + 32);
		f.seek(sz - 8);
		uint8_t t8[8]; f.read(t8, 8);
		uint64_t const idxoff = get_be64(t8);
		if (idxoff > sz - 8 || sz - 8 - idxoff != 16 * h.nblocks) throw IoError("malformed index in " + fn);
		size_t const hdmax = (size_t)std::min<uint64_t>(idxoff, 40 + 2 * (2 + 3 * 65536));
		std::vector<uint8_t> hd(hdmax);
		f.seek(0); f.read(hd.data(), hdmax);
		const uint8_t * p = hd.data() + 40;
		read_table(p, hd.data() + hdmax, h.sym, 256, fn);
		read_table(p, hd.data() + hdmax, h.len, RL_LENBINS, fn);
		payload_off = (uint64_t)(p - hd.data());
		payload_off = (payload_off + 7) & ~7ull;
		if (payload_off > idxoff || (idxoff - payload_off) % 8) throw IoError("malformed payload in " + fn);
		payload_words = (idxoff - payload_off) / 8;
		std::vector<uint8_t> idx(16 * h.nblocks);
		f.seek(idxoff); f.read(idx.data(), idx.size());
		woff.resize(h.nblocks); soff.resize(h.nblocks);
		for (uint64_t b = 0; b < h.nblocks; ++b) { woff[b] = get_be64(idx.data() + 16 * b); soff[b] = get_be64(idx.data() + 16 * b + 8); }
		dsym.reset(new HuffDecoder(h.sym)); dlen.reset(new HuffDecoder(h.len));
	}

	uint64_t block_runs(uint64_t b) const { return std::min<uint64_t>(h.runs_per_block, h.nruns - b * h.runs_per_block); }
	uint64_t block_words(uint64_t b) const { return (b + 1 < h.nblocks ? woff[b + 1] : payload_words) - woff[b]; }

	// decode one block into runs
	void decode_block(File & f, uint64_t b, std::vector<uint8_t> & buf, std::vector<std::pair<uint8_t, uint64_t>> & runs) const {
		uint64_t const bytes = 8 * block_words(b);
		buf.resize(bytes);
		f.seek(payload_off + 8 * woff[b]);
		f.read(buf.data(), bytes);
		decode_block_mem(buf.data(), bytes, b, runs);
	}
	void decode_block_mem(const uint8_t * d, uint64_t bytes, uint64_t b, std::vector<std::pair<uint8_t, uint64_t>> & runs) const {
		BitReader br(d, bytes);
		uint64_t const nr = block_runs(b);
		runs.resize(nr);
		for (uint64_t k = 0; k < nr; ++k) {
			uint32_t const s = br.huff(*dsym);
			uint64_t l = br.huff(*dlen);
			if (l == 0) { unsigned const nb = (unsigned)br.bits(6) + 1; l = br.bits(nb); }
			runs[k] = std::make_pair((uint8_t)s, l);
		}
	}
};

struct RlDecoder::Impl {
	std::vector<std::unique_ptr<RlFile>> files;
	size_t fi = 0;
	std::unique_ptr<File> fh;
	uint64_t blk = 0;
	std::vector<uint8_t> buf;
	std::vector<std::pair<uint8_t, uint64_t>> runs;
	size_t ri = 0;
	uint64_t skip = 0;      // symbols to drop from the front (offset inside the first run)
	bool have_put = false;
	std::pair<int64_t, uint64_t> cur = std::make_pair((int64_t)-1, (uint64_t)0);

	bool load_next_block() {
		while (fi < files.size()) {
			RlFile const & F = *files[fi];
			if (blk < F.h.nblocks) {
				if (!fh) fh.reset(new File(F.fn, "rb"));
				F.decode_block(*fh, blk++, buf, runs);
				ri = 0;
				return true;
			}
			++fi; blk = 0; fh.reset();
		}
		return false;
	}
};

RlDecoder::RlDecoder(std::vector<std::string> const & fns, uint64_t offset, uint64_t) : impl(new Impl) {
	for (auto const & f : fns) impl->files.emplace_back(new RlFile(f));
	// position on the block that holds symbol `offset` (uses the block index)
	uint64_t off = offset;
	while (impl->fi < impl->files.size() && off >= impl->files[impl->fi]->h.n) { off -= impl->files[impl->fi]->h.n; ++impl->fi; }
	if (impl->fi < impl->files.size()) {
		RlFile const & F = *impl->files[impl->fi];
		uint64_t const b = (uint64_t)(std::upper_bound(F.soff.begin(), F.soff.end(), off) - F.soff.begin());
		impl->blk = b ? b - 1 : 0;
		uint64_t drop = F.h.nblocks ? off - F.soff[impl->blk] : 0;
		if (impl->load_next_block()) {
			while (impl->ri < impl->runs.size() && drop >= impl->runs[impl->ri].second) drop -= impl->runs[impl->ri++].second;
			impl->skip = drop;
		}
	}
}
RlDecoder::~RlDecoder() {}

std::pair<int64_t, uint64_t> RlDecoder::decodeRun() {
	if (impl->cur.second) { auto r = impl->cur; impl->cur.second = 0; return r; }
	while (impl->ri >= impl->runs.size()) if (!impl->load_next_block()) return std::make_pair((int64_t)-1, (uint64_t)0);
	auto const & r = impl->runs[impl->ri++];
	uint64_t l = r.second;
	if (impl->skip) { l -= impl->skip; impl->skip = 0; }
	return std::make_pair((int64_t)r.first, l);
}

int64_t RlDecoder::decode() {
	if (!impl->cur.second) {
		impl->cur = decodeRun();
		if (impl->cur.first < 0) { impl->cur.second = 0; return -1; }
	}
	--impl->cur.second;
	return impl->cur.first;
}

uint64_t RlDecoder::getLength(std::vector<std::string> const & files, uint64_t) {
	uint64_t n = 0;
	for (auto const & f : files) n += RlFile(f).h.n;
	return n;
}

std::vector<uint8_t> RlDecoder::decodeAll(std::vector<std::string> const & fns, uint64_t numthreads) {
	std::vector<std::unique_ptr<RlFile>> files;
	uint64_t n = 0;
	std::vector<uint64_t> base;
	for (auto const & f : fns) { files.emplace_back(new RlFile(f)); base.push_back(n); n += files.back()->h.n; }
	std::vector<uint8_t> out(n);
	if (!numthreads) numthreads = 1;
	for (size_t fi = 0; fi < files.size(); ++fi) {
		RlFile const & F = *files[fi];
		uint64_t const nb = F.h.nblocks;
		uint64_t const nt = std::max<uint64_t>(1, std::min<uint64_t>(numthreads, nb));
		std::vector<std::thread> th;
		std::vector<std::string> errs(nt);
		for (uint64_t t = 0; t < nt; ++t) th.emplace_back([&, t]() {
			try {
				File f(F.fn, "rb");
				std::vector<uint8_t> buf; std::vector<std::pair<uint8_t, uint64_t>> runs;
				for (uint64_t b = nb * t / nt; b < nb * (t + 1) / nt; ++b) {
					F.decode_block(f, b, buf, runs);
					uint64_t o = base[fi] + F.soff[b];
					for (auto const & r : runs) {
						if (o + r.second > n) throw IoError("run-length stream longer than its header says: " + F.fn);
						if (r.second < 8) { for (uint64_t x = 0; x < r.second; ++x) out[o + x] = r.first; } else memset(out.data() + o, r.first, r.second);
						o += r.second;
					}
				}
			} catch (std::exception const & ex) { errs[t] = ex.what(); }
		});
		for (auto & x : th) x.join();
		for (auto const & e : errs) if (!e.empty()) throw IoError(e);
	}
	return out;
}

uint64_t RlDecoder::getBlockSymHistograms(std::string const & bwtfn, std::string const & outfn, int64_t minsym, int64_t maxsym, uint64_t numthreads) {
	if (maxsym < minsym || minsym < 0 || maxsym > 255) throw IoError("getBlockSymHistograms: symbol range must lie in 0..255");
	RlFile const F(bwtfn);
	uint64_t const nb = F.h.nblocks, ns = (uint64_t)(maxsym - minsym + 1);
	std::vector<uint64_t> H(nb * ns, 0); // first the counts of the block itself
	uint64_t const nt = std::max<uint64_t>(1, std::min<uint64_t>(numthreads ? numthreads : 1, nb));
	std::vector<std::thread> th;
	std::vector<std::string> errs(nt);
	for (uint64_t t = 0; t < nt; ++t) th.emplace_back([&, t]() {
		try {
			File f(F.fn, "rb");
			std::vector<uint8_t> buf; std::vector<std::pair<uint8_t, uint64_t>> runs;
			for (uint64_t b = nb * t / nt; b < nb * (t + 1) / nt; ++b) {
				F.decode_block(f, b, buf, runs);
				for (auto const & r : runs) {
					if ((int64_t)r.first < minsym || (int64_t)r.first > maxsym) throw IoError("getBlockSymHistograms: symbol outside [minsym,maxsym] in " + F.fn);
					H[b * ns + (uint64_t)(r.first - minsym)] += r.second;
				}
			}
		} catch (std::exception const & ex) { errs[t] = ex.what(); }
	});
	for (auto & x : th) x.join();
	for (auto const & e : errs) if (!e.empty()) throw IoError(e);
	// exclusive prefix sums down the blocks, written big-endian
	std::vector<uint64_t> run(ns, 0);
	std::vector<uint8_t> o;
	o.reserve(8 * nb * ns);
	for (uint64_t b = 0; b < nb; ++b)
		for (uint64_t c = 0; c < ns; ++c) { put_be64(o, run[c]); run[c] += H[b * ns + c]; }
	write_file(outfn, o.data(), o.size());
	return nb;
}

uint64_t RlDecoder::rankm(std::string const & bwtfn, std::string const & sparserankfn, int64_t minsym, int64_t maxsym, int64_t sym, uint64_t i) {
	if (sym < minsym || sym > maxsym) throw IoError("rankm: symbol outside [minsym,maxsym]");
	RlFile const F(bwtfn);
	if (i > F.h.n) throw IoError("rankm: position behind the end of the sequence");
	if (F.h.nblocks == 0) return 0;
	uint64_t const ns = (uint64_t)(maxsym - minsym + 1);
	uint64_t b = (uint64_t)(std::upper_bound(F.soff.begin(), F.soff.end(), i) - F.soff.begin());
	b = b ? b - 1 : 0;
	File sr(sparserankfn, "rb");
	sr.seek(8 * (b * ns + (uint64_t)(sym - minsym)));
	uint8_t v8[8];
	sr.read(v8, 8);
	uint64_t r = get_be64(v8);
	File f(F.fn, "rb");
	std::vector<uint8_t> buf; std::vector<std::pair<uint8_t, uint64_t>> runs;
	F.decode_block(f, b, buf, runs);
	uint64_t left = i - F.soff[b];
	for (auto const & q : runs) {
		if (!left) break;
		uint64_t const use = std::min<uint64_t>(left, q.second);
		if ((int64_t)q.first == sym) r += use;
		left -= use;
	}
	return r;
}

// ---- compactstream container [layout unpinned] ---------------------------------------------
static void put_le64(std::vector<uint8_t> & o, uint64_t v) { for (int i = 0; i < 8; ++i) o.push_back((uint8_t)(v >> (8 * i))); }
static uint64_t get_le64(const uint8_t * p) { uint64_t v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | p[i]; return v; }
struct CompactWriter::Impl {
	File f;
	unsigned b;
	uint64_t n = 0, acc = 0;
	unsigned fill = 0; // bits held in acc
	std::vector<uint8_t> out;
	bool open = true;
	Impl(std::string const & fn, unsigned bits) : f(fn, "wb"), b(bits) {}
	void word() { // native little-endian uint64, the first symbol in its top bits
		for (int i = 0; i < 8; ++i) out.push_back((uint8_t)(acc >> (8 * i)));
		acc = 0; fill = 0;
		if (out.size() >= (1u << 20)) { f.write(out.data(), out.size()); out.clear(); }
	}
};
CompactWriter::CompactWriter(std::string const & fn, unsigned bits) : impl(new Impl(fn, bits)) {
	if (bits < 1 || bits > 8) throw IoError("compact container: bits per symbol must be 1..8");
	std::vector<uint8_t> h;
	put_le64(h, bits); put_le64(h, 0); put_le64(h, 0); put_le64(h, 0);
	impl->f.write(h.data(), h.size());
}
CompactWriter::~CompactWriter() {
	try { flush(); } catch (...) {}
}
uint64_t CompactWriter::size() const { return impl->n; }
void CompactWriter::write(const uint8_t * syms, size_t n) {
	Impl & w = *impl;
	if (!w.open) throw IoError("compact container: write after flush on " + w.f.fn);
	unsigned const b = w.b;
	for (size_t i = 0; i < n; ++i) {
		uint64_t const v = syms[i];
		if (v >> b) throw IoError("compact container: symbol does not fit the bits per symbol");
		unsigned const room = 64 - w.fill;
		if (b <= room) {
			w.acc |= v << (room - b);
			w.fill += b;
			if (w.fill == 64) w.word();
		} else { // the symbol straddles two words
			unsigned const lo = b - room;
			w.acc |= v >> lo;
			w.word();
			w.acc = (v & ((1ull << lo) - 1)) << (64 - lo);
			w.fill = lo;
		}
	}
	w.n += n;
}
void CompactWriter::flush() {
	Impl & w = *impl;
	if (!w.open) return;
	w.open = false;
	if (w.fill) w.word();
	w.f.write(w.out.data(), w.out.size());
	w.out.clear();
	uint64_t const words = (w.n * w.b + 63) / 64;
	std::vector<uint8_t> h;
	put_le64(h, w.b); put_le64(h, w.n); put_le64(h, words); put_le64(h, words);
	w.f.seek(0);
	w.f.write(h.data(), h.size());
	w.f.close();
}

struct CompactReader::Impl {
	File f;
	uint64_t n = 0, pos = 0;
	unsigned b = 0;
	bool le_words = false; // words stored as native little-endian uint64 (else: big-endian byte stream)
	std::vector<uint8_t> buf;
	explicit Impl(std::string const & fn) : f(fn, "rb") {}
};
CompactReader::CompactReader(std::string const & fn) : impl(new Impl(fn)) {
	uint64_t const fsz = file_size(fn);
	if (fsz < 32) throw IoError("compact file too short: " + fn);
	uint8_t h[32];
	impl->f.read(h, 32);
	// bits per symbol is 1..8 in exactly one of the two byte orders (formats.h)
	uint64_t b = get_be64(h);
	impl->n = get_be64(h + 8);
	if (b < 1 || b > 8) { b = get_le64(h); impl->n = get_le64(h + 8); impl->le_words = true; }
	if (b < 1 || b > 8) throw IoError("compact file: unsupported bits per symbol in " + fn);
	if (impl->n > (fsz - 32) * 8 / b || (impl->le_words && (impl->n * b + 63) / 64 * 8 > fsz - 32)) throw IoError("compact file: truncated: " + fn);
	impl->b = (unsigned)b;
}
CompactReader::~CompactReader() {}
uint64_t CompactReader::size() const { return impl->n; }
unsigned CompactReader::bits() const { return impl->b; }
size_t CompactReader::read(uint8_t * out, size_t want) {
	Impl & r = *impl;
	size_t const n = (size_t)std::min<uint64_t>(want, r.n - r.pos);
	if (!n) return 0;
	uint64_t const bit0 = r.pos * r.b, bit1 = (r.pos + n) * r.b;
	// whole 64-bit words around the wanted bits, so that the little-endian flip stays inside the buffer
	uint64_t const byte0 = (bit0 >> 6) << 3, byte1 = ((bit1 + 63) >> 6) << 3;
	uint64_t const have = std::min<uint64_t>(byte1, file_size(r.f.fn) - 32) - byte0;
	r.buf.assign((size_t)(byte1 - byte0) + 8, 0);
	r.f.seek(32 + byte0);
	r.f.read(r.buf.data(), (size_t)have);
	unsigned const b = r.b;
	uint32_t const mask = (1u << b) - 1u;
	size_t const flip = r.le_words ? 7 : 0;
	for (size_t i = 0; i < n; ++i) {
		uint64_t const bit = bit0 + (uint64_t)i * b - (byte0 << 3);
		size_t const by = (size_t)(bit >> 3);
		uint32_t const two = ((uint32_t)r.buf[by ^ flip] << 8) | r.buf[(by + 1) ^ flip];
		out[i] = (uint8_t)((two >> (16 - (unsigned)(bit & 7) - b)) & mask);
	}
	r.pos += n;
	return n;
}

// ---- key=value arguments -------------------------------------------------------------------
ArgInfo::ArgInfo(int argc, char ** argv) {
	progname = argc ? argv[0] : "";
	for (int i = 1; i < argc; ++i) {
		std::string const a = argv[i];
		size_t const eq = a.find('=');
		if (a == "-h" || a == "--help") help = true;
		else if (eq != std::string::npos && eq > 0) kv[a.substr(0, eq)] = a.substr(eq + 1);
		else rest.push_back(a);
	}
}
std::string ArgInfo::get(std::string const & k, std::string const & def) const {
	auto it = kv.find(k);
	return it == kv.end() ? def : it->second;
}
uint64_t ArgInfo::parse_unit_number(std::string const & s) {
	if (s.empty()) throw IoError("empty number");
	size_t i = 0;
	uint64_t v = 0;
	while (i < s.size() && s[i] >= '0' && s[i] <= '9') { v = v * 10 + (uint64_t)(s[i] - '0'); ++i; }
	if (i == 0) throw IoError("cannot parse number " + s);
	if (i < s.size()) {
		switch (s[i]) {
			case 'k': case 'K': v <<= 10; break;
			case 'm': case 'M': v <<= 20; break;
			case 'g': case 'G': v <<= 30; break;
			case 't': case 'T': v <<= 40; break;
			default: throw IoError("unknown unit in number " + s);
		}
	}
	return v;
}
uint64_t ArgInfo::getu(std::string const & k, uint64_t def) const {
	auto it = kv.find(k);
	return it == kv.end() ? def : parse_unit_number(it->second);
}

} // namespace b3m
