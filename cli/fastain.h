// Byte source (plain or gzip, file or stdin) and a FASTA record reader for the input converters
// (what libmaus2::lz::BufferedGzipStream + libmaus2::fastx::StreamFastAReaderWrapper provide to
// /root/reference/src/fagzToCompact4.cpp:140-160).  Host only; zlib does the inflating.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <zlib.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace b3mcli {

class ByteSource {
public:
	// fn empty = stdin
	ByteSource(std::string const & fn, bool gz) : name(fn.empty() ? "<stdin>" : fn), buf(1u << 16) {
		if (gz) {
			g = fn.empty() ? gzdopen(0, "rb") : gzopen(fn.c_str(), "rb");
			if (!g) throw std::runtime_error("cannot open " + name + " for reading");
			gzbuffer(g, 1u << 18);
		} else {
			f = fn.empty() ? stdin : fopen(fn.c_str(), "rb");
			if (!f) throw std::runtime_error("cannot open " + name + " for reading");
		}
	}
	~ByteSource() {
		if (g) gzclose(g);
		if (f && f != stdin) fclose(f);
	}
	ByteSource(ByteSource const &) = delete;
	ByteSource & operator=(ByteSource const &) = delete;
	// up to n bytes; 0 at the end of the data
	size_t read(void * p, size_t n) {
		if (g) {
			int const r = gzread(g, p, (unsigned)n);
			if (r < 0) {
				int e = 0;
				const char * m = gzerror(g, &e);
				throw std::runtime_error("gzip error in " + name + ": " + (m ? m : "?"));
			}
			return (size_t)r;
		}
		size_t const r = fread(p, 1, n, f);
		if (r < n && ferror(f)) throw std::runtime_error("read error on " + name);
		return r;
	}
	int get() {
		if (pos == fill) {
			fill = read(buf.data(), buf.size());
			pos = 0;
			if (!fill) return -1;
		}
		return buf[pos++];
	}
	void unget() { --pos; } // valid after a successful get()
private:
	std::string name;
	gzFile g = nullptr;
	FILE * f = nullptr;
	std::vector<uint8_t> buf;
	size_t pos = 0, fill = 0;
};

struct FastaRecord {
	std::string sid;      // header line without '>'
	std::string spattern; // sequence lines joined, white space removed
};

class FastaReader {
public:
	explicit FastaReader(ByteSource & s) : src(s) {}
	bool next(FastaRecord & r) {
		int c;
		// find the next header
		while ((c = src.get()) >= 0 && c != '>') {
			if (c != '\n' && c != '\r' && c != ' ' && c != '\t')
				throw std::runtime_error("FASTA: data before the first '>' header");
		}
		if (c < 0) return false;
		r.sid.clear();
		r.spattern.clear();
		while ((c = src.get()) >= 0 && c != '\n') if (c != '\r') r.sid.push_back((char)c);
		bool bol = true;
		while ((c = src.get()) >= 0) {
			if (c == '>' && bol) { src.unget(); break; }
			if (c == '\n') { bol = true; continue; }
			bol = false;
			if (c == '\r' || c == ' ' || c == '\t') continue;
			r.spattern.push_back((char)c);
		}
		return true;
	}
private:
	ByteSource & src;
};

// ---- small helpers shared by the fagzToCompact* converters ---------------------------------
// 1025 -> "1k 1" (/root/reference/src/fagzToCompact4.cpp:31-58)
inline std::string format_bytes(uint64_t n) {
	static const char * units[] = {"", "k", "m", "g", "t", "p", "e", "z", "y"};
	std::vector<std::string> parts;
	for (unsigned u = 0; n; ++u, n /= 1024) parts.push_back(std::to_string(n % 1024) + units[u]);
	std::string s;
	for (size_t i = parts.size(); i-- > 0;) { s += parts[i]; if (i) s += " "; }
	return s;
}
inline std::string basename_of(std::string const & s) {
	size_t const p = s.rfind('/');
	return p == std::string::npos ? s : s.substr(p + 1);
}
inline std::string strip_after_dot(std::string const & s) { return s.substr(0, s.find('.')); }
inline std::string clip_off(std::string const & s, std::string const & suffix) {
	if (s.size() >= suffix.size() && !s.compare(s.size() - suffix.size(), suffix.size(), suffix)) return s.substr(0, s.size() - suffix.size());
	return s;
}
inline std::string common_prefix(std::vector<std::string> const & v) {
	if (v.empty()) return std::string();
	std::string p = v[0];
	for (size_t i = 1; i < v.size(); ++i) {
		size_t k = 0;
		while (k < p.size() && k < v[i].size() && p[k] == v[i][k]) ++k;
		p.resize(k);
	}
	return p;
}
// positional file names plus the non-empty lines of the file named by inputfilenames=
inline std::vector<std::string> input_names(std::vector<std::string> const & rest, std::string const & listfn) {
	std::vector<std::string> names = rest;
	if (!listfn.empty()) {
		FILE * f = fopen(listfn.c_str(), "r");
		if (!f) throw std::runtime_error("cannot open " + listfn);
		std::string line;
		int c;
		while ((c = fgetc(f)) != EOF) {
			if (c == '\n') { if (!line.empty()) names.push_back(line); line.clear(); }
			else if (c != '\r') line.push_back((char)c);
		}
		if (!line.empty()) names.push_back(line);
		fclose(f);
	}
	return names;
}

} // namespace b3mcli
