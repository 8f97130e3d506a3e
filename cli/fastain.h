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

} // namespace b3mcli
