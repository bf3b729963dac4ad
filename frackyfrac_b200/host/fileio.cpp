// File input / output for the host side: stand-ins for aio.Open / aio.Create of
// fluhus/gostuff v1.0.1 (called at frcfrc/frcfrc.go:93,102; the module's source is
// not under /root/reference, so the suffix rule below is "parity unpinned" —
// SURVEY.md §8c).  The codec follows the file suffix: ".gz" = gzip (zlib),
// ".zst" = zstandard (libzstd.so.1 through dlopen — the image ships the library
// without its header), anything else = plain bytes.
//
// Compressed output is written as a sequence of independent gzip members / zstd
// frames, one per block, compressed by `threads` workers and written in order
// (at 10^9 pairs/s a single deflate stream would be the end-to-end
// bottleneck).  Concatenated members are a valid stream for gzip, zcat, Go's
// compress/gzip (multistream is its default) and zstd alike; the DEcompressed
// bytes are what parity is defined on.
#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "hostlib.hpp"

namespace frchost {
namespace {

bool has_suffix(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

enum Codec { kPlain, kGzip, kZstd };
Codec codec_of(const std::string& path) {
  return has_suffix(path, ".gz") ? kGzip : has_suffix(path, ".zst") ? kZstd : kPlain;
}

// ---- zstd through dlopen (simple + streaming-decompress API, ABI-stable since 1.0)
struct ZBuf { void* p; size_t size, pos; };
struct Zstd {
  size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
  size_t (*bound)(size_t) = nullptr;
  unsigned (*is_error)(size_t) = nullptr;
  const char* (*error_name)(size_t) = nullptr;
  void* (*create_d)() = nullptr;
  size_t (*free_d)(void*) = nullptr;
  size_t (*decompress_stream)(void*, ZBuf*, ZBuf*) = nullptr;
};
const Zstd& zstd() {
  static const Zstd z = [] {
    Zstd r;
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("zstd: ") + dlerror());
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) throw std::runtime_error(std::string("zstd: missing symbol ") + n);
      return p;
    };
    r.compress = reinterpret_cast<decltype(r.compress)>(sym("ZSTD_compress"));
    r.bound = reinterpret_cast<decltype(r.bound)>(sym("ZSTD_compressBound"));
    r.is_error = reinterpret_cast<decltype(r.is_error)>(sym("ZSTD_isError"));
    r.error_name = reinterpret_cast<decltype(r.error_name)>(sym("ZSTD_getErrorName"));
    r.create_d = reinterpret_cast<decltype(r.create_d)>(sym("ZSTD_createDStream"));
    r.free_d = reinterpret_cast<decltype(r.free_d)>(sym("ZSTD_freeDStream"));
    r.decompress_stream = reinterpret_cast<decltype(r.decompress_stream)>(sym("ZSTD_decompressStream"));
    return r;
  }();
  return z;
}

std::string read_raw(const std::string& path) {
  FILE* f = path.empty() ? stdin : fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("open " + path + ": " + strerror(errno));
  std::string out;
  std::vector<char> buf(1 << 20);
  size_t r;
  while ((r = fread(buf.data(), 1, buf.size(), f)) > 0) out.append(buf.data(), r);
  const bool bad = ferror(f);
  if (!path.empty()) fclose(f);
  if (bad) throw std::runtime_error("read " + path + ": " + strerror(errno));
  return out;
}

std::string gunzip(const std::string& in, const std::string& path) {
  std::string out;
  if (in.empty()) throw std::runtime_error(path + ": EOF");  // gzip.NewReader on an empty file
  z_stream zs{};
  if (inflateInit2(&zs, 15 + 16) != Z_OK) throw std::runtime_error("gzip: inflateInit failed");
  std::vector<unsigned char> buf(1 << 20);
  size_t off = 0;  // zlib counts avail_in in 32 bits: feed the input in windows
  int rc = Z_OK;
  bool done = false, truncated = false;
  while (!done) {
    if (zs.avail_in == 0) {
      if (off >= in.size()) { truncated = true; break; }
      const size_t take = std::min<size_t>(in.size() - off, size_t{1} << 30);
      zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data() + off));
      zs.avail_in = static_cast<uInt>(take);
      off += take;
    }
    zs.next_out = buf.data();
    zs.avail_out = static_cast<uInt>(buf.size());
    rc = inflate(&zs, Z_NO_FLUSH);
    out.append(reinterpret_cast<char*>(buf.data()), buf.size() - zs.avail_out);
    if (rc == Z_STREAM_END) {
      if (zs.avail_in == 0 && off >= in.size()) done = true;  // last member
      else inflateReset(&zs);                                 // next member of a multi-member file
    } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
      break;
    }
  }
  const std::string msg = zs.msg ? zs.msg : "";
  inflateEnd(&zs);
  if (!done) throw std::runtime_error(path + ": gzip: " + (truncated || msg.empty() ? "unexpected EOF" : msg));
  return out;
}

std::string unzstd(const std::string& in, const std::string& path) {
  const Zstd& z = zstd();
  void* ds = z.create_d();
  if (!ds) throw std::runtime_error("zstd: out of memory");
  std::string out;
  std::vector<char> buf(1 << 20);
  ZBuf src{const_cast<char*>(in.data()), in.size(), 0};
  size_t rc = 0;
  while (src.pos < src.size) {
    ZBuf dst{buf.data(), buf.size(), 0};
    rc = z.decompress_stream(ds, &dst, &src);
    if (z.is_error(rc)) { std::string m = z.error_name(rc); z.free_d(ds); throw std::runtime_error(path + ": zstd: " + m); }
    out.append(buf.data(), dst.pos);
  }
  // drain what the decoder still holds for the last frame
  while (rc != 0) {
    ZBuf dst{buf.data(), buf.size(), 0};
    rc = z.decompress_stream(ds, &dst, &src);
    if (z.is_error(rc)) { std::string m = z.error_name(rc); z.free_d(ds); throw std::runtime_error(path + ": zstd: " + m); }
    out.append(buf.data(), dst.pos);
    if (dst.pos == 0 && rc != 0) { z.free_d(ds); throw std::runtime_error(path + ": zstd: unexpected EOF"); }
  }
  z.free_d(ds);
  return out;
}

void gzip_member(const char* p, size_t n, int level, std::string& out) {
  z_stream zs{};
  if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
    throw std::runtime_error("gzip: deflateInit failed");
  out.resize(deflateBound(&zs, static_cast<uLong>(n)) + 32);
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(p));
  zs.avail_in = static_cast<uInt>(n);
  zs.next_out = reinterpret_cast<Bytef*>(&out[0]);
  zs.avail_out = static_cast<uInt>(out.size());
  const int rc = deflate(&zs, Z_FINISH);
  const size_t produced = out.size() - zs.avail_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) throw std::runtime_error("gzip: deflate failed");
  out.resize(produced);
}

void zstd_frame(const char* p, size_t n, int level, std::string& out) {
  const Zstd& z = zstd();
  out.resize(z.bound(n));
  const size_t rc = z.compress(&out[0], out.size(), p, n, level);
  if (z.is_error(rc)) throw std::runtime_error(std::string("zstd: ") + z.error_name(rc));
  out.resize(rc);
}

constexpr size_t kBlock = size_t{4} << 20;  // uncompressed bytes per member / frame

}  // namespace

std::string read_file(const std::string& path) {
  std::string raw = read_raw(path);
  switch (codec_of(path)) {
    case kGzip: return gunzip(raw, path);
    case kZstd: return unzstd(raw, path);
    default: return raw;
  }
}

struct Writer::Impl {
  FILE* f = nullptr;
  bool own = false;
  Codec codec = kPlain;
  int threads = 1, level = 0;
  bool wrote_member = false;
  std::string pending;
  std::string path;

  void put(const char* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error(std::string("write: ") + strerror(errno));
  }
  // compresses pending[0, upto) as ceil(upto / kBlock) members and writes them in order
  void flush_blocks(size_t upto) {
    const size_t nb = (upto + kBlock - 1) / kBlock;
    if (nb == 0) return;
    std::vector<std::string> outs(nb);
    std::vector<std::string> errs(nb);
    auto work = [&](size_t b) {
      const size_t lo = b * kBlock, hi = std::min(upto, lo + kBlock);
      try {
        if (codec == kGzip) gzip_member(pending.data() + lo, hi - lo, level, outs[b]);
        else zstd_frame(pending.data() + lo, hi - lo, level, outs[b]);
      } catch (const std::exception& e) { errs[b] = e.what(); }
    };
    const size_t nt = std::min<size_t>(static_cast<size_t>(threads), nb);
    if (nt <= 1) {
      for (size_t b = 0; b < nb; ++b) work(b);
    } else {
      std::vector<std::thread> th;
      for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&, t] { for (size_t b = t; b < nb; b += nt) work(b); });
      for (auto& x : th) x.join();
    }
    for (size_t b = 0; b < nb; ++b) {
      if (!errs[b].empty()) throw std::runtime_error(errs[b]);
      put(outs[b].data(), outs[b].size());
    }
    wrote_member = true;
    pending.erase(0, upto);
  }
};

Writer::Writer(const std::string& path, int threads) : p_(new Impl) {
  p_->path = path;
  p_->codec = codec_of(path);
  p_->threads = threads < 1 ? 1 : threads;
  p_->level = p_->codec == kGzip ? Z_DEFAULT_COMPRESSION : 3;  // the defaults of gzip.NewWriter / zstd
  if (p_->codec == kZstd) zstd();                               // fail before any output if the library is missing
  p_->f = path.empty() ? stdout : fopen(path.c_str(), "wb");
  p_->own = !path.empty();
  if (!p_->f) { std::string m = "open " + path + ": " + strerror(errno); delete p_; p_ = nullptr; throw std::runtime_error(m); }
}

Writer::~Writer() {
  if (!p_) return;
  if (p_->own && p_->f) fclose(p_->f);
  delete p_;
}

void Writer::write(const char* data, size_t n) {
  if (p_->codec == kPlain) { p_->put(data, n); return; }
  p_->pending.append(data, n);
  const size_t whole = p_->pending.size() / kBlock * kBlock;
  if (whole >= kBlock * static_cast<size_t>(p_->threads)) p_->flush_blocks(whole);
}

void Writer::close() {
  if (!p_->f) return;
  if (p_->codec != kPlain) {
    // an empty stream still gets its header and trailer, as gzip.Writer.Close does
    if (!p_->pending.empty()) p_->flush_blocks(p_->pending.size());
    else if (!p_->wrote_member) { p_->pending.assign(""); std::string m;
      if (p_->codec == kGzip) gzip_member("", 0, p_->level, m); else zstd_frame("", 0, p_->level, m);
      p_->put(m.data(), m.size()); }
  }
  const bool bad = p_->own ? fclose(p_->f) != 0 : fflush(p_->f) != 0;
  p_->f = nullptr;
  if (bad) throw std::runtime_error(std::string("write: ") + strerror(errno));
}

}  // namespace frchost

// C entry points for the Python tests (no GPU needed).
extern "C" {
const char* frch_last_error();
void frch_set_error(const char* m);

// Reads (and decodes by suffix) a whole file; malloc'd buffer, free with frch_free.
char* frch_read_file(const char* path, size_t* len) {
  try {
    std::string s = frchost::read_file(path);
    char* p = static_cast<char*>(malloc(s.size() + 1));
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
  } catch (const std::exception& e) { frch_set_error(e.what()); return nullptr; }
}

// Writes `n_chunks` chunks through a Writer (codec by suffix); 0 on success.
int frch_write_file(const char* path, const char* const* chunks, const size_t* sizes, int n_chunks, int threads) {
  try {
    frchost::Writer w(path, threads);
    for (int i = 0; i < n_chunks; ++i) w.write(chunks[i], sizes[i]);
    w.close();
    return 0;
  } catch (const std::exception& e) { frch_set_error(e.what()); return 1; }
}
}
