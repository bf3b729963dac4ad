#include "hostlib.hpp"

#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string_view>
#include <thread>

namespace frchost {

// ------------------------------------------------------------------- Newick
namespace {
// character classes of the Newick scanner: 1 = structural delimiter, 2 = white space
struct NwkLut {
  uint8_t t[256] = {};
  constexpr NwkLut() {
    for (char c : {'(', ')', ',', ':', ';', '[', ']'}) t[static_cast<unsigned char>(c)] = 1;
    for (char c : {' ', '\t', '\n', '\r', '\f', '\v'}) t[static_cast<unsigned char>(c)] = 2;
  }
};
constexpr NwkLut kNwk;
inline bool is_space(unsigned char c) { return kNwk.t[c] == 2; }
inline bool nwk_token_end(unsigned char c) { return kNwk.t[c] != 0; }  // delimiter or space
}  // namespace

FlatTree parse_newick(const char* s, size_t len) {
  FlatTree t;
  auto add = [&](int32_t parent) {
    t.parent.push_back(parent);
    t.length.push_back(0.0);
    t.name.emplace_back();
    t.n_children.push_back(0);
    if (parent >= 0) t.n_children[parent]++;
    return static_cast<int32_t>(t.parent.size()) - 1;
  };
  // A node gets its id when it is opened, which is pre-order: "(" opens the
  // first child of the current node, "," opens the next sibling.
  {
    // one node per ',' or '(' (+ the root): reserving spares the name vector its reallocation moves
    size_t nodes = 1;
    for (size_t k = 0; k < len; ++k) nodes += (s[k] == ',') | (s[k] == '(');
    t.parent.reserve(nodes); t.length.reserve(nodes); t.name.reserve(nodes); t.n_children.reserve(nodes);
  }
  int32_t cur = add(-1);
  size_t i = 0;
  bool done = false;
  while (i < len && !done) {
    unsigned char c = static_cast<unsigned char>(s[i]);
    if (is_space(c)) { ++i; continue; }
    switch (c) {
      case '[':
        while (i < len && s[i] != ']') ++i;
        if (i < len) ++i;
        break;
      case '(': cur = add(cur); ++i; break;
      case ',':
        if (t.parent[cur] < 0) throw std::runtime_error("newick: ',' outside parentheses");
        cur = add(t.parent[cur]); ++i; break;
      case ')':
        if (t.parent[cur] < 0) throw std::runtime_error("newick: unbalanced ')'");
        cur = t.parent[cur]; ++i; break;
      case ';': done = true; ++i; break;
      case ':': {
        ++i;
        while (i < len && is_space(static_cast<unsigned char>(s[i]))) ++i;
        size_t st = i;
        while (i < len && !nwk_token_end(static_cast<unsigned char>(s[i]))) ++i;
        // plain decimals (every tree a program writes) through from_chars: correctly rounded like strtod,
        // no allocation, no locale; whatever it does not consume entirely takes the strtod path as before
        double d = 0.0;
        const unsigned char c0 = i > st ? static_cast<unsigned char>(s[st]) : 0;
        const unsigned char c1 = i > st + 1 ? static_cast<unsigned char>(s[st + 1]) : 0;
        auto r = std::from_chars(s + st, s + i, d);
        const bool plain = (std::isdigit(c0) || c0 == '.' || (c0 == '-' && (std::isdigit(c1) || c1 == '.'))) &&
                           r.ec == std::errc() && r.ptr == s + i;
        if (!plain) {
          std::string tok(s + st, i - st);
          char* end = nullptr;
          d = strtod(tok.c_str(), &end);
          if (tok.empty() || *end) throw std::runtime_error("newick: bad branch length \"" + tok + "\"");
        }
        t.length[cur] = d;
        break;
      }
      case '\'': {
        std::string nm;
        ++i;
        for (;;) {
          if (i >= len) throw std::runtime_error("newick: unterminated quoted label");
          if (s[i] == '\'') {
            if (i + 1 < len && s[i + 1] == '\'') ++i; else { ++i; break; }
          }
          nm.push_back(s[i++]);
        }
        t.name[cur] = nm;
        break;
      }
      default: {
        size_t st = i;
        while (i < len && !nwk_token_end(static_cast<unsigned char>(s[i]))) ++i;
        t.name[cur].assign(s + st, i - st);
      }
    }
  }
  if (!done) throw std::runtime_error("no tree in the given file");
  if (cur != 0) throw std::runtime_error("newick: unbalanced '('");
  return t;
}

// ------------------------------------------------------------------- tables
// Map-free, parallel restatement of parser.ParseAbundance / ParseSparseAbundance
// (parser/parser.go:21-140): rows are independent, so the text is cut at line boundaries into one
// chunk per worker (the reference hands rows to `ngoroutines` workers through ppln.Serial the same
// way, parser.go:24,87), every worker parses its rows straight into CSR fragments without a
// per-sample map, a per-token string or a regexp, and the fragments are joined in file order.
// Every accept / reject rule and the text of the first error (lowest row) are kept.
namespace {
// Go regexp \s = [\t\n\f\r ]; parser.go:17 tokenises with \S+.
struct SpaceLut {
  bool t[256] = {};
  constexpr SpaceLut() { t['\t'] = t['\n'] = t['\f'] = t['\r'] = t[' '] = true; }
};
constexpr SpaceLut kSpace;
inline bool go_space(unsigned char c) { return kSpace.t[c]; }

// Syntax of strconv.ParseFloat beyond plain decimals (strconv/atof.go: special, readFloat, underscoreOK):
//   [+-] inf | infinity | nan (any case; nan unsigned)
//   [+-] digits [. digits] [e|E [+-] digits]          at least one mantissa digit, exponent digits required
//   [+-] 0x hexdigits [. hexdigits] p|P [+-] digits    the p exponent is REQUIRED
//   underscores may separate digits (and follow the base prefix), nowhere else ("1_000", not "1__0", "_1", "1_")
// Returns false for anything else (what strtod would still take: "0x1A", "nan(1)", "1e", leading space ...);
// on success `clean` holds the token without its underscores, ready for strtod.
bool go_float_syntax(const char* s, size_t n, std::string& clean) {
  auto lower = [](char c) { return static_cast<char>(c | 0x20); };
  size_t i = 0;
  if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
  const size_t body = i;
  auto ieq = [&](const char* w) {
    size_t k = 0;
    for (; w[k]; ++k)
      if (body + k >= n || lower(s[body + k]) != w[k]) return false;
    return body + k == n;
  };
  if (ieq("inf") || ieq("infinity") || (body == 0 && ieq("nan"))) { clean.assign(s, n); return true; }
  bool hex = false;
  if (i + 1 < n && s[i] == '0' && lower(s[i + 1]) == 'x') { hex = true; i += 2; }
  // underscore placement: '^' start, '0' digit or base prefix, '_' underscore, '!' anything else
  char saw = hex ? '0' : '^';
  bool digits = false, dot = false, exp = false, underscores = false;
  for (; i < n; ++i) {
    const char c = s[i];
    const bool dig = (c >= '0' && c <= '9') || (hex && !exp && lower(c) >= 'a' && lower(c) <= 'f');
    if (dig) { digits = true; saw = '0'; continue; }
    if (c == '_') { if (saw != '0') return false; saw = '_'; underscores = true; continue; }
    if (saw == '_') return false;
    saw = '!';
    if (c == '.' && !dot && !exp) { dot = true; continue; }
    if (!exp && digits && (hex ? lower(c) == 'p' : lower(c) == 'e')) {
      exp = true;
      if (i + 1 < n && (s[i + 1] == '+' || s[i + 1] == '-')) ++i;
      if (i + 1 >= n || s[i + 1] < '0' || s[i + 1] > '9') return false;  // exponent digits required
      continue;
    }
    return false;
  }
  if (!digits || saw == '_' || (hex && !exp)) return false;
  clean.assign(s, n);
  if (underscores) clean.erase(std::remove(clean.begin(), clean.end(), '_'), clean.end());
  return true;
}

// strconv.ParseFloat(s, 64).  Fast path: std::from_chars (correctly rounded, no allocation, no
// locale) for plain decimals; every other spelling is checked against Go's syntax (go_float_syntax)
// and then converted by strtod.
bool parse_float(const char* s, size_t n, double* out) {
  if (n == 0) return false;
  if (n == 1 && s[0] >= '0' && s[0] <= '9') { *out = s[0] - '0'; return true; }  // dense tables are mostly "0"
  {
    double v;
    auto r = std::from_chars(s, s + n, v);  // general format; rejects a leading '+', underscores, hex
    if (r.ec == std::errc() && r.ptr == s + n && !(s[0] == 'i' || s[0] == 'I' || s[0] == 'n' || s[0] == 'N') &&
        !(n > 1 && (s[1] == 'i' || s[1] == 'I' || s[1] == 'n' || s[1] == 'N'))) {
      *out = v;
      return true;
    }
    if (r.ec == std::errc::result_out_of_range && r.ptr == s + n) {
      // Go: 1e999 is an error (value out of range), 1e-999 parses to 0
      bool neg_exp = false;
      for (size_t q = 0; q + 1 < n; ++q)
        if ((s[q] == 'e' || s[q] == 'E') && s[q + 1] == '-') neg_exp = true;
      if (!neg_exp) return false;
    }
  }
  std::string clean;
  if (!go_float_syntax(s, n, clean)) return false;
  char* end = nullptr;
  errno = 0;
  const double v = strtod(clean.c_str(), &end);
  if (end == clean.c_str() || *end) return false;
  if (errno == ERANGE && std::isinf(v)) return false;  // Go: value out of range
  *out = v;
  return true;
}

// name -> id, open addressing (FNV-1a, linear probing); ids in order of first appearance.
struct Interner {
  struct Slot { uint64_t h; int32_t id; };
  std::vector<Slot> slots = std::vector<Slot>(1024, Slot{0, -1});
  std::vector<std::string>* names;
  static uint64_t hash(const char* s, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= static_cast<unsigned char>(s[i]); h *= 1099511628211ull; }
    return h | 1;
  }
  void grow() {
    std::vector<Slot> old(slots.size() * 2, Slot{0, -1});
    old.swap(slots);
    for (const Slot& o : old)
      if (o.id >= 0) {
        size_t k = o.h & (slots.size() - 1);
        while (slots[k].id >= 0) k = (k + 1) & (slots.size() - 1);
        slots[k] = o;
      }
  }
  int32_t get(const char* s, size_t n) {
    const uint64_t h = hash(s, n);
    size_t k = h & (slots.size() - 1);
    while (slots[k].id >= 0) {
      if (slots[k].h == h) {
        const std::string& nm = (*names)[slots[k].id];
        if (nm.size() == n && memcmp(nm.data(), s, n) == 0) return slots[k].id;
      }
      k = (k + 1) & (slots.size() - 1);
    }
    const int32_t id = static_cast<int32_t>(names->size());
    names->emplace_back(s, n);
    slots[k] = Slot{h, id};
    if (names->size() * 2 > slots.size()) grow();
    return id;
  }
};

constexpr size_t kMaxRow = size_t(1) << 25;  // sc.Buffer(nil, 1<<25), parser.go:145

struct Fragment {                      // what one worker produced for its run of rows
  std::vector<std::string> species;    // sparse: names in order of first appearance in this chunk
  std::vector<int64_t> row_len;
  std::vector<int32_t> sp;
  std::vector<double> val;
  int64_t err_row = -1;                // 1-based row number of the first failing row (-1: none)
  std::string err;
};

// Parses the rows of text[begin, end) (begin at a line start); row numbers start at first_row + 1.
void parse_rows(const char* text, size_t begin, size_t end, bool sparse, const std::vector<int32_t>& hdr,
                int64_t first_row, Fragment* f) {
  Interner in;
  in.names = &f->species;
  std::vector<int64_t> stamp, pos;  // per species: local sample that last set it, and where
  if (!sparse) { stamp.assign(hdr.empty() ? 0 : *std::max_element(hdr.begin(), hdr.end()) + 1, 0); pos = stamp; }
  size_t p = begin;
  int64_t rowno = first_row, cur = 0;
  auto fail = [&](const std::string& msg) { f->err_row = rowno; f->err = msg; };
  auto verr = [&](int64_t k, const std::string& what) {
    fail("row #" + std::to_string(rowno) + ": value #" + std::to_string(k) + ": " + what);
  };
  while (p < end) {
    size_t e = p;
    while (e < end && text[e] != '\n') ++e;
    size_t le = e;
    if (le > p && text[le - 1] == '\r') --le;  // bufio.ScanLines drops one trailing \r
    ++rowno;
    if (le - p >= kMaxRow) { fail("bufio.Scanner: token too long"); return; }
    const char* row = text + p;
    const size_t rl = le - p;
    p = e + 1;
    ++cur;
    const size_t row_begin = f->sp.size();
    // dense: parseRow checks the token count before any value (parser.go:61-64); here the row is
    // tokenised once, a value error is held back until the count is known to be right
    std::string held;
    int64_t k = 0;
    size_t i = 0;
    while (i < rl) {
      while (i < rl && go_space(static_cast<unsigned char>(row[i]))) ++i;
      size_t st = i;
      while (i < rl && !go_space(static_cast<unsigned char>(row[i]))) ++i;
      if (i == st) break;
      ++k;
      const char* tok = row + st;
      const size_t tl = i - st;
      double v;
      int32_t sp;
      if (sparse) {
        size_t last = std::string::npos;  // split on the LAST colon (parser.go:129-140)
        for (size_t q = tl; q-- > 0;) if (tok[q] == ':') { last = q; break; }
        if (last == std::string::npos) { verr(k, "no colon in \"" + std::string(tok, tl) + "\""); return; }
        if (last == 0) { verr(k, "empty species name"); return; }
        if (!parse_float(tok + last + 1, tl - last - 1, &v)) {
          verr(k, "strconv.ParseFloat: parsing \"" + std::string(tok + last + 1, tl - last - 1) + "\": invalid syntax");
          return;
        }
        if (std::isnan(v) || std::isinf(v) || v < 0) { verr(k, "bad value: " + std::to_string(v)); return; }
        if (v == 0) { verr(k, "zeros are not allowed in sparse format"); return; }
        sp = in.get(tok, last);
        if (static_cast<size_t>(sp) >= stamp.size()) { stamp.resize(f->species.size() * 2 + 16, 0); pos.resize(stamp.size(), 0); }
      } else {
        if (tl == 1 && tok[0] == '0') continue;  // zeros are dropped (parser.go:75-77)
        if (!held.empty() || k > static_cast<int64_t>(hdr.size())) continue;  // only counting from here on
        if (!parse_float(tok, tl, &v)) {
          held = "row #" + std::to_string(rowno) + ": value #" + std::to_string(k) +
                 ": strconv.ParseFloat: parsing \"" + std::string(tok, tl) + "\": invalid syntax";
          continue;
        }
        if (std::isnan(v) || std::isinf(v) || v < 0) {
          held = "row #" + std::to_string(rowno) + ": value #" + std::to_string(k) + ": bad value: " + std::to_string(v);
          continue;
        }
        if (v == 0) continue;
        sp = hdr[k - 1];
      }
      if (stamp[sp] == cur) {
        f->val[pos[sp]] = v;  // m[name] = f : last assignment wins (parser.go:78,124)
      } else {
        stamp[sp] = cur;
        pos[sp] = static_cast<int64_t>(f->sp.size());
        f->sp.push_back(sp);
        f->val.push_back(v);
      }
    }
    if (!sparse) {
      if (k != static_cast<int64_t>(hdr.size())) {
        fail("row #" + std::to_string(rowno) + ": has " + std::to_string(k) + " values, expected " +
             std::to_string(hdr.size()));
        return;
      }
      if (!held.empty()) { fail(held); return; }
    }
    f->row_len.push_back(static_cast<int64_t>(f->sp.size() - row_begin));
  }
}

template <class F>
void parallel_for(int n, F&& fn) {
  if (n <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < n; ++t) th.emplace_back([&fn, t] { fn(t); });
  fn(0);
  for (auto& x : th) x.join();
}
}  // namespace

Table parse_table(const char* text, size_t len, bool sparse, int threads) {
  Table t;
  std::vector<int32_t> hdr;
  size_t body = 0;
  int64_t first_row = 0;
  if (!sparse) {  // the first row names the species (parser.go:33-38)
    Interner in;
    in.names = &t.species;
    size_t e = 0;
    while (e < len && text[e] != '\n') ++e;
    size_t le = e;
    if (le > 0 && text[le - 1] == '\r') --le;
    if (le >= kMaxRow) throw std::runtime_error("bufio.Scanner: token too long");
    size_t i = 0;
    while (i < le) {
      while (i < le && go_space(static_cast<unsigned char>(text[i]))) ++i;
      size_t st = i;
      while (i < le && !go_space(static_cast<unsigned char>(text[i]))) ++i;
      if (i > st) hdr.push_back(in.get(text + st, i - st));
    }
    if (len == 0) return t;
    if (hdr.empty()) throw std::runtime_error("row #1 has 0 values");
    body = std::min(len, e + 1);
    first_row = 1;
  }
  // chunks of whole rows, about equal in bytes
  int T = std::max(1, threads);
  T = static_cast<int>(std::min<size_t>(T, (len - body) / (1 << 16) + 1));
  std::vector<size_t> cut(T + 1, len);
  cut[0] = body;
  for (int k = 1; k < T; ++k) {
    size_t c = body + (len - body) * k / T;
    c = std::max(c, cut[k - 1]);
    while (c < len && c > body && text[c - 1] != '\n') ++c;
    cut[k] = c;
  }
  // global row numbers: rows before each chunk
  std::vector<int64_t> rows_before(T + 1, first_row);
  {
    std::vector<int64_t> cnt(T, 0);
    parallel_for(T, [&](int k) {
      int64_t c = 0;
      const char* p = text + cut[k];
      const char* e = text + cut[k + 1];
      while (p < e) {
        const void* q = memchr(p, '\n', e - p);
        ++c;
        if (!q) break;
        p = static_cast<const char*>(q) + 1;
      }
      cnt[k] = c;
    });
    for (int k = 0; k < T; ++k) rows_before[k + 1] = rows_before[k] + cnt[k];
  }
  std::vector<Fragment> frag(T);
  parallel_for(T, [&](int k) { parse_rows(text, cut[k], cut[k + 1], sparse, hdr, rows_before[k], &frag[k]); });
  for (int k = 0; k < T; ++k)
    if (frag[k].err_row >= 0) throw std::runtime_error(frag[k].err);  // chunks are in file order: lowest row first
  // join: species ids of a sparse table are renumbered in order of first appearance in the file
  std::vector<std::vector<int32_t>> remap(T);
  if (sparse) {
    Interner in;
    in.names = &t.species;
    for (int k = 0; k < T; ++k) {
      remap[k].resize(frag[k].species.size());
      for (size_t s = 0; s < frag[k].species.size(); ++s)
        remap[k][s] = in.get(frag[k].species[s].data(), frag[k].species[s].size());
    }
  }
  std::vector<int64_t> nnz_before(T + 1, 0), nrow_before(T + 1, 0);
  for (int k = 0; k < T; ++k) {
    nnz_before[k + 1] = nnz_before[k] + static_cast<int64_t>(frag[k].sp.size());
    nrow_before[k + 1] = nrow_before[k] + static_cast<int64_t>(frag[k].row_len.size());
  }
  t.sp.resize(nnz_before[T]);
  t.val.resize(nnz_before[T]);
  t.row_ptr.assign(nrow_before[T] + 1, 0);
  parallel_for(T, [&](int k) {
    const Fragment& f = frag[k];
    int32_t* dsp = t.sp.data() + nnz_before[k];
    if (sparse) for (size_t q = 0; q < f.sp.size(); ++q) dsp[q] = remap[k][f.sp[q]];
    else if (!f.sp.empty()) memcpy(dsp, f.sp.data(), f.sp.size() * sizeof(int32_t));
    if (!f.val.empty()) memcpy(t.val.data() + nnz_before[k], f.val.data(), f.val.size() * sizeof(double));
    int64_t at = nnz_before[k];
    for (size_t r = 0; r < f.row_len.size(); ++r) {
      at += f.row_len[r];
      t.row_ptr[nrow_before[k] + r + 1] = at;
    }
  });
  return t;
}

// ------------------------------------------------------------------ resolve
Csr resolve(const Table& t, const FlatTree& tree, int threads) {
  const int32_t B = static_cast<int32_t>(tree.parent.size());
  // treeNames(): any node name validates (unifrac.go:70-76); only leaves carry
  // abundance (unifrac.go:38-43).
  std::unordered_map<std::string_view, std::vector<int32_t>> leaves_of;
  leaves_of.reserve(static_cast<size_t>(B) * 2);
  for (int32_t v = 0; v < B; ++v) {
    auto& l = leaves_of[std::string_view(tree.name[v])];
    if (tree.n_children[v] == 0) l.push_back(v);
  }
  std::vector<const std::vector<int32_t>*> sp_leaves(t.species.size(), nullptr);
  for (size_t s = 0; s < t.species.size(); ++s) {
    auto it = leaves_of.find(std::string_view(t.species[s]));
    if (it != leaves_of.end()) sp_leaves[s] = &it->second;
  }
  const int64_t N = t.n_samples();
  Csr c;
  c.row_ptr.assign(N + 1, 0);
  // pass 1: validate + count (first error by sample order), pass 2: fill
  int T = std::max(1, std::min<int>(threads, static_cast<int>(N / 64 + 1)));
  std::vector<int64_t> bad_row(T, -1), bad_k(T, -1);
  parallel_for(T, [&](int k) {
    for (int64_t r = N * k / T; r < N * (k + 1) / T; ++r) {
      int64_t n = 0;
      for (int64_t e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) {
        const auto* l = sp_leaves[t.sp[e]];
        if (!l) { if (bad_row[k] < 0) { bad_row[k] = r; bad_k[k] = e; } continue; }
        n += static_cast<int64_t>(l->size());
      }
      c.row_ptr[r + 1] = n;
    }
  });
  for (int k = 0; k < T; ++k)
    if (bad_row[k] >= 0) {
      char vb[40];
      format_go(t.val[bad_k[k]], vb);
      throw std::runtime_error("sample #" + std::to_string(bad_row[k] + 1) + " has value " + vb + " for species \"" +
                               t.species[t.sp[bad_k[k]]] + "\" which is not in the tree");
    }
  for (int64_t r = 0; r < N; ++r) c.row_ptr[r + 1] += c.row_ptr[r];
  c.col.resize(c.row_ptr[N]);
  c.val.resize(c.row_ptr[N]);
  parallel_for(T, [&](int k) {
    for (int64_t r = N * k / T; r < N * (k + 1) / T; ++r) {
      int64_t at = c.row_ptr[r];
      for (int64_t e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e)
        for (int32_t leaf : *sp_leaves[t.sp[e]]) { c.col[at] = leaf; c.val[at] = t.val[e]; ++at; }
    }
  });
  return c;
}

// ---------------------------------------------------------------- formatting
// fmt %v on a float64 = strconv.FormatFloat(f, 'g', -1, 64): shortest digits
// that round-trip; %e layout when the decimal exponent is < -4 or >= 6 (the
// shortest-%g rule: fmt.Println(1e6) prints 1e+06), exponent of at least two
// digits.  All finite UniFrac results lie in [0, 1], so only the lower
// threshold is ever exercised.
int format_go(double v, char* buf) {
  if (std::isnan(v)) { memcpy(buf, "NaN", 4); return 3; }
  if (std::isinf(v)) { memcpy(buf, v > 0 ? "+Inf" : "-Inf", 5); return 4; }
  if (v == 0) {
    if (std::signbit(v)) { memcpy(buf, "-0", 3); return 2; }
    memcpy(buf, "0", 2);
    return 1;
  }
  char sci[40];
  auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);  // shortest round-trip
  *r.ptr = 0;
  // sci = [-]d[.ddd]e[+-]XX
  char digits[24] = {'0'};
  int nd = 0;
  const char* q = sci;
  bool neg = false;
  if (*q == '-') { neg = true; ++q; }
  for (; *q && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
  int x = atoi(q + 1);
  int o = 0;
  if (neg) buf[o++] = '-';
  if (x < -4 || x >= 6) {
    buf[o++] = digits[0];
    if (nd > 1) { buf[o++] = '.'; for (int i = 1; i < nd; ++i) buf[o++] = digits[i]; }
    buf[o++] = 'e';
    buf[o++] = x < 0 ? '-' : '+';
    int ax = x < 0 ? -x : x;
    if (ax < 10) buf[o++] = '0';
    auto rr = std::to_chars(buf + o, buf + o + 8, ax);
    o = static_cast<int>(rr.ptr - buf);
  } else if (x < 0) {
    buf[o++] = '0'; buf[o++] = '.';
    for (int i = 0; i < -x - 1; ++i) buf[o++] = '0';
    for (int i = 0; i < nd; ++i) buf[o++] = digits[i];
  } else {
    for (int i = 0; i <= x; ++i) buf[o++] = i < nd ? digits[i] : '0';
    if (nd > x + 1) { buf[o++] = '.'; for (int i = x + 1; i < nd; ++i) buf[o++] = digits[i]; }
  }
  buf[o] = 0;
  return o;
}

template <class T>
static void append_lines_t(const T* d, int64_t n, std::string& out) {
  // formatted in blocks on the stack: one append per ~200 values instead of one per line
  char b[8192];
  size_t o = 0;
  for (int64_t k = 0; k < n; ++k) {
    if (o + 40 > sizeof b) { out.append(b, o); o = 0; }
    o += static_cast<size_t>(format_go(static_cast<double>(d[k]), b + o));  // (float: widened exactly, as Go's float64(f))
    b[o++] = '\n';
  }
  out.append(b, o);
}

void append_lines(const double* d, int64_t n, std::string& out) { append_lines_t(d, n, out); }

template <class T>
static void format_parts_parallel_t(const T* d, int64_t n, int threads, std::vector<std::string>& parts) {
  if (threads < 1) threads = 1;
  if (n < 4096) threads = 1;
  threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, n / 2048)));
  if (static_cast<int>(parts.size()) < threads) parts.resize(threads);
  for (auto& p : parts) p.clear();  // (capacity is kept: a caller that reuses `parts` allocates once)
  if (threads == 1) { append_lines_t(d, n, parts[0]); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    const int64_t b = n * t / threads, e = n * (t + 1) / threads;
    th.emplace_back([&, t, b, e] {
      parts[t].reserve(static_cast<size_t>(e - b) * 20);
      append_lines_t(d + b, e - b, parts[t]);
    });
  }
  for (auto& x : th) x.join();
}

void format_parts_parallel(const double* d, int64_t n, int threads, std::vector<std::string>& parts) {
  format_parts_parallel_t(d, n, threads, parts);
}

void format_parts_parallel_f32(const float* d, int64_t n, int64_t first_index, const int64_t* ex_index,
                               const double* ex_value, int64_t n_ex, int threads, std::vector<std::string>& parts) {
  if (n_ex <= 0) { format_parts_parallel_t(d, n, threads, parts); return; }
  // (rare: distances below fp32's range travel beside the band as doubles)
  std::vector<double> wide(d, d + n);
  for (int64_t k = 0; k < n_ex; ++k) {
    const int64_t at = ex_index[k] - first_index;
    if (at >= 0 && at < n) wide[at] = ex_value[k];
  }
  format_parts_parallel_t(wide.data(), n, threads, parts);
}

void format_lines_parallel(const double* d, int64_t n, int threads, std::string& out) {
  if (threads <= 1 || n < 4096) { append_lines(d, n, out); return; }
  std::vector<std::string> parts;
  format_parts_parallel(d, n, threads, parts);
  size_t total = out.size();
  for (auto& p : parts) total += p.size();
  out.reserve(total);
  for (auto& p : parts) out.append(p);
}

// sprspr's toSparse (sprspr/sprspr.go:19-36): one line per sample, "name:%g" joined by tabs.  The reference
// ranges over a Go map (random order, its test sorts the fields before comparing: sprspr_test.go:26-31); here the
// fields come in first-assignment order, i.e. the header's column order.
void to_sparse_lines(const Table& t, std::string& out) {
  char num[40];
  for (int64_t s = 0; s < t.n_samples(); ++s) {
    for (int64_t k = t.row_ptr[s]; k < t.row_ptr[s + 1]; ++k) {
      if (k > t.row_ptr[s]) out.push_back('\t');
      out.append(t.species[t.sp[k]]);
      out.push_back(':');
      out.append(num, static_cast<size_t>(format_go(t.val[k], num)));  // %g and %v print a float64 alike
    }
    out.push_back('\n');
  }
}

}  // namespace frchost

// -------------------------------------------------------------------- C ABI
// For the Python tests of the host logic (tests run it without a GPU).
extern "C" {

struct frch_tree { frchost::FlatTree t; };
struct frch_table { frchost::Table t; };
struct frch_csr { frchost::Csr c; };

static thread_local std::string g_err;
const char* frch_last_error() { return g_err.c_str(); }
void frch_set_error(const char* m) { g_err = m; }

frch_tree* frch_tree_parse(const char* text, size_t len) {
  try { return new frch_tree{frchost::parse_newick(text, len)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void frch_tree_free(frch_tree* t) { delete t; }
int32_t frch_tree_size(const frch_tree* t) { return static_cast<int32_t>(t->t.parent.size()); }
const int32_t* frch_tree_parent(const frch_tree* t) { return t->t.parent.data(); }
const double* frch_tree_length(const frch_tree* t) { return t->t.length.data(); }
const char* frch_tree_name(const frch_tree* t, int32_t v) { return t->t.name[v].c_str(); }

frch_table* frch_table_parse_mt(const char* text, size_t len, int sparse, int threads) {
  try { return new frch_table{frchost::parse_table(text, len, sparse != 0, threads)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
frch_csr* frch_resolve_mt(const frch_table* tab, const frch_tree* tree, int threads) {
  try { return new frch_csr{frchost::resolve(tab->t, tree->t, threads)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
frch_table* frch_table_parse(const char* text, size_t len, int sparse) {
  try { return new frch_table{frchost::parse_table(text, len, sparse != 0)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void frch_table_free(frch_table* t) { delete t; }
int64_t frch_table_samples(const frch_table* t) { return t->t.n_samples(); }
const int64_t* frch_table_row_ptr(const frch_table* t) { return t->t.row_ptr.data(); }
const int32_t* frch_table_species_ids(const frch_table* t) { return t->t.sp.data(); }
const double* frch_table_values(const frch_table* t) { return t->t.val.data(); }
const char* frch_table_species_name(const frch_table* t, int32_t id) { return t->t.species[id].c_str(); }

frch_csr* frch_resolve(const frch_table* tab, const frch_tree* tree) {
  try { return new frch_csr{frchost::resolve(tab->t, tree->t)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void frch_csr_free(frch_csr* c) { delete c; }
int64_t frch_csr_nnz(const frch_csr* c) { return static_cast<int64_t>(c->c.col.size()); }
const int64_t* frch_csr_row_ptr(const frch_csr* c) { return c->c.row_ptr.data(); }
const int32_t* frch_csr_col(const frch_csr* c) { return c->c.col.data(); }
const double* frch_csr_val(const frch_csr* c) { return c->c.val.data(); }

int frch_format_go(double v, char* buf) { return frchost::format_go(v, buf); }

// Formats n values, one per line; returns a malloc'd buffer (free with frch_free) and its length.
char* frch_format_lines(const double* d, int64_t n, int threads, size_t* len) {
  std::string out;
  frchost::format_lines_parallel(d, n, threads, out);
  char* p = static_cast<char*>(malloc(out.size() + 1));
  memcpy(p, out.data(), out.size());
  p[out.size()] = 0;
  *len = out.size();
  return p;
}
// The CLI's fp32 route: a float32 band (frc_next_f32) + its exceptions -> the same text.
char* frch_format_lines_f32(const float* d, int64_t n, int64_t first_index, const int64_t* ex_index, const double* ex_value,
                            int64_t n_ex, int threads, size_t* len) {
  std::vector<std::string> parts;
  frchost::format_parts_parallel_f32(d, n, first_index, ex_index, ex_value, n_ex, threads, parts);
  size_t total = 0;
  for (const std::string& x : parts) total += x.size();
  char* p = static_cast<char*>(malloc(total + 1));
  size_t o = 0;
  for (const std::string& x : parts) { memcpy(p + o, x.data(), x.size()); o += x.size(); }
  p[total] = 0;
  *len = total;
  return p;
}

void frch_free(void* p) { free(p); }

// Dense table text -> sparse table text (sprspr); malloc'd, free with frch_free; NULL + frch_last_error on bad input.
char* frch_to_sparse(const char* text, size_t len, int threads, size_t* out_len) {
  try {
    frchost::Table t = frchost::parse_table(text, len, /*sparse=*/false, threads);
    std::string out;
    frchost::to_sparse_lines(t, out);
    char* p = static_cast<char*>(malloc(out.size() + 1));
    memcpy(p, out.data(), out.size());
    p[out.size()] = 0;
    *out_len = out.size();
    return p;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

}  // extern "C"
