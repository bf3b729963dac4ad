#include "hostlib.hpp"

#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace frchost {

// ------------------------------------------------------------------- Newick
namespace {
bool nwk_delim(unsigned char c) {
  return c == '(' || c == ')' || c == ',' || c == ':' || c == ';' || c == '[' || c == ']';
}
bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
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
        while (i < len && !nwk_delim(static_cast<unsigned char>(s[i])) && !is_space(static_cast<unsigned char>(s[i]))) ++i;
        std::string tok(s + st, i - st);
        char* end = nullptr;
        double d = strtod(tok.c_str(), &end);
        if (tok.empty() || *end) throw std::runtime_error("newick: bad branch length \"" + tok + "\"");
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
        while (i < len && !nwk_delim(static_cast<unsigned char>(s[i])) && !is_space(static_cast<unsigned char>(s[i]))) ++i;
        t.name[cur].assign(s + st, i - st);
      }
    }
  }
  if (!done) throw std::runtime_error("no tree in the given file");
  if (cur != 0) throw std::runtime_error("newick: unbalanced '('");
  return t;
}

// ------------------------------------------------------------------- tables
namespace {
// Go regexp \s = [\t\n\f\r ]; parser.go:17 tokenises with \S+.
bool go_space(unsigned char c) { return c == '\t' || c == '\n' || c == '\f' || c == '\r' || c == ' '; }

bool parse_float(const char* s, size_t n, double* out) {
  if (n == 0) return false;
  unsigned char c0 = static_cast<unsigned char>(s[0]);
  if (!(std::isdigit(c0) || c0 == '+' || c0 == '-' || c0 == '.' || c0 == 'i' || c0 == 'I' || c0 == 'n' || c0 == 'N'))
    return false;
  std::string tok(s, n);
  char* end = nullptr;
  errno = 0;
  double v = strtod(tok.c_str(), &end);
  if (end == tok.c_str() || *end) return false;
  if (errno == ERANGE && std::isinf(v)) return false;  // Go: value out of range
  *out = v;
  return true;
}

struct Interner {
  std::unordered_map<std::string, int32_t> ids;
  std::vector<std::string>* names;
  int32_t get(const char* s, size_t n) {
    std::string k(s, n);
    auto it = ids.find(k);
    if (it != ids.end()) return it->second;
    int32_t id = static_cast<int32_t>(names->size());
    names->push_back(k);
    ids.emplace(std::move(k), id);
    return id;
  }
};

constexpr size_t kMaxRow = size_t(1) << 25;  // sc.Buffer(nil, 1<<25), parser.go:145
}  // namespace

Table parse_table(const char* text, size_t len, bool sparse) {
  Table t;
  Interner in;
  in.names = &t.species;
  std::vector<int32_t> hdr;
  bool have_hdr = false;
  std::vector<int64_t> stamp, pos;  // per species: sample that last set it, and where
  size_t p = 0;
  int64_t rowno = 0;
  auto err = [&](int64_t k, const std::string& what) {
    return std::runtime_error("row #" + std::to_string(rowno) + ": value #" + std::to_string(k) + ": " + what);
  };
  while (p < len) {
    size_t e = p;
    while (e < len && text[e] != '\n') ++e;
    size_t le = e;
    if (le > p && text[le - 1] == '\r') --le;  // bufio.ScanLines drops one trailing \r
    if (le - p >= kMaxRow) throw std::runtime_error("bufio.Scanner: token too long");
    ++rowno;
    const char* row = text + p;
    const size_t rl = le - p;
    p = e + 1;
    if (!sparse && !have_hdr) {
      size_t i = 0;
      while (i < rl) {
        while (i < rl && go_space(static_cast<unsigned char>(row[i]))) ++i;
        size_t st = i;
        while (i < rl && !go_space(static_cast<unsigned char>(row[i]))) ++i;
        if (i > st) hdr.push_back(in.get(row + st, i - st));
      }
      if (hdr.empty()) throw std::runtime_error("row #1 has 0 values");
      have_hdr = true;
      continue;
    }
    const int64_t cur = t.n_samples() + 1;
    const size_t row_begin = t.sp.size();
    int64_t k = 0;
    size_t i = 0;
    if (!sparse) {  // parseRow checks the count before any value (parser.go:61-64)
      int64_t cnt = 0;
      size_t q = 0;
      while (q < rl) {
        while (q < rl && go_space(static_cast<unsigned char>(row[q]))) ++q;
        size_t st = q;
        while (q < rl && !go_space(static_cast<unsigned char>(row[q]))) ++q;
        if (q > st) ++cnt;
      }
      if (cnt != static_cast<int64_t>(hdr.size()))
        throw std::runtime_error("row #" + std::to_string(rowno) + ": has " + std::to_string(cnt) +
                                 " values, expected " + std::to_string(hdr.size()));
    }
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
        for (size_t q = 0; q < tl; ++q) if (tok[q] == ':') last = q;
        if (last == std::string::npos) throw err(k, "no colon in \"" + std::string(tok, tl) + "\"");
        if (last == 0) throw err(k, "empty species name");
        if (!parse_float(tok + last + 1, tl - last - 1, &v))
          throw err(k, "strconv.ParseFloat: parsing \"" + std::string(tok + last + 1, tl - last - 1) + "\": invalid syntax");
        if (std::isnan(v) || std::isinf(v) || v < 0) throw err(k, "bad value: " + std::to_string(v));
        if (v == 0) throw err(k, "zeros are not allowed in sparse format");
        sp = in.get(tok, last);
      } else {
        if (!parse_float(tok, tl, &v))
          throw err(k, "strconv.ParseFloat: parsing \"" + std::string(tok, tl) + "\": invalid syntax");
        if (std::isnan(v) || std::isinf(v) || v < 0) throw err(k, "bad value: " + std::to_string(v));
        if (v == 0) continue;
        sp = hdr[k - 1];
      }
      if (static_cast<size_t>(sp) >= stamp.size()) { stamp.resize(t.species.size(), 0); pos.resize(t.species.size(), 0); }
      if (stamp[sp] == cur) {
        t.val[pos[sp]] = v;  // m[name] = f : last assignment wins (parser.go:78,124)
      } else {
        stamp[sp] = cur;
        pos[sp] = static_cast<int64_t>(t.sp.size());
        t.sp.push_back(sp);
        t.val.push_back(v);
      }
    }
    (void)row_begin;
    t.row_ptr.push_back(static_cast<int64_t>(t.sp.size()));
  }
  return t;
}

// ------------------------------------------------------------------ resolve
Csr resolve(const Table& t, const FlatTree& tree) {
  const int32_t B = static_cast<int32_t>(tree.parent.size());
  // treeNames(): any node name validates (unifrac.go:70-76); only leaves carry
  // abundance (unifrac.go:38-43).
  std::unordered_map<std::string, std::vector<int32_t>> leaves_of;
  leaves_of.reserve(B * 2);
  for (int32_t v = 0; v < B; ++v) {
    auto& l = leaves_of[tree.name[v]];
    if (tree.n_children[v] == 0) l.push_back(v);
  }
  std::vector<const std::vector<int32_t>*> sp_leaves(t.species.size(), nullptr);
  std::vector<char> known(t.species.size(), 0);
  for (size_t s = 0; s < t.species.size(); ++s) {
    auto it = leaves_of.find(t.species[s]);
    if (it != leaves_of.end()) { known[s] = 1; sp_leaves[s] = &it->second; }
  }
  Csr c;
  c.row_ptr.push_back(0);
  for (int64_t r = 0; r < t.n_samples(); ++r) {
    for (int64_t k = t.row_ptr[r]; k < t.row_ptr[r + 1]; ++k) {
      const int32_t sp = t.sp[k];
      if (!known[sp]) {
        char vb[40];
        format_go(t.val[k], vb);
        throw std::runtime_error("sample #" + std::to_string(r + 1) + " has value " + vb + " for species \"" +
                                 t.species[sp] + "\" which is not in the tree");
      }
      for (int32_t leaf : *sp_leaves[sp]) { c.col.push_back(leaf); c.val.push_back(t.val[k]); }
    }
    c.row_ptr.push_back(static_cast<int64_t>(c.col.size()));
  }
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

void append_lines(const double* d, int64_t n, std::string& out) {
  char b[40];
  for (int64_t k = 0; k < n; ++k) {
    int l = format_go(d[k], b);
    b[l] = '\n';
    out.append(b, static_cast<size_t>(l) + 1);
  }
}

void format_lines_parallel(const double* d, int64_t n, int threads, std::string& out) {
  if (threads < 1) threads = 1;
  if (threads == 1 || n < 4096) { append_lines(d, n, out); return; }
  threads = static_cast<int>(std::min<int64_t>(threads, n / 2048));
  std::vector<std::string> parts(threads);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    const int64_t b = n * t / threads, e = n * (t + 1) / threads;
    th.emplace_back([&, t, b, e] {
      parts[t].reserve(static_cast<size_t>(e - b) * 20);
      append_lines(d + b, e - b, parts[t]);
    });
  }
  size_t total = out.size();
  for (int t = 0; t < threads; ++t) { th[t].join(); total += parts[t].size(); }
  out.reserve(total);
  for (auto& p : parts) out.append(p);
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

frch_tree* frch_tree_parse(const char* text, size_t len) {
  try { return new frch_tree{frchost::parse_newick(text, len)}; }
  catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void frch_tree_free(frch_tree* t) { delete t; }
int32_t frch_tree_size(const frch_tree* t) { return static_cast<int32_t>(t->t.parent.size()); }
const int32_t* frch_tree_parent(const frch_tree* t) { return t->t.parent.data(); }
const double* frch_tree_length(const frch_tree* t) { return t->t.length.data(); }
const char* frch_tree_name(const frch_tree* t, int32_t v) { return t->t.name[v].c_str(); }

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
void frch_free(void* p) { free(p); }

}  // extern "C"
