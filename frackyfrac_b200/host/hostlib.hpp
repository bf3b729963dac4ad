// Host side above the C ABI, in C++ because the reference host is compiled Go
// and this image has no Go toolchain.  It mirrors what frackyfrac keeps on the
// host: the Newick reader (third-party biostuff/formats/newick, used at
// frcfrc/frcfrc.go:109-114), the dense / sparse table readers
// (parser/parser.go:21-140), species validation (frcfrc/unifrac.go:70-93),
// the flattening enumerateNodes performs (unifrac.go:127-133), and the
// fmt.Fprintln writer (frcfrc.go:58-62).  No distance arithmetic lives here.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace frchost {

struct FlatTree {
  std::vector<int32_t> parent;     // pre-order ids, root = 0, parent[0] = -1
  std::vector<double> length;      // node.Distance (0 when absent)
  std::vector<std::string> name;   // "" when absent
  std::vector<int32_t> n_children;
};

// First tree of the text (frcfrc.go:109-114).  Throws std::runtime_error.
FlatTree parse_newick(const char* text, size_t len);

// One map per sample, as parser.ParseAbundance / ParseSparseAbundance deliver
// them, with species names interned.  Entry order = first assignment.
struct Table {
  std::vector<std::string> species;              // id -> name
  std::vector<int64_t> row_ptr{0};               // per sample
  std::vector<int32_t> sp;                       // species id
  std::vector<double> val;
  int64_t n_samples() const { return static_cast<int64_t>(row_ptr.size()) - 1; }
};
// `threads` workers parse runs of whole rows in parallel (the reference's ngoroutines, parser.go:21,85).
Table parse_table(const char* text, size_t len, bool sparse, int threads = 1);

// validateSpecies + name -> leaf resolution.  Output is the frc_csr_t payload:
// every leaf carrying the name gets the value (unifrac.go:38-43), names that
// only match internal nodes are dropped.  Throws with the reference's message
// (unifrac.go:85-88) when a species is not in the tree.
struct Csr {
  std::vector<int64_t> row_ptr;
  std::vector<int32_t> col;
  std::vector<double> val;
};
Csr resolve(const Table& t, const FlatTree& tree, int threads = 1);

// sprspr/sprspr.go:19-36: the table as sparse-format text, one "name:%g" field per non-zero entry.
void to_sparse_lines(const Table& t, std::string& out);

// Go's fmt %v for float64, no newline.  Returns the length written (<= 32).
int format_go(double v, char* buf);

// One distance per line, exactly as `fmt.Fprintln(w, f)`; appends to `out`.
void append_lines(const double* d, int64_t n, std::string& out);

// Same bytes, formatted by `threads` workers over contiguous slices and joined in order
// (the reference formats on one goroutine, frcfrc.go:58-62; at 10^9 pairs/s that is the
// end-to-end bottleneck).
void format_lines_parallel(const double* d, int64_t n, int threads, std::string& out);

// The same bytes as consecutive pieces, parts[0] + parts[1] + ...: the writer hands each piece to the file
// as it is, which spares a single-threaded concatenation of the whole band (250 MB of text at cfg2).
// Unused trailing entries of `parts` are left empty; reusing `parts` across calls reuses its buffers.
void format_parts_parallel(const double* d, int64_t n, int threads, std::vector<std::string>& parts);
// The same for a float32 band of the fast paths (frc_next_f32): every value is widened exactly (Go's float64(f))
// and printed as that double, i.e. the bytes `frc_next` + format_parts_parallel would give, without the widening
// pass over memory.  ex_* are the band's fp32 exceptions (frc_chunk_exceptions; usually none).
void format_parts_parallel_f32(const float* d, int64_t n, int64_t first_index, const int64_t* ex_index,
                               const double* ex_value, int64_t n_ex, int threads, std::vector<std::string>& parts);

// aio.Open stand-in (frcfrc.go:93,109): the whole file, decoded by suffix (".gz", ".zst"); "" = stdin.
std::string read_file(const std::string& path);

// aio.Create stand-in (frcfrc.go:100-106): codec by suffix; compressed output is produced by `threads`
// workers as independent members / frames written in order.  "" = stdout.  Throws on I/O errors.
class Writer {
 public:
  Writer(const std::string& path, int threads);
  ~Writer();
  Writer(const Writer&) = delete;
  Writer& operator=(const Writer&) = delete;
  void write(const char* data, size_t n);
  void close();

 private:
  struct Impl;
  Impl* p_;
};

}  // namespace frchost
