// C++ stand-in for the Go `sprspr` binary (sprspr/sprspr.go): dense abundance table on stdin, sparse format on
// stdout, same banner, same error exit (common.ExitIfError: "ERROR: ..." and status 2).  The dense reader is the
// one the UniFrac CLI uses (parser.ParseAbundance semantics, parser/parser.go:21-82); the reference parses with
// 2 goroutines (sprspr.go:21), which only changes speed.
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>

#include "hostlib.hpp"

int main() {
  fputs("SparseySparse converts dense format abundance tables to sparse format.\n\n"
        "Usage:\nsprspr < INPUT_FILE > OUTPUT_FILE\n\nReading standard input...\n", stderr);
  try {
    const std::string in = frchost::read_file("");
    const frchost::Table t = frchost::parse_table(in.data(), in.size(), /*sparse=*/false, 2);
    std::string out;
    frchost::to_sparse_lines(t, out);
    if (fwrite(out.data(), 1, out.size(), stdout) != out.size() || fflush(stdout) != 0) {
      fputs("ERROR: write failed\n", stderr);
      return 2;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "ERROR: %s\n", e.what());
    return 2;
  }
  return 0;
}
