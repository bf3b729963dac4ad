/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the reference hot path (fluhus/frackyfrac, frcfrc/unifrac.go)
 * used only as the parity checker (tests/, __graft_entry__.smoke()) and as the
 * timed CPU baseline (bench.py cpu_baseline / --impl reference).  Nothing under
 * frackyfrac_b200/ may include, link or call this.
 *
 * Parity status: PINNED against the reference's own known-answer tests
 * (frcfrc/unifrac_test.go:12-74) and its six CLI golden comparisons
 * (testdata/run.sh:3-16 vs testdata/{uwtd1,uwtd2,wtd}.want) — see
 * tests/test_oracle_kat.py.  The Go reference itself cannot be built in this
 * image (no Go toolchain, dependencies not vendored), so there is no oracle/_ref.
 */
#ifndef UNIFRAC_ORACLE_H
#define UNIFRAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tree orc_tree;     /* pointer tree, as newick.Node            */
typedef struct orc_table orc_table;   /* []map[string]float64, one map per sample */

/* Last error message of the calling thread ("" if none). */
const char *orc_last_error(void);

/* ---- parsing (restates the third-party newick reader and parser/parser.go) ---- */
orc_tree *orc_tree_parse(const char *text, size_t len);
void orc_tree_free(orc_tree *t);
int64_t orc_tree_num_nodes(const orc_tree *t);

/* Dense (parser.go:21-81) or sparse (parser.go:85-140) abundance table. */
orc_table *orc_table_parse(const char *text, size_t len, int sparse);
void orc_table_free(orc_table *t);
int64_t orc_table_num_samples(const orc_table *t);
int64_t orc_table_sample_size(const orc_table *t, int64_t sample);
/* Entry `k` of sample's map (insertion order of first occurrence). */
const char *orc_table_entry_name(const orc_table *t, int64_t sample, int64_t k);
double orc_table_entry_value(const orc_table *t, int64_t sample, int64_t k);

/* Build inputs without text (used for the large synthetic baselines):
 * a tree from pre-order parent/length arrays (leaves are named "L<id>"), and a
 * table from CSR rows whose columns are leaf node ids of that tree. */
orc_tree *orc_tree_from_flat(int32_t n_nodes, const int32_t *parent, const double *length);
orc_table *orc_table_from_csr(int64_t n_samples, const int64_t *row_ptr,
                              const int32_t *leaf_id, const double *val);

/* unifrac.go:80-93.  Returns 0 if every species of every sample names a tree
 * node, else -1 and sets the error text to the reference's message. */
int orc_validate_species(const orc_table *tab, const orc_tree *tree);

/* unifrac.go:97-124 + :209-228.  Writes n(n-1)/2 distances in IterPairs order
 * (common/common.go:21-31).  `nthreads` = -p.  `normalize`: 1 = default
 * (sort + divide by all-node total); 0 = the reference's -l exactly as coded
 * (post-order lists are NOT re-sorted, see the .c file); 2 = -l as documented
 * (sorted lists, raw values) — what the CUDA engine implements for -l.
 * `out` must hold n(n-1)/2 doubles.  Returns 0 / -1. */
int orc_unifrac(const orc_table *tab, const orc_tree *tree, int weighted, int normalize,
                int nthreads, double *out);

/* Same, but only pairs (i, j<i) with i in [row_begin, row_end) — a bounded
 * sample of the workload for timing.  `out` holds the corresponding slice. */
int orc_unifrac_rows(const orc_table *tab, const orc_tree *tree, int weighted, int normalize,
                     int nthreads, int64_t row_begin, int64_t row_end, double *out,
                     double *embed_seconds, double *pair_seconds);

/* Per-sample sparse node list after embedding (+ optional normalisation):
 * exposes unifrac.go:32-67 so the device embedding can be compared directly.
 * Returns nnz and fills ids/vals (capacity cap) for `sample`. */
int64_t orc_flat_nodes(const orc_table *tab, const orc_tree *tree, int normalize, int64_t sample,
                       int64_t cap, int64_t *ids, double *vals);

/* Pre-order flattening of the tree (unifrac.go:127-133, :117-120). */
int orc_tree_flatten(const orc_tree *t, int32_t *parent, double *length);

/* fmt.Fprintln(w, f) for a float64 (frcfrc.go:59): Go %v.  Writes a NUL
 * terminated string without the newline; returns its length. */
int orc_format_go(double v, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
