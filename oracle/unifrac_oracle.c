/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See unifrac_oracle.h.
 *
 * CPU restatement of fluhus/frackyfrac's UniFrac path, written from the
 * behaviour of:
 *   frcfrc/unifrac.go:32-53    abundanceToFlatNodes  -> embed_rec()
 *   frcfrc/unifrac.go:56-67    normalizeFlatNodes    -> normalize_flat()
 *   frcfrc/unifrac.go:70-93    validateSpecies       -> orc_validate_species()
 *   frcfrc/unifrac.go:97-124   unifrac               -> orc_unifrac_rows()
 *   frcfrc/unifrac.go:127-133  enumerateNodes        -> tree_index()
 *   frcfrc/unifrac.go:144-171  unifracDistUnweighted -> dist_unweighted()
 *   frcfrc/unifrac.go:174-205  unifracDistWeighted   -> dist_weighted()
 *   frcfrc/unifrac.go:209-228  unifracDists          -> pair loop in worker()
 *   common/common.go:21-31     IterPairs             -> row-major (i, j<i)
 *   parser/parser.go:21-140    dense / sparse tables -> orc_table_parse()
 *   frcfrc/frcfrc.go:58-62     fmt.Fprintln(w, f)    -> orc_format_go()
 * The Newick reader restates the grammar the third-party
 * github.com/fluhus/biostuff v1.0.0 formats/newick package is used for here
 * (frcfrc/frcfrc.go:109-114): only `(child,...)name:length;` with plain decimal
 * lengths is pinned by the reference's tests; quoted labels and [comments] are
 * accepted but "parity unpinned".
 *
 * Data structures deliberately mirror the reference (16-byte {int64,double}
 * flat nodes, id-sorted, merge-join per pair, double len[] gather) so that the
 * timed CPU baseline measures the reference's algorithm, not a different one.
 */
#define _GNU_SOURCE
#include "unifrac_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

/* ------------------------------------------------------------------ errors */
static __thread char g_err[512];
const char *orc_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); } while (0)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* -------------------------------------------------------- string hash map */
typedef struct {
  char **keys;
  int32_t *vals;
  size_t cap, n;
} strmap;

static uint64_t fnv1a(const char *s, size_t n) {
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
  return h;
}
static void strmap_init(strmap *m, size_t cap) {
  m->cap = 16;
  while (m->cap < cap * 2) m->cap <<= 1;
  m->keys = calloc(m->cap, sizeof *m->keys);
  m->vals = malloc(m->cap * sizeof *m->vals);
  m->n = 0;
}
static void strmap_free(strmap *m, int free_keys) {
  if (free_keys) for (size_t i = 0; i < m->cap; i++) free(m->keys[i]);
  free(m->keys); free(m->vals);
}
static int32_t strmap_get(const strmap *m, const char *s, size_t n) {
  size_t i = fnv1a(s, n) & (m->cap - 1);
  while (m->keys[i]) {
    if (strlen(m->keys[i]) == n && memcmp(m->keys[i], s, n) == 0) return m->vals[i];
    i = (i + 1) & (m->cap - 1);
  }
  return -1;
}
static void strmap_grow(strmap *m);
/* Inserts (copying the key) if absent; returns the stored value. */
static int32_t strmap_put(strmap *m, const char *s, size_t n, int32_t v) {
  if ((m->n + 1) * 2 > m->cap) strmap_grow(m);
  size_t i = fnv1a(s, n) & (m->cap - 1);
  while (m->keys[i]) {
    if (strlen(m->keys[i]) == n && memcmp(m->keys[i], s, n) == 0) return m->vals[i];
    i = (i + 1) & (m->cap - 1);
  }
  m->keys[i] = strndup(s, n);
  m->vals[i] = v;
  m->n++;
  return v;
}
static void strmap_grow(strmap *m) {
  strmap b;
  b.cap = m->cap * 2; b.n = 0;
  b.keys = calloc(b.cap, sizeof *b.keys);
  b.vals = malloc(b.cap * sizeof *b.vals);
  for (size_t i = 0; i < m->cap; i++) if (m->keys[i]) {
    size_t j = fnv1a(m->keys[i], strlen(m->keys[i])) & (b.cap - 1);
    while (b.keys[j]) j = (j + 1) & (b.cap - 1);
    b.keys[j] = m->keys[i]; b.vals[j] = m->vals[i]; b.n++;
  }
  free(m->keys); free(m->vals);
  *m = b;
}

/* ------------------------------------------------------------------- tree */
typedef struct node {
  char *name;
  double dist;
  struct node **children;
  int32_t nchild, cap;
  int32_t id;      /* pre-order index (unifrac.go:127-133) */
  int32_t species; /* column of this node's name in the table, resolved per run */
  struct node *parent;
} node;

struct orc_tree {
  node *root;
  int64_t n;
  node **pre; /* nodes in pre-order */
};

static node *node_new(node *parent) {
  node *x = calloc(1, sizeof *x);
  x->parent = parent;
  x->name = NULL;
  if (parent) {
    if (parent->nchild == parent->cap) {
      parent->cap = parent->cap ? parent->cap * 2 : 2;
      parent->children = realloc(parent->children, (size_t)parent->cap * sizeof(node *));
    }
    parent->children[parent->nchild++] = x;
  }
  return x;
}

/* Pre-order numbering, parent before children, children in file order. */
static void tree_index(orc_tree *t) {
  int64_t n = 0, cap = 1024, sp = 0;
  node **stack = malloc((size_t)cap * sizeof *stack);
  node **pre = malloc((size_t)cap * sizeof *pre);
  int64_t pcap = cap;
  stack[sp++] = t->root;
  while (sp) {
    node *x = stack[--sp];
    if (n == pcap) { pcap *= 2; pre = realloc(pre, (size_t)pcap * sizeof *pre); }
    x->id = (int32_t)n;
    pre[n++] = x;
    if (sp + x->nchild > cap) {
      while (sp + x->nchild > cap) cap *= 2;
      stack = realloc(stack, (size_t)cap * sizeof *stack);
    }
    for (int32_t c = x->nchild - 1; c >= 0; c--) stack[sp++] = x->children[c];
  }
  free(stack);
  t->pre = pre;
  t->n = n;
}

void orc_tree_free(orc_tree *t) {
  if (!t) return;
  if (t->pre) {
    for (int64_t i = 0; i < t->n; i++) {
      free(t->pre[i]->name); free(t->pre[i]->children); free(t->pre[i]);
    }
    free(t->pre);
  }
  free(t);
}
int64_t orc_tree_num_nodes(const orc_tree *t) { return t->n; }

static int is_newick_delim(int c) {
  return c == '(' || c == ')' || c == ',' || c == ':' || c == ';' || c == '[' || c == ']';
}

orc_tree *orc_tree_parse(const char *s, size_t len) {
  g_err[0] = 0;
  orc_tree *t = calloc(1, sizeof *t);
  node *cur = node_new(NULL);
  t->root = cur;
  size_t i = 0;
  int done = 0;
  /* State: we are positioned "at" node cur, before its children/label. */
  while (i < len && !done) {
    unsigned char c = (unsigned char)s[i];
    if (isspace(c)) { i++; continue; }
    if (c == '[') { /* comment */
      while (i < len && s[i] != ']') i++;
      if (i < len) i++;
      continue;
    }
    if (c == '(') { cur = node_new(cur); i++; continue; }
    if (c == ',') {
      if (!cur->parent) { FAIL("newick: ',' outside parentheses at byte %zu", i); goto bad; }
      cur = node_new(cur->parent); i++; continue;
    }
    if (c == ')') {
      if (!cur->parent) { FAIL("newick: unbalanced ')' at byte %zu", i); goto bad; }
      cur = cur->parent; i++; continue;
    }
    if (c == ';') { done = 1; i++; break; }
    if (c == ':') {
      i++;
      while (i < len && isspace((unsigned char)s[i])) i++;
      char buf[128];
      size_t k = 0;
      while (i < len && !is_newick_delim((unsigned char)s[i]) && !isspace((unsigned char)s[i]) &&
             k + 1 < sizeof buf) buf[k++] = s[i++];
      buf[k] = 0;
      char *end;
      double d = strtod(buf, &end);
      if (k == 0 || *end) { FAIL("newick: bad branch length %s", buf); goto bad; }
      cur->dist = d;
      continue;
    }
    /* label */
    if (c == '\'') {
      size_t cap = 16, k = 0;
      char *b = malloc(cap);
      i++;
      for (;;) {
        if (i >= len) { free(b); FAIL("newick: unterminated quoted label"); goto bad; }
        if (s[i] == '\'') {
          if (i + 1 < len && s[i + 1] == '\'') { i++; } else { i++; break; }
        }
        if (k + 2 > cap) { cap *= 2; b = realloc(b, cap); }
        b[k++] = s[i++];
      }
      b[k] = 0;
      free(cur->name);
      cur->name = b;
      continue;
    }
    {
      size_t st = i;
      while (i < len && !is_newick_delim((unsigned char)s[i]) && !isspace((unsigned char)s[i])) i++;
      free(cur->name);
      cur->name = strndup(s + st, i - st);
    }
  }
  if (!done) { FAIL("newick: missing ';'"); goto bad; }
  if (cur != t->root) { FAIL("newick: unbalanced '('"); goto bad; }
  tree_index(t);
  for (int64_t k = 0; k < t->n; k++) if (!t->pre[k]->name) t->pre[k]->name = strdup("");
  return t;
bad:
  tree_index(t);
  orc_tree_free(t);
  return NULL;
}

orc_tree *orc_tree_from_flat(int32_t n, const int32_t *parent, const double *length) {
  g_err[0] = 0;
  if (n < 1 || parent[0] != -1) { FAIL("flat tree: node 0 must be the root"); return NULL; }
  orc_tree *t = calloc(1, sizeof *t);
  node **all = malloc((size_t)n * sizeof *all);
  for (int32_t v = 0; v < n; v++) {
    if (v > 0 && (parent[v] < 0 || parent[v] >= v)) {
      FAIL("flat tree: parent[%d]=%d is not a smaller id", v, parent[v]);
      for (int32_t u = 0; u < v; u++) { free(all[u]->name); free(all[u]->children); free(all[u]); }
      free(all); free(t);
      return NULL;
    }
    all[v] = node_new(v ? all[parent[v]] : NULL);
    all[v]->dist = length[v];
  }
  for (int32_t v = 0; v < n; v++) {
    char b[32];
    if (all[v]->nchild == 0) snprintf(b, sizeof b, "L%d", v); else b[0] = 0;
    all[v]->name = strdup(b);
  }
  t->root = all[0];
  free(all);
  tree_index(t);
  /* ids were assigned in ascending order with parents first, so pre-order
   * re-indexing must reproduce them if and only if the input was pre-order. */
  return t;
}

int orc_tree_flatten(const orc_tree *t, int32_t *parent, double *length) {
  for (int64_t i = 0; i < t->n; i++) {
    parent[i] = t->pre[i]->parent ? t->pre[i]->parent->id : -1;
    length[i] = t->pre[i]->dist;
  }
  return 0;
}

/* ------------------------------------------------------------------ table */
typedef struct { int32_t sp; double val; } entry;
typedef struct { entry *e; int64_t n, cap; } sample;

struct orc_table {
  strmap names;      /* species name -> column */
  char **name_of;    /* column -> name */
  int64_t nsp, spcap;
  sample *s;
  int64_t n, cap;
};

static int32_t table_intern(orc_table *t, const char *s, size_t n) {
  int32_t v = strmap_get(&t->names, s, n);
  if (v >= 0) return v;
  v = (int32_t)t->nsp;
  strmap_put(&t->names, s, n, v);
  if (t->nsp == t->spcap) {
    t->spcap = t->spcap ? t->spcap * 2 : 64;
    t->name_of = realloc(t->name_of, (size_t)t->spcap * sizeof *t->name_of);
  }
  t->name_of[t->nsp++] = strndup(s, n);
  return v;
}
static sample *table_add_sample(orc_table *t) {
  if (t->n == t->cap) {
    t->cap = t->cap ? t->cap * 2 : 64;
    t->s = realloc(t->s, (size_t)t->cap * sizeof *t->s);
  }
  sample *x = &t->s[t->n++];
  memset(x, 0, sizeof *x);
  return x;
}
/* m[name] = v : last assignment wins (parser.go:78,124). `seen` maps
 * species -> position in this sample (stamp-checked). */
static void sample_set(sample *x, int32_t sp, double v, int64_t *pos, int64_t *stamp, int64_t cur) {
  if (stamp[sp] == cur) { x->e[pos[sp]].val = v; return; }
  if (x->n == x->cap) {
    x->cap = x->cap ? x->cap * 2 : 8;
    x->e = realloc(x->e, (size_t)x->cap * sizeof *x->e);
  }
  stamp[sp] = cur; pos[sp] = x->n;
  x->e[x->n].sp = sp; x->e[x->n].val = v; x->n++;
}

void orc_table_free(orc_table *t) {
  if (!t) return;
  for (int64_t i = 0; i < t->n; i++) free(t->s[i].e);
  free(t->s);
  for (int64_t i = 0; i < t->nsp; i++) free(t->name_of[i]);
  free(t->name_of);
  strmap_free(&t->names, 1);
  free(t);
}
int64_t orc_table_num_samples(const orc_table *t) { return t->n; }
int64_t orc_table_sample_size(const orc_table *t, int64_t s) { return t->s[s].n; }
const char *orc_table_entry_name(const orc_table *t, int64_t s, int64_t k) {
  return t->name_of[t->s[s].e[k].sp];
}
double orc_table_entry_value(const orc_table *t, int64_t s, int64_t k) { return t->s[s].e[k].val; }

/* Go regexp \s is [\t\n\f\r ] (parser.go:17 splits on \S+). */
static int go_space(int c) { return c == '\t' || c == '\n' || c == '\f' || c == '\r' || c == ' '; }

/* The spellings strconv.ParseFloat accepts (strconv/atof.go: special(), readFloat(), underscoreOK()):
 * inf / infinity / nan in any case (inf signed), decimal mantissa with optional e-exponent, hex mantissa with a
 * REQUIRED p-exponent, underscores only between digits or right after the 0x prefix.  Copies the token without
 * underscores into buf (n + 1 bytes) and returns 0, or -1 for anything Go calls a syntax error. */
static int go_float_token(const char *s, size_t n, char *buf) {
  size_t i = 0, k = 0;
  if (i < n && (s[i] == '+' || s[i] == '-')) buf[k++] = s[i++];
  size_t rest = n - i;
  if ((rest == 3 && !strncasecmp(s + i, "inf", 3)) || (rest == 8 && !strncasecmp(s + i, "infinity", 8)) ||
      (i == 0 && rest == 3 && !strncasecmp(s, "nan", 3))) {
    memcpy(buf, s, n); buf[n] = 0;
    return 0;
  }
  int hex = 0;
  if (i + 1 < n && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) { hex = 1; buf[k++] = s[i++]; buf[k++] = s[i++]; }
  int prev_digit = hex, prev_us = 0, digits = 0, dot = 0, in_exp = 0;
  for (; i < n; i++) {
    char c = s[i];
    int dig = isdigit((unsigned char)c) || (hex && !in_exp && isxdigit((unsigned char)c));
    if (c == '_') {
      if (!prev_digit) return -1;
      prev_us = 1; prev_digit = 0;
      continue;
    }
    if (dig) { digits = 1; prev_digit = 1; prev_us = 0; buf[k++] = c; continue; }
    if (prev_us) return -1;
    prev_digit = 0;
    if (c == '.' && !dot && !in_exp) { dot = 1; buf[k++] = c; continue; }
    if (!in_exp && digits && (hex ? (c == 'p' || c == 'P') : (c == 'e' || c == 'E'))) {
      in_exp = 1; buf[k++] = c;
      if (i + 1 < n && (s[i + 1] == '+' || s[i + 1] == '-')) buf[k++] = s[++i];
      if (i + 1 >= n || !isdigit((unsigned char)s[i + 1])) return -1;
      continue;
    }
    return -1;
  }
  if (!digits || prev_us || (hex && !in_exp)) return -1;
  buf[k] = 0;
  return 0;
}

/* strconv.ParseFloat(tok, 64): nil error required. */
static int parse_float(const char *s, size_t n, double *out) {
  char small[64], *buf = small;
  if (n == 0) return -1;
  if (n + 1 > sizeof small) buf = malloc(n + 1);
  int ok = go_float_token(s, n, buf);
  if (ok == 0) {
    char *end;
    errno = 0;
    double v = strtod(buf, &end);
    if (*end || end == buf) ok = -1;
    /* Go reports overflow as an error; underflow to 0 is not one. */
    else if (errno == ERANGE && isinf(v)) ok = -1;
    else *out = v;
  }
  if (buf != small) free(buf);
  return ok;
}

#define MAX_ROW (1u << 25) /* sc.Buffer(nil, 1<<25), parser.go:145 */

orc_table *orc_table_parse(const char *text, size_t len, int sparse) {
  g_err[0] = 0;
  orc_table *t = calloc(1, sizeof *t);
  strmap_init(&t->names, 1024);
  int64_t *pos = NULL, *stamp = NULL, stcap = 0;
  int32_t *hdr = NULL; int64_t nhdr = -1;
  size_t p = 0;
  int64_t rowno = 0;
  while (p < len) {
    /* bufio.ScanLines: up to '\n', dropping one trailing '\r'; a final line
     * without newline is still a line; nothing after the last '\n' is not. */
    size_t e = p;
    while (e < len && text[e] != '\n') e++;
    size_t le = e;
    if (le > p && text[le - 1] == '\r') le--;
    if (le - p >= MAX_ROW) { FAIL("bufio.Scanner: token too long"); goto bad; }
    rowno++;
    const char *row = text + p; size_t rl = le - p;
    p = e + 1;

    if (!sparse && nhdr < 0) {
      int64_t cap = 16; nhdr = 0;
      hdr = malloc((size_t)cap * sizeof *hdr);
      size_t i = 0;
      while (i < rl) {
        while (i < rl && go_space((unsigned char)row[i])) i++;
        size_t st = i;
        while (i < rl && !go_space((unsigned char)row[i])) i++;
        if (i > st) {
          if (nhdr == cap) { cap *= 2; hdr = realloc(hdr, (size_t)cap * sizeof *hdr); }
          hdr[nhdr++] = table_intern(t, row + st, i - st);
        }
      }
      if (nhdr == 0) { FAIL("row #1 has 0 values"); goto bad; }
      continue;
    }
    sample *x = table_add_sample(t);
    int64_t cur = t->n; /* stamp value for this sample */
    int64_t k = 0;
    size_t i = 0;
    while (i < rl) {
      while (i < rl && go_space((unsigned char)row[i])) i++;
      size_t st = i;
      while (i < rl && !go_space((unsigned char)row[i])) i++;
      if (i == st) break;
      k++;
      const char *tok = row + st; size_t tl = i - st;
      double v;
      int32_t sp;
      if (sparse) {
        /* splitSparse: split on the LAST ':' (parser.go:129-140) */
        size_t last = (size_t)-1;
        for (size_t q = 0; q < tl; q++) if (tok[q] == ':') last = q;
        if (last == (size_t)-1) { FAIL("row #%lld: value #%lld: no colon in \"%.*s\"", (long long)rowno, (long long)k, (int)tl, tok); goto bad; }
        if (last == 0) { FAIL("row #%lld: value #%lld: empty species name", (long long)rowno, (long long)k); goto bad; }
        if (parse_float(tok + last + 1, tl - last - 1, &v)) { FAIL("row #%lld: value #%lld: cannot parse \"%.*s\"", (long long)rowno, (long long)k, (int)(tl - last - 1), tok + last + 1); goto bad; }
        if (isnan(v) || isinf(v) || v < 0) { FAIL("row #%lld: value #%lld: bad value: %f", (long long)rowno, (long long)k, v); goto bad; }
        if (v == 0) { FAIL("row #%lld: value #%lld: zeros are not allowed in sparse format", (long long)rowno, (long long)k); goto bad; }
        sp = table_intern(t, tok, last);
      } else {
        if (k > nhdr) continue; /* counted below */
        if (parse_float(tok, tl, &v)) { FAIL("row #%lld: value #%lld: cannot parse \"%.*s\"", (long long)rowno, (long long)k, (int)tl, tok); goto bad; }
        if (isnan(v) || isinf(v) || v < 0) { FAIL("row #%lld: value #%lld: bad value: %f", (long long)rowno, (long long)k, v); goto bad; }
        if (v == 0) continue;
        sp = hdr[k - 1];
      }
      if (t->nsp > stcap) {
        int64_t nc = stcap ? stcap : 64;
        while (nc < t->nsp) nc *= 2;
        pos = realloc(pos, (size_t)nc * sizeof *pos);
        stamp = realloc(stamp, (size_t)nc * sizeof *stamp);
        for (int64_t q = stcap; q < nc; q++) stamp[q] = 0;
        stcap = nc;
      }
      sample_set(x, sp, v, pos, stamp, cur);
    }
    if (!sparse && k != nhdr) {
      /* the length check precedes value parsing in parseRow (parser.go:61-64) */
      FAIL("row #%lld: has %lld values, expected %lld", (long long)rowno, (long long)k, (long long)nhdr);
      goto bad;
    }
  }
  free(pos); free(stamp); free(hdr);
  return t;
bad:
  free(pos); free(stamp); free(hdr);
  orc_table_free(t);
  return NULL;
}

orc_table *orc_table_from_csr(int64_t n, const int64_t *row_ptr, const int32_t *leaf, const double *val) {
  g_err[0] = 0;
  orc_table *t = calloc(1, sizeof *t);
  strmap_init(&t->names, 1024);
  int64_t *pos = NULL, *stamp = NULL, stcap = 0;
  for (int64_t s = 0; s < n; s++) {
    sample *x = table_add_sample(t);
    for (int64_t k = row_ptr[s]; k < row_ptr[s + 1]; k++) {
      char b[32];
      int bl = snprintf(b, sizeof b, "L%d", leaf[k]);
      int32_t sp = table_intern(t, b, (size_t)bl);
      if (t->nsp > stcap) {
        int64_t nc = stcap ? stcap : 64;
        while (nc < t->nsp) nc *= 2;
        pos = realloc(pos, (size_t)nc * sizeof *pos);
        stamp = realloc(stamp, (size_t)nc * sizeof *stamp);
        for (int64_t q = stcap; q < nc; q++) stamp[q] = 0;
        stcap = nc;
      }
      sample_set(x, sp, val[k], pos, stamp, s + 1);
    }
  }
  free(pos); free(stamp);
  return t;
}

/* -------------------------------------------------------------- validation */
int orc_validate_species(const orc_table *tab, const orc_tree *tree) {
  g_err[0] = 0;
  strmap names; /* treeNames(): every node's name, internal ones too */
  strmap_init(&names, (size_t)tree->n);
  for (int64_t i = 0; i < tree->n; i++) strmap_put(&names, tree->pre[i]->name, strlen(tree->pre[i]->name), 1);
  int rc = 0;
  for (int64_t s = 0; s < tab->n && !rc; s++)
    for (int64_t k = 0; k < tab->s[s].n; k++) {
      const char *nm = tab->name_of[tab->s[s].e[k].sp];
      if (strmap_get(&names, nm, strlen(nm)) < 0) {
        char vb[64];
        orc_format_go(tab->s[s].e[k].val, vb, sizeof vb);
        FAIL("sample #%lld has value %s for species \"%s\" which is not in the tree",
             (long long)(s + 1), vb, nm);
        rc = -1;
        break;
      }
    }
  strmap_free(&names, 1);
  return rc;
}

/* --------------------------------------------------------------- embedding */
typedef struct { int64_t id; double abnd; } flat_node; /* unifrac.go:137-140 */
typedef struct { flat_node *v; int64_t n, cap; } flat_list;

static void flat_push(flat_list *l, int64_t id, double a) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 64;
    l->v = realloc(l->v, (size_t)l->cap * sizeof *l->v);
  }
  l->v[l->n].id = id; l->v[l->n].abnd = a; l->n++;
}

/* abundanceToFlatNodes (unifrac.go:32-53), flatNodeOptimization = true. */
static double embed_rec(const double *abnd, const node *x, flat_list *out) {
  double sum = 0.0;
  for (int32_t c = 0; c < x->nchild; c++) sum += embed_rec(abnd, x->children[c], out);
  if (x->nchild == 0) {
    double a = x->species >= 0 ? abnd[x->species] : 0.0;
    if (a > 0) sum += a;
  }
  if (sum > 0) flat_push(out, x->id, sum);
  return sum;
}

static int cmp_flat(const void *a, const void *b) {
  int64_t x = ((const flat_node *)a)->id, y = ((const flat_node *)b)->id;
  return (x > y) - (x < y);
}
/* normalizeFlatNodes (unifrac.go:56-67). */
static void normalize_flat(flat_list *l) {
  qsort(l->v, (size_t)l->n, sizeof *l->v, cmp_flat);
  double sum = 0.0;
  for (int64_t i = 0; i < l->n; i++) sum += l->v[i].abnd;
  for (int64_t i = 0; i < l->n; i++) l->v[i].abnd /= sum;
}

/* ---------------------------------------------------------------- distances */
static double dist_unweighted(const flat_list *A, const flat_list *B, const double *td) {
  const flat_node *a = A->v, *b = B->v;
  double result = 0.0, common = 0.0;
  int64_t i = 0, j = 0;
  while (i < A->n && j < B->n) {
    if (a[i].id < b[j].id) { result += td[a[i].id]; i++; continue; }
    if (a[i].id > b[j].id) { result += td[b[j].id]; j++; continue; }
    common += td[a[i].id];
    i++; j++;
  }
  for (; i < A->n; i++) result += td[a[i].id];
  for (; j < B->n; j++) result += td[b[j].id];
  result /= (result + common);
  return result;
}

static double dist_weighted(const flat_list *A, const flat_list *B, const double *td) {
  const flat_node *a = A->v, *b = B->v;
  double numer = 0.0, denom = 0.0;
  int64_t i = 0, j = 0;
  while (i < A->n && j < B->n) {
    if (a[i].id < b[j].id) {
      numer += td[a[i].id] * a[i].abnd;
      denom += td[a[i].id] * a[i].abnd;
      i++; continue;
    }
    if (a[i].id > b[j].id) {
      numer += td[b[j].id] * b[j].abnd;
      denom += td[b[j].id] * b[j].abnd;
      j++; continue;
    }
    numer += td[a[i].id] * fabs(a[i].abnd - b[j].abnd);
    denom += td[a[i].id] * (a[i].abnd + b[j].abnd);
    i++; j++;
  }
  for (; i < A->n; i++) { numer += td[a[i].id] * a[i].abnd; denom += td[a[i].id] * a[i].abnd; }
  for (; j < B->n; j++) { numer += td[b[j].id] * b[j].abnd; denom += td[b[j].id] * b[j].abnd; }
  return numer / denom;
}

/* ------------------------------------------------------------------ driver */
typedef struct {
  const orc_table *tab;
  const orc_tree *tree;
  int weighted, normalize;
  flat_list *sets;
  const double *td;
  int64_t row_begin, row_end;
  double *out;
  atomic_llong next;
  int phase; /* 0 embed, 1 pairs */
} job;

static void *worker(void *arg) {
  job *jb = arg;
  if (jb->phase == 0) {
    double *abnd = calloc((size_t)(jb->tab->nsp ? jb->tab->nsp : 1), sizeof *abnd);
    for (;;) {
      int64_t s = atomic_fetch_add(&jb->next, 1);
      if (s >= jb->tab->n) break;
      const sample *x = &jb->tab->s[s];
      for (int64_t k = 0; k < x->n; k++) abnd[x->e[k].sp] = x->e[k].val;
      embed_rec(abnd, jb->tree->root, &jb->sets[s]);
      if (jb->normalize == 1) normalize_flat(&jb->sets[s]);
      else if (jb->normalize == 2) qsort(jb->sets[s].v, (size_t)jb->sets[s].n, sizeof(flat_node), cmp_flat);
      /* normalize == 0: the reference's -l leaves the list in post-order */
      for (int64_t k = 0; k < x->n; k++) abnd[x->e[k].sp] = 0.0;
    }
    free(abnd);
  } else {
    int64_t base = jb->row_begin * (jb->row_begin - 1) / 2;
    for (;;) {
      /* rows handed out from the longest downwards for balance; results land
       * at their IterPairs index, so output order is independent of this. */
      int64_t r = atomic_fetch_add(&jb->next, 1);
      int64_t i = jb->row_end - 1 - r;
      if (i < jb->row_begin) break;
      double *o = jb->out + (i * (i - 1) / 2 - base);
      for (int64_t j = 0; j < i; j++)
        o[j] = jb->weighted ? dist_weighted(&jb->sets[i], &jb->sets[j], jb->td)
                            : dist_unweighted(&jb->sets[i], &jb->sets[j], jb->td);
    }
  }
  return NULL;
}

static void run_phase(job *jb, int nthreads) {
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, (size_t)1 << 30); /* embed_rec recurses to tree depth */
  pthread_t *th = malloc((size_t)nthreads * sizeof *th);
  atomic_store(&jb->next, 0);
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], &at, worker, jb);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
  pthread_attr_destroy(&at);
}

/* Resolve each leaf's name against the table's columns: abnd[tree.Name]. */
static void resolve_species(const orc_table *tab, const orc_tree *tree) {
  for (int64_t i = 0; i < tree->n; i++) {
    node *x = tree->pre[i];
    x->species = strmap_get(&tab->names, x->name, strlen(x->name));
  }
}

/* `normalize` selects how unifrac.go:108-110 is restated:
 *   1  default: normalizeFlatNodes (sort by id, divide by the all-node total).
 *   0  the reference's -l, FAITHFULLY: the sort lives inside
 *      normalizeFlatNodes (unifrac.go:57), so with -l the lists stay in the
 *      post-order emission order of abundanceToFlatNodes and the two-pointer
 *      merge-join (unifrac.go:178-203) runs over lists that are not id-sorted:
 *      it mis-pairs nodes and the result is not a UniFrac distance.  The
 *      reference's tests never run -l, so this is unpinned reference behaviour.
 *   2  -l as documented ("leave abundance values unnormalized"): lists sorted
 *      by id, values left raw.  This is what the CUDA engine computes for
 *      normalize=0; the divergence from mode 0 is stated in DESIGN.md and
 *      demonstrated by tests/test_oracle_kat.py::test_l_flag_reference_quirk. */

int orc_unifrac_rows(const orc_table *tab, const orc_tree *tree, int weighted, int normalize,
                     int nthreads, int64_t row_begin, int64_t row_end, double *out,
                     double *embed_seconds, double *pair_seconds) {
  g_err[0] = 0;
  if (nthreads < 1) { FAIL("bad number of threads: %d", nthreads); return -1; }
  if (row_begin < 0 || row_end > tab->n || row_begin > row_end) { FAIL("bad row range"); return -1; }
  resolve_species(tab, tree);
  job jb;
  memset(&jb, 0, sizeof jb);
  jb.tab = tab; jb.tree = tree; jb.weighted = weighted; jb.normalize = normalize;
  jb.sets = calloc((size_t)(tab->n ? tab->n : 1), sizeof *jb.sets);
  double *td = malloc((size_t)tree->n * sizeof *td);
  for (int64_t i = 0; i < tree->n; i++) td[i] = tree->pre[i]->dist;
  jb.td = td;
  double t0 = now_s();
  jb.phase = 0;
  run_phase(&jb, nthreads);
  double t1 = now_s();
  jb.phase = 1; jb.row_begin = row_begin; jb.row_end = row_end; jb.out = out;
  run_phase(&jb, nthreads);
  double t2 = now_s();
  if (embed_seconds) *embed_seconds = t1 - t0;
  if (pair_seconds) *pair_seconds = t2 - t1;
  for (int64_t s = 0; s < tab->n; s++) free(jb.sets[s].v);
  free(jb.sets); free(td);
  return 0;
}

int orc_unifrac(const orc_table *tab, const orc_tree *tree, int weighted, int normalize,
                int nthreads, double *out) {
  return orc_unifrac_rows(tab, tree, weighted, normalize, nthreads, 0, tab->n, out, NULL, NULL);
}

int64_t orc_flat_nodes(const orc_table *tab, const orc_tree *tree, int normalize, int64_t s,
                       int64_t cap, int64_t *ids, double *vals) {
  resolve_species(tab, tree);
  double *abnd = calloc((size_t)(tab->nsp ? tab->nsp : 1), sizeof *abnd);
  const sample *x = &tab->s[s];
  for (int64_t k = 0; k < x->n; k++) abnd[x->e[k].sp] = x->e[k].val;
  flat_list l = {0};
  embed_rec(abnd, tree->root, &l);
  if (normalize == 1) normalize_flat(&l);
  else if (normalize == 2) qsort(l.v, (size_t)l.n, sizeof(flat_node), cmp_flat);
  int64_t n = l.n;
  for (int64_t k = 0; k < n && k < cap; k++) { ids[k] = l.v[k].id; vals[k] = l.v[k].abnd; }
  free(l.v); free(abnd);
  return n;
}

/* --------------------------------------------------------------- formatting */
/* Go: fmt.Fprintln(w, f) -> %v -> strconv.FormatFloat(f, 'g', -1, 64):
 * shortest digits that round-trip; %e form when exp < -4 || exp >= 6 (the
 * shortest-%g rule of strconv), exponent with at least two digits. */
int orc_format_go(double v, char *buf, size_t cap) {
  if (isnan(v)) return snprintf(buf, cap, "NaN");
  if (isinf(v)) return snprintf(buf, cap, v > 0 ? "+Inf" : "-Inf");
  if (v == 0) return snprintf(buf, cap, signbit(v) ? "-0" : "0");
  char e[40];
  int p;
  for (p = 0; p < 17; p++) {
    snprintf(e, sizeof e, "%.*e", p, v);
    if (strtod(e, NULL) == v) break;
  }
  /* e = [-]d[.ddd]e[+-]XX */
  char digits[24]; int nd = 0; int neg = 0;
  const char *q = e;
  if (*q == '-') { neg = 1; q++; }
  for (; *q && *q != 'e'; q++) if (*q != '.') digits[nd++] = *q;
  int x = atoi(q + 1);
  while (nd > 1 && digits[nd - 1] == '0') nd--; /* %.*e never pads here, but be safe */
  char out[64]; int o = 0;
  if (neg) out[o++] = '-';
  if (x < -4 || x >= 6) {
    out[o++] = digits[0];
    if (nd > 1) { out[o++] = '.'; for (int i = 1; i < nd; i++) out[o++] = digits[i]; }
    out[o++] = 'e';
    out[o++] = x < 0 ? '-' : '+';
    int ax = x < 0 ? -x : x;
    if (ax < 10) out[o++] = '0';
    o += snprintf(out + o, sizeof out - (size_t)o, "%d", ax);
  } else if (x < 0) {
    out[o++] = '0'; out[o++] = '.';
    for (int i = 0; i < -x - 1; i++) out[o++] = '0';
    for (int i = 0; i < nd; i++) out[o++] = digits[i];
  } else {
    for (int i = 0; i <= x; i++) out[o++] = i < nd ? digits[i] : '0';
    if (nd > x + 1) { out[o++] = '.'; for (int i = x + 1; i < nd; i++) out[o++] = digits[i]; }
  }
  out[o] = 0;
  return snprintf(buf, cap, "%s", out);
}

/* ---------------------------------------------------------------------- CLI */
#ifdef ORACLE_MAIN
/* frcfrc's flag surface (frcfrc/frcfrc.go:18-27,70-88) over the oracle, so the
 * reference's testdata/run.sh can be replayed against it. */
static char *slurp(const char *path, size_t *len) {
  FILE *f = path ? fopen(path, "rb") : stdin;
  if (!f) return NULL;
  size_t cap = 1 << 16, n = 0;
  char *b = malloc(cap);
  for (;;) {
    if (n == cap) { cap *= 2; b = realloc(b, cap); }
    size_t r = fread(b + n, 1, cap - n, f);
    if (!r) break;
    n += r;
  }
  if (path) fclose(f);
  *len = n;
  return b;
}
#define DIE(...) do { fprintf(stderr, "ERROR: "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(2); } while (0)
int main(int argc, char **argv) {
  const char *fin = NULL, *fout = NULL, *ftree = NULL;
  int wgt = 0, sparse = 0, nt = 1, nnorm = 0;
  if (argc == 1) { fprintf(stderr, "usage: frcfrc_oracle -t tree [-i in] [-o out] [-w] [-s] [-p n] [-l]\n"); return 0; }
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-i") && i + 1 < argc) fin = argv[++i];
    else if (!strcmp(argv[i], "-o") && i + 1 < argc) fout = argv[++i];
    else if (!strcmp(argv[i], "-t") && i + 1 < argc) ftree = argv[++i];
    else if (!strcmp(argv[i], "-p") && i + 1 < argc) nt = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-w")) wgt = 1;
    else if (!strcmp(argv[i], "-s")) sparse = 1;
    else if (!strcmp(argv[i], "-l")) nnorm = 1;
    else DIE("bad flag %s", argv[i]);
  }
  if (!ftree) DIE("please provide a tree file with -t");
  if (nt < 1) DIE("bad number of threads: %d", nt);
  if (nnorm && !wgt) DIE("-l can only be used with weighted unifrac");
  size_t tl, il;
  char *tt = slurp(ftree, &tl);
  if (!tt) DIE("open %s: %s", ftree, strerror(errno));
  orc_tree *tree = orc_tree_parse(tt, tl);
  if (!tree) DIE("%s", orc_last_error());
  char *it = slurp(fin, &il);
  if (!it) DIE("open %s: %s", fin, strerror(errno));
  orc_table *tab = orc_table_parse(it, il, sparse);
  if (!tab) DIE("%s", orc_last_error());
  if (orc_validate_species(tab, tree)) DIE("%s", orc_last_error());
  int64_t n = orc_table_num_samples(tab);
  int64_t np = n * (n - 1) / 2;
  double *out = malloc((size_t)(np > 0 ? np : 1) * sizeof *out);
  if (orc_unifrac(tab, tree, wgt, !nnorm, nt, out)) DIE("%s", orc_last_error());
  FILE *fo = fout ? fopen(fout, "wb") : stdout;
  if (!fo) DIE("create %s: %s", fout, strerror(errno));
  char b[64];
  for (int64_t k = 0; k < np; k++) { orc_format_go(out[k], b, sizeof b); fputs(b, fo); fputc('\n', fo); }
  if (fout) fclose(fo);
  return 0;
}
#endif
