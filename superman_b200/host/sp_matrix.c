/* Matrix input and preprocessing in C: the int/ float/ double/ text reader, CRS+CCS conversion,
 * SortOrder / SkipOrder and the grid-graph generator.
 *
 * Mirrors, for the drop-in boundary:
 *   header sniff + ReadMatrix<T>          main.cu:494-498, util.h:343-358
 *   matrix2compressed<T>                  util.h:522-551
 *   matrix2compressed_sortOrder<T>        util.h:553-619
 *   matrix2compressed_skipOrder<T>        util.h:621-684
 *   gridGraph2compressed                  util.h:403-520
 * Element type is always double here (int and float files are exactly representable); the file's
 * declared type is kept in sp_matrix.type because RunAlgo's launch geometry and the result line
 * do not depend on it any more, but callers may want to know.
 * Deliberate fixes: the dense buffer is zero-initialised; entries outside
 * [0, nov) are ignored instead of written out of bounds; nnz is counted, not trusted from the header.
 */
#define _POSIX_C_SOURCE 200809L
#include "superman_b200.h"
#include "sp_sched.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void matrix_zero(sp_matrix *m) { memset(m, 0, sizeof(*m)); }

void sp_matrix_free(sp_matrix *m) {
  if (!m) return;
  free(m->mat); free(m->cptrs); free(m->rows); free(m->cvals);
  free(m->rptrs); free(m->cols); free(m->rvals);
  matrix_zero(m);
}

static int alloc_dense(sp_matrix *m, int nov) {
  if (nov < 1 || nov > SP_MAX_NOV) {
    sp_set_error("matrix order %d outside [1, %d]", nov, SP_MAX_NOV);
    return SP_ELIMIT;
  }
  m->nov = nov;
  m->mat = (double *)calloc((size_t)nov * nov, sizeof(double));
  if (!m->mat) { sp_set_error("out of memory"); return SP_ENOMEM; }
  return SP_OK;
}

int sp_matrix_from_dense(const double *mat, int nov, sp_matrix *out) {
  if (!mat || !out) { sp_set_error("null argument"); return SP_EINVAL; }
  matrix_zero(out);
  int rc = alloc_dense(out, nov);
  if (rc != SP_OK) return rc;
  memcpy(out->mat, mat, (size_t)nov * nov * sizeof(double));
  out->type = SP_TYPE_DOUBLE;
  return SP_OK;
}

/* One data line "i j val".  Mirrors `iss >> i >> j >> val` with val of the file's type: an int
 * file reads an integer prefix ("3.7" -> 3), a float file rounds to float first. */
static int parse_triple(const char *line, int type, int *i, int *j, double *val) {
  char *end;
  errno = 0;
  long a = strtol(line, &end, 10);
  if (end == line) return 0;
  const char *p = end;
  long b = strtol(p, &end, 10);
  if (end == p) return 0;
  p = end;
  if (type == SP_TYPE_INT) {
    long v = strtol(p, &end, 10);
    if (end == p) return 0;
    *val = (double)v;
  } else {
    double v = strtod(p, &end);
    if (end == p) return 0;
    *val = (type == SP_TYPE_FLOAT) ? (double)(float)v : v;
  }
  *i = (int)a; *j = (int)b;
  return 1;
}

/* MatrixMarket coordinate files (the revised front-end's reader,
 * revised_perman/read_matrix.hpp:11-157 + mmio banner): `%%MatrixMarket matrix coordinate
 * {real|integer|pattern} {general|symmetric}`, comment lines starting with %, `rows cols entries`,
 * then 1-based `i j [val]` lines; `pattern` files and the -b flag give 1 for every listed entry;
 * `symmetric` files mirror the off-diagonal entries.  The matrix must be square. */
static int read_mtx(FILE *f, const char *path, const char *banner, int binary, sp_matrix *out) {
  char obj[32] = "", fmt[32] = "", field[32] = "", sym[32] = "";
  if (sscanf(banner, "%%%%MatrixMarket %31s %31s %31s %31s", obj, fmt, field, sym) < 4) {
    sp_set_error("%s: malformed MatrixMarket banner", path);
    return SP_EIO;
  }
  for (char *p = field; *p; ++p) *p = (char)tolower((unsigned char)*p);
  for (char *p = sym; *p; ++p) *p = (char)tolower((unsigned char)*p);
  for (char *p = fmt; *p; ++p) *p = (char)tolower((unsigned char)*p);
  if (strcmp(fmt, "coordinate") != 0) { sp_set_error("%s: only coordinate MatrixMarket files are supported", path); return SP_EIO; }
  const int pattern = strcmp(field, "pattern") == 0;
  const int integer = strcmp(field, "integer") == 0;
  if (!pattern && !integer && strcmp(field, "real") != 0 && strcmp(field, "double") != 0) {
    sp_set_error("%s: unsupported MatrixMarket field `%s`", path, field);
    return SP_EIO;
  }
  const int symmetric = strcmp(sym, "symmetric") == 0;
  if (!symmetric && strcmp(sym, "general") != 0) { sp_set_error("%s: unsupported MatrixMarket symmetry `%s`", path, sym); return SP_EIO; }
  char *line = NULL;
  size_t cap = 0;
  int rc = SP_OK, rows = 0, cols = 0, entries = 0, have_size = 0;
  while (getline(&line, &cap, f) >= 0) {
    const char *p = line;
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == '%' || *p == '\n' || *p == '\r' || *p == '\0') continue;
    if (!have_size) {
      if (sscanf(p, "%d %d %d", &rows, &cols, &entries) < 3) { sp_set_error("%s: missing size line", path); rc = SP_EIO; break; }
      if (rows != cols) { sp_set_error("%s: matrix is %d x %d, the permanent needs a square matrix", path, rows, cols); rc = SP_EINVAL; break; }
      if ((rc = alloc_dense(out, rows)) != SP_OK) break;
      out->type = (pattern || integer) ? SP_TYPE_INT : SP_TYPE_DOUBLE;
      out->header_nnz = entries;
      have_size = 1;
      continue;
    }
    char *end;
    long i = strtol(p, &end, 10);
    if (end == p) continue;
    p = end;
    long j = strtol(p, &end, 10);
    if (end == p) continue;
    double v = 1.0;
    if (!pattern) {
      p = end;
      v = strtod(p, &end);
      if (end == p) continue;
    }
    if (binary || pattern) v = 1.0;
    --i; --j;
    if (i < 0 || j < 0 || i >= rows || j >= rows) continue;
    out->mat[(size_t)i * rows + j] = v;
    if (symmetric && i != j) out->mat[(size_t)j * rows + i] = v;
  }
  free(line);
  if (rc == SP_OK && !have_size) { sp_set_error("%s: missing size line", path); rc = SP_EIO; }
  return rc;
}

int sp_matrix_read(const char *path, int binary, sp_matrix *out) {
  if (!path || !out) { sp_set_error("null argument"); return SP_EINVAL; }
  matrix_zero(out);
  FILE *f = fopen(path, "r");
  if (!f) { sp_set_error("cannot open %s: %s", path, strerror(errno)); return SP_EIO; }
  char *line = NULL;
  size_t cap = 0;
  int rc = SP_OK;
  if (getline(&line, &cap, f) < 0) {
    sp_set_error("%s: empty file", path);
    rc = SP_EIO;
    goto done;
  }
  if (strncmp(line, "%%MatrixMarket", 14) == 0) {
    rc = read_mtx(f, path, line, binary, out);
    goto done;
  }
  int nov = 0, nnz = 0;
  char tname[32] = "";
  if (sscanf(line, "%d %d %31s", &nov, &nnz, tname) < 3) {
    sp_set_error("%s: header must be `nov nnz {int|float|double}`", path);
    rc = SP_EIO;
    goto done;
  }
  int type;
  if (strcmp(tname, "int") == 0) type = SP_TYPE_INT;
  else if (strcmp(tname, "float") == 0) type = SP_TYPE_FLOAT;
  else if (strcmp(tname, "double") == 0) type = SP_TYPE_DOUBLE;
  else { sp_set_error("%s: unknown element type `%s`", path, tname); rc = SP_EIO; goto done; }
  if ((rc = alloc_dense(out, nov)) != SP_OK) goto done;
  out->type = type;
  out->header_nnz = nnz;
  while (getline(&line, &cap, f) >= 0) {
    int i, j;
    double v;
    if (!parse_triple(line, type, &i, &j, &v)) continue;          /* erroneous line: skipped */
    if (i < 0 || j < 0 || i >= nov || j >= nov) continue;
    out->mat[(size_t)i * nov + j] = binary ? 1.0 : v;
  }
done:
  free(line);
  fclose(f);
  if (rc != SP_OK) sp_matrix_free(out);
  return rc;
}

/* CRS (rptrs, cols, rvals) and CCS (cptrs, rows, cvals) of the entries > 0, both in ascending
 * index order inside a row / column. */
static int build_compressed(sp_matrix *m) {
  const int n = m->nov;
  int count = 0;
  for (size_t e = 0; e < (size_t)n * n; ++e) count += (m->mat[e] > 0);
  free(m->cptrs); free(m->rows); free(m->cvals); free(m->rptrs); free(m->cols); free(m->rvals);
  const size_t c1 = (size_t)(count > 0 ? count : 1);
  m->cptrs = (int *)malloc((size_t)(n + 1) * sizeof(int));
  m->rptrs = (int *)malloc((size_t)(n + 1) * sizeof(int));
  m->rows = (int *)malloc(c1 * sizeof(int));
  m->cols = (int *)malloc(c1 * sizeof(int));
  m->cvals = (double *)malloc(c1 * sizeof(double));
  m->rvals = (double *)malloc(c1 * sizeof(double));
  if (!m->cptrs || !m->rptrs || !m->rows || !m->cols || !m->cvals || !m->rvals) {
    sp_set_error("out of memory");
    return SP_ENOMEM;
  }
  int r = 0, c = 0;
  for (int a = 0; a < n; ++a) {
    m->rptrs[a] = r;
    m->cptrs[a] = c;
    for (int b = 0; b < n; ++b) {
      const double row_entry = m->mat[(size_t)a * n + b];   /* A[a][b] */
      const double col_entry = m->mat[(size_t)b * n + a];   /* A[b][a] */
      if (row_entry > 0) { m->cols[r] = b; m->rvals[r] = row_entry; ++r; }
      if (col_entry > 0) { m->rows[c] = b; m->cvals[c] = col_entry; ++c; }
    }
  }
  m->rptrs[n] = r;
  m->cptrs[n] = c;
  m->nnz = count;
  return SP_OK;
}

/* mat <- mat[rowperm[r]][colperm[c]] */
static int permute_dense(sp_matrix *m, const int *rowperm, const int *colperm) {
  const int n = m->nov;
  double *fresh = (double *)malloc((size_t)n * n * sizeof(double));
  if (!fresh) { sp_set_error("out of memory"); return SP_ENOMEM; }
  for (int r = 0; r < n; ++r) {
    const double *src = m->mat + (size_t)(rowperm ? rowperm[r] : r) * n;
    for (int c = 0; c < n; ++c) fresh[(size_t)r * n + c] = src[colperm ? colperm[c] : c];
  }
  free(m->mat);
  m->mat = fresh;
  return SP_OK;
}

/* SortOrder: columns ascending by their number of positive entries, ties in original order (what
 * the reference's 0/1 qsort comparator does on glibc).  The reference also
 * rewrites `mat` from the CCS it built, which drops entries <= 0; so do we. */
static int order_sort(sp_matrix *m) {
  const int n = m->nov;
  int *perm = (int *)malloc((size_t)n * sizeof(int));
  int *deg = (int *)calloc((size_t)n, sizeof(int));
  if (!perm || !deg) { free(perm); free(deg); sp_set_error("out of memory"); return SP_ENOMEM; }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) deg[c] += (m->mat[(size_t)r * n + c] > 0);
  /* counting sort by degree: stable by construction */
  int pos = 0;
  for (int d = 0; d <= n; ++d)
    for (int c = 0; c < n; ++c)
      if (deg[c] == d) perm[pos++] = c;
  for (size_t e = 0; e < (size_t)n * n; ++e)
    if (!(m->mat[e] > 0)) m->mat[e] = 0.0;
  int rc = permute_dense(m, NULL, perm);
  free(perm);
  free(deg);
  return rc;
}

/* SkipOrder: repeatedly take the unplaced column of minimum remaining degree (lowest index on
 * ties; degrees count entries != 0 in rows not yet placed; the reference's sentinel INT8_MAX = 127
 * also caps the comparison, util.h:643-650), then place the not-yet-placed rows it touches, in
 * ascending order, each one lowering the degree of the other columns it touches. */
static int order_skip(sp_matrix *m) {
  const int n = m->nov;
  enum { PLACED = 127 };
  int *colperm = (int *)malloc((size_t)n * sizeof(int));
  int *rowperm = (int *)malloc((size_t)n * sizeof(int));
  int *deg = (int *)calloc((size_t)n, sizeof(int));
  char *row_done = (char *)calloc((size_t)n, 1);
  if (!colperm || !rowperm || !deg || !row_done) {
    free(colperm); free(rowperm); free(deg); free(row_done);
    sp_set_error("out of memory");
    return SP_ENOMEM;
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) deg[c] += (m->mat[(size_t)r * n + c] != 0);
  int nrows = 0;
  for (int step = 0; step < n; ++step) {
    int pick = -1, best = PLACED;
    for (int c = 0; c < n; ++c)
      if (deg[c] < best) { best = deg[c]; pick = c; }
    if (pick < 0) {                       /* only columns of degree >= 127 left (n > 127) */
      for (int c = 0; c < n && pick < 0; ++c)
        if (deg[c] != PLACED) pick = c;
      if (pick < 0) pick = 0;
    }
    deg[pick] = PLACED;
    colperm[step] = pick;
    for (int r = 0; r < n; ++r) {
      if (row_done[r] || m->mat[(size_t)r * n + pick] == 0) continue;
      row_done[r] = 1;
      rowperm[nrows++] = r;
      for (int c = 0; c < n; ++c)
        if (m->mat[(size_t)r * n + c] != 0 && deg[c] != PLACED) deg[c]--;
    }
  }
  for (int r = 0; r < n; ++r)             /* all-zero rows: the reference leaves them undefined */
    if (!row_done[r]) rowperm[nrows++] = r;
  int rc = permute_dense(m, rowperm, colperm);
  free(colperm); free(rowperm); free(deg); free(row_done);
  return rc;
}

int sp_matrix_compress(sp_matrix *m, int preprocessing) {
  if (!m || !m->mat) { sp_set_error("matrix not loaded"); return SP_EINVAL; }
  int rc = SP_OK;
  if (preprocessing == 1) rc = order_sort(m);
  else if (preprocessing == 2) rc = order_skip(m);
  if (rc != SP_OK) return rc;
  return build_compressed(m);
}

/* Bipartite biadjacency matrix of the m x n grid graph: vertices are coloured like a
 * checkerboard; the vertex at (i, j) of the `row x col` layout (col = the even dimension) gets
 * number x = i*(col/2) + j/2 inside its colour class, and mat[black][white] = 1 for every grid
 * edge.  Same numbering as gridGraph2compressed (util.h:403-520), so the 0/1 pattern -- and
 * therefore every trial of the approximators -- is identical.  Both dimensions odd -> -1. */
int sp_matrix_grid(int gm, int gn, sp_matrix *out) {
  if (!out) { sp_set_error("null argument"); return SP_EINVAL; }
  matrix_zero(out);
  if (gm < 1 || gn < 1) { sp_set_error("grid dimensions must be positive"); return SP_EINVAL; }
  if ((gm & 1) && (gn & 1)) {
    sp_set_error("one of the grid dimensions should be even");
    return SP_EINVAL;
  }
  const int col = (gm % 2 == 0) ? gm : gn;
  const int row = (gm % 2 == 0) ? gn : gm;
  const int half = col / 2;
  const int nov = gm * gn / 2;
  int rc = alloc_dense(out, nov);
  if (rc != SP_OK) return rc;
  out->type = SP_TYPE_INT;
  for (int i = 0; i < row; ++i) {
    for (int j = 0; j < col; ++j) {
      const int x = i * half + j / 2;
      const int black = ((i + j) % 2 == 0);
      int nb[4], k = 0;
      if (x - half >= 0) nb[k++] = x - half;            /* vertical neighbours */
      if (x + half < nov) nb[k++] = x + half;
      nb[k++] = x;                                      /* horizontal neighbour sharing x */
      if ((j % 2 == 0) && j > 0) nb[k++] = x - 1;
      if ((j % 2 == 1) && j < col - 1) nb[k++] = x + 1;
      for (int e = 0; e < k; ++e) {
        if (black) out->mat[(size_t)x * nov + nb[e]] = 1.0;
        else       out->mat[(size_t)nb[e] * nov + x] = 1.0;
      }
    }
  }
  rc = build_compressed(out);
  if (rc != SP_OK) { sp_matrix_free(out); return rc; }
  out->header_nnz = out->nnz;
  return SP_OK;
}
