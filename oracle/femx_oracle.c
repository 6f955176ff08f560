/*
 * femx_oracle.c — CPU restatement of the reference's assembly path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cuda-fem_b200/, the C-ABI
 * library) may link, import or execute this file.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker or the reported CPU baseline.
 *
 * Parity status
 *   - 2-D P1 Poisson (the only form the reference implements): PINNED.  The
 *     element values are checked in tests/ against (1) the reference's own GiNaC
 *     output strings (fea_test_sm_sym_sparse2.cu:188-205, evaluated verbatim by
 *     tests/golden/make_golden.py) and (2) the known-answer triplets of the
 *     2x2 mesh; mesh and pattern against the reference's compiled host code
 *     (oracle/_ref, built by oracle/build_ref.py) when /root/reference exists.
 *   - mass term, 3-D tets, elasticity: the reference has no implementation →
 *     "parity unpinned" by the reference; this file is the definition, checked
 *     against closed forms (sympy-free identities: row sums, volume, rigid modes).
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference checkout).  Serial, plain C99, fp64 throughout (orc_*_f32 variants
 * round inputs/outputs the way the fp32 reference kernel does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- quadrature literals: fea_symbolic_nvrtc_sparse.cpp:380-383 (SURVEY Q9:
 * 8-digit constants; triT is its own literal, not 1-r-s). */
static const double TRI_W[7] = {0.06296959, 0.06619708, 0.06296959, 0.06619708,
                                0.06296959, 0.06619708, 0.11250000};
static const double TRI_R[7] = {0.10128651, 0.47014206, 0.79742699, 0.47014206,
                                0.10128651, 0.05971587, 0.33333333};
static const double TRI_S[7] = {0.10128651, 0.05971587, 0.10128651, 0.47014206,
                                0.79742699, 0.47014206, 0.33333333};
/* 4-point degree-2 tetrahedron rule (no reference counterpart) */
#define TET_A 0.5854101966249685
#define TET_B 0.1381966011250105
static const double TET_W[4] = {1.0 / 24.0, 1.0 / 24.0, 1.0 / 24.0, 1.0 / 24.0};
static const double TET_R[4] = {TET_A, TET_B, TET_B, TET_B};
static const double TET_S[4] = {TET_B, TET_A, TET_B, TET_B};
static const double TET_T[4] = {TET_B, TET_B, TET_A, TET_B};

void orc_tri_rule(double* w, double* r, double* s) {
  memcpy(w, TRI_W, sizeof TRI_W);
  memcpy(r, TRI_R, sizeof TRI_R);
  memcpy(s, TRI_S, sizeof TRI_S);
}

/* ---- RectangleMesh::generate, fea_symbolic_nvrtc_sparse.cpp:170-216.
 * Nodes row-major idx = i*(nCol+1)+j, x = x0+j*stepx, y = y0+i*stepy; boundary
 * flag; per cell two triangles (n1,n1+1,n3) and (n1+1, n3+1, n3). */
void orc_rect_mesh(double x0, double x1, double y0, double y1, int64_t nRow,
                   int64_t nCol, double* X, double* Y, int32_t* flag, int32_t* conn) {
  double stepx = (x1 - x0) / (double)nCol;
  double stepy = (y1 - y0) / (double)nRow;
  for (int64_t i = 0; i <= nRow; i++) {
    double y = y0 + (double)i * stepy;
    for (int64_t j = 0; j <= nCol; j++) {
      int64_t n = i * (nCol + 1) + j;
      if (X) X[n] = x0 + (double)j * stepx;
      if (Y) Y[n] = y;
      if (flag) flag[n] = (i == 0 || i == nRow || j == 0 || j == nCol) ? 1 : 0;
    }
  }
  if (!conn) return;
  int64_t e = 0;
  for (int64_t i = 0; i < nRow; i++)
    for (int64_t j = 0; j < nCol; j++) {
      int64_t n1 = i * (nCol + 1) + j, n3 = (i + 1) * (nCol + 1) + j;
      conn[3 * e + 0] = (int32_t)n1;
      conn[3 * e + 1] = (int32_t)(n1 + 1);
      conn[3 * e + 2] = (int32_t)n3;
      e++;
      conn[3 * e + 0] = (int32_t)(n1 + 1);
      conn[3 * e + 1] = (int32_t)(n3 + 1);
      conn[3 * e + 2] = (int32_t)n3;
      e++;
    }
}

/* Kuhn 6-tet box (new).  For permutation p of the axes the
 * tet is base, base+e_p0, base+e_p0+e_p1, base+(1,1,1); odd permutations have
 * their two middle vertices swapped so every tet has positive orientation. */
static const int KUHN_PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2},
                                    {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
static const int KUHN_ODD[6] = {0, 1, 1, 0, 0, 1};

void orc_box_mesh(double x0, double x1, double y0, double y1, double z0, double z1,
                  int64_t nx, int64_t ny, int64_t nz, double* X, double* Y, double* Z,
                  int32_t* conn) {
  double hx = (x1 - x0) / (double)nx, hy = (y1 - y0) / (double)ny,
         hz = (z1 - z0) / (double)nz;
  for (int64_t k = 0; k <= nz; k++)
    for (int64_t j = 0; j <= ny; j++)
      for (int64_t i = 0; i <= nx; i++) {
        int64_t n = (k * (ny + 1) + j) * (nx + 1) + i;
        if (X) X[n] = x0 + (double)i * hx;
        if (Y) Y[n] = y0 + (double)j * hy;
        if (Z) Z[n] = z0 + (double)k * hz;
      }
  if (!conn) return;
  int64_t e = 0;
  for (int64_t k = 0; k < nz; k++)
    for (int64_t j = 0; j < ny; j++)
      for (int64_t i = 0; i < nx; i++)
        for (int p = 0; p < 6; p++) {
          int64_t c[3] = {i, j, k};
          int64_t v[4];
          v[0] = (c[2] * (ny + 1) + c[1]) * (nx + 1) + c[0];
          for (int a = 0; a < 3; a++) {
            c[KUHN_PERM[p][a]] += 1;
            v[a + 1] = (c[2] * (ny + 1) + c[1]) * (nx + 1) + c[0];
          }
          /* positive orientation for det[x1-x4, x2-x4, x3-x4] */
          if (!KUHN_ODD[p]) {
            int64_t tmp = v[1];
            v[1] = v[2];
            v[2] = tmp;
          }
          for (int a = 0; a < 4; a++) conn[4 * e + a] = (int32_t)v[a];
          e++;
        }
}

/* ---- element matrices -----------------------------------------------------
 * Semantics of WeakForm::build (fea_symbolic_nvrtc_sparse.cpp:333-345) and of
 * the kernel's quadrature sum (:459-463, 473-477):
 *   Ae[li][lj] = sum_q w_q * ( a(u = phi_lj, v = phi_li) * jac )(r_q, s_q)
 * with phi = (r, s, 1-r-s) (:264-269), jac = dfx/dr dfy/ds - dfy/dr dfx/ds
 * (:281-289) for fx = x1 r + x2 s + x3 (1-r-s), and
 *   dr/dx = (y2-y3)/jac, dr/dy = (x3-x2)/jac  (:68-82)
 *   ds/dx = (y3-y1)/jac, ds/dy = (x1-x3)/jac  (:87-101).
 * jac is SIGNED: a clockwise element yields a negated matrix, as the reference. */
enum { ORC_POISSON = 1, ORC_POISSON_MASS = 2, ORC_MASS = 3, ORC_ELASTICITY = 4 };

static void tri_grads(const double* x, const double* y, double* jac, double g[3][2]) {
  double j = (x[0] - x[2]) * (y[1] - y[2]) - (y[0] - y[2]) * (x[1] - x[2]);
  *jac = j;
  g[0][0] = (y[1] - y[2]) / j;
  g[0][1] = (x[2] - x[1]) / j;
  g[1][0] = (y[2] - y[0]) / j;
  g[1][1] = (x[0] - x[2]) / j;
  g[2][0] = -(g[0][0] + g[1][0]); /* phi3 = 1 - r - s */
  g[2][1] = -(g[0][1] + g[1][1]);
}

static void tet_grads(const double* x, const double* y, const double* z, double* jac,
                      double g[4][3]) {
  /* affine map X = x1 r + x2 s + x3 t + x4 (1-r-s-t); J[c][a] = d X_c / d ref_a */
  double J[3][3] = {{x[0] - x[3], x[1] - x[3], x[2] - x[3]},
                    {y[0] - y[3], y[1] - y[3], y[2] - y[3]},
                    {z[0] - z[3], z[1] - z[3], z[2] - z[3]}};
  double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  *jac = det;
  /* inverse = adj/det; grad(ref_a)[c] = Jinv[a][c] */
  double inv[3][3];
  inv[0][0] = c00 / det;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  inv[1][0] = c01 / det;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  inv[2][0] = c02 / det;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  for (int a = 0; a < 3; a++)
    for (int c = 0; c < 3; c++) g[a][c] = inv[a][c];
  for (int c = 0; c < 3; c++) g[3][c] = -(g[0][c] + g[1][c] + g[2][c]);
}

/* Element matrix, n = nn*nd, row-major Ae[li*n+lj].  xyz: nn coords per axis.
 * nq/qw/qr/qs/qt: quadrature (NULL → defaults above). */
void orc_element_matrix(int form, int dim, int nd, const double* params,
                        const double* x, const double* y, const double* z, int nq,
                        const double* qw, const double* qr, const double* qs,
                        const double* qt, double* Ae) {
  int nn = dim + 1, n = nn * nd;
  double jac, g[4][3] = {{0}};
  if (dim == 2) {
    double g2[3][2];
    tri_grads(x, y, &jac, g2);
    for (int a = 0; a < 3; a++) {
      g[a][0] = g2[a][0];
      g[a][1] = g2[a][1];
    }
    if (!qw) { nq = 7; qw = TRI_W; qr = TRI_R; qs = TRI_S; }
  } else {
    tet_grads(x, y, z, &jac, g);
    if (!qw) { nq = 4; qw = TET_W; qr = TET_R; qs = TET_S; qt = TET_T; }
  }
  double cmass = (form == ORC_POISSON_MASS) ? (params && params[0] != 0.0 ? params[0] : 1.0)
                                             : 1.0;
  double lam = params ? params[0] : 0.0, mu = params ? params[1] : 0.0;
  for (int i = 0; i < n * n; i++) Ae[i] = 0.0;
  for (int q = 0; q < nq; q++) {
    double phi[4];
    phi[0] = qr[q];
    phi[1] = qs[q];
    if (dim == 2) {
      phi[2] = 1.0 - qr[q] - qs[q];
    } else {
      phi[2] = qt[q];
      phi[3] = 1.0 - qr[q] - qs[q] - qt[q];
    }
    for (int li = 0; li < n; li++)
      for (int lj = 0; lj < n; lj++) {
        int a = li / nd, c = li % nd, b = lj / nd, d = lj % nd;
        double gg = 0.0;
        for (int k = 0; k < dim; k++) gg += g[b][k] * g[a][k];
        double val = 0.0;
        switch (form) {
          case ORC_POISSON: val = gg; break;
          case ORC_POISSON_MASS: val = gg + cmass * phi[b] * phi[a]; break;
          case ORC_MASS: val = phi[b] * phi[a]; break;
          case ORC_ELASTICITY:
            val = lam * g[a][c] * g[b][d] + mu * ((c == d ? gg : 0.0) + g[a][d] * g[b][c]);
            break;
        }
        Ae[li * n + lj] += qw[q] * (val * jac);
      }
  }
}

/* ---- load vector.  WeakForm::build also generates rhs[j] = f*phi_j*jac
 * (fea_symbolic_nvrtc_sparse.cpp:346-351; recorded output fea_symbolic.cu:335,339,343) and then
 * discards it.  b[dof] = sum_e sum_q w_q f(x_q) phi_li(q) jac, scatter-add in element order.
 *   kind 0: constant source, f = fvec[comp]         (scalar forms: fvec[0])
 *   kind 1: the reference's f = -2 (x^2 + y^2) + 36  (fea_symbolic_nvrtc_sparse.cpp:495), 2-D scalar */
void orc_assemble_rhs(int kind, int dim, int nd, const double* fvec, int64_t n_elems,
                      const int32_t* conn, const double* X, const double* Y, const double* Z,
                      int64_t n_dofs, double* b, double* b_elem) {
  int nn = dim + 1;
  const double *qw, *qr, *qs, *qt = NULL;
  int nq;
  if (dim == 2) { nq = 7; qw = TRI_W; qr = TRI_R; qs = TRI_S; }
  else { nq = 4; qw = TET_W; qr = TET_R; qs = TET_S; qt = TET_T; }
  for (int64_t i = 0; i < n_dofs; i++) b[i] = 0.0;
  for (int64_t e = 0; e < n_elems; e++) {
    double x[4], y[4], z[4] = {0, 0, 0, 0}, jac, g[4][3];
    for (int a = 0; a < nn; a++) {
      int32_t nd_ = conn[nn * e + a];
      x[a] = X[nd_]; y[a] = Y[nd_];
      if (dim == 3) z[a] = Z[nd_];
    }
    if (dim == 2) { double g2[3][2]; tri_grads(x, y, &jac, g2); }
    else tet_grads(x, y, z, &jac, g);
    for (int a = 0; a < nn; a++)
      for (int c = 0; c < nd; c++) {
        double acc = 0.0;
        for (int q = 0; q < nq; q++) {
          double phi[4];
          phi[0] = qr[q]; phi[1] = qs[q];
          if (dim == 2) phi[2] = 1.0 - qr[q] - qs[q];
          else { phi[2] = qt[q]; phi[3] = 1.0 - qr[q] - qs[q] - qt[q]; }
          double f;
          if (kind == 0) f = fvec[c];
          else {
            double xq = 0.0, yq = 0.0;
            for (int k = 0; k < nn; k++) { xq += x[k] * phi[k]; yq += y[k] * phi[k]; }
            f = -2.0 * (xq * xq + yq * yq) + 36.0;
          }
          acc += qw[q] * (f * phi[a] * jac);
        }
        if (b_elem) b_elem[(e * nn + a) * nd + c] = acc;
        b[(int64_t)nd * conn[nn * e + a] + c] += acc;
      }
  }
}

/* ---- COO surface: slot = e*n*n + li*n + lj, rowA = dof(li), colA = dof(lj)
 * (fea_symbolic_nvrtc_sparse.cpp:444-445, 473-477).  Coordinates are node-indexed
 * (X[node]); dof = nd*node + comp. */
void orc_assemble_coo(int form, int dim, int nd, const double* params, int64_t n_elems,
                      const int32_t* conn, const double* X, const double* Y,
                      const double* Z, double* A, int32_t* rowA, int32_t* colA) {
  int nn = dim + 1, n = nn * nd;
  double Ae[144];
  for (int64_t e = 0; e < n_elems; e++) {
    double x[4], y[4], z[4] = {0, 0, 0, 0};
    for (int a = 0; a < nn; a++) {
      int32_t nd_ = conn[nn * e + a];
      x[a] = X[nd_];
      y[a] = Y[nd_];
      if (dim == 3) z[a] = Z[nd_];
    }
    orc_element_matrix(form, dim, nd, params, x, y, z, 0, NULL, NULL, NULL, NULL, Ae);
    for (int li = 0; li < n; li++)
      for (int lj = 0; lj < n; lj++) {
        int64_t slot = e * n * n + li * n + lj;
        if (A) A[slot] = Ae[li * n + lj];
        if (rowA) rowA[slot] = nd * conn[nn * e + li / nd] + li % nd;
        if (colA) colA[slot] = nd * conn[nn * e + lj / nd] + lj % nd;
      }
  }
}

/* ---- pattern: Mesh::getNeighborNodesList, fea_symbolic_nvrtc_sparse2.cpp:181-210.
 * Per node the ascending duplicate-free set of all nodes sharing an element,
 * self included.  Two-pass: returns node-level row_ptr (n_nodes+1) and, when
 * col_idx != NULL, the columns.  Call with col_idx == NULL first to size. */
static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

int64_t orc_pattern(int nn, int64_t n_nodes, int64_t n_elems, const int32_t* conn,
                    int64_t* row_ptr, int32_t* col_idx) {
  /* node → element adjacency (counting sort), then per node sort+unique */
  int64_t* ptr = (int64_t*)calloc((size_t)n_nodes + 1, sizeof(int64_t));
  for (int64_t k = 0; k < n_elems * nn; k++) ptr[conn[k] + 1]++;
  for (int64_t i = 0; i < n_nodes; i++) ptr[i + 1] += ptr[i];
  int32_t* adj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_elems * nn + 1));
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_nodes + 1));
  memcpy(cur, ptr, sizeof(int64_t) * (size_t)n_nodes);
  for (int64_t e = 0; e < n_elems; e++)
    for (int a = 0; a < nn; a++) adj[cur[conn[nn * e + a]]++] = (int32_t)e;
  int64_t cap = 64, nnz = 0;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  row_ptr[0] = 0;
  for (int64_t i = 0; i < n_nodes; i++) {
    int64_t m = (ptr[i + 1] - ptr[i]) * nn;
    if (m > cap) {
      cap = 2 * m;
      tmp = (int32_t*)realloc(tmp, sizeof(int32_t) * (size_t)cap);
    }
    int64_t c = 0;
    for (int64_t k = ptr[i]; k < ptr[i + 1]; k++)
      for (int a = 0; a < nn; a++) tmp[c++] = conn[nn * adj[k] + a];
    qsort(tmp, (size_t)c, sizeof(int32_t), cmp_i32);
    int64_t u = 0;
    for (int64_t k = 0; k < c; k++)
      if (k == 0 || tmp[k] != tmp[k - 1]) {
        if (col_idx) col_idx[nnz + u] = tmp[k];
        u++;
      }
    nnz += u;
    row_ptr[i + 1] = nnz;
  }
  free(tmp);
  free(cur);
  free(adj);
  free(ptr);
  return nnz;
}

/* Reference padded layout: len[i], idx[i*width+j]
 * (fea_symbolic_nvrtc_sparse2.cpp:205-207), zero fill (:642-644). */
void orc_csr_to_ell_pattern(int64_t n_nodes, const int64_t* row_ptr,
                            const int32_t* col_idx, int width, int32_t* len,
                            int32_t* idx) {
  for (int64_t i = 0; i < n_nodes; i++) {
    int64_t l = row_ptr[i + 1] - row_ptr[i];
    len[i] = (int32_t)l;
    for (int j = 0; j < width; j++)
      idx[i * width + j] = j < l ? col_idx[row_ptr[i] + j] : 0;
  }
}

/* dof-level CSR from the node-level pattern: row nd*i+c holds columns
 * nd*col+d ascending (SURVEY §8a definition for nd > 1). */
void orc_expand_pattern(int nd, int64_t n_nodes, const int64_t* row_ptr,
                        const int32_t* col_idx, int64_t* drow_ptr, int32_t* dcol_idx) {
  int64_t k = 0;
  drow_ptr[0] = 0;
  for (int64_t i = 0; i < n_nodes; i++)
    for (int c = 0; c < nd; c++) {
      for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; p++)
        for (int d = 0; d < nd; d++) {
          if (dcol_idx) dcol_idx[k] = nd * col_idx[p] + d;
          k++;
        }
      drow_ptr[nd * i + c + 1] = k;
    }
}

/* ---- numeric pass: serial scatter-add in element order into the prebuilt
 * dof-level CSR (what fea_kernel's linear search + atomicAdd computes,
 * fea_symbolic_nvrtc_sparse2.cpp:533-544, minus the race; loop style of the
 * author's commented host loop fea_kernal.cu:193-214). */
void orc_assemble_csr(int form, int dim, int nd, const double* params, int64_t n_elems,
                      const int32_t* conn, const double* X, const double* Y,
                      const double* Z, const int64_t* drow_ptr, const int32_t* dcol_idx,
                      int64_t n_rows, double* values) {
  int nn = dim + 1, n = nn * nd;
  double Ae[144];
  memset(values, 0, sizeof(double) * (size_t)drow_ptr[n_rows]);
  for (int64_t e = 0; e < n_elems; e++) {
    double x[4], y[4], z[4] = {0, 0, 0, 0};
    for (int a = 0; a < nn; a++) {
      int32_t nd_ = conn[nn * e + a];
      x[a] = X[nd_];
      y[a] = Y[nd_];
      if (dim == 3) z[a] = Z[nd_];
    }
    orc_element_matrix(form, dim, nd, params, x, y, z, 0, NULL, NULL, NULL, NULL, Ae);
    for (int li = 0; li < n; li++) {
      int64_t gi = (int64_t)nd * conn[nn * e + li / nd] + li % nd;
      int64_t lo = drow_ptr[gi], hi = drow_ptr[gi + 1];
      for (int lj = 0; lj < n; lj++) {
        int32_t gj = nd * conn[nn * e + lj / nd] + lj % nd;
        int64_t a = lo, b = hi; /* binary search; the reference searches linearly */
        while (a < b) {
          int64_t m = (a + b) >> 1;
          if (dcol_idx[m] < gj) a = m + 1; else b = m;
        }
        values[a] += Ae[li * n + lj];
      }
    }
  }
}

/* ---- Dirichlet conditions by symmetric elimination on the dof-level CSR (the
 * reference sets Node::flag, fea_test.cu:100-103, and never uses it). */
void orc_apply_dirichlet(int64_t n_rows, const int64_t* row_ptr, const int32_t* col_idx,
                         const int32_t* flag, const double* g, double* values, double* rhs) {
  for (int64_t j = 0; j < n_rows; j++) {
    double corr = 0.0;
    for (int64_t k = row_ptr[j]; k < row_ptr[j + 1]; k++) {
      int32_t i = col_idx[k];
      if (flag[j]) values[k] = (i == j) ? 1.0 : 0.0;
      else if (flag[i]) { corr += values[k] * g[i]; values[k] = 0.0; }
    }
    if (rhs) rhs[j] = flag[j] ? g[j] : rhs[j] - corr;
  }
}

/* ---- validation helpers: y = A x, and unpreconditioned CG (SURVEY cfg5). */
void orc_spmv(int64_t n_rows, const int64_t* row_ptr, const int32_t* col_idx,
              const double* values, const double* x, double* y) {
  for (int64_t i = 0; i < n_rows; i++) {
    double s = 0.0;
    for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) s += values[k] * x[col_idx[k]];
    y[i] = s;
  }
}

/* Returns the number of iterations done; res[it] = ||r_it||_2 (it = 0..iters). */
int orc_cg(int64_t n, const int64_t* row_ptr, const int32_t* col_idx,
           const double* values, const double* b, double* x, int iters, double* res) {
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  double* Ap = (double*)malloc(sizeof(double) * (size_t)n);
  double rr = 0.0;
  for (int64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; rr += r[i] * r[i]; }
  if (res) res[0] = sqrt(rr);
  int it = 0;
  for (; it < iters; it++) {
    orc_spmv(n, row_ptr, col_idx, values, p, Ap);
    double pAp = 0.0;
    for (int64_t i = 0; i < n; i++) pAp += p[i] * Ap[i];
    if (pAp == 0.0) break;
    double alpha = rr / pAp, rr2 = 0.0;
    for (int64_t i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rr2 += r[i] * r[i]; }
    double beta = rr2 / rr;
    rr = rr2;
    if (res) res[it + 1] = sqrt(rr);
    for (int64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
  }
  free(r); free(p); free(Ap);
  return it;
}
