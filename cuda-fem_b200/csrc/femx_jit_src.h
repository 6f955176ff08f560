// CUDA source templates compiled at run time by NVRTC (sm_100a cubin).
//
// The generator (femx_form.cpp) prepends a block of #defines:
//   FEMX_REAL, NN, ND, DIM, FEMX_TILE_NODES, FEMX_PROLOGUE,
//   FEMX_ROW_<li>(R,S,T,U,W)  — one quadrature-point update of matrix row li,
//   FEMX_QUAD(M)              — M(r,s,t,u,w) instantiated for every point.
// This plays the part of the reference's codeTemplate string
// (fea_symbolic_nvrtc_sparse.cpp:379-481) but is organised by matrix ROW so
// that the owner of a CSR row can evaluate just the entries it stores.
#pragma once

static const char* const kFemxJitCommon = R"FEMX(
typedef FEMX_REAL real;
typedef long long i64;

// pow with the constant exponents GiNaC prints (pow(x,2.0)); folds to x*x.
template <class T>
__device__ __forceinline__ T femx_pow(T a, double e) {
  if (e == 2.0) return a * a;
  if (e == 3.0) return a * a * a;
  if (e == -1.0) return T(1.0) / a;
  if (e == -2.0) return T(1.0) / (a * a);
  return (T)pow((double)a, e);
}
namespace std {
template <class T>
__device__ __forceinline__ T femx_pow(T a, double e) { return ::femx_pow(a, e); }
}
#define pow femx_pow
#define powf femx_pow

// Reciprocal for the emitter's prologue: hardware seed (rcp.approx.ftz.f64, >= 20 good bits) and two
// Newton steps -> relative error far below 1 ulp of a double for normal inputs, branch-free
// (the compiler's IEEE division adds a slow-path test and a fifth correction FMA per element).
// FEMX_RCP3: one third-order step instead, x(1 + e + e^2) with e = 1 - a x: error e^3 <= 2^-60, one fma less.
#ifndef FEMX_HOST_EMU  // (tests compile the lattice kernel for the host: tests/test_lattice_host.py)
__device__ __forceinline__ double femx_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
#if FEMX_RCP3
  x = fma(x, fma(e, e, e), x);
#else
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  x = fma(x, e, x);
#endif
  return x;
}
__device__ __forceinline__ float femx_rcp(float a) { return 1.0f / a; }
// A product that must stay a product: the built-in forms spell every fused multiply-add out
// (fma) and wrap the remaining products in femx_mul, so that no contraction decision is left to
// the compiler and the specialised and the generic numeric pass round identically.
__device__ __forceinline__ double femx_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float femx_mul(float a, float b) { return __fmul_rn(a, b); }
#endif
// rowinfo[i].y = #incidences | FEMX_ROW_SPEC | own position << 24 (femx_internal.h)
#define FEMX_NP_MASK 0x3fffff
#define FEMX_ROW_SPEC (1 << 23)
// column offsets (column - own node) of the pattern's stencil class, passed by value at launch
struct femx_soff { int v[24]; };

#define NDOF (NN * ND)
// local index of the j-th other vertex of an incidence at local vertex li (even permutation; see femx_internal.h)
#define FEMX_OTH(li, j) (NN == 4 ? ((li) ^ ((j) + 1)) : (((li) + 1 + (j)) % 3))

// Row `li` of the element matrix: out[lj] = sum_q w_q * integrand(li, lj).
__device__ __forceinline__ void femx_row(const int li, const real* cx, const real* cy,
                                         const real* cz, real* out) {
  const real x1 = cx[0], x2 = cx[1], x3 = cx[2];
  const real y1 = cy[0], y2 = cy[1], y3 = cy[2];
#if DIM == 3
  const real x4 = cx[3], y4 = cy[3];
  const real z1 = cz[0], z2 = cz[1], z3 = cz[2], z4 = cz[3];
#endif
  FEMX_PROLOGUE
#pragma unroll
  for (int j = 0; j < NDOF; ++j) out[j] = real(0);
  switch (li) {
    FEMX_ROW_CASES
  }
}
)FEMX";

// ---- kernel ABI #1: COO triplets -------------------------------------------
// One thread per (element, local row).  Slot order e*n*n + li*n + lj and the
// (row = dof of li, col = dof of lj) orientation are the reference's
// (fea_symbolic_nvrtc_sparse.cpp:444-445, 473-477).  A warp writes one
// contiguous run of 32*n values.
static const char* const kFemxJitCoo = R"FEMX(
extern "C" __global__ void __launch_bounds__(256)
femx_coo(const int* __restrict__ conn, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
         const int expanded, real* __restrict__ A, int* __restrict__ rowA,
         int* __restrict__ colA, const i64 n_elems) {
  const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_elems * NDOF) return;
  const i64 e = tid / NDOF;
  const int li = (int)(tid - e * NDOF);
  int nodes[NN];
#pragma unroll
  for (int a = 0; a < NN; ++a) nodes[a] = conn ? __ldg(conn + e * NN + a) : 0;
  real cx[NN], cy[NN], cz[NN];
#pragma unroll
  for (int a = 0; a < NN; ++a) {
    const i64 p = expanded ? (e * NN + a) : (i64)nodes[a] * cs;
    cx[a] = __ldg(X + p);
    cy[a] = __ldg(Y + p);
#if DIM == 3
    cz[a] = __ldg(Z + p);
#else
    cz[a] = real(0);
#endif
  }
  real out[NDOF];
  femx_row(li, cx, cy, cz, out);
  const i64 base = tid * NDOF;
  const int gi = ND * nodes[li / ND] + li % ND;
#pragma unroll
  for (int lj = 0; lj < NDOF; ++lj) {
    if (A) A[base + lj] = out[lj];
    if (rowA) rowA[base + lj] = gi;
    if (colA) colA[base + lj] = ND * nodes[lj / ND] + lj % ND;
  }
}
)FEMX";

// ---- kernel ABI #1, scalar forms: one thread per ELEMENT -------------------------------
// Geometry is evaluated once per element, all n*n entries are produced by straight-line code
// (no divergent row switch), staged in shared memory in slot order and written with three TMA
// bulk stores per 128-element tile (values, rows, columns): the global stores are perfectly
// coalesced and cost no LSU work.  Same slot order / orientation as femx_coo.
static const char* const kFemxJitCooElem = R"FEMX(
#define FEMX_COO_TILE 128
#define N2 (NDOF * NDOF)
#define FEMX_COO_STORE(LI)                                                     \
  _Pragma("unroll") for (int lj = 0; lj < NDOF; ++lj) {                        \
    sA[t * N2 + (LI) * NDOF + lj] = out[lj];                                   \
    sR[t * N2 + (LI) * NDOF + lj] = ND * nodes[(LI) / ND] + (LI) % ND;         \
    sC[t * N2 + (LI) * NDOF + lj] = ND * nodes[lj / ND] + lj % ND;             \
  }

extern "C" __global__ void __launch_bounds__(FEMX_COO_TILE)
femx_coo(const int* __restrict__ conn, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
         const int expanded, real* __restrict__ A, int* __restrict__ rowA,
         int* __restrict__ colA, const i64 n_elems) {
  extern __shared__ __align__(128) unsigned char femx_smem[];
  real* sA = reinterpret_cast<real*>(femx_smem);
  int* sR = reinterpret_cast<int*>(sA + FEMX_COO_TILE * N2);
  int* sC = sR + FEMX_COO_TILE * N2;
  const i64 e0 = (i64)blockIdx.x * FEMX_COO_TILE;
  const int ne_t = (int)min((i64)FEMX_COO_TILE, n_elems - e0);
  const int t = threadIdx.x;
  if (t < ne_t) {
    const i64 e = e0 + t;
    int nodes[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) nodes[a] = conn ? __ldg(conn + e * NN + a) : 0;
    real cx[NN], cy[NN], cz[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
      const i64 p = expanded ? (e * NN + a) : (i64)nodes[a] * cs;
      cx[a] = __ldg(X + p);
      cy[a] = __ldg(Y + p);
      cz[a] = DIM == 3 ? __ldg(Z + p) : real(0);
    }
    const real x1 = cx[0], x2 = cx[1], x3 = cx[2];
    const real y1 = cy[0], y2 = cy[1], y3 = cy[2];
#if DIM == 3
    const real x4 = cx[3], y4 = cy[3];
    const real z1 = cz[0], z2 = cz[1], z3 = cz[2], z4 = cz[3];
#endif
    FEMX_PROLOGUE
    FEMX_COO_ROWS
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const i64 s0 = e0 * N2;  // first slot of the tile
  if (ne_t == FEMX_COO_TILE) {
    // full tile: sizes and addresses are multiples of 16 bytes
    if (t == 0) {
      if (A)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(A + s0), "r"((unsigned)__cvta_generic_to_shared(sA)), "r"((unsigned)(FEMX_COO_TILE * N2 * sizeof(real))) : "memory");
      if (rowA)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(rowA + s0), "r"((unsigned)__cvta_generic_to_shared(sR)), "r"((unsigned)(FEMX_COO_TILE * N2 * 4)) : "memory");
      if (colA)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(colA + s0), "r"((unsigned)__cvta_generic_to_shared(sC)), "r"((unsigned)(FEMX_COO_TILE * N2 * 4)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    for (int j = t; j < ne_t * N2; j += FEMX_COO_TILE) {
      if (A) A[s0 + j] = sA[j];
      if (rowA) rowA[s0 + j] = sR[j];
      if (colA) colA[s0 + j] = sC[j];
    }
  }
}
)FEMX";

// ---- kernel ABI #2: deterministic numeric pass into CSR --------------------
// A CTA owns FEMX_TILE_NODES consecutive node rows (x ND dof rows).  Thread
// (node, c) walks the node's incident elements in ascending element order,
// evaluates matrix row li = ND*local_node + c of each, and adds the entries
// into ITS OWN segment of a shared-memory image of the tile's CSR values; the
// image is then written out with fully coalesced stores.  Every CSR value is
// produced by exactly one thread in a fixed order: no atomics, bitwise
// reproducible (replaces the linear search + global atomicAdd of
// fea_symbolic_nvrtc_sparse2.cpp:533-544).
//
// Scatter map (built once by the symbolic pass): for incidence `it` of a row,
// code = positions (7 bits each) of the element's NN nodes inside the row's
// sorted column list | li << 28.  Codes are stored SELL-32: slice s = rows
// [32s, 32s+32), entry (row, it) at slice_ptr[s] + 32*it + row%32, so a warp
// reads one contiguous 128-byte line per incidence.  Node ids come from the
// tile's column list staged in shared memory, so connectivity is not re-read.
static const char* const kFemxJitCsrRow = R"FEMX(
#ifndef FEMX_HOST_EMU
// predicated read-only global load: keeps the gathers of the software pipeline
// branch-free so that they are issued before the current incidence is evaluated
// (when the predicate is off the result is never used)
__device__ __forceinline__ double femx_ldg_if(const double* p, int pred) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.f64 %0, [%1];\n\t}"
               : "=d"(v) : "l"(p), "r"(pred));
  return v;
}
__device__ __forceinline__ float femx_ldg_if(const float* p, int pred) {
  float v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.f32 %0, [%1];\n\t}"
               : "=f"(v) : "l"(p), "r"(pred));
  return v;
}
// read-only global load that stays where it is written: the speculative gathers at kernel entry must
// not be sunk below the branch on the row's metadata (the compiler does that to a plain __ldg: two
// serialised memory round trips instead of one)
__device__ __forceinline__ double femx_ldg_pinned(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float femx_ldg_pinned(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int2 femx_ldg_pinned(const int2* p) {
  int2 v;
  asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int femx_ldg_pinned(const int* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
#endif
#if FEMX_UNIT_STRIDE
#define FEMX_CS 1
#else
#define FEMX_CS cs
#endif
#if FEMX_MIN_BLOCKS > 0
#define FEMX_BOUNDS __launch_bounds__(FEMX_TILE_NODES, FEMX_MIN_BLOCKS)
#else
#define FEMX_BOUNDS __launch_bounds__(FEMX_TILE_NODES)
#endif
#define FEMX_EPV (16 / (int)sizeof(real))  // values per 16 bytes
// header of the stencil-class kernel's shared memory: warp masks (128 B) + per-row value ends
#define FEMX_SPEC_HDR (128 + ((FEMX_TILE_NODES * 4 + 127) / 128) * 128)

// Scatter code of one incidence (row node = local node li of element e):
//   bits  0-6, 7-13, 14-20 : positions in the row's column list of the OTHER vertices, in the order
//                            FEMX_OTH(li, j): cyclic (li+1+j)%3 for triangles, li ^ (j+1) for tetrahedra
//                            (always an even permutation of the element: same signed Jacobian)
//   bits 21-23             : first-touch flags of those positions (this incidence is the first of
//                            the row to contribute there: store, do not add — the value image
//                            needs no zero fill)
//   bits 28-29             : li
// rowinfo[i] = { row_ptr[i], #incidences | position of the row's own node << 24 }.
//
// Data movement of one tile (FEMX_TILE_NODES node rows):
//   in : scatter codes + column list  -- two TMA bulk copies (cp.async.bulk, one thread,
//        mbarrier completion), no per-thread load/store instructions
//   out: the tile's CSR values        -- one TMA bulk store from the shared-memory image
//   gathered node coordinates come through L1 (read-only path), software-pipelined.
#define FEMX_FIRST(J) ((code >> (21 + (J))) & 1u)

// ---- one row by the generic incidence loop -------------------------------------------------
//   sc    the row's scatter codes, consecutive incidences 32 words apart (SELL-32)
//   scol  the row's sorted column list (LOCAL node ids)
//   pelem the row's e*NN+li list (SELL-32; element-expanded coordinates only)
//   srow  the row's value segment (private to the thread), rstride = row length * ND
// Tile mode passes shared-memory copies of codes and columns (staged by TMA), row-list mode reads
// them straight from global memory; the arithmetic — and therefore every bit of the result — is the same.
#ifndef FEMX_HOST_EMU
__device__ __forceinline__ void femx_generic_row(const int2 r0, const int np, const int rstride, real* srow,
                                                 const unsigned* sc, const int* scol, const int* pelem,
                                                 const real* __restrict__ X, const real* __restrict__ Y,
                                                 const real* __restrict__ Z, const i64 cs) {
  unsigned code = sc[0];
  const int ps = (int)((unsigned)r0.y >> 24) * ND;  // own column: the same for every incidence
  real ox[NN - 1], oy[NN - 1], oz[NN - 1];
#if FEMX_EXPANDED
  // element-expanded coordinates X[NN*e + a] (the reference's layout, SURVEY Q17)
  real sx, sy, sz = real(0);
  {
    const int ea = __ldg(pelem);  // e*NN + li
    const int li = (code >> 28) & 3;
    const int e0 = ea - li;
    sx = __ldg(X + ea); sy = __ldg(Y + ea);
    if (DIM == 3) sz = __ldg(Z + ea);
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) {
      const int b = FEMX_OTH(li, j);
      ox[j] = __ldg(X + e0 + b); oy[j] = __ldg(Y + e0 + b);
      oz[j] = DIM == 3 ? __ldg(Z + e0 + b) : real(0);
    }
  }
#else
  // the row's own node is a vertex of every incident element: it stays in registers
  const i64 pself = (i64)scol[(unsigned)r0.y >> 24] * FEMX_CS;
  const real sx = __ldg(X + pself), sy = __ldg(Y + pself), sz = DIM == 3 ? __ldg(Z + pself) : real(0);
#pragma unroll
  for (int j = 0; j < NN - 1; ++j) {
    const i64 p = (i64)scol[(code >> (7 * j)) & 127] * FEMX_CS;
    ox[j] = __ldg(X + p); oy[j] = __ldg(Y + p);
    oz[j] = DIM == 3 ? __ldg(Z + p) : real(0);
  }
#endif
  real dacc[ND * ND];  // the diagonal block (own column) accumulates in registers
#pragma unroll
  for (int d = 0; d < ND * ND; ++d) dacc[d] = real(0);
#if FEMX_MIDGATHER && !FEMX_EXPANDED
  // One coordinate buffer: the gathers of incidence it+1 are issued in the MIDDLE of incidence
  // it, right after its geometry prologue has consumed the coordinates (saves the second
  // buffer's registers; the loads fly during the entry evaluation and the scatter).
#pragma unroll 1
  for (int it = 0; it < np; ++it) {
    const int more = it + 1 < np;
    sc += more ? 32 : 0;
    const unsigned ncd = *sc;
    int nidx[NN - 1];
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) nidx[j] = scol[(ncd >> (7 * j)) & 127];
    int po[NN - 1];
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) po[j] = ((code >> (7 * j)) & 127) * ND;
#define FEMX_GATHER_NEXT                                                        \
    _Pragma("unroll") for (int j = 0; j < NN - 1; ++j) {                    \
      const i64 p_ = (i64)nidx[j] * FEMX_CS;                                \
      ox[j] = femx_ldg_if(X + p_, more); oy[j] = femx_ldg_if(Y + p_, more); \
      if (DIM == 3) oz[j] = femx_ldg_if(Z + p_, more);                      \
    }
    switch (FEMX_ROTINV ? 0 : (int)((code >> 28) & 3)) {
      FEMX_CSR_CASES
    }
#undef FEMX_GATHER_NEXT
    code = ncd;
  }
#else
#define FEMX_GATHER_NEXT
#pragma unroll FEMX_UNROLL
  for (int it = 0; it < np; ++it) {
    // ---- software pipeline: the gathers of incidence it+1 are issued first
    const int more = it + 1 < np;
    sc += more ? 32 : 0;
    const unsigned ncd = *sc;
    real nox[NN - 1], noy[NN - 1], noz[NN - 1];
#if FEMX_EXPANDED
    pelem += more ? 32 : 0;
    real nsx, nsy, nsz = real(0);
    {
      const int ea = __ldg(pelem);
      const int nli = (ncd >> 28) & 3;
      const int e0 = ea - nli;
      nsx = femx_ldg_if(X + ea, more); nsy = femx_ldg_if(Y + ea, more);
      if (DIM == 3) nsz = femx_ldg_if(Z + ea, more);
#pragma unroll
      for (int j = 0; j < NN - 1; ++j) {
        const int b = FEMX_OTH(nli, j);
        nox[j] = femx_ldg_if(X + e0 + b, more); noy[j] = femx_ldg_if(Y + e0 + b, more);
        noz[j] = DIM == 3 ? femx_ldg_if(Z + e0 + b, more) : real(0);
      }
    }
#else
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) {
      const i64 p = (i64)scol[(ncd >> (7 * j)) & 127] * FEMX_CS;
      nox[j] = femx_ldg_if(X + p, more); noy[j] = femx_ldg_if(Y + p, more);
      noz[j] = DIM == 3 ? femx_ldg_if(Z + p, more) : real(0);
    }
#endif
    // ---- evaluate incidence it: row li*ND + c of the element matrix
    int po[NN - 1];
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) po[j] = ((code >> (7 * j)) & 127) * ND;
    switch (FEMX_ROTINV ? 0 : (int)((code >> 28) & 3)) {
      FEMX_CSR_CASES
    }
    code = ncd;
#pragma unroll
    for (int j = 0; j < NN - 1; ++j) { ox[j] = nox[j]; oy[j] = noy[j]; oz[j] = noz[j]; }
#if FEMX_EXPANDED
    sx = nsx; sy = nsy; sz = nsz;
#endif
  }
#undef FEMX_GATHER_NEXT
#endif
#if FEMX_ROWSUM
  {  // stiffness rows sum to zero: diagonal = cj * (sum of the incident Jacobians) - (sum of the off-diagonal values)
    real S_ = real(0);
    const int rl_ = rstride / ND, sp_ = ps / ND;
    for (int k = 0; k < rl_; ++k)
      if (k != sp_) S_ += srow[k];
    dacc[0] = fma(FEMX_CJ, dacc[0], -S_);
  }
#endif
#pragma unroll
  for (int c = 0; c < ND; ++c)
#pragma unroll
    for (int d = 0; d < ND; ++d) srow[c * rstride + ps + d] = dacc[c * ND + d];
}

#endif  // !FEMX_HOST_EMU
)FEMX";

// the tile / stencil-class kernel (appended to kFemxJitCsrRow)
static const char* const kFemxJitCsr = R"FEMX(
extern "C" __global__ void FEMX_BOUNDS
femx_csr(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr,
         const int* __restrict__ col_loc, const unsigned* __restrict__ sell_code,
         const int* __restrict__ sell_elem, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
         real* __restrict__ vals, const int n_rows, const int row_node0, const int node_max,
         const femx_soff soff, const int* __restrict__ rowlist, const int n_list, const int seg) {
  extern __shared__ __align__(128) unsigned char femx_smem[];
#if FEMX_SPEC
  // ---- pattern with a stencil class.  The first CTAs take the rows OUTSIDE the class (the boundary
  // rows of a structured mesh; compacted list, one thread each, generic incidence loop, codes and
  // columns read straight from global memory, no barriers); the others take tiles of class rows.
  // The two kinds write disjoint parts of `vals`.
  const int n_lb = (n_list + FEMX_TILE_NODES - 1) / FEMX_TILE_NODES;
#if FEMX_LISTLAST
  const int n_tb = (n_rows + FEMX_TILE_NODES - 1) / FEMX_TILE_NODES;
  const int lb = (int)blockIdx.x - n_tb, tb = blockIdx.x;
#else
  const int lb = (int)blockIdx.x < n_lb ? (int)blockIdx.x : -1, tb = (int)blockIdx.x - n_lb;
#endif
  if (lb >= 0) {
    const int k_ = lb * FEMX_TILE_NODES + threadIdx.x;
    if (k_ >= n_list) return;
    const int row = __ldg(rowlist + k_);
    const int2 r0 = __ldg(&rowinfo[row]);
    const int rlen = __ldg(&rowinfo[row + 1].x) - r0.x;
    const int np = r0.y & FEMX_NP_MASK;
    real* srow = reinterpret_cast<real*>(femx_smem) + (size_t)threadIdx.x * seg;
    if (np > 0) {
      const int sp = __ldg(slice_ptr + (row >> 5)) + (row & 31);
      femx_generic_row(r0, np, rlen * ND, srow, sell_code + sp, col_loc + r0.x, sell_elem + sp, X, Y, Z, cs);
      real* dst = vals + (i64)r0.x * (ND * ND);
      for (int j = 0; j < rlen * (ND * ND); ++j) dst[j] = srow[j];
    }
    return;
  }
  const int i0 = tb * FEMX_TILE_NODES;
#else
  const int i0 = blockIdx.x * FEMX_TILE_NODES;
#endif
  const int nt = min(FEMX_TILE_NODES, n_rows - i0);
#if FEMX_SPEC
  // Speculative gathers, issued before any metadata has arrived: a row of the stencil class finds its
  // columns at (own node + constant offset), so their coordinates need neither the row pointers nor
  // the column list.  (Clamped: for a row outside the class the values are simply not used.)
  const int node_ = row_node0 + i0 + min((int)threadIdx.x, nt - 1);
  // (the row metadata is pinned too: left to the scheduler, its loads end up behind the first use of the
  //  coordinates — a second serialised round trip)
  const int ln = threadIdx.x;  // one thread per node row (all ND dof rows of the node)
  const int rowc = i0 + min(ln, nt - 1);
  const int2 r0 = femx_ldg_pinned(&rowinfo[rowc]);
  const int rnext = femx_ldg_pinned(&rowinfo[rowc + 1].x);
  const int base = femx_ldg_pinned(&rowinfo[i0].x);
  FEMX_SPEC_LOAD
#else
  const int base = __ldg(&rowinfo[i0].x);
  const int cntn = __ldg(&rowinfo[i0 + nt].x) - base;  // node-level nonzeros of the tile
  const int cnt = cntn * (ND * ND);
  const int ln = threadIdx.x;  // one thread per node row (all ND dof rows of the node)
  const int rowc = i0 + min(ln, nt - 1);
  const int2 r0 = __ldg(&rowinfo[rowc]);
  const int rnext = __ldg(&rowinfo[rowc + 1].x);
#endif
  const i64 vb = (i64)base * (ND * ND);              // first value index of the tile
  const int vph = (int)(vb & (FEMX_EPV - 1));         // phase of the value run (elements)
#if FEMX_SPEC
  // ---- tile of class rows: the straight-line body generated for the class's scatter codes; nothing
  // is staged in (no codes, no column list, no mbarrier).  smem: [class mask of each warp (128 B) | end
  // of each row's values | values], the values placed with the 16-byte phase of their global address
  // (aligned bulk stores).
  unsigned* s_mask = reinterpret_cast<unsigned*>(femx_smem);
  int* s_end = reinterpret_cast<int*>(femx_smem + 128);   // s_end[i]: end of row i's values, relative to the tile
  real* s_vals = reinterpret_cast<real*>(femx_smem + FEMX_SPEC_HDR) + vph;
  // The body runs for EVERY thread, before the row's metadata is looked at: a branch on `mine` ahead of it
  // would let the compiler sink the entry gathers below that branch (two serialised memory round trips).
  // Rows outside the class compute on clamped, meaningless coordinates and store nothing.
  const bool mine = ln < nt && (r0.y & FEMX_ROW_SPEC);
  {
    real* srow = s_vals + (r0.x - base) * (ND * ND);
    FEMX_SPEC_BODY
  }
  {
    const unsigned wm = __ballot_sync(0xffffffffu, mine);
    if ((ln & 31) == 0) s_mask[ln >> 5] = wm;
    if (ln < nt) s_end[ln] = (rnext - base) * (ND * ND);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // ---- write the tile: one bulk store per maximal run of class rows, issued by the run's first row
  // (the rows in between belong to the row-list CTAs and must not be touched)
  if (mine && (ln == 0 || !((s_mask[(ln - 1) >> 5] >> ((ln - 1) & 31)) & 1u))) {
    int e = ln + 1;  // first row after the run
    while (e < nt) {
      const unsigned z = (~s_mask[e >> 5]) >> (e & 31);
      if (z) { e += __ffs(z) - 1; break; }
      e = (e | 31) + 1;
    }
    e = min(e, nt);
    const int a0 = (r0.x - base) * (ND * ND);
    const int n = s_end[e - 1] - a0;
    real* dst = vals + vb + a0;
    const real* src = s_vals + a0;
    const int head = min(n, (FEMX_EPV - (int)((vb + a0) & (FEMX_EPV - 1))) & (FEMX_EPV - 1));
    const int mid = (n - head) & ~(FEMX_EPV - 1);
    if (mid > 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(dst + head), "r"((unsigned)__cvta_generic_to_shared(src + head)), "r"((unsigned)(mid * sizeof(real))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    for (int j = 0; j < head; ++j) dst[j] = src[j];                    // ragged ends: < 16 bytes each
    for (int j = head + mid; j < n; ++j) dst[j] = src[j];
    if (mid > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
#else
  const int sbase = __ldg(slice_ptr + (i0 >> 5));      // the tile's slices are contiguous
  const int ncode = __ldg(slice_ptr + ((i0 + nt + 31) >> 5)) - sbase;
  // smem: [mbarrier | codes | values | columns].  Values and columns are placed with the
  // same 16-byte phase as their global addresses so that bulk copies are aligned.
  const int cph = base & 3;                           // phase of the column run (ints)
  unsigned* s_code = reinterpret_cast<unsigned*>(femx_smem + 128);
  real* s_vals = reinterpret_cast<real*>(s_code + ncode) + vph;
  const int vspan = (vph + cnt + FEMX_EPV - 1) & ~(FEMX_EPV - 1);
  int* s_cols = reinterpret_cast<int*>(s_vals - vph + vspan) + cph;
  // per-row metadata (above) is issued before the staging wait so that its latency overlaps the bulk copies
  const int spg = __ldg(slice_ptr + (rowc >> 5));
  const unsigned bar = (unsigned)__cvta_generic_to_shared(femx_smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#if FEMX_EXPANDED
    const unsigned cbytes = 0;
#else
    const unsigned cbytes = (unsigned)(((cph + cntn + 3) & ~3) * 4);
#endif
    const unsigned kbytes = (unsigned)ncode * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kbytes + cbytes) : "memory");
    if (kbytes)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((unsigned)__cvta_generic_to_shared(s_code)), "l"(sell_code + sbase), "r"(kbytes), "r"(bar) : "memory");
    if (cbytes)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((unsigned)__cvta_generic_to_shared(s_cols - cph)), "l"(col_loc + (base - cph)), "r"(cbytes), "r"(bar) : "memory");
  }
  __syncthreads();  // barrier initialised
  {
    unsigned done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
    } while (!done);
  }
  if (ln < nt) {
    const int row = i0 + ln;
    const int rlen = rnext - r0.x;
    const int off = r0.x - base;
    const int sp = spg + (row & 31);
    const int np = r0.y & FEMX_NP_MASK;
    if (np > 0)
      femx_generic_row(r0, np, rlen * ND, s_vals + off * (ND * ND), s_code + (sp - sbase), s_cols + off,
                       sell_elem + sp, X, Y, Z, cs);
  }
  // ---- write the tile: generic-proxy writes -> async proxy, then one bulk store
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  {
    real* dst = vals + vb;
    const int head = min(cnt, (FEMX_EPV - vph) & (FEMX_EPV - 1));   // elements before the first 16-byte boundary
    const int mid = (cnt - head) & ~(FEMX_EPV - 1);                  // elements in whole 16-byte units
    if (threadIdx.x == 0 && mid > 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(dst + head), "r"((unsigned)__cvta_generic_to_shared(s_vals + head)), "r"((unsigned)(mid * sizeof(real))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // ragged head / tail (fewer than 16 bytes each)
    const int ntail = cnt - head - mid;
    if ((int)threadIdx.x < head) dst[threadIdx.x] = s_vals[threadIdx.x];
    else if ((int)threadIdx.x - head < ntail)   // head, ntail < 16 bytes each: threads 0..6 at most
      dst[mid + threadIdx.x] = s_vals[mid + threadIdx.x];
    if (threadIdx.x == 0 && mid > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
#endif
}
)FEMX";

static const char* const kFemxJitLattice = R"FEMX(
// ---- kernel ABI #2, lattice mode: every element evaluated ONCE per tile --------------------
// (appended to kFemxJitCsrRow; macros FEMX_LT_* generated by femx_lattice.cpp.)
// A CTA owns (FEMX_LT_TX-1) x (FEMX_LT_TY-1) node columns of a lattice mesh and marches through node planes
// [k0, k1).  Thread (ix, iy) holds the column with lower-corner node (i, j): per plane it evaluates the cell
// (i, j, kc) — its P elements, each ONCE — reduces them to one value per cell edge and one Jacobian sum per
// corner, adds what the cell below left for the shared plane (register carry), and publishes the sums other
// columns need ("fields") to shared memory.  The values of row (i, j, kc) are then gathered: each off-diagonal
// entry is the sum of the fields of the cells around its edge, the diagonal follows from the zero row sum of
// the stiffness part.  Rows go to a shared-memory image and leave through one bulk store per run of consecutive
// class rows.  Column ix = 0 / iy = 0 is the halo (cells only, no rows).
// The columns of a CTA are not kept in lock step: publishing plane kc ARRIVES on an mbarrier, the gather of that
// plane WAITS on it one cell later (after the arithmetic of cell kc+1) — by then the phase has normally completed,
// so warps drift up to one plane apart and one warp's gather overlaps another's arithmetic.  Field buffers are 2 /
// 3 planes deep accordingly (femx_lattice.cpp).  A line of the tile lies inside one warp (FEMX_LT_TX divides 32),
// so a run of rows is completed, fenced and stored warp-locally (__syncwarp, no CTA barrier).
// Rows outside the stencil class (the mesh boundary) are taken by femx_rowlist (generic incidence loop).
struct femx_lat {
  int cnx, cny, cnz;  // cells per axis
  int sy, sz;         // node strides (x stride 1)
  int node0;          // node id of lattice node (0,0,0)
  int klo, khi;       // node planes to produce: [klo, khi]
  int ntx, nty;       // tiles per plane
  int kc;             // node planes per CTA
};
#define LT_NT FEMX_TILE_NODES
#ifndef FEMX_HOST_EMU
#define FEMX_TID ((int)threadIdx.x)
#define FEMX_BID ((int)blockIdx.x)
#define FEMX_LT_KERNEL extern "C" __global__ void __launch_bounds__(LT_NT, FEMX_LT_MINB)
__device__ __forceinline__ void femx_lt_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void femx_lt_bulk_store(real* dst, const real* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void femx_lt_bulk_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void femx_lt_prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void femx_lt_bar_init(void* bar, int count) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}
__device__ __forceinline__ void femx_lt_bar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void femx_lt_bar_wait(void* bar, int parity) {
  unsigned done;
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void femx_lt_syncwarp() { __syncwarp(); }
__device__ __forceinline__ unsigned femx_lt_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
#endif

#define LT_LOAD_PLANE(K, A, B, C, D)                                                          \
  {                                                                                           \
    const i64 p_ = (i64)(nb + (K) * lat.sz) * FEMX_CS, q_ = p_ + (i64)lat.sy * FEMX_CS;       \
    cx##A = __ldg(X + p_); cx##B = __ldg(X + p_ + FEMX_CS);                                   \
    cx##C = __ldg(X + q_); cx##D = __ldg(X + q_ + FEMX_CS);                                   \
    cy##A = __ldg(Y + p_); cy##B = __ldg(Y + p_ + FEMX_CS);                                   \
    cy##C = __ldg(Y + q_); cy##D = __ldg(Y + q_ + FEMX_CS);                                   \
    cz##A = __ldg(Z + p_); cz##B = __ldg(Z + p_ + FEMX_CS);                                   \
    cz##C = __ldg(Z + q_); cz##D = __ldg(Z + q_ + FEMX_CS);                                   \
  }
#define LT_PREFETCH_PLANE(K)                                                                   \
  {                                                                                           \
    const i64 p_ = (i64)(nb + (K) * lat.sz) * FEMX_CS, q_ = p_ + (i64)lat.sy * FEMX_CS;       \
    femx_lt_prefetch(X + p_); femx_lt_prefetch(X + p_ + FEMX_CS); femx_lt_prefetch(X + q_); femx_lt_prefetch(X + q_ + FEMX_CS); \
    femx_lt_prefetch(Y + p_); femx_lt_prefetch(Y + p_ + FEMX_CS); femx_lt_prefetch(Y + q_); femx_lt_prefetch(Y + q_ + FEMX_CS); \
    femx_lt_prefetch(Z + p_); femx_lt_prefetch(Z + p_ + FEMX_CS); femx_lt_prefetch(Z + q_); femx_lt_prefetch(Z + q_ + FEMX_CS); \
  }

// Gather + store of one node plane: `mine` rows take their values from the field buffers (GS / GD of the plane, GP of
// the plane below), place them in the image with the 16-byte phase of their global address, and the first row of
// every run of consecutive class rows (found by ballot inside the line's warp) issues one bulk store for the run.
#define LT_GATHER_PLANE(MINE, R0, GS, GD, GP)                                                                     \
  {                                                                                                               \
    femx_lt_bulk_wait(); /* this warp's bulk stores of the previous plane have read the image */                  \
    femx_lt_syncwarp();                                                                                           \
    real* lt_row = lt_img + t * FEMX_LT_RLEN + (((R0).x - t * FEMX_LT_RLEN) & (FEMX_EPV - 1));                     \
    if (MINE) { FEMX_LT_GATHER(GS, GD, GP) }                                                                      \
    femx_lt_fence();                                                                                              \
    const unsigned lt_m = femx_lt_ballot(MINE);                                                                   \
    const unsigned lt_line = (lt_m >> ((t & 31) - ix)) & ((FEMX_LT_TX == 32) ? 0xffffffffu : ((1u << FEMX_LT_TX) - 1u)); \
    if ((MINE) && (ix == 0 || !((lt_line >> (ix - 1)) & 1u))) {                                                   \
      const unsigned z_ = ~(lt_line >> ix);                                                                       \
      const int len = z_ ? __ffs(z_) - 1 : 32 - ix;                                                               \
      const int n_ = len * FEMX_LT_RLEN;                                                                          \
      real* dst = vals + (i64)(R0).x;                                                                             \
      const real* src = lt_row;                                                                                   \
      const int head = min(n_, (FEMX_EPV - (int)((R0).x & (FEMX_EPV - 1))) & (FEMX_EPV - 1));                     \
      const int mid = (n_ - head) & ~(FEMX_EPV - 1);                                                              \
      if (mid > 0) femx_lt_bulk_store(dst + head, src + head, (unsigned)(mid * sizeof(real)));                    \
      for (int q = 0; q < head; ++q) dst[q] = src[q]; /* ragged ends: < 16 bytes each */                          \
      for (int q = head + mid; q < n_; ++q) dst[q] = src[q];                                                      \
    }                                                                                                             \
  }

#ifndef FEMX_HOST_EMU
// ---- the rows OUTSIDE the class (the mesh boundary): compacted list, one thread each, generic incidence loop.
// A kernel of its own (launched behind the lattice pass on the same stream): it needs 40 registers, not the 168 of the
// lattice pass — as CTAs of that kernel these rows held a sixth of its SM slots while stalling on dependent gathers.
extern "C" __global__ void __launch_bounds__(128)
femx_rowlist(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr,
             const int* __restrict__ col_loc, const unsigned* __restrict__ sell_code,
             const int* __restrict__ sell_elem, const real* __restrict__ X,
             const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
             real* __restrict__ vals, const int* __restrict__ rowlist, const int n_list, const int seg) {
  extern __shared__ __align__(128) unsigned char femx_smem[];
  const int k_ = blockIdx.x * 128 + threadIdx.x;
  if (k_ >= n_list) return;
  const int row = __ldg(rowlist + k_);
  const int2 r0 = __ldg(&rowinfo[row]);
  const int rlen = __ldg(&rowinfo[row + 1].x) - r0.x;
  const int np = r0.y & FEMX_NP_MASK;
  real* srow = reinterpret_cast<real*>(femx_smem) + (size_t)threadIdx.x * seg;
  if (np > 0) {
    const int sp = __ldg(slice_ptr + (row >> 5)) + (row & 31);
    femx_generic_row(r0, np, rlen * ND, srow, sell_code + sp, col_loc + r0.x, sell_elem + sp, X, Y, Z, cs);
    real* dst = vals + (i64)r0.x * (ND * ND);
    for (int j = 0; j < rlen * (ND * ND); ++j) dst[j] = srow[j];
  }
}
#endif

FEMX_LT_KERNEL
femx_csr(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr,
         const int* __restrict__ col_loc, const unsigned* __restrict__ sell_code,
         const int* __restrict__ sell_elem, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
         real* __restrict__ vals, const int n_rows, const int row_node0, const femx_lat lat) {
#ifndef FEMX_HOST_EMU
  extern __shared__ __align__(128) unsigned char femx_smem[];
#else
  unsigned char* femx_smem = femx_emu_smem();
#endif
  const int b = FEMX_BID;
  const int t = FEMX_TID;
  const int tx = b % lat.ntx, ty = (b / lat.ntx) % lat.nty, tz = b / (lat.ntx * lat.nty);
  const int ix = t % FEMX_LT_TX, iy = t / FEMX_LT_TX;
  const int iu = tx * (FEMX_LT_TX - 1) + ix, ju = ty * (FEMX_LT_TY - 1) + iy;  // the column: lower-corner node (iu, ju)
  // columns beyond the lattice (and the padding threads of the last warp) evaluate a clamped duplicate, unused
  const int i = min(iu, lat.cnx - 1), j = min(ju, lat.cny - 1);
  const bool own_col = ix >= 1 && iy >= 1 && iy < FEMX_LT_TY && iu < lat.cnx && ju < lat.cny;
  const int k0 = lat.klo + tz * lat.kc, k1 = min(k0 + lat.kc, lat.khi + 1);
  void* lt_bar = femx_smem;                                                                   // mbarrier (128-byte header)
  real* lt_FS = reinterpret_cast<real*>(femx_smem + 128) + t;                                 // 2-deep fields [f][2][LT_NT]
  real* lt_FD = lt_FS + 2 * FEMX_LT_NS * LT_NT;                                               // 3-deep fields [f][3][LT_NT]
  real* lt_img = reinterpret_cast<real*>(femx_smem + 128) + FEMX_LT_NSLOT * LT_NT;            // value image [LT_NT][RLEN] (+ phase slack)
  femx_lt_bar_init(lt_bar, LT_NT);
  const int nb = lat.node0 + i + j * lat.sy;  // node id of the column at plane 0
  // corner c = dx | dy << 1 | dz << 2 of the current cell: c0..c3 on node plane kc, c4..c7 on plane kc + 1
  real cx0, cx1, cx2, cx3, cx4, cx5, cx6, cx7, cy0, cy1, cy2, cy3, cy4, cy5, cy6, cy7, cz0, cz1, cz2, cz3, cz4, cz5, cz6, cz7;
  LT_LOAD_PLANE(k0 - 1, 0, 1, 2, 3)
#if FEMX_LT_PF
  LT_PREFETCH_PLANE(k0)
#else
  LT_LOAD_PLANE(k0, 4, 5, 6, 7)
#endif
  FEMX_LT_CARRY_DECL
  int2 r0p = make_int2(0, 0);   // metadata of the row one plane below (gathered one cell late)
  bool minep = false;
  int b2 = 0, b3 = 0;           // buffer of plane kc: (kc - k0 + 1) & 1 and % 3
  FEMX_LT_LOOP_PRAGMA
  for (int kc = k0 - 1; kc < k1; ++kc) {
#if FEMX_LT_PF
    // the top plane was prefetched into L1 one cell ago; the plane after it is requested now
    LT_LOAD_PLANE(kc + 1, 4, 5, 6, 7)
    if (kc + 1 < k1) LT_PREFETCH_PLANE(kc + 2)
#endif
    // metadata of row (i, j, kc): needed when the plane is gathered, issued before the cell
    const int row = nb + kc * lat.sz - row_node0;
    const bool rowok = own_col && kc >= k0 && row >= 0 && row < n_rows;
    int2 r0 = make_int2(0, 0);
    if (rowok) r0 = __ldg(&rowinfo[row]);
    FEMX_LT_EDGES
    // roll the planes; the loads of the next top plane fly during the cell's arithmetic
    cx0 = cx4; cx1 = cx5; cx2 = cx6; cx3 = cx7; cy0 = cy4; cy1 = cy5; cy2 = cy6; cy3 = cy7;
    cz0 = cz4; cz1 = cz5; cz2 = cz6; cz3 = cz7;
#if !FEMX_LT_PF
    if (kc + 1 < k1) LT_LOAD_PLANE(kc + 2, 4, 5, 6, 7)
#endif
    FEMX_LT_CELL
    FEMX_LT_FIELDS
    if (kc >= k0) {
      // every column has published plane kc - 1 (arrived one cell ago): gather it while slower warps still compute
      femx_lt_bar_wait(lt_bar, (kc - k0) & 1);
      if (kc > k0) {
        const int p3 = b3 == 0 ? 2 : b3 - 1, q3 = p3 == 0 ? 2 : p3 - 1;   // 3-deep buffers of planes kc-1, kc-2
        LT_GATHER_PLANE(minep, r0p, lt_FS + (b2 ^ 1) * LT_NT, lt_FD + p3 * LT_NT, lt_FD + q3 * LT_NT)
      }
      FEMX_LT_PUBLISH(lt_FS + b2 * LT_NT, lt_FD + b3 * LT_NT)
    } else {
      FEMX_LT_PUBLISH_UP(lt_FD + b3 * LT_NT)
    }
    femx_lt_bar_arrive(lt_bar);
    r0p = r0;
    minep = rowok && (r0.y & FEMX_ROW_SPEC);
    b2 ^= 1;
    b3 = b3 == 2 ? 0 : b3 + 1;
  }
  {  // the last plane
    femx_lt_bar_wait(lt_bar, (k1 - k0) & 1);
    const int p3 = b3 == 0 ? 2 : b3 - 1, q3 = p3 == 0 ? 2 : p3 - 1;
    LT_GATHER_PLANE(minep, r0p, lt_FS + (b2 ^ 1) * LT_NT, lt_FD + p3 * LT_NT, lt_FD + q3 * LT_NT)
  }
  femx_lt_bulk_wait();
}
)FEMX";

// ---- load vector: one thread per node row, ND components --------------------------------
// Same incidence walk as femx_csr (SELL-32 codes, ascending element order) but the result is
// one value per dof, so there is no shared-memory image: codes and columns are read straight
// from global memory (coalesced / L1).
static const char* const kFemxJitRhs = R"FEMX(
extern "C" __global__ void __launch_bounds__(128)
femx_rhs(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr,
         const int* __restrict__ col_loc, const unsigned* __restrict__ sell_code,
         const int* __restrict__ sell_elem, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs, const int expanded,
         real* __restrict__ rhs, const int n_rows) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int2 r0 = __ldg(&rowinfo[row]);
  const int* cols = col_loc + r0.x;
  const int sp = __ldg(slice_ptr + (row >> 5)) + (row & 31);
  real racc[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) racc[d] = real(0);
  const int np = r0.y & FEMX_NP_MASK;
  const int self_pos = (int)((unsigned)r0.y >> 24);
  for (int it = 0; it < np; ++it) {
    const unsigned code = __ldg(sell_code + sp + it * 32);
    const int li = (code >> 28) & 3;
    real SX, SY, SZ = real(0), OX[NN - 1], OY[NN - 1], OZ[NN - 1];
    if (expanded) {
      const int ea = __ldg(sell_elem + sp + it * 32);
      const int e0 = ea - li;
      SX = __ldg(X + ea); SY = __ldg(Y + ea);
      if (DIM == 3) SZ = __ldg(Z + ea);
#pragma unroll
      for (int j = 0; j < NN - 1; ++j) {
        const int b = FEMX_OTH(li, j);
        OX[j] = __ldg(X + e0 + b); OY[j] = __ldg(Y + e0 + b);
        OZ[j] = DIM == 3 ? __ldg(Z + e0 + b) : real(0);
      }
    } else {
      const i64 ps = (i64)__ldg(cols + self_pos) * cs;
      SX = __ldg(X + ps); SY = __ldg(Y + ps);
      if (DIM == 3) SZ = __ldg(Z + ps);
#pragma unroll
      for (int j = 0; j < NN - 1; ++j) {
        const i64 p = (i64)__ldg(cols + ((code >> (7 * j)) & 127)) * cs;
        OX[j] = __ldg(X + p); OY[j] = __ldg(Y + p);
        OZ[j] = DIM == 3 ? __ldg(Z + p) : real(0);
      }
    }
    switch (FEMX_ROTINV ? 0 : li) {
      FEMX_RHS_CASES
    }
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) rhs[(i64)row * ND + d] = racc[d];
}
)FEMX";
