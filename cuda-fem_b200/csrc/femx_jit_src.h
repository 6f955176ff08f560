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

#define NDOF (NN * ND)

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
// Scatter map: pair_code[k] holds, for incidence k of the row, the positions
// (7 bits each) of the element's NN nodes inside the row's sorted column list
// and the local index li (bits 28-29).  The node ids themselves come from the
// column list, so connectivity is not re-read.
static const char* const kFemxJitCsr = R"FEMX(
extern "C" __global__ void __launch_bounds__(FEMX_TILE_NODES * ND)
femx_csr(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx,
         const unsigned* __restrict__ pair_code, const int* __restrict__ pair_elem,
         const real* __restrict__ X, const real* __restrict__ Y,
         const real* __restrict__ Z, const i64 cs, const int expanded,
         real* __restrict__ vals, const int n_rows, const int col_base) {
  extern __shared__ __align__(16) unsigned char femx_smem[];
  real* s_vals = reinterpret_cast<real*>(femx_smem);
  const int i0 = blockIdx.x * FEMX_TILE_NODES;
  const int nt = min(FEMX_TILE_NODES, n_rows - i0);
  const int base = __ldg(&rowinfo[i0].x);
  const int cnt = (__ldg(&rowinfo[i0 + nt].x) - base) * (ND * ND);
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) s_vals[j] = real(0);
  __syncthreads();
  const int ln = threadIdx.x / ND;
  const int c = threadIdx.x - ln * ND;
  if (ln < nt) {
    const int2 r0 = __ldg(&rowinfo[i0 + ln]);
    const int2 r1 = __ldg(&rowinfo[i0 + ln + 1]);
    const int rlen = r1.x - r0.x;
    real* srow = s_vals + (r0.x - base) * (ND * ND) + c * rlen * ND;
    const int* cols = col_idx + r0.x;
    for (int k = r0.y; k < r1.y; ++k) {
      const unsigned code = __ldg(pair_code + k);
      int pos[NN];
#pragma unroll
      for (int a = 0; a < NN; ++a) pos[a] = (code >> (7 * a)) & 127;
      const int li = (code >> 28) & 3;
      real cx[NN], cy[NN], cz[NN];
      if (expanded) {
        const i64 e = __ldg(pair_elem + k) / NN;
#pragma unroll
        for (int a = 0; a < NN; ++a) {
          cx[a] = __ldg(X + e * NN + a);
          cy[a] = __ldg(Y + e * NN + a);
#if DIM == 3
          cz[a] = __ldg(Z + e * NN + a);
#else
          cz[a] = real(0);
#endif
        }
      } else {
#pragma unroll
        for (int a = 0; a < NN; ++a) {
          const i64 p = (i64)(__ldg(cols + pos[a]) - col_base) * cs;
          cx[a] = __ldg(X + p);
          cy[a] = __ldg(Y + p);
#if DIM == 3
          cz[a] = __ldg(Z + p);
#else
          cz[a] = real(0);
#endif
        }
      }
      real out[NDOF];
      femx_row(li * ND + c, cx, cy, cz, out);
#pragma unroll
      for (int a = 0; a < NN; ++a)
#pragma unroll
        for (int d = 0; d < ND; ++d) srow[pos[a] * ND + d] += out[a * ND + d];
    }
  }
  __syncthreads();
  real* dst = vals + (i64)base * (ND * ND);
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) dst[j] = s_vals[j];
}
)FEMX";
