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
// Scatter map (built once by the symbolic pass): for incidence `it` of a row,
// code = positions (7 bits each) of the element's NN nodes inside the row's
// sorted column list | li << 28.  Codes are stored SELL-32: slice s = rows
// [32s, 32s+32), entry (row, it) at slice_ptr[s] + 32*it + row%32, so a warp
// reads one contiguous 128-byte line per incidence.  Node ids come from the
// tile's column list staged in shared memory, so connectivity is not re-read.
static const char* const kFemxJitCsr = R"FEMX(
// predicated read-only global load: keeps the gathers of the software pipeline
// branch-free so that they are issued before the current incidence is evaluated
__device__ __forceinline__ double femx_ldg_if(const double* p, int pred) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.nc.f64 %0, [%1];\n\t}"
               : "=d"(v) : "l"(p), "r"(pred));
  return v;
}
__device__ __forceinline__ float femx_ldg_if(const float* p, int pred) {
  float v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.global.nc.f32 %0, [%1];\n\t}"
               : "=f"(v) : "l"(p), "r"(pred));
  return v;
}
#if FEMX_UNIT_STRIDE
#define FEMX_CS 1
#else
#define FEMX_CS cs
#endif

// Scatter code of one incidence (row node = local node li of element e):
//   bits  0-6, 7-13, 14-20 : positions in the row's column list of the OTHER vertices,
//                            in cyclic order (li+1)%NN, (li+2)%NN, ...
//   bits 21-27             : position of the row's own node (the diagonal)
//   bits 28-29             : li
extern "C" __global__ void __launch_bounds__(FEMX_TILE_NODES * ND)
femx_csr(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr,
         const int* __restrict__ col_loc, const unsigned* __restrict__ sell_code,
         const int* __restrict__ sell_elem, const real* __restrict__ X,
         const real* __restrict__ Y, const real* __restrict__ Z, const i64 cs,
         real* __restrict__ vals, const int n_rows) {
  extern __shared__ __align__(128) unsigned char femx_smem[];
  const int i0 = blockIdx.x * FEMX_TILE_NODES;
  const int nt = min(FEMX_TILE_NODES, n_rows - i0);
  const int base = __ldg(&rowinfo[i0].x);
  const int cntn = __ldg(&rowinfo[i0 + nt].x) - base;  // node-level nonzeros of the tile
  const int cnt = cntn * (ND * ND);
  const int sbase = __ldg(slice_ptr + (i0 >> 5));      // the tile's slices are contiguous
  const int ncode = __ldg(slice_ptr + ((i0 + nt + 31) >> 5)) - sbase;
  // smem: [codes | values | columns]; codes first so that 16-byte async copies are aligned
  unsigned* s_code = reinterpret_cast<unsigned*>(femx_smem);
  real* s_vals = reinterpret_cast<real*>(s_code + ncode);
  int* s_cols = reinterpret_cast<int*>(s_vals + cnt);
  // Stage the tile's streamed inputs (scatter codes, column list) with asynchronous
  // global->shared copies: every copy of the tile is in flight at once, no register
  // staging, one memory latency per tile instead of one per loop trip.
  {
    const unsigned sdst = (unsigned)__cvta_generic_to_shared(s_code);
    const unsigned* gsrc = sell_code + sbase;  // 128-byte aligned (slices are 32-entry multiples)
    for (int j = threadIdx.x * 4; j < ncode; j += blockDim.x * 4)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + j * 4), "l"(gsrc + j) : "memory");
#if !FEMX_EXPANDED
    const unsigned cdst = (unsigned)__cvta_generic_to_shared(s_cols);
    const int* csrc = col_loc + base;
    for (int j = threadIdx.x; j < cntn; j += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cdst + j * 4), "l"(csrc + j) : "memory");
#endif
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) s_vals[j] = real(0);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const int ln = threadIdx.x / ND;
  const int c = threadIdx.x - ln * ND;
  if (ln < nt) {
    const int row = i0 + ln;
    const int2 r0 = __ldg(&rowinfo[row]);
    const int rlen = __ldg(&rowinfo[row + 1].x) - r0.x;
    const int off = r0.x - base;
    real* srow = s_vals + off * (ND * ND) + c * rlen * ND;
    const int sp = __ldg(slice_ptr + (row >> 5)) + (row & 31);
    const unsigned* sc = s_code + (sp - sbase);
    const int np = r0.y;
    if (np > 0) {
      unsigned code = sc[0];
      real ox[NN - 1], oy[NN - 1], oz[NN - 1];
#if FEMX_EXPANDED
      // element-expanded coordinates X[NN*e + a] (the reference's layout, SURVEY Q17)
      const int* pelem = sell_elem + sp;
      real sx, sy, sz = real(0);
      {
        const int ea = __ldg(pelem);  // e*NN + li
        const int li = (code >> 28) & 3;
        const int e0 = ea - li;
        sx = __ldg(X + ea); sy = __ldg(Y + ea);
        if (DIM == 3) sz = __ldg(Z + ea);
#pragma unroll
        for (int j = 0; j < NN - 1; ++j) {
          int b = li + 1 + j; b -= b >= NN ? NN : 0;
          ox[j] = __ldg(X + e0 + b); oy[j] = __ldg(Y + e0 + b);
          oz[j] = DIM == 3 ? __ldg(Z + e0 + b) : real(0);
        }
      }
#else
      // the row's own node is a vertex of every incident element: it stays in registers
      const int* scol = s_cols + off;
      const i64 pself = (i64)scol[(code >> 21) & 127] * FEMX_CS;
      const real sx = __ldg(X + pself), sy = __ldg(Y + pself), sz = DIM == 3 ? __ldg(Z + pself) : real(0);
#pragma unroll
      for (int j = 0; j < NN - 1; ++j) {
        const i64 p = (i64)scol[(code >> (7 * j)) & 127] * FEMX_CS;
        ox[j] = __ldg(X + p); oy[j] = __ldg(Y + p);
        oz[j] = DIM == 3 ? __ldg(Z + p) : real(0);
      }
#endif
      real dacc[ND];  // the diagonal block row (own column) accumulates in registers
#pragma unroll
      for (int d = 0; d < ND; ++d) dacc[d] = real(0);
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
            int b = nli + 1 + j; b -= b >= NN ? NN : 0;
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
        switch (((code >> 28) & 3) * ND + c) {
          FEMX_CSR_CASES
        }
        if (!more) {
          const int ps = ((code >> 21) & 127) * ND;
#pragma unroll
          for (int d = 0; d < ND; ++d) srow[ps + d] = dacc[d];
        }
        code = ncd;
#pragma unroll
        for (int j = 0; j < NN - 1; ++j) { ox[j] = nox[j]; oy[j] = noy[j]; oz[j] = noz[j]; }
#if FEMX_EXPANDED
        sx = nsx; sy = nsy; sz = nsz;
#endif
      }
    }
  }
  __syncthreads();
  real* dst = vals + (i64)base * (ND * ND);
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) dst[j] = s_vals[j];
}
)FEMX";
