// Device-side structured mesh generators and the element-expanded coordinate layout.
#include "femx_internal.h"

namespace {

// RectangleMesh::generate (fea_symbolic_nvrtc_sparse.cpp:170-216) without the
// per-node `new`: node (i,j) → index i*(nCol+1)+j, x = x0 + j*stepx, y = y0 + i*stepy.
template <class T>
__global__ void rect_nodes(double x0, double y0, double stepx, double stepy, int64_t nRow, int64_t nCol,
                           int64_t row_lo, int64_t n_local, T* __restrict__ X, T* __restrict__ Y,
                           int* __restrict__ flag) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_local) return;
  int64_t i = row_lo + n / (nCol + 1), j = n % (nCol + 1);
  // mul then add, no FMA contraction: bit-identical to the reference's host arithmetic
  if (X) X[n] = (T)__dadd_rn(x0, __dmul_rn((double)j, stepx));
  if (Y) Y[n] = (T)__dadd_rn(y0, __dmul_rn((double)i, stepy));
  if (flag) flag[n] = (i == 0 || i == nRow || j == 0 || j == nCol) ? 1 : 0;
}

// one thread per cell: triangles (n1, n1+1, n3) and (n1+1, n3+1, n3)
__global__ void rect_cells(int64_t nCol, int64_t n_cells, int* __restrict__ conn) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  int64_t i = c / nCol, j = c % nCol;
  int n1 = (int)(i * (nCol + 1) + j), n3 = (int)((i + 1) * (nCol + 1) + j);
  int* o = conn + 6 * c;
  o[0] = n1; o[1] = n1 + 1; o[2] = n3;
  o[3] = n1 + 1; o[4] = n3 + 1; o[5] = n3;
}

template <class T>
__global__ void box_nodes(double x0, double y0, double z0, double hx, double hy, double hz, int64_t nx,
                          int64_t ny, int64_t k_lo, int64_t n_local, T* __restrict__ X, T* __restrict__ Y,
                          T* __restrict__ Z) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_local) return;
  int64_t i = n % (nx + 1), j = (n / (nx + 1)) % (ny + 1), k = k_lo + n / ((nx + 1) * (ny + 1));
  if (X) X[n] = (T)__dadd_rn(x0, __dmul_rn((double)i, hx));
  if (Y) Y[n] = (T)__dadd_rn(y0, __dmul_rn((double)j, hy));
  if (Z) Z[n] = (T)__dadd_rn(z0, __dmul_rn((double)k, hz));
}

// Kuhn split, one thread per tet; permutation table and orientation fix as in
// DESIGN.md (even permutations swap their two middle vertices).

// one thread per tetrahedron, one 16-byte store; 32-bit arithmetic (the caller has checked 4 * n_tets < 2^31) and the
// permutation table packed into an immediate (a constant-memory table indexed per thread serialises)
__global__ void box_cells(int64_t nx64, int64_t ny64, int64_t n_tets, int* __restrict__ conn) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)n_tets) return;
  const unsigned nx = (unsigned)nx64, ny = (unsigned)ny64;
  const unsigned cell = t / 6u, p = t - cell * 6u;
  const unsigned line = cell / nx, cz = line / ny;
  unsigned c[3] = {cell - line * nx, line - cz * ny, cz};
  // axis order of Kuhn tetrahedron p, 2 bits per step: {0,1,2} {0,2,1} {1,0,2} {1,2,0} {2,0,1} {2,1,0}; odd permutations: p = 1, 2, 5
  const unsigned long long kPerm = 0x24ull | (0x18ull << 6) | (0x21ull << 12) | (0x09ull << 18) | (0x12ull << 24) | (0x06ull << 30);
  const unsigned perm = (unsigned)(kPerm >> (6 * p)) & 63u;
  int v[4];
  v[0] = (int)((c[2] * (ny + 1) + c[1]) * (nx + 1) + c[0]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const unsigned ax = (perm >> (2 * a)) & 3u;
    c[0] += ax == 0; c[1] += ax == 1; c[2] += ax == 2;
    v[a + 1] = (int)((c[2] * (ny + 1) + c[1]) * (nx + 1) + c[0]);
  }
  if (!((0x26u >> p) & 1u)) { int tmp = v[1]; v[1] = v[2]; v[2] = tmp; }
  reinterpret_cast<int4*>(conn)[t] = make_int4(v[0], v[1], v[2], v[3]);
}

template <class T>
__global__ void expand_k(int64_t total, const int* __restrict__ conn, const T* __restrict__ node,
                         T* __restrict__ elem) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < total) elem[k] = node[conn[k]];
}

inline unsigned nb(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

int femx_mesh_rectangle(femx_ctx* ctx, double x0, double x1, double y0, double y1, int64_t nRow, int64_t nCol,
                        int64_t row_lo, int64_t row_hi, int dtype, void* d_x, void* d_y, int32_t* d_flag,
                        int32_t* d_conn, void* stream) {
  if (!ctx) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_mesh_rectangle: ctx is NULL");
  if (nRow < 1 || nCol < 1 || row_lo < 0 || row_hi > nRow || row_lo > row_hi)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_mesh_rectangle: bad sizes");
  int64_t n_local = (row_hi - row_lo + 1) * (nCol + 1);
  int64_t n_cells = (row_hi - row_lo) * nCol;
  if (n_local >= (1LL << 31) - 1 || 6 * n_cells >= (1LL << 31) - 1)
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_mesh_rectangle: slab exceeds 32-bit indexing");
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  double stepx = (x1 - x0) / (double)nCol, stepy = (y1 - y0) / (double)nRow;
  if (d_x || d_y || d_flag) {
    if (dtype == FEMX_F64)
      rect_nodes<double><<<nb(n_local), 256, 0, st>>>(x0, y0, stepx, stepy, nRow, nCol, row_lo, n_local,
                                                      (double*)d_x, (double*)d_y, d_flag);
    else
      rect_nodes<float><<<nb(n_local), 256, 0, st>>>(x0, y0, stepx, stepy, nRow, nCol, row_lo, n_local,
                                                     (float*)d_x, (float*)d_y, d_flag);
  }
  if (d_conn && n_cells > 0) rect_cells<<<nb(n_cells), 256, 0, st>>>(nCol, n_cells, d_conn);
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_mesh_box(femx_ctx* ctx, double x0, double x1, double y0, double y1, double z0, double z1, int64_t nx,
                  int64_t ny, int64_t nz, int64_t k_lo, int64_t k_hi, int dtype, void* d_x, void* d_y, void* d_z,
                  int32_t* d_conn, void* stream) {
  if (!ctx) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_mesh_box: ctx is NULL");
  if (nx < 1 || ny < 1 || nz < 1 || k_lo < 0 || k_hi > nz || k_lo > k_hi)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_mesh_box: bad sizes");
  int64_t n_local = (k_hi - k_lo + 1) * (nx + 1) * (ny + 1);
  int64_t n_tets = 6 * (k_hi - k_lo) * nx * ny;
  if (n_local >= (1LL << 31) - 1 || 4 * n_tets >= (1LL << 31) - 1)
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_mesh_box: slab exceeds 32-bit indexing");
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  double hx = (x1 - x0) / (double)nx, hy = (y1 - y0) / (double)ny, hz = (z1 - z0) / (double)nz;
  if (d_x || d_y || d_z) {
    if (dtype == FEMX_F64)
      box_nodes<double><<<nb(n_local), 256, 0, st>>>(x0, y0, z0, hx, hy, hz, nx, ny, k_lo, n_local, (double*)d_x,
                                                     (double*)d_y, (double*)d_z);
    else
      box_nodes<float><<<nb(n_local), 256, 0, st>>>(x0, y0, z0, hx, hy, hz, nx, ny, k_lo, n_local, (float*)d_x,
                                                    (float*)d_y, (float*)d_z);
  }
  if (d_conn && n_tets > 0) {
    if ((uintptr_t)d_conn % 16)
      return femx_fail(ctx, FEMX_ERR_INVALID, "femx_mesh_box: d_conn must be 16-byte aligned");
    box_cells<<<nb(n_tets), 256, 0, st>>>(nx, ny, n_tets, d_conn);
  }
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_mesh_expand(femx_ctx* ctx, int dtype, int nn, int64_t n_elems, const int32_t* d_conn, const void* d_node,
                     void* d_elem, void* stream) {
  if (!ctx || !d_conn || !d_node || !d_elem)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_mesh_expand: NULL argument");
  int64_t total = n_elems * nn;
  if (total == 0) return FEMX_OK;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (dtype == FEMX_F64)
    expand_k<double><<<nb(total), 256, 0, (cudaStream_t)stream>>>(total, d_conn, (const double*)d_node, (double*)d_elem);
  else
    expand_k<float><<<nb(total), 256, 0, (cudaStream_t)stream>>>(total, d_conn, (const float*)d_node, (float*)d_elem);
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

}  // extern "C"
