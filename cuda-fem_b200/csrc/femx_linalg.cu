// Validation layer: CSR SpMV on the pattern's native layout and the fused
// vector kernels of an unpreconditioned CG (no reference counterpart).
// All reductions use a fixed grid and a fixed tree → bitwise reproducible.
#include <algorithm>
#include "femx_internal.h"

namespace {

// One thread per dof row (i, c).  values are laid out as the dof-level CSR:
// row start = nd*nd*row_ptr[i] + c*nd*len, entry (p, d) at p*nd + d.
template <class T>
__global__ void spmv_k(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int row0, int n_rows, int nd,
                       const T* __restrict__ vals, const T* __restrict__ x, long long x_base, T* __restrict__ y) {
  int64_t t = (int64_t)row0 * nd + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_rows * nd) return;
  int i = (int)(t / nd), c = (int)(t - (int64_t)i * nd);
  int lo = rowinfo[i].x, len = rowinfo[i + 1].x - lo;
  const T* v = vals + (long long)lo * nd * nd + (long long)c * nd * len;
  double s = 0.0;
  for (int p = 0; p < len; ++p) {
    long long col = (long long)col_idx[lo + p] * nd - x_base;  // x_base already holds -nd*col_base
    for (int d = 0; d < nd; ++d) s += (double)v[p * nd + d] * (double)x[col + d];
  }
  y[t] = (T)s;
}

// Scalar operators (nd == 1): one CTA per tile of node rows.  The tile's values and column list are
// contiguous in the CSR arrays, so they are streamed into shared memory with fully coalesced
// asynchronous copies (cp.async, all in flight at once); each thread then walks its own row out of
// shared memory and only the x gathers go through L1/L2.  (Thread-per-row straight from global
// memory reads each row with a 15-element stride between lanes.)
// column offsets (column - own node) of the pattern's stencil class, passed by value
struct spmv_cls { int len; int off[24]; };

template <class T, int TILE>
__global__ void __launch_bounds__(TILE) spmv_tile_k(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx,
                                                     int row0, int n_rows, const T* __restrict__ vals, const T* __restrict__ x,
                                                     long long x_off, T* __restrict__ y, const spmv_cls cls, int row_node0,
                                                     int nblk_a, int row0_b, int n_rows_b) {
  extern __shared__ __align__(16) unsigned char sm[];
  // rows [row0, n_rows) in the first nblk_a CTAs, rows [row0_b, n_rows_b) in the others (the two ghost-reading ends of a slab
  // in ONE launch)
  const bool second = (int)blockIdx.x >= nblk_a;
  const int i0 = second ? row0_b + ((int)blockIdx.x - nblk_a) * TILE : row0 + blockIdx.x * TILE;
  if (second) n_rows = n_rows_b;
  const int nt = min(TILE, n_rows - i0);
  const int base = rowinfo[i0].x;
  const int cnt = rowinfo[i0 + nt].x - base;
  T* s_v = reinterpret_cast<T*>(sm);
  int* s_c = reinterpret_cast<int*>(s_v + cnt + (cnt & 1));
  const unsigned vdst = (unsigned)__cvta_generic_to_shared(s_v), cdst = (unsigned)__cvta_generic_to_shared(s_c);
  int lo = 0, len = 0;
  bool in_cls = true;
  if ((int)threadIdx.x < nt) {
    const int2 r0 = rowinfo[i0 + threadIdx.x];
    lo = r0.x - base;
    len = rowinfo[i0 + threadIdx.x + 1].x - base - lo;
    in_cls = cls.len > 0 && (r0.y & FEMX_ROW_SPEC);
  }
  // A tile made of stencil-class rows only (every interior tile of a structured mesh) needs no column list: the
  // columns of a class row are its own node + the class's constant offsets — 8 instead of 12 bytes per nonzero.
  const bool all_cls = __syncthreads_and(in_cls);
  for (int j = threadIdx.x; j < cnt; j += TILE) {
    if (sizeof(T) == 8)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(vdst + j * 8), "l"(vals + base + j) : "memory");
    else
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(vdst + j * 4), "l"(vals + base + j) : "memory");
    if (!all_cls) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cdst + j * 4), "l"(col_idx + base + j) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if ((int)threadIdx.x < nt) {
    double s = 0.0;
    if (all_cls) {
      const T* xr = x + ((long long)(row_node0 + i0 + (int)threadIdx.x) - x_off);
#pragma unroll 4
      for (int p = 0; p < len; ++p) s += (double)s_v[lo + p] * (double)__ldg(xr + cls.off[p]);
    } else {
      for (int p = 0; p < len; ++p) s += (double)s_v[lo + p] * (double)__ldg(x + ((long long)s_c[lo + p] - x_off));
    }
    y[i0 + threadIdx.x] = (T)s;
  }
}

template <class T>
__global__ void __launch_bounds__(256) dot2_partial(int64_t n, const T* __restrict__ a, const T* __restrict__ b,
                                                    const T* __restrict__ c, const T* __restrict__ d,
                                                    double* __restrict__ part) {
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    s0 += (double)a[i] * (double)b[i];
    if (c) s1 += (double)c[i] * (double)d[i];
  }
  __shared__ double sh0[256], sh1[256];
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh0[threadIdx.x] += sh0[threadIdx.x + o]; sh1[threadIdx.x] += sh1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[blockIdx.x] = sh0[0]; part[FEMX_DOT_BLOCKS + blockIdx.x] = sh1[0]; }
}

__global__ void __launch_bounds__(256) dot2_final(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  __shared__ double sh0[256], sh1[256];
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { s0 += part[i]; s1 += part[FEMX_DOT_BLOCKS + i]; }
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh0[threadIdx.x] += sh0[threadIdx.x + o]; sh1[threadIdx.x] += sh1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh0[0]; out[1] = sh1[0]; }
}

template <class T>
__global__ void axpy_ratio_k(int64_t n, const double* __restrict__ num, const double* __restrict__ den, double sign,
                             const T* __restrict__ x, T* __restrict__ y) {
  const double alpha = sign * (*num) / (*den);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (T)((double)y[i] + alpha * (double)x[i]);
}

template <class T>
__global__ void xpby_ratio_k(int64_t n, const double* __restrict__ num, const double* __restrict__ den,
                             const T* __restrict__ r, T* __restrict__ p) {
  const double beta = (*num) / (*den);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (T)((double)r[i] + beta * (double)p[i]);
}

// one thread per dof row; columns are LOCAL node ids, so flag/g are indexed by local dof
template <class T>
__global__ void dirichlet_k(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int n_rows, int nd,
                            int row_begin, const int* __restrict__ flag, const T* __restrict__ g,
                            T* __restrict__ vals, T* __restrict__ rhs) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_rows * nd) return;
  int i = (int)(t / nd), c = (int)(t - (int64_t)i * nd);
  int lo = rowinfo[i].x, len = rowinfo[i + 1].x - lo;
  T* v = vals + (long long)lo * nd * nd + (long long)c * nd * len;
  const long long self = (long long)(row_begin + i) * nd + c;  // local dof of this row
  const bool fixed = flag[self] != 0;
  T corr = T(0);
  for (int p = 0; p < len; ++p) {
    const long long colb = (long long)col_idx[lo + p] * nd;
    for (int d = 0; d < nd; ++d) {
      const long long col = colb + d;
      if (fixed) {
        v[p * nd + d] = col == self ? T(1) : T(0);
      } else if (flag[col] != 0) {
        corr += v[p * nd + d] * g[col];
        v[p * nd + d] = T(0);
      }
    }
  }
  if (rhs) rhs[t] = fixed ? g[self] : rhs[t] - corr;
}

inline unsigned nb(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// y[rows row_lo..row_hi) = A x for a range of NODE rows (the multi-GPU layer multiplies the rows that read no ghost
// column while the halo is still travelling)
int femx_spmv_range2(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                     int64_t row_lo, int64_t row_hi, int64_t row_lo2, int64_t row_hi2, void* stream);

int femx_spmv_range(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                    int64_t row_lo, int64_t row_hi, void* stream) {
  return femx_spmv_range2(p, dtype, d_values, d_x, x_base, d_y, row_lo, row_hi, 0, 0, stream);
}

// two row ranges in one launch (either may be empty)
int femx_spmv_range2(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                     int64_t row_lo, int64_t row_hi, int64_t row_lo2, int64_t row_hi2, void* stream) {
  if (!p || !d_values || !d_x || !d_y) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_spmv: NULL argument");
  if (row_lo < 0 || row_hi > p->n_rows || row_lo > row_hi || row_lo2 < 0 || row_hi2 > p->n_rows || row_lo2 > row_hi2)
    return femx_fail(p->ctx, FEMX_ERR_INVALID, "femx_spmv: bad row range");
  if (row_hi == row_lo && row_hi2 == row_lo2) return FEMX_OK;
  if (row_hi == row_lo) { row_lo = row_lo2; row_hi = row_hi2; row_lo2 = row_hi2 = 0; }
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  const int64_t n = (row_hi - row_lo) * p->nd;
  const long long xb = (long long)x_base - (long long)p->nd * p->col_base;
  const size_t rs = dtype == FEMX_F64 ? 8 : 4;
  // (max_tile_nnz is taken over 128-row tiles starting at multiples of 128; a window that starts elsewhere lies in two of
  // them — but never holds more than 128 of the longest rows, which on a structured slab is the smaller bound: twice the
  // shared memory halved the resident CTAs of every rank whose first interior row is not a multiple of 128)
  const size_t win_nnz = (row_lo % 128 || row_lo2 % 128) ? std::min<size_t>((size_t)p->max_tile_nnz * 2, (size_t)p->max_row * 128)
                                                         : (size_t)p->max_tile_nnz;
  const size_t smem = (win_nnz + 2) * (rs + 4);
  if (p->nd == 1 && smem <= 200 * 1024 && p->tile_nodes == 128) {
    // tile-staged kernel (128-row tiles)
    const int nblk_a = (int)((row_hi - row_lo + 127) / 128), nblk_b = (int)((row_hi2 - row_lo2 + 127) / 128);
    const unsigned blocks = (unsigned)(nblk_a + nblk_b);
    spmv_cls cls = {};
    if (p->spec_np > 0 && p->spec_rlen <= 24) {
      cls.len = p->spec_rlen;
      for (int k = 0; k < p->spec_rlen; ++k) cls.off[k] = p->spec_off[k];
    }
    const int row_node0 = (int)p->row_begin;
    if (dtype == FEMX_F64) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(spmv_tile_k<double, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      spmv_tile_k<double, 128><<<blocks, 128, smem, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)row_lo, (int)row_hi,
                                                                             (const double*)d_values, (const double*)d_x, xb,
                                                                             (double*)d_y, cls, row_node0, nblk_a, (int)row_lo2, (int)row_hi2);
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(spmv_tile_k<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      spmv_tile_k<float, 128><<<blocks, 128, smem, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)row_lo, (int)row_hi,
                                                                            (const float*)d_values, (const float*)d_x, xb,
                                                                            (float*)d_y, cls, row_node0, nblk_a, (int)row_lo2, (int)row_hi2);
    }
  } else if (dtype == FEMX_F64)
    spmv_k<double><<<nb(n), 256, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)row_lo, (int)row_hi, p->nd,
                                                            (const double*)d_values, (const double*)d_x, xb, (double*)d_y);
  else
    spmv_k<float><<<nb(n), 256, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)row_lo, (int)row_hi, p->nd,
                                                           (const float*)d_values, (const float*)d_x, xb, (float*)d_y);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  if (!(p->nd == 1 && smem <= 200 * 1024 && p->tile_nodes == 128) && row_hi2 > row_lo2)
    return femx_spmv_range2(p, dtype, d_values, d_x, x_base, d_y, row_lo2, row_hi2, 0, 0, stream);
  return FEMX_OK;
}

extern "C" {

int femx_spmv_rows(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                   int64_t row_lo, int64_t row_hi, void* stream) {
  return femx_spmv_range(p, dtype, d_values, d_x, x_base, d_y, row_lo, row_hi, stream);
}

int femx_spmv(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
              void* stream) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_spmv: NULL argument");
  return femx_spmv_range(p, dtype, d_values, d_x, x_base, d_y, 0, p->n_rows, stream);
}

int femx_apply_dirichlet(const femx_pattern* p, int dtype, const int32_t* d_flag, const void* d_g, void* d_values,
                         void* d_rhs, void* stream) {
  if (!p || !d_flag || !d_g || !d_values)
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_apply_dirichlet: NULL argument");
  if (p->n_rows == 0) return FEMX_OK;
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  int64_t n = p->n_rows * p->nd;
  if (dtype == FEMX_F64)
    dirichlet_k<double><<<nb(n), 256, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)p->n_rows, p->nd,
                                                                 (int)p->row_begin, d_flag, (const double*)d_g,
                                                                 (double*)d_values, (double*)d_rhs);
  else
    dirichlet_k<float><<<nb(n), 256, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)p->n_rows, p->nd,
                                                                (int)p->row_begin, d_flag, (const float*)d_g,
                                                                (float*)d_values, (float*)d_rhs);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_dot2(femx_ctx* ctx, int dtype, int64_t n, const void* d_a, const void* d_b, const void* d_c, const void* d_d,
              double* d_out, void* stream) {
  if (!ctx || !d_a || !d_b || !d_out || (d_c && !d_d))
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_dot2: NULL argument");
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  // partial sums: stream-ordered scratch per call (two calls on different streams of one ctx do not share it)
  double* scratch = nullptr;
  FEMX_CUDA_OK(ctx, ctx->pool ? cudaMallocFromPoolAsync((void**)&scratch, sizeof(double) * 2 * FEMX_DOT_BLOCKS, ctx->pool, st)
                              : cudaMallocAsync((void**)&scratch, sizeof(double) * 2 * FEMX_DOT_BLOCKS, st));
  int blocks = (int)std::min<int64_t>(FEMX_DOT_BLOCKS, std::max<int64_t>(1, (n + 255) / 256));
  if (dtype == FEMX_F64)
    dot2_partial<double><<<blocks, 256, 0, st>>>(n, (const double*)d_a, (const double*)d_b, (const double*)d_c,
                                                 (const double*)d_d, scratch);
  else
    dot2_partial<float><<<blocks, 256, 0, st>>>(n, (const float*)d_a, (const float*)d_b, (const float*)d_c,
                                                (const float*)d_d, scratch);
  dot2_final<<<1, 256, 0, st>>>(scratch, blocks, d_out);
  cudaFreeAsync(scratch, st);
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_axpy_ratio(femx_ctx* ctx, int dtype, int64_t n, const double* d_num, const double* d_den, double sign,
                    const void* d_x, void* d_y, void* stream) {
  if (!ctx || !d_num || !d_den || !d_x || !d_y) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_axpy_ratio: NULL argument");
  if (n == 0) return FEMX_OK;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (dtype == FEMX_F64)
    axpy_ratio_k<double><<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, d_num, d_den, sign, (const double*)d_x, (double*)d_y);
  else
    axpy_ratio_k<float><<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, d_num, d_den, sign, (const float*)d_x, (float*)d_y);
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_xpby_ratio(femx_ctx* ctx, int dtype, int64_t n, const double* d_num, const double* d_den, const void* d_r,
                    void* d_p, void* stream) {
  if (!ctx || !d_num || !d_den || !d_r || !d_p) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_xpby_ratio: NULL argument");
  if (n == 0) return FEMX_OK;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (dtype == FEMX_F64)
    xpby_ratio_k<double><<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, d_num, d_den, (const double*)d_r, (double*)d_p);
  else
    xpby_ratio_k<float><<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, d_num, d_den, (const float*)d_r, (float*)d_p);
  FEMX_CUDA_OK(ctx, cudaGetLastError());
  return FEMX_OK;
}

}  // extern "C"
