// The symbolic pass: CSR pattern + scatter map on the device.
//
// Replaces the reference's host-side Mesh::getNeighborNodesList
// (fea_symbolic_nvrtc_sparse2.cpp:181-210: one std::set<int> per node, 9*NE
// red-black-tree inserts) with
//   1. histogram of node incidences            (integer atomics: result exact)
//   2. exclusive scan                          → pair_ptr
//   3. bucket fill + per-row sort by element   → pair_elem   (ascending ⇒ deterministic)
//   4. per-row merge of the incident elements' nodes into a sorted unique list → row length
//   5. exclusive scan                          → row_ptr
//   6. column fill + per-incidence position codes (the element-slot → CSR-offset map)
// The output is bit-identical to the reference's sorted std::set rows.
#include <algorithm>
#include <array>
#include <climits>
#include <cstring>
#include <cstdio>
#include <vector>

#include "femx_internal.h"

#define FEMX_MAX_ROW 128  // node-level row length limit (7-bit positions in pair_code)

namespace {

// (cudaFreeAsync(NULL) is an error that would stay behind as the last error: clean-up paths free what exists)
inline void free_async(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }

// ------------------------------------------------------------------ scan ---
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int woff = wid ? warp_sums[wid - 1] : 0;
  *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return woff + x - v;
}

// phase 1: per-tile sums (64-bit so that overflow of the grand total is detected)
__global__ void scan_tile_sums(const int* __restrict__ in, int64_t n, long long* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// phase 2: one block scans the tile sums in place (exclusive), writes the total
__global__ void scan_sums(long long* __restrict__ sums, int nt, long long* __restrict__ total) {
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < nt; b += SCAN_THREADS) {
    int i = b + threadIdx.x;
    long long v = i < nt ? sums[i] : 0;
    // simple Hillis-Steele in shared memory (nt is small)
    __shared__ long long buf[SCAN_THREADS];
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
      long long y = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < nt) sums[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry += buf[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// phase 3: exclusive scan inside each tile + tile offset; out has n+1 entries
__global__ void scan_apply(const int* __restrict__ in, int64_t n, const long long* __restrict__ sums,
                           int* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_exclusive_scan(s, &tot) + (int)sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = ex;
}

// out[0..n] = exclusive scan of in[0..n); *h_total = grand total (host)
int exclusive_scan(femx_ctx* ctx, const int* d_in, int64_t n, int* d_out, long long* h_total,
                   cudaStream_t st, long long* d_total = nullptr) {
  if (n == 0) {
    FEMX_CUDA_OK(ctx, cudaMemsetAsync(d_out, 0, sizeof(int), st));
    if (d_total) FEMX_CUDA_OK(ctx, cudaMemsetAsync(d_total, 0, sizeof(long long), st));
    else *h_total = 0;
    return FEMX_OK;
  }
  int nt = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  long long* d_sums = nullptr;
  FEMX_CUDA_OK(ctx, ctx->pool ? cudaMallocFromPoolAsync((void**)&d_sums, sizeof(long long) * (nt + 1), ctx->pool, st)
                              : cudaMallocAsync((void**)&d_sums, sizeof(long long) * (nt + 1), st));
  scan_tile_sums<<<nt, SCAN_THREADS, 0, st>>>(d_in, n, d_sums);
  scan_sums<<<1, SCAN_THREADS, 0, st>>>(d_sums, nt, d_sums + nt);
  scan_apply<<<nt, SCAN_THREADS, 0, st>>>(d_in, n, d_sums, d_out);
  cudaError_t e;
  if (d_total) {  // the caller collects the total later, with its own synchronisation
    e = cudaMemcpyAsync(d_total, d_sums + nt, sizeof(long long), cudaMemcpyDeviceToDevice, st);
  } else {
    e = cudaMemcpyAsync(h_total, d_sums + nt, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  free_async(d_sums, st);
  FEMX_CUDA_OK(ctx, e);
  return FEMX_OK;
}

// ------------------------------------------------------------ incidences ---
__global__ void count_pairs(const int* __restrict__ conn, int64_t total, int nn, int row_begin, int row_end,
                            int n_nodes, int* __restrict__ cnt, int* __restrict__ err) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  int node = conn[k];
  if (node < 0 || node >= n_nodes) { atomicOr(err, 1); return; }
  // a vertex listed twice in one element: degenerate, and the numeric pass relies on the
  // vertices of an element occupying distinct value slots
  {
    const int64_t e0 = (k / nn) * nn;
    for (int64_t q = e0; q < k; ++q)
      if (conn[q] == node) atomicOr(err, 4);
  }
  if (node >= row_begin && node < row_end) atomicAdd(&cnt[node - row_begin], 1);
}

__global__ void fill_pairs(const int* __restrict__ conn, int64_t total, int row_begin, int row_end,
                           const int* __restrict__ pair_ptr, int* __restrict__ cursor,
                           int* __restrict__ pair_elem) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  int node = conn[k];
  if (node >= row_begin && node < row_end) {
    int r = node - row_begin;
    int slot = atomicAdd(&cursor[r], 1);
    pair_elem[pair_ptr[r] + slot] = (int)k;  // k = e*nn + li
  }
}

// per row: insertion sort of its incidences (ascending e*nn+li) — the bucket
// order left by the atomics is arbitrary, the sorted order is not.
__global__ void sort_pairs(const int* __restrict__ pair_ptr, int n_rows, int* __restrict__ pair_elem) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int lo = pair_ptr[r], hi = pair_ptr[r + 1];
  for (int i = lo + 1; i < hi; ++i) {
    int v = pair_elem[i], j = i - 1;
    while (j >= lo && pair_elem[j] > v) { pair_elem[j + 1] = pair_elem[j]; --j; }
    pair_elem[j + 1] = v;
  }
}

// Sorted duplicate-free union of the nodes of a row's incident elements
// (= the std::set of getNeighborNodesList).  Returns the length, or -1 if it
// exceeds FEMX_MAX_ROW.
template <int NN>
__device__ __forceinline__ int build_row(const int* __restrict__ conn, const int* __restrict__ pair_elem,
                                         int lo, int hi, int* list) {
  int len = 0;
  for (int k = lo; k < hi; ++k) {
    const int e = pair_elem[k] / NN;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
      const int node = conn[(int64_t)e * NN + a];
      int p = len;
      while (p > 0 && list[p - 1] > node) --p;
      if (p > 0 && list[p - 1] == node) continue;
      if (len == FEMX_MAX_ROW) return -1;
      for (int q = len; q > p; --q) list[q] = list[q - 1];
      list[p] = node;
      ++len;
    }
  }
  return len;
}

template <int NN>
__global__ void row_lengths(const int* __restrict__ conn, const int* __restrict__ pair_ptr,
                            const int* __restrict__ pair_elem, int n_rows, int* __restrict__ rowlen,
                            int* __restrict__ err, int* __restrict__ max_row) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int list[FEMX_MAX_ROW];
  int len = build_row<NN>(conn, pair_elem, pair_ptr[r], pair_ptr[r + 1], list);
  if (len < 0) { atomicOr(err, 2); len = 0; }
  rowlen[r] = len;
  atomicMax(max_row, len);
}

template <int NN>
__global__ void row_fill(const int* __restrict__ conn, const int* __restrict__ pair_ptr,
                         const int* __restrict__ pair_elem, const int* __restrict__ row_ptr, int n_rows,
                         int row_begin, int* __restrict__ col_idx, unsigned* __restrict__ pair_code,
                         int2* __restrict__ rowinfo) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) { rowinfo[r] = make_int2(row_ptr[r], 0); return; }
  int list[FEMX_MAX_ROW];
  const int lo = pair_ptr[r], hi = pair_ptr[r + 1];
  const int len = build_row<NN>(conn, pair_elem, lo, hi, list);
  const int rp = row_ptr[r];
  for (int p = 0; p < len; ++p) col_idx[rp + p] = list[p];
  // position of the row's own node (local id = row_begin + r) in its sorted column list
  int self_pos = 0;
  {
    const int self = row_begin + r;
    int a0 = 0, b0 = len;
    while (a0 < b0) {
      int m = (a0 + b0) >> 1;
      if (list[m] < self) a0 = m + 1; else b0 = m;
    }
    self_pos = a0;
  }
  rowinfo[r] = make_int2(rp, (hi - lo) | (self_pos << 24));
  unsigned seen[FEMX_MAX_ROW / 32] = {0u, 0u, 0u, 0u};  // value slots already touched by an earlier incidence
  for (int k = lo; k < hi; ++k) {
    const int pe = pair_elem[k];
    const int e = pe / NN, li = pe - e * NN;
    // bits 7*j: position in the sorted row of vertex femx_oth(NN, li, j); bits 21+j: first-touch flag
    // of that position (no earlier incidence of the row contributes there); bits 28-29: li
    unsigned code = (unsigned)li << 28;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
      if (a == li) continue;
      const int node = conn[(int64_t)e * NN + a];
      int a0 = 0, b0 = len;  // binary search in the sorted list
      while (a0 < b0) {
        int m = (a0 + b0) >> 1;
        if (list[m] < node) a0 = m + 1; else b0 = m;
      }
      const int j = NN == 4 ? (a ^ li) - 1 : (a - li - 1 + 3) % 3;
      code |= (unsigned)a0 << (7 * j);
      if (!(seen[a0 >> 5] & (1u << (a0 & 31)))) {
        code |= 1u << (21 + j);
        seen[a0 >> 5] |= 1u << (a0 & 31);
      }
    }
    pair_code[k] = code;
  }
}

// SELL-32 layout of the scatter map: slice = 32 consecutive rows, padded to the
// slice's longest incidence list.
__global__ void slice_sizes(const int* __restrict__ pair_ptr, int n_rows, int n_slices, int* __restrict__ size) {
  int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_slices) return;
  int r = s * 32 + (threadIdx.x & 31);
  int np = r < n_rows ? pair_ptr[r + 1] - pair_ptr[r] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np = max(np, __shfl_xor_sync(0xffffffffu, np, o));
  if ((threadIdx.x & 31) == 0) size[s] = np * 32;
}

__global__ void to_sell(const int* __restrict__ pair_ptr, const int* __restrict__ slice_ptr, int n_rows,
                        const int* __restrict__ pair_elem, const unsigned* __restrict__ pair_code,
                        int* __restrict__ sell_elem, unsigned* __restrict__ sell_code) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int lo = pair_ptr[r], np = pair_ptr[r + 1] - lo;
  const int sp = slice_ptr[r >> 5] + (r & 31);
  for (int it = 0; it < np; ++it) {
    sell_elem[sp + it * 32] = pair_elem[lo + it];
    sell_code[sp + it * 32] = pair_code[lo + it];
  }
}

__global__ void tile_max(const int2* __restrict__ rowinfo, const int* __restrict__ slice_ptr, int n_rows, int tile,
                         int* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i0 = (int64_t)t * tile;
  if (i0 >= n_rows) return;
  int i1 = (int)min((int64_t)n_rows, i0 + tile);
  atomicMax(out, rowinfo[i1].x - rowinfo[i0].x);
  atomicMax(out + 1, slice_ptr[(i1 + 31) >> 5] - slice_ptr[i0 >> 5]);
}

// ------------------------------------------------------- stencil classes ---
// Rows whose (incidence count, row length, own position, scatter-code sequence) coincide form a
// stencil class: the numeric pass does exactly the same thing for each of them, only on different
// nodes.  On a structured mesh every interior row is in one class.
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h += v + 0x9e3779b97f4a7c15ull;  // splitmix64 step
  h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ull;
  h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
  return h ^ (h >> 31);
}

__global__ void row_class_hash(const int* __restrict__ pair_ptr, const unsigned* __restrict__ pair_code,
                               const int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int row_begin,
                               int n_rows, unsigned long long* __restrict__ hash) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int lo = pair_ptr[r], hi = pair_ptr[r + 1];
  const int2 ri = rowinfo[r];
  const int rlen = rowinfo[r + 1].x - ri.x;
  unsigned long long h = mix64(0x66656d78ull, (unsigned long long)(hi - lo) | ((unsigned long long)rlen << 24) |
                                                  ((unsigned long long)((unsigned)ri.y >> 24) << 40));
  for (int k = lo; k < hi; ++k) h = mix64(h, pair_code[k]);
  for (int k = 0; k < rlen; ++k) h = mix64(h, (unsigned)(col_idx[ri.x + k] - (row_begin + r)));
  hash[r] = h;
}

// sample rows: one per stride, at a pseudo-random place inside it (an even spacing aliases with the line
// length of a structured mesh: every sample then lands in the same mesh column, e.g. on the boundary)
__global__ void sample_rows(const unsigned long long* __restrict__ hash, long long n_rows, int n_samples,
                            unsigned long long* __restrict__ out_hash, int* __restrict__ out_row) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_samples) return;
  const long long stride = n_rows / n_samples;
  const long long r = (long long)i * stride + (long long)(mix64(0x73616d70ull, (unsigned long long)i) % (unsigned long long)stride);
  out_hash[i] = hash[r];
  out_row[i] = (int)r;
}

// flags the rows of the class of row `ref` (full comparison, the hash only filters) and counts them
__global__ void mark_class(const int* __restrict__ pair_ptr, const unsigned* __restrict__ pair_code,
                           int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int n_rows,
                           const unsigned long long* __restrict__ hash, int ref, int* __restrict__ count) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool match = false;
  if (r < n_rows && hash[r] == hash[ref]) {
    const int lo = pair_ptr[r], np = pair_ptr[r + 1] - lo;
    const int rlo = pair_ptr[ref], rnp = pair_ptr[ref + 1] - rlo;
    const int2 a = rowinfo[r], b = rowinfo[ref];
    match = np == rnp && rowinfo[r + 1].x - a.x == rowinfo[ref + 1].x - b.x &&
            ((unsigned)a.y >> 24) == ((unsigned)b.y >> 24);
    for (int k = 0; match && k < np; ++k) match = pair_code[lo + k] == pair_code[rlo + k];
    // ... and the same column OFFSETS from the row's own node (col - row is what the kernel adds)
    const int rlen = rowinfo[ref + 1].x - b.x;
    for (int k = 0; match && k < rlen; ++k) match = col_idx[a.x + k] - r == col_idx[b.x + k] - ref;
  }
  const unsigned m = __ballot_sync(0xffffffffu, match);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
  // `ref` itself is flagged last (by the host) so that it stays comparable while this kernel runs
  if (match && r != ref) rowinfo[r].y |= FEMX_ROW_SPEC;
}

__global__ void mark_ref(int2* __restrict__ rowinfo, int ref) { rowinfo[ref].y |= FEMX_ROW_SPEC; }

// rows outside the class: flag (for the scan), then compaction in ascending order
__global__ void other_flags(const int2* __restrict__ rowinfo, int n_rows, int* __restrict__ flag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) flag[r] = (rowinfo[r].y & FEMX_ROW_SPEC) ? 0 : 1;
}

__global__ void other_fill(const int2* __restrict__ rowinfo, int n_rows, const int* __restrict__ pos,
                           int* __restrict__ list, int* __restrict__ max_len) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || (rowinfo[r].y & FEMX_ROW_SPEC)) return;
  list[pos[r]] = r;
  atomicMax(max_len, rowinfo[r + 1].x - rowinfo[r].x);
}

// ---------------------------------------------------------------- exports ---
__global__ void export_csr_k(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int n_rows,
                             int nd, int col_base, const int* __restrict__ l2g, long long* __restrict__ rp64,
                             int* __restrict__ rp32, int* __restrict__ dcol) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // dof row
  int64_t nrows_d = (int64_t)n_rows * nd;
  if (t > nrows_d) return;
  if (t == nrows_d) {
    long long tot = (long long)rowinfo[n_rows].x * nd * nd;
    if (rp64) rp64[t] = tot;
    if (rp32) rp32[t] = (int)tot;
    return;
  }
  int i = (int)(t / nd), c = (int)(t - (int64_t)i * nd);
  int lo = rowinfo[i].x, len = rowinfo[i + 1].x - lo;
  long long start = (long long)lo * nd * nd + (long long)c * nd * len;
  if (rp64) rp64[t] = start;
  if (rp32) rp32[t] = (int)start;
  if (dcol)
    for (int p = 0; p < len; ++p) {
      int col = l2g ? l2g[col_idx[lo + p]] : col_idx[lo + p] + col_base;
      for (int d = 0; d < nd; ++d) dcol[start + (long long)p * nd + d] = col * nd + d;
    }
}

__global__ void export_ell_k(const int2* __restrict__ rowinfo, const int* __restrict__ col_idx, int n_rows,
                             int width, int col_base, int* __restrict__ len_out, int* __restrict__ idx) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int lo = rowinfo[r].x, len = rowinfo[r + 1].x - lo;
  if (len_out) len_out[r] = len;
  if (idx)
    for (int j = 0; j < width; ++j) idx[(int64_t)r * width + j] = j < len ? col_idx[lo + j] + col_base : 0;
}

template <class T>
__global__ void csr_to_ell_k(const int2* __restrict__ rowinfo, int n_rows, int width,
                             const T* __restrict__ vals, T* __restrict__ ell) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int lo = rowinfo[r].x, len = rowinfo[r + 1].x - lo;
  for (int j = 0; j < width; ++j) ell[(int64_t)r * width + j] = j < len ? vals[lo + j] : T(0);
}

// arrays owned by the pattern object: from the context's stream-ordered pool (a rebuilt pattern recycles the memory of a
// destroyed one instead of paying cudaMalloc / cudaFree of gigabytes), plain cudaMalloc when there is no pool
template <class T>
int dev_alloc(femx_ctx* ctx, T** p, int64_t n, int64_t* bytes, cudaStream_t st) {
  size_t b = sizeof(T) * (size_t)(n > 0 ? n : 1);
  cudaError_t e = ctx->pool ? cudaMallocFromPoolAsync((void**)p, b, ctx->pool, st) : cudaMalloc((void**)p, b);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return femx_fail(ctx, FEMX_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", b, cudaGetErrorString(e));
  }
  if (bytes) *bytes += (int64_t)b;
  return FEMX_OK;
}

// temporaries of the symbolic pass come from the stream-ordered pool (no device-wide sync per free)
template <class T>
int tmp_alloc(femx_ctx* ctx, T** p, int64_t n, cudaStream_t st) {
  size_t b = sizeof(T) * (size_t)(n > 0 ? n : 1);
  cudaError_t e = ctx->pool ? cudaMallocFromPoolAsync((void**)p, b, ctx->pool, st) : cudaMallocAsync((void**)p, b, st);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return femx_fail(ctx, FEMX_ERR_NOMEM, "cudaMallocAsync(%zu bytes) failed: %s", b, cudaGetErrorString(e));
  }
  return FEMX_OK;
}

inline unsigned nblocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// the rows outside the class, compacted (ascending): the numeric pass runs them in a launch of their own
int build_other_rows(femx_ctx* ctx, femx_pattern* p, long long class_rows, cudaStream_t st) {
  const int64_t nr = p->n_rows;
  int *d_flag = nullptr, *d_pos = nullptr, *d_count = nullptr;
  int rc = tmp_alloc(ctx, &d_flag, nr + 1, st);
  if (rc == FEMX_OK) rc = tmp_alloc(ctx, &d_pos, nr + 1, st);
  if (rc == FEMX_OK) rc = tmp_alloc(ctx, &d_count, 1, st);
  long long n_other = 0;
  if (rc == FEMX_OK) {
    other_flags<<<nblocks(nr, 256), 256, 0, st>>>(p->d_rowinfo, (int)nr, d_flag);
    rc = exclusive_scan(ctx, d_flag, nr, d_pos, &n_other, st);
  }
  if (rc == FEMX_OK && n_other != nr - class_rows)
    rc = femx_fail(ctx, FEMX_ERR_CUDA, "stencil class: %lld rows outside the class, expected %lld", n_other,
                   (long long)(nr - class_rows));
  if (rc == FEMX_OK) rc = dev_alloc(ctx, &p->d_other_rows, n_other, &p->bytes, st);
  if (rc == FEMX_OK) {
    cudaMemsetAsync(d_count, 0, sizeof(int), st);
    other_fill<<<nblocks(nr, 256), 256, 0, st>>>(p->d_rowinfo, (int)nr, d_pos, p->d_other_rows, d_count);
    int mx = 0;
    cudaError_t e = cudaMemcpyAsync(&mx, d_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = femx_fail(ctx, FEMX_ERR_CUDA, "stencil class: row list failed: %s", cudaGetErrorString(e));
    p->n_other = n_other;
    p->max_row_other = mx;
  }
  free_async(d_flag, st);
  free_async(d_pos, st);
  free_async(d_count, st);
  return rc;
}

// Finds the dominant stencil class among evenly spaced sample rows, flags its rows / whole tiles in
// rowinfo and keeps a host copy of the class's scatter codes for the JIT (femx_form.cpp).
int detect_stencil_class(femx_ctx* ctx, femx_pattern* p, const int* d_pair_ptr, const unsigned* d_pair_code,
                         cudaStream_t st) {
  const int64_t nr = p->n_rows;
  unsigned long long* d_hash = nullptr;
  int* d_count = nullptr;
  int rc = tmp_alloc(ctx, &d_hash, nr, st);
  if (rc == FEMX_OK) rc = tmp_alloc(ctx, &d_count, 1, st);
  auto done = [&](int code) {
    free_async(d_hash, st);
    free_async(d_count, st);
    return code;
  };
  if (rc != FEMX_OK) return done(rc);
#define SC_CUDA(call)                                                                                \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return done(femx_fail(ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)));  \
  } while (0)
  row_class_hash<<<nblocks(nr, 128), 128, 0, st>>>(d_pair_ptr, d_pair_code, p->d_rowinfo, p->d_col_idx,
                                                   (int)p->row_begin, (int)nr, d_hash);
  const int S = (int)std::min<int64_t>(nr, 256);
  std::vector<unsigned long long> hs(S);
  std::vector<int> hrow(S);
  {
    unsigned long long* d_sh = nullptr;
    int* d_sr = nullptr;
    rc = tmp_alloc(ctx, &d_sh, S, st);
    if (rc == FEMX_OK) rc = tmp_alloc(ctx, &d_sr, S, st);
    if (rc != FEMX_OK) { free_async(d_sh, st); free_async(d_sr, st); return done(rc); }
    sample_rows<<<nblocks(S, 128), 128, 0, st>>>(d_hash, (long long)nr, S, d_sh, d_sr);
    cudaError_t e1 = cudaMemcpyAsync(hs.data(), d_sh, sizeof(unsigned long long) * S, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaMemcpyAsync(hrow.data(), d_sr, sizeof(int) * S, cudaMemcpyDeviceToHost, st);
    cudaError_t e3 = cudaStreamSynchronize(st);
    free_async(d_sh, st);
    free_async(d_sr, st);
    SC_CUDA(e1); SC_CUDA(e2); SC_CUDA(e3);
  }
  int best = -1, best_cnt = 0;
  for (int i = 0; i < S; ++i) {
    int c = 0;
    for (int j = 0; j < S; ++j) c += hs[j] == hs[i];
    if (c > best_cnt) { best_cnt = c; best = i; }
  }
  if (best < 0 || best_cnt * 4 < S) return done(FEMX_OK);  // no class covers a quarter of the samples
  const int64_t ref = hrow[best];
  int2 ri[2];
  int pp[2];
  SC_CUDA(cudaMemcpyAsync(ri, p->d_rowinfo + ref, sizeof ri, cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaMemcpyAsync(pp, d_pair_ptr + ref, sizeof pp, cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  const int np = pp[1] - pp[0], rlen = ri[1].x - ri[0].x, self = (int)((unsigned)ri[0].y >> 24);
  if (np < 1 || np > FEMX_SPEC_MAX_NP || rlen > (p->nn == 4 ? 16 : FEMX_SPEC_MAX_RLEN)) return done(FEMX_OK);
  std::vector<uint32_t> codes(np);
  std::vector<int32_t> offs(rlen);
  SC_CUDA(cudaMemcpyAsync(codes.data(), d_pair_code + pp[0], sizeof(uint32_t) * np, cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaMemcpyAsync(offs.data(), p->d_col_idx + ri[0].x, sizeof(int32_t) * rlen, cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
  mark_class<<<nblocks(nr, 128), 128, 0, st>>>(d_pair_ptr, d_pair_code, p->d_rowinfo, p->d_col_idx, (int)nr, d_hash,
                                               (int)ref, d_count);
  mark_ref<<<1, 1, 0, st>>>(p->d_rowinfo, (int)ref);
  int cnt = 0;
  SC_CUDA(cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  SC_CUDA(cudaGetLastError());
  rc = build_other_rows(ctx, p, cnt, st);
  if (rc != FEMX_OK) return done(rc);
#undef SC_CUDA
  p->spec_np = np; p->spec_rlen = rlen; p->spec_self = self; p->spec_rows = cnt;
  p->spec_codes = codes;
  for (auto& o : offs) o -= (int32_t)(p->row_begin + ref);  // column (local node id) minus the row's own node
  p->spec_off = offs;
  // the kernel cache key covers what the generated code depends on (codes, sizes) — not the offsets,
  // which are launch parameters: one cubin serves every mesh size with this stencil
  unsigned long long kh = 1469598103934665603ull;
  for (int k = 0; k < np; ++k) kh = (kh ^ codes[k]) * 1099511628211ull;
  char key[96];
  snprintf(key, sizeof key, "%016llx_%d_%d_%d", kh, np, rlen, self);
  p->spec_key = key;
  return done(FEMX_OK);
}


// ------------------------------------------------------------------ lattice ---
// Is the mesh a lattice (femx_internal.h: femx_lattice)?  Hypothesis from the first two cells, the extents
// from the first place where translation invariance breaks, then every element is checked.
struct lat_desc {
  int P, nn, cnx, cny;
  long long sy, sz;
  int node0;
  int off[32];  // node offset of vertex a of element t from the cell's lower corner
};

// The extents in one launch of one block: out[0] = cells per line = the smallest c in [1, nsx) with
// conn[c * P * nn] - conn[0] != c (nsx if there is none); for tetrahedra, if that is a valid line length,
// out[2] = lines per plane = the smallest c in [1, nsy) with conn[c * cnx * P * nn] - conn[0] != c * sy (nsy if none).
// Both searches are bounded by the strides (a line has fewer cells than sy nodes, a plane fewer lines than sz / sy + 1).
__global__ void lattice_extents(const int* __restrict__ conn, long long n_cells, int P, int nn, long long sy, long long sz,
                                int dim, int* __restrict__ out) {
  __shared__ int best;
  const long long nsx = min(n_cells, sy);
  if (threadIdx.x == 0) best = (int)nsx;
  __syncthreads();
  for (long long c = threadIdx.x + 1; c < nsx; c += blockDim.x)
    if ((long long)conn[c * P * nn] - (long long)conn[0] != c) atomicMin(&best, (int)c);
  __syncthreads();
  const long long cnx = best;
  __syncthreads();
  if (threadIdx.x == 0) out[0] = (int)cnx;
  if (dim != 3 || cnx < 2 || n_cells % cnx || cnx + 1 > sy) return;
  const long long n_lines = n_cells / cnx, nsy = min(n_lines, sz / sy + 1);
  if (threadIdx.x == 0) best = (int)nsy;
  __syncthreads();
  for (long long c = threadIdx.x + 1; c < nsy; c += blockDim.x)
    if ((long long)conn[c * cnx * P * nn] - (long long)conn[0] != c * sy) atomicMin(&best, (int)c);
  __syncthreads();
  if (threadIdx.x == 0) out[2] = best;
}

// every element against the template; V = 4: one 16-byte load per tetrahedron (conn 16-byte aligned), else scalar loads.
// 32-bit arithmetic throughout (n_elems * nn < 2^31, node ids < 2^31): a 64-bit division per thread made this kernel
// instruction-bound at a third of the memory rate.
template <int V>
__global__ void lattice_verify(const int* __restrict__ conn, unsigned n_elems, lat_desc d, int* __restrict__ bad) {
  constexpr int U = 4;                                   // elements per thread: all loads in flight before the first check
  const unsigned e0 = blockIdx.x * (blockDim.x * U) + threadIdx.x;
  int4 v[U];
  if (V == 4) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned e = e0 + u * blockDim.x;
      v[u] = e < n_elems ? __ldg(reinterpret_cast<const int4*>(conn) + e) : make_int4(0, 0, 0, 0);
    }
  }
  bool ok = true;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = e0 + u * blockDim.x;
    if (e >= n_elems) break;
    const unsigned c = e / (unsigned)d.P;
    const int t = (int)(e - c * (unsigned)d.P);
    const unsigned line = c / (unsigned)d.cnx, ci = c - line * (unsigned)d.cnx;
    const unsigned ck = line / (unsigned)d.cny, cj = line - ck * (unsigned)d.cny;
    const int base = (int)((unsigned)d.node0 + ci + cj * (unsigned)d.sy + ck * (unsigned)d.sz);
    if (V == 4) {
      ok = ok && v[u].x == base + d.off[t * 4] && v[u].y == base + d.off[t * 4 + 1] && v[u].z == base + d.off[t * 4 + 2] &&
           v[u].w == base + d.off[t * 4 + 3];
    } else {
      for (int a = 0; a < d.nn; ++a) ok = ok && conn[(size_t)e * d.nn + a] == base + d.off[t * d.nn + a];
    }
  }
  if (__syncthreads_or(!ok) && threadIdx.x == 0) atomicAdd(bad, 1);
}

// The same check for tetrahedra with one thread block per piece of a cell LINE: the line's position is decoded once per
// block, the cell of an element by a multiplication (magic = ceil(2^34 / P), exact below 2^29 elements per line), the
// template offsets sit in shared memory (indexed per thread, the kernel-parameter copy would serialise).  The generic
// kernel above spends 3 integer divisions per element and runs at a third of the memory rate.
template <int U>
__global__ void lattice_verify_lines(const int4* __restrict__ conn, lat_desc d, unsigned line_elems, unsigned long long magic,
                                     int* __restrict__ bad) {
  __shared__ int off[32];
  if (threadIdx.x < 32) off[threadIdx.x] = d.off[threadIdx.x];
  const unsigned line = blockIdx.x;
  const unsigned ck = line / (unsigned)d.cny, cj = line - ck * (unsigned)d.cny;
  const int base = (int)((unsigned)d.node0 + cj * (unsigned)d.sy + ck * (unsigned)d.sz);
  const int4* __restrict__ src = conn + (size_t)line * line_elems;
  const unsigned x0 = blockIdx.y * (blockDim.x * U) + threadIdx.x;
  int4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned x = x0 + u * blockDim.x;
    v[u] = x < line_elems ? __ldg(src + x) : make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  bool ok = true;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned x = x0 + u * blockDim.x;
    if (x < line_elems) {
      const unsigned ci = (unsigned)(((unsigned long long)x * magic) >> 34);
      const int t = (int)(x - ci * (unsigned)d.P), b = base + (int)ci;
      ok = ok && v[u].x == b + off[t * 4] && v[u].y == b + off[t * 4 + 1] && v[u].z == b + off[t * 4 + 2] &&
           v[u].w == b + off[t * 4 + 3];
    }
  }
  if (__syncthreads_or(!ok) && threadIdx.x == 0) atomicAdd(bad, 1);
}

// ... and for any element type (triangles; unaligned connectivity): one thread per 32-bit word of a line
template <int U>
__global__ void lattice_verify_words(const int* __restrict__ conn, lat_desc d, unsigned line_words, unsigned long long magic_nn,
                                     unsigned long long magic_p, int* __restrict__ bad) {
  __shared__ int off[32];
  if (threadIdx.x < 32) off[threadIdx.x] = d.off[threadIdx.x];
  const unsigned line = blockIdx.x;
  const unsigned ck = line / (unsigned)d.cny, cj = line - ck * (unsigned)d.cny;
  const int base = (int)((unsigned)d.node0 + cj * (unsigned)d.sy + ck * (unsigned)d.sz);
  const int* __restrict__ src = conn + (size_t)line * line_words;
  const unsigned x0 = blockIdx.y * (blockDim.x * U) + threadIdx.x;
  int v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned x = x0 + u * blockDim.x;
    v[u] = x < line_words ? __ldg(src + x) : 0;
  }
  __syncthreads();
  bool ok = true;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned x = x0 + u * blockDim.x;
    if (x < line_words) {
      const unsigned e = (unsigned)(((unsigned long long)x * magic_nn) >> 34), a = x - e * (unsigned)d.nn;
      const unsigned ci = (unsigned)(((unsigned long long)e * magic_p) >> 34), t = e - ci * (unsigned)d.P;
      ok = ok && v[u] == base + (int)ci + off[t * d.nn + a];
    }
  }
  if (__syncthreads_or(!ok) && threadIdx.x == 0) atomicAdd(bad, 1);
}

template <int U>
void launch_verify_lines(const int* d_conn, const lat_desc& d, long long n_lines, unsigned line_elems, int* d_bad, cudaStream_t st) {
  const unsigned long long magic = ((1ULL << 34) + d.P - 1) / d.P;
  dim3 grid((unsigned)n_lines, (line_elems + 256 * U - 1) / (256 * U));
  lattice_verify_lines<U><<<grid, 256, 0, st>>>(reinterpret_cast<const int4*>(d_conn), d, line_elems, magic, d_bad);
}

// check_class: the lattice is only recorded when the pattern's class rows are exactly its interior nodes
// (general symbolic pass); the lattice-templated pass defines the class itself and passes false
int detect_lattice(femx_ctx* ctx, femx_pattern* p, const int* d_conn, cudaStream_t st, bool check_class) {
  p->lat = femx_lattice();
  const int nn = p->nn, dim = nn - 1;
  const int64_t ne = p->n_elems;
  if (ne < 2 || ne * nn >= (1LL << 31)) return FEMX_OK;
  int head[64];
  const int nh = (int)std::min<int64_t>(ne * nn, 64);
  int* d_tmp = nullptr;
  int rc = tmp_alloc(ctx, &d_tmp, 4, st);
  if (rc != FEMX_OK) return rc;
  auto done = [&](int code) { free_async(d_tmp, st); return code; };
#define LT_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return done(femx_fail(ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)));    \
  } while (0)
  LT_CUDA(cudaMemcpyAsync(head, d_conn, sizeof(int) * nh, cudaMemcpyDeviceToHost, st));
  LT_CUDA(cudaStreamSynchronize(st));
  // period: the smallest P such that cell 1 is cell 0 shifted by one node
  int P = 0;
  for (int q = 1; q <= 8 && 2 * q * nn <= nh && !P; ++q) {
    bool ok = true;
    for (int k = 0; k < q * nn; ++k) ok = ok && head[q * nn + k] - head[k] == 1;
    if (ok) P = q;
  }
  if (!P || ne % P) return done(FEMX_OK);
  // strides from the distinct node offsets of cell 0: {0, 1, sy, sy+1 [, sz, sz+1, sz+sy, sz+sy+1]}
  int b0 = head[0];
  for (int k = 1; k < P * nn; ++k) b0 = std::min(b0, head[k]);
  std::vector<long long> u;
  for (int k = 0; k < P * nn; ++k) u.push_back(head[k] - b0);
  std::sort(u.begin(), u.end());
  u.erase(std::unique(u.begin(), u.end()), u.end());
  if ((int)u.size() != (1 << dim) || u[1] != 1 || u[3] != u[2] + 1 || u[2] < 3) return done(FEMX_OK);
  const long long sy = u[2];
  long long sz = 0;
  if (dim == 3) {
    sz = u[4];
    if (u[5] != sz + 1 || u[6] != sz + sy || u[7] != sz + sy + 1) return done(FEMX_OK);
  }
  femx_lattice L;
  L.dim = dim; L.P = P; L.node0 = b0; L.s[0] = 1; L.s[1] = sy; L.s[2] = sz;
  lat_desc d = {};
  d.P = P; d.nn = nn; d.sy = sy; d.sz = sz; d.node0 = b0;
  for (int t = 0; t < P; ++t)
    for (int a = 0; a < nn; ++a) {
      const long long off = head[t * nn + a] - b0;
      const int c = (int)(std::find(u.begin(), u.end(), off) - u.begin());
      L.corner[t][a] = (unsigned char)c;   // u is sorted: index = dx | dy << 1 | dz << 2
      d.off[t * nn + a] = (int)off;
    }
  // extents
  const long long n_cells = ne / P;
  int h_tmp[4] = {0, 0, 0, 0};
  LT_CUDA(cudaMemsetAsync(d_tmp, 0, sizeof(int) * 4, st));
  lattice_extents<<<1, 1024, 0, st>>>(d_conn, n_cells, P, nn, sy, sz, dim, d_tmp);
  LT_CUDA(cudaMemcpyAsync(h_tmp, d_tmp, sizeof h_tmp, cudaMemcpyDeviceToHost, st));
  LT_CUDA(cudaStreamSynchronize(st));
  const long long cnx = h_tmp[0];   // (== the search bound when no break was found below it: then the mesh is one line)
  if (cnx < 2 || n_cells % cnx || cnx + 1 > sy) return done(FEMX_OK);
  const long long n_lines = n_cells / cnx;
  long long cny = n_lines, cnz = 1;
  if (dim == 3) {
    const long long nsy = std::min<long long>(n_lines, sz / sy + 1);
    cny = h_tmp[2] == nsy && nsy < n_lines ? 0 : h_tmp[2];   // no break below the bound although lines remain: not a lattice
    if (cny < 1 || n_lines % cny || cny * sy + cnx + 1 > sz) return done(FEMX_OK);
    cnz = n_lines / cny;
  }
  L.cn[0] = (int)cnx; L.cn[1] = (int)cny; L.cn[2] = (int)cnz;
  if (b0 + cnx + cny * sy + cnz * sz >= p->n_nodes) return done(FEMX_OK);
  d.cnx = (int)cnx; d.cny = (int)cny;
  const long long line_elems = cnx * P;
  if (nn == 4 && (uintptr_t)d_conn % 16 == 0 && line_elems >= 256 && line_elems < (1LL << 29) && P * nn <= 32 &&
      (line_elems + 2047) / 2048 <= 65535) {
    // elements per thread: the most that leaves the fewest idle threads in a line's last block
    int U = 1;
    long long waste = -1;
    for (int u = 1; u <= 8; ++u) {
      const long long w = (line_elems + 256 * u - 1) / (256 * u) * (256 * u) - line_elems;
      if (waste < 0 || w <= waste) { waste = w; U = u; }
    }
    switch (U) {
      case 1: launch_verify_lines<1>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 2: launch_verify_lines<2>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 3: launch_verify_lines<3>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 4: launch_verify_lines<4>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 5: launch_verify_lines<5>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 6: launch_verify_lines<6>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      case 7: launch_verify_lines<7>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
      default: launch_verify_lines<8>(d_conn, d, n_lines, (unsigned)line_elems, d_tmp + 1, st); break;
    }
  } else if (line_elems * nn >= 256 && line_elems * nn < (1LL << 29) && P * nn <= 32 && (line_elems * nn + 2047) / 2048 <= 65535) {
    const unsigned words = (unsigned)(line_elems * nn);
    const unsigned long long m_nn = ((1ULL << 34) + nn - 1) / nn, m_p = ((1ULL << 34) + P - 1) / P;
    dim3 grid((unsigned)n_lines, (words + 2047) / 2048);
    lattice_verify_words<8><<<grid, 256, 0, st>>>(d_conn, d, words, m_nn, m_p, d_tmp + 1);
  } else if (nn == 4 && (uintptr_t)d_conn % 16 == 0) lattice_verify<4><<<nblocks(ne, 1024), 256, 0, st>>>(d_conn, (unsigned)ne, d, d_tmp + 1);
  else lattice_verify<1><<<nblocks(ne, 1024), 256, 0, st>>>(d_conn, (unsigned)ne, d, d_tmp + 1);
  LT_CUDA(cudaMemcpyAsync(h_tmp, d_tmp + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  LT_CUDA(cudaStreamSynchronize(st));
  LT_CUDA(cudaGetLastError());
  if (h_tmp[0] != 0) return done(FEMX_OK);
  // the class rows must be exactly the lattice-interior nodes among the owned rows (every one of them is then
  // written by the lattice pass, every other row by the row list)
  long long interior = 0;
  for (long long k = 1; k < (dim == 3 ? cnz : 2); ++k)
    for (long long j = 1; j < cny; ++j) {
      const long long n0 = b0 + 1 + j * sy + (dim == 3 ? k * sz : 0), n1 = n0 + cnx - 1;  // nodes [n0, n1) of this line
      const long long lo = std::max<long long>(n0, p->row_begin), hi = std::min<long long>(n1, p->row_end);
      if (hi > lo) interior += hi - lo;
    }
  if (check_class && interior != p->spec_rows) return done(FEMX_OK);
  for (int t = 0; t < P; ++t)  // (degenerate template elements are left to the general pass, which reports them)
    for (int a = 0; a < nn; ++a)
      for (int b = a + 1; b < nn; ++b)
        if (L.corner[t][a] == L.corner[t][b]) return done(FEMX_OK);
  L.ok = true;
  p->lat = L;
  p->lat_rows = interior;
  return done(FEMX_OK);
#undef LT_CUDA
}


// --------------------------------------------- lattice-templated symbolic pass ---
// On a lattice mesh (every element verified against the cell template: detect_lattice) the rows of the pattern are
// translates of at most 3^dim templates — one per (low face / interior / high face) position along each axis.  The
// templates are produced on the host by running the SAME row rules (ascending incidences, sorted duplicate-free
// columns, position codes with first-touch flags) on a replica of at most 2 cells per axis; the device then writes
// rowinfo, columns and the SELL scatter map of every row from its template at memory speed, instead of the
// histogram / bucket / per-row sort-merge pipeline.  Output identical to the general pass, bit for bit (tests).
#define LT_MAX_NP 32
#define LT_MAX_RLEN 27
struct lat_row_tmpl {
  int np, rlen, self;
  signed char col[LT_MAX_RLEN][3];   // column displacement (dx, dy, dz), ascending node id
  signed char cell[LT_MAX_NP][3];    // cell of incidence k relative to the node: (sx, sy, sz) in {-1, 0}
  unsigned char t[LT_MAX_NP], li[LT_MAX_NP];
  unsigned code[LT_MAX_NP];
  int off[LT_MAX_RLEN];              // col[k] as a node offset of THIS lattice (filled by build_from_lattice)
};
struct lat_geom {
  int dim, P, nn;
  int cn[3];
  long long s[3];
  long long node0;
  int row_begin, n_rows;
  int dom;  // dominant class (flagged FEMX_ROW_SPEC), -1 = none
  int w32;  // the prefix sums of the whole lattice fit 32 bits (device: lat_prefix32)
};

// Prefix sums of a per-class weight over the lattice's nodes in node order, in closed form: whole planes + whole lines
// + the started line.  With w = row length this IS row_ptr, with w = [class == dominant] the rank of a row among the
// rows outside the class — no length array, no device-wide scan.
struct lat_pref {
  long long line[9], plane[3], total, base;   // base: the prefix at row_begin (subtracted by the callers)
  int w[27];
};

__host__ __device__ inline int lat_axis_class(long long pos, int cn) { return pos == 0 ? 0 : (pos == cn ? 2 : 1); }
// number of positions below pos (0 <= pos <= cn + 1) whose class along the axis is cls
__host__ __device__ inline long long lat_below(long long pos, int cn, int cls) {
  if (cls == 0) return pos > 0;
  if (cls == 2) return pos > cn;
  const long long m = pos - 1 < 0 ? 0 : pos - 1;
  return m > cn - 1 ? cn - 1 : m;
}

// node offset q = node - node0 -> clamped lattice position; returns the class (-1: not a lattice node)
// (node ids and strides are below 2^31: 32-bit divisions on the device)
__host__ __device__ inline int lat_locate(const lat_geom& g, long long q64, long long* ijk) {
  if (q64 < 0) { ijk[0] = ijk[1] = ijk[2] = 0; return -1; }
  unsigned q = (unsigned)q64, k = 0;
  if (g.dim == 3) { k = q / (unsigned)g.s[2]; q -= k * (unsigned)g.s[2]; }
  unsigned j = q / (unsigned)g.s[1], i = q - j * (unsigned)g.s[1];
  bool in = true;
  if (g.dim == 3 && k > (unsigned)g.cn[2]) { k = g.cn[2] + 1; j = 0; i = 0; in = false; }
  if (j > (unsigned)g.cn[1]) { j = g.cn[1] + 1; i = 0; in = false; }
  if (i > (unsigned)g.cn[0]) { i = g.cn[0] + 1; in = false; }
  ijk[0] = i; ijk[1] = j; ijk[2] = k;
  if (!in) return -1;
  return lat_axis_class(i, g.cn[0]) + 3 * lat_axis_class(j, g.cn[1]) + (g.dim == 3 ? 9 * lat_axis_class(k, g.cn[2]) : 0);
}

// sum of w over the lattice nodes in front of the (clamped) position ijk
__host__ __device__ inline long long lat_prefix(const lat_geom& g, const lat_pref& P, const long long* ijk) {
  const long long i = ijk[0], j = ijk[1], k = ijk[2];
  if (g.dim == 3 && k > g.cn[2]) return P.total;
  const int ck = g.dim == 3 ? lat_axis_class(k, g.cn[2]) : 0;
  long long s = 0;
  if (g.dim == 3)
    for (int c = 0; c < 3; ++c) s += lat_below(k, g.cn[2], c) * P.plane[c];
  for (int c = 0; c < 3; ++c) s += lat_below(j, g.cn[1], c) * P.line[c + 3 * ck];
  if (j <= g.cn[1]) {
    const int cj = lat_axis_class(j, g.cn[1]);
    for (int c = 0; c < 3; ++c) s += lat_below(i, g.cn[0], c) * P.w[c + 3 * cj + 9 * ck];
  }
  return s;
}

// the same sum in 32-bit arithmetic (the caller has checked that P.total fits): a third of the integer multiplies
__device__ __forceinline__ int lat_below32(int pos, int cn, int cls) {
  return cls == 0 ? (pos > 0) : (cls == 2 ? (pos > cn) : min(max(pos - 1, 0), cn - 1));
}
__device__ __forceinline__ int lat_prefix32(const lat_geom& g, const lat_pref& P, const long long* ijk) {
  const int i = (int)ijk[0], j = (int)ijk[1], k = (int)ijk[2];
  if (g.dim == 3 && k > g.cn[2]) return (int)P.total;
  const int ck = g.dim == 3 ? lat_axis_class(k, g.cn[2]) : 0;
  int s = 0;
  if (g.dim == 3) {
#pragma unroll
    for (int c = 0; c < 3; ++c) s += lat_below32(k, g.cn[2], c) * (int)P.plane[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) s += lat_below32(j, g.cn[1], c) * (int)P.line[c + 3 * ck];
  if (j <= g.cn[1]) {
    const int cj = lat_axis_class(j, g.cn[1]);
#pragma unroll
    for (int c = 0; c < 3; ++c) s += lat_below32(i, g.cn[0], c) * P.w[c + 3 * cj + 9 * ck];
  }
  return s;
}

lat_pref lat_pref_make(const lat_geom& g, const int* w) {
  lat_pref P = {};
  auto cnt = [&](int d, int c) -> long long { return c == 1 ? g.cn[d] - 1 : 1; };
  for (int c = 0; c < 27; ++c) P.w[c] = w[c];
  for (int ck = 0; ck < 3; ++ck) {
    for (int cj = 0; cj < 3; ++cj) {
      long long l = 0;
      for (int c = 0; c < 3; ++c) l += cnt(0, c) * w[c + 3 * cj + 9 * ck];
      P.line[cj + 3 * ck] = l;
      P.plane[ck] += cnt(1, cj) * l;
    }
  }
  P.total = P.plane[0];
  if (g.dim == 3) P.total = cnt(2, 0) * P.plane[0] + cnt(2, 1) * P.plane[1] + cnt(2, 2) * P.plane[2];
  long long ijk[3];
  lat_locate(g, (long long)g.row_begin - g.node0, ijk);
  P.base = lat_prefix(g, P, ijk);
  return P;
}

// slice s of the padded scatter map: 32 x the largest incidence count among its 32 rows.  One thread per slice: a slice
// that lies inside one lattice line (all but cn[0] / 32 of them) has at most three classes, known from its two ends.
__global__ void lat_slice_sizes_k(lat_geom g, const lat_row_tmpl* __restrict__ T, int n_slices, int* __restrict__ size) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices) return;
  const int r0 = s * 32, r1 = min(r0 + 32, g.n_rows);
  long long ijk[3];
  const int c0 = lat_locate(g, (long long)g.row_begin + r0 - g.node0, ijk);
  int np = 0;
  if (c0 >= 0 && ijk[0] + (r1 - r0 - 1) <= g.cn[0]) {
    const int i0 = (int)ijk[0], i1 = i0 + (r1 - r0 - 1);
    const int base = c0 - lat_axis_class(i0, g.cn[0]);   // 3 cj + 9 ck
    if (i0 == 0) np = max(np, T[base].np);
    if (i1 == g.cn[0]) np = max(np, T[base + 2].np);
    if (max(i0, 1) <= min(i1, g.cn[0] - 1)) np = max(np, T[base + 1].np);
  } else {
    for (int r = r0; r < r1; ++r) {
      const int c = lat_locate(g, (long long)g.row_begin + r - g.node0, ijk);
      if (c >= 0) np = max(np, T[c].np);
    }
  }
  size[s] = np * 32;
}

__global__ void lat_row_fill_k(lat_geom g, lat_pref PL, lat_pref PD, const lat_row_tmpl* __restrict__ T,
                               const int* __restrict__ slice_ptr, int2* __restrict__ rowinfo, int* __restrict__ col_idx,
                               unsigned* __restrict__ sell_code, int* __restrict__ sell_elem, int* __restrict__ other_rows,
                               int map_only) {
  // map_only = 0: rowinfo + columns of every row, the list of rows outside the dominant class and their scatter map
  // (the class rows' map is read by no kernel of the default path and is written on demand: femx_pattern_complete_map);
  // map_only = 1: the scatter map of the class rows
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int rc = min(r, g.n_rows);                                // (lanes behind the last row: zero-length rows at nnz)
  const long long node = (long long)g.row_begin + rc;
  long long ijk[3];
  int c = lat_locate(g, node - g.node0, ijk);
  if (r >= g.n_rows) c = -1;
  if (!map_only) {
    // rowinfo: one 8-byte store per row (row_ptr in closed form); columns: the 32 rows of a warp own ONE contiguous piece
    // of col_idx, written 128 bytes per instruction (entry p belongs to the last row j of the warp with row_ptr[j] <= p)
    const int rp = g.w32 ? lat_prefix32(g, PL, ijk) - (int)PL.base : (int)(lat_prefix(g, PL, ijk) - PL.base);
    if (r <= g.n_rows)
      rowinfo[r] = make_int2(rp, c < 0 ? 0 : (T[c].np | (c == g.dom ? FEMX_ROW_SPEC : 0) | (T[c].self << 24)));
    if (r < g.n_rows && c != g.dom && other_rows)
      other_rows[r - (g.w32 ? lat_prefix32(g, PD, ijk) - (int)PD.base : (int)(lat_prefix(g, PD, ijk) - PD.base))] = r;
    const int p_begin = __shfl_sync(0xffffffffu, rp, 0);
    const int p_end = __shfl_sync(0xffffffffu, rp + (c < 0 ? 0 : T[c].rlen), 31);
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const int row0 = g.row_begin + (r - lane);
    if (c0 >= 0 && T[c0].rlen > 1 && __all_sync(0xffffffffu, c == c0)) {
      // 32 rows of one class (the usual warp): position -> (row, entry) by one multiplication
      const int len = T[c0].rlen;
      const unsigned magic = 0xffffffffu / (unsigned)len + 1u;     // floor(idx / len) = umulhi(idx, magic), idx < 32 * 27
      const int* __restrict__ off = T[c0].off;
      for (int idx = lane; idx < 32 * len; idx += 32) {
        const int j = (int)__umulhi((unsigned)idx, magic);
        col_idx[p_begin + idx] = row0 + j + off[idx - j * len];
      }
    } else {
      for (int base = p_begin; base < p_end; base += 32) {
        const int p = base + lane;
        int j = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          const int v = __shfl_sync(0xffffffffu, rp, j + s);
          if (v <= p) j += s;
        }
        const int rpj = __shfl_sync(0xffffffffu, rp, j), cj = __shfl_sync(0xffffffffu, c, j);
        if (p < p_end && cj >= 0) col_idx[p] = row0 + j + T[cj].off[p - rpj];
      }
    }
    if (c < 0 || c == g.dom) return;
  } else if (c < 0 || c != g.dom) {
    return;
  }
  const lat_row_tmpl& t = T[c];
  const int sp = slice_ptr[r >> 5] + (r & 31);
  for (int k = 0; k < t.np; ++k) {
    const long long cell = (ijk[0] + t.cell[k][0]) + (long long)g.cn[0] * ((ijk[1] + t.cell[k][1]) + (long long)g.cn[1] * (ijk[2] + t.cell[k][2]));
    sell_elem[sp + k * 32] = (int)((cell * g.P + t.t[k]) * g.nn + t.li[k]);
    sell_code[sp + k * 32] = t.code[k];
  }
}

// the row rules of row_fill / build_row on a replica mesh, on the host
void lattice_templates(const femx_lattice& L, int nn, std::vector<lat_row_tmpl>* out) {
  const int dim = L.dim;
  int m[3] = {std::min(L.cn[0], 2), std::min(L.cn[1], 2), dim == 3 ? std::min(L.cn[2], 2) : 1};
  const int rs[3] = {1, m[0] + 1, (m[0] + 1) * (m[1] + 1)};
  auto node_of = [&](int i, int j, int k) { return i + j * rs[1] + (dim == 3 ? k * rs[2] : 0); };
  std::vector<std::vector<int>> conn;  // replica elements, cell-major like the lattice
  std::vector<std::array<int, 4>> ecell;
  for (int ck = 0; ck < (dim == 3 ? m[2] : 1); ++ck)
    for (int cj = 0; cj < m[1]; ++cj)
      for (int ci = 0; ci < m[0]; ++ci)
        for (int t = 0; t < L.P; ++t) {
          std::vector<int> e(nn);
          for (int a = 0; a < nn; ++a) {
            const int c = L.corner[t][a];
            e[a] = node_of(ci + (c & 1), cj + ((c >> 1) & 1), ck + ((c >> 2) & 1));
          }
          conn.push_back(e);
          ecell.push_back({ci, cj, ck, t});
        }
  out->assign(27, lat_row_tmpl());
  for (int az = 0; az < (dim == 3 ? 3 : 1); ++az)
    for (int ay = 0; ay < 3; ++ay)
      for (int ax = 0; ax < 3; ++ax) {
        lat_row_tmpl& T = (*out)[ax + 3 * ay + 9 * az];
        memset(&T, 0, sizeof T);
        // representative position of the class along each axis; classes that do not occur stay empty
        int pos[3];
        const int a3[3] = {ax, ay, az};
        bool exists = true;
        for (int d = 0; d < 3; ++d) {
          const int md = d < dim ? m[d] : 0;
          if (a3[d] == 0) pos[d] = 0;
          else if (a3[d] == 2) pos[d] = md;
          else { pos[d] = 1; if (md < 2) exists = false; }
          if (d >= dim && a3[d] != 0) exists = false;
        }
        if (!exists) continue;
        const int n = node_of(pos[0], pos[1], pos[2]);
        std::vector<int> inc, list;  // incidences e*nn + li ascending; sorted duplicate-free columns
        for (size_t e = 0; e < conn.size(); ++e)
          for (int a = 0; a < nn; ++a)
            if (conn[e][a] == n) inc.push_back((int)e * nn + a);
        for (int pe : inc)
          for (int a = 0; a < nn; ++a) list.push_back(conn[pe / nn][a]);
        std::sort(list.begin(), list.end());
        list.erase(std::unique(list.begin(), list.end()), list.end());
        T.np = (int)inc.size();
        T.rlen = (int)list.size();
        T.self = (int)(std::lower_bound(list.begin(), list.end(), n) - list.begin());
        for (int k = 0; k < T.rlen; ++k) {
          const int q = list[k];
          const int qi = q % rs[1], qj = (q / rs[1]) % (m[1] + 1), qk = dim == 3 ? q / rs[2] : 0;
          T.col[k][0] = (signed char)(qi - pos[0]); T.col[k][1] = (signed char)(qj - pos[1]); T.col[k][2] = (signed char)(qk - pos[2]);
        }
        std::vector<char> seen(list.size(), 0);
        for (int k = 0; k < T.np; ++k) {
          const int e = inc[k] / nn, li = inc[k] % nn;
          unsigned code = (unsigned)li << 28;
          for (int a = 0; a < nn; ++a) {
            if (a == li) continue;
            const int p0 = (int)(std::lower_bound(list.begin(), list.end(), conn[e][a]) - list.begin());
            const int j = nn == 4 ? (a ^ li) - 1 : (a - li - 1 + 3) % 3;
            code |= (unsigned)p0 << (7 * j);
            if (!seen[p0]) { code |= 1u << (21 + j); seen[p0] = 1; }
          }
          T.code[k] = code;
          T.t[k] = (unsigned char)ecell[e][3];
          T.li[k] = (unsigned char)li;
          T.cell[k][0] = (signed char)(ecell[e][0] - pos[0]); T.cell[k][1] = (signed char)(ecell[e][1] - pos[1]);
          T.cell[k][2] = (signed char)(ecell[e][2] - pos[2]);
        }
      }
}


int build_from_lattice(femx_ctx* ctx, femx_pattern* p, cudaStream_t st) {
  const femx_lattice& L = p->lat;
  const int nn = p->nn, dim = L.dim;
  const int64_t nr = p->n_rows;
  std::vector<lat_row_tmpl> T;
  lattice_templates(L, nn, &T);
  // dominant class: the one with the most nodes in the lattice (the interior, unless an axis has a single cell)
  int dom = -1;
  long long best = 0;
  for (int c = 0; c < 27; ++c) {
    if (T[c].np == 0) continue;
    const int a3[3] = {c % 3, (c / 3) % 3, c / 9};
    long long cnt = 1;
    for (int d = 0; d < dim; ++d) cnt *= a3[d] == 1 ? std::max(L.cn[d] - 1, 0) : 1;
    if (cnt > best) { best = cnt; dom = c; }
  }
  const bool want_class = p->nd == 1 && ctx->knobs.spec != 0 && dom >= 0 && T[dom].np <= FEMX_SPEC_MAX_NP &&
                          T[dom].rlen <= (nn == 4 ? 16 : FEMX_SPEC_MAX_RLEN);
  lat_geom g = {};
  g.dim = dim; g.P = L.P; g.nn = nn;
  for (int d = 0; d < 3; ++d) { g.cn[d] = L.cn[d]; g.s[d] = L.s[d]; }
  g.node0 = L.node0; g.row_begin = (int)p->row_begin; g.n_rows = (int)nr;
  g.dom = want_class ? dom : -1;
  // everything the general pass gets from scans and reductions follows from the class counts of the row range
  long long cnt[27], nnz = 0, tot_pairs = 0, n_dom = 0;
  int max_row = 0, max_other = 0;
  auto count_classes = [&]() {
    long long e_ijk[3];
    lat_locate(g, (long long)g.row_begin + nr - g.node0, e_ijk);
    for (int c = 0; c < 27; ++c) {
      int w[27] = {};
      w[c] = 1;
      const lat_pref P = lat_pref_make(g, w);
      cnt[c] = T[c].np ? lat_prefix(g, P, e_ijk) - P.base : 0;
    }
  };
  count_classes();
  if (want_class && cnt[dom] * 4 < nr) g.dom = -1;   // (the general pass wants a quarter of its sample rows)
  for (int c = 0; c < 27; ++c) {
    if (cnt[c] <= 0) continue;
    nnz += cnt[c] * T[c].rlen;
    tot_pairs += cnt[c] * T[c].np;
    max_row = std::max(max_row, T[c].rlen);
    if (c == g.dom) n_dom = cnt[c];
    else max_other = std::max(max_other, T[c].rlen);
  }
  if (nnz >= (1LL << 31) - 1 || tot_pairs >= (1LL << 31) - 1)
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: %lld node-level nonzeros / %lld incidences exceed 32-bit offsets", nnz, tot_pairs);
  int w_len[27], w_dom[27];
  for (int c = 0; c < 27; ++c) { w_len[c] = T[c].rlen; w_dom[c] = c == g.dom; }
  const lat_pref PL = lat_pref_make(g, w_len), PD = lat_pref_make(g, w_dom);
  g.w32 = PL.total < (1LL << 31) - 1 && PD.total < (1LL << 31) - 1;
  for (auto& t : T)
    for (int k = 0; k < t.rlen; ++k) t.off[k] = (int)(t.col[k][0] + t.col[k][1] * L.s[1] + t.col[k][2] * L.s[2]);

  lat_row_tmpl* d_T = nullptr;
  int *d_ssize = nullptr, *d_flags = nullptr;
  long long* d_tot = nullptr;
  int rc = FEMX_OK;
  auto cleanup = [&]() { free_async(d_T, st); free_async(d_ssize, st); free_async(d_flags, st); free_async(d_tot, st); };
#define LB_TRY(x) do { rc = (x); if (rc != FEMX_OK) { cleanup(); return rc; } } while (0)
#define LB_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) { cleanup(); return femx_fail(ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } \
  } while (0)
  const int64_t n_slices = (nr + 31) / 32;
  LB_TRY(tmp_alloc(ctx, &d_T, 27, st));
  LB_TRY(tmp_alloc(ctx, &d_flags, 4, st));
  LB_TRY(tmp_alloc(ctx, &d_ssize, n_slices + 1, st));
  LB_TRY(tmp_alloc(ctx, &d_tot, 1, st));
  LB_CUDA(cudaMemcpyAsync(d_T, T.data(), sizeof(lat_row_tmpl) * 27, cudaMemcpyHostToDevice, st));
  LB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, st));
  p->max_row = max_row;
  p->n_pairs = (int64_t)tot_pairs;
  p->nnz_node = nnz;
  const long long n_other = g.dom >= 0 && n_dom > 0 ? nr - n_dom : 0;
  LB_TRY(dev_alloc(ctx, &p->d_slice_ptr, n_slices + 1, &p->bytes, st));
  // the padded scatter map is allocated by its bound (every slice as long as the longest row's incidence list: on a
  // lattice that is what all but the boundary slices are); the exact size arrives with the final synchronisation
  int max_np = 0;
  for (int c = 0; c < 27; ++c)
    if (cnt[c] > 0) max_np = std::max(max_np, T[c].np);
  long long n_sell = n_slices * 32 * max_np;
  if (n_sell >= (1LL << 31) - 1)
    return (cleanup(), femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: padded scatter map (%lld) exceeds 32-bit offsets", n_sell));
  if (n_slices > 0) lat_slice_sizes_k<<<nblocks(n_slices, 256), 256, 0, st>>>(g, d_T, (int)n_slices, d_ssize);
  LB_TRY(exclusive_scan(ctx, d_ssize, n_slices, p->d_slice_ptr, nullptr, st, d_tot));
  LB_TRY(dev_alloc(ctx, &p->d_rowinfo, nr + 1, &p->bytes, st));
  LB_TRY(dev_alloc(ctx, &p->d_col_idx, nnz + 8, &p->bytes, st));
  LB_CUDA(cudaMemsetAsync(p->d_col_idx + nnz, 0, sizeof(int) * 8, st));
  LB_TRY(dev_alloc(ctx, &p->d_sell_code, n_sell, &p->bytes, st));
  LB_TRY(dev_alloc(ctx, &p->d_sell_elem, n_sell, &p->bytes, st));
  if (n_other > 0) LB_TRY(dev_alloc(ctx, &p->d_other_rows, n_other, &p->bytes, st));
  // (no zero fill: the padding of a slice is copied by the generic pass's bulk loads but never used)
  lat_row_fill_k<<<nblocks(nr + 1, 128), 128, 0, st>>>(g, PL, PD, d_T, p->d_slice_ptr, p->d_rowinfo, p->d_col_idx,
                                                       p->d_sell_code, p->d_sell_elem, n_other > 0 ? p->d_other_rows : nullptr, 0);
  if (nr > 0) {
    int64_t ntiles = (nr + p->tile_nodes - 1) / p->tile_nodes;
    tile_max<<<nblocks(ntiles, 128), 128, 0, st>>>(p->d_rowinfo, p->d_slice_ptr, (int)nr, p->tile_nodes, d_flags + 2);
  }
  // the scatter map of the class rows is completed on demand (femx_pattern_complete_map); keep what that needs
  p->map_complete = !(g.dom >= 0 && n_dom > 0);
  if (!p->map_complete) {
    rc = dev_alloc(ctx, (lat_row_tmpl**)&p->d_lat_tmpl, 27, &p->bytes, st);
    if (rc != FEMX_OK) { cleanup(); return rc; }
    LB_CUDA(cudaMemcpyAsync(p->d_lat_tmpl, d_T, sizeof(lat_row_tmpl) * 27, cudaMemcpyDeviceToDevice, st));   // (stream-ordered before d_T is freed)
    p->lat_dom = g.dom;
  }
  int h_flags[4];
  long long sell_exact = 0;
  LB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st));
  LB_CUDA(cudaMemcpyAsync(&sell_exact, d_tot, sizeof sell_exact, cudaMemcpyDeviceToHost, st));
  LB_CUDA(cudaStreamSynchronize(st));
  LB_CUDA(cudaGetLastError());
  if (sell_exact > n_sell) { cleanup(); return femx_fail(ctx, FEMX_ERR_CUDA, "lattice pass: scatter map %lld above its bound %lld", sell_exact, n_sell); }
  p->n_sell = sell_exact;
  p->max_tile_nnz = h_flags[2];
  p->max_tile_codes = h_flags[3];
  if (g.dom >= 0 && n_dom > 0) {
    const lat_row_tmpl& D = T[g.dom];
    p->spec_np = D.np; p->spec_rlen = D.rlen; p->spec_self = D.self; p->spec_rows = n_dom;
    p->spec_codes.assign(D.code, D.code + D.np);
    p->spec_off.resize(D.rlen);
    for (int k = 0; k < D.rlen; ++k) p->spec_off[k] = (int32_t)(D.col[k][0] + D.col[k][1] * L.s[1] + D.col[k][2] * L.s[2]);
    unsigned long long kh = 1469598103934665603ull;
    for (int k = 0; k < D.np; ++k) kh = (kh ^ D.code[k]) * 1099511628211ull;
    char key[96];
    snprintf(key, sizeof key, "%016llx_%d_%d_%d", kh, D.np, D.rlen, D.self);
    p->spec_key = key;
    p->n_other = n_other;
    p->max_row_other = max_other;
    // the element-once numeric pass takes the class rows only when they are the lattice-interior nodes
    const int interior = dim == 3 ? 13 : 4;
    p->lat_rows = g.dom == interior ? n_dom : 0;
  } else {
    p->lat_rows = 0;
  }
  cleanup();
  return FEMX_OK;
#undef LB_TRY
#undef LB_CUDA
}

}  // namespace

// Patterns built by the lattice-templated pass leave out the scatter map of the class rows (3.1 of the 4.4 GB on cfg3):
// the stencil-class and lattice numeric passes never read it.  The passes that do — the generic incidence loop over
// all rows (unstructured fallback switched on by option, element-expanded coordinates, custom strings with contraction)
// and the load vector — call this first; it runs once per pattern.
int femx_pattern_complete_map(const femx_pattern* cp, void* stream) {
  femx_pattern* p = const_cast<femx_pattern*>(cp);
  if (!p || p->map_complete) return FEMX_OK;
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  const femx_lattice& L = p->lat;
  lat_geom g = {};
  g.dim = L.dim; g.P = L.P; g.nn = p->nn;
  for (int d = 0; d < 3; ++d) { g.cn[d] = L.cn[d]; g.s[d] = L.s[d]; }
  g.node0 = L.node0; g.row_begin = (int)p->row_begin; g.n_rows = (int)p->n_rows;
  g.dom = p->lat_dom;
  const lat_pref none = {};
  lat_row_fill_k<<<nblocks(p->n_rows + 1, 128), 128, 0, (cudaStream_t)stream>>>(g, none, none, (const lat_row_tmpl*)p->d_lat_tmpl, p->d_slice_ptr,
                                                                                  p->d_rowinfo, p->d_col_idx, p->d_sell_code, p->d_sell_elem, nullptr, 1);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  p->map_complete = true;
  return FEMX_OK;
}

namespace {

}  // namespace

extern "C" {

int femx_pattern_build(femx_ctx* ctx, int nn, int nd, int64_t n_nodes, int64_t n_elems,
                       const int32_t* d_conn, int64_t row_begin, int64_t row_end, int64_t col_base,
                       void* stream, femx_pattern** out) {
  if (!ctx || !out) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_pattern_build: NULL argument");
  *out = nullptr;
  if (nn != 3 && nn != 4)
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: nn=%d (only 3 or 4)", nn);
  if (nd < 1 || nd > 3) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_pattern_build: nd=%d", nd);
  if (n_nodes < 0 || n_elems < 0 || row_begin < 0 || row_end < row_begin || row_end > n_nodes)
    return femx_fail(ctx, FEMX_ERR_INVALID,
                     "femx_pattern_build: bad sizes (n_nodes=%lld n_elems=%lld rows=[%lld,%lld))",
                     (long long)n_nodes, (long long)n_elems, (long long)row_begin, (long long)row_end);
  if (n_elems > 0 && !d_conn) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_pattern_build: d_conn is NULL");
  if (n_nodes >= (1LL << 31) - 1 || n_elems * nn >= (1LL << 31) - 1 ||
      (col_base + n_nodes) * nd >= (1LL << 31) - 1)
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED,
                     "femx_pattern_build: sizes exceed 32-bit indexing on one device; shard the mesh");
  cudaStream_t st = (cudaStream_t)stream;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));

  femx_pattern* p = new femx_pattern();
  p->ctx = ctx; p->nn = nn; p->nd = nd; p->n_nodes = n_nodes; p->n_elems = n_elems;
  p->row_begin = row_begin; p->row_end = row_end; p->col_base = col_base;
  p->n_rows = row_end - row_begin;
  p->tile_nodes = femx_tile_nodes_for(nd, ctx->knobs);
  const int64_t nr = p->n_rows, total = n_elems * nn;
  // Lattice meshes (every element a translate of the cell template — verified): rows from templates, at memory speed
  if (ctx->knobs.lattice_pattern != 0 && nr > 0 && n_elems > 0) {
    int rc0 = detect_lattice(ctx, p, d_conn, st, false);
    if (rc0 != FEMX_OK) { femx_pattern_destroy(p); return rc0; }
    if (p->lat.ok) {
      rc0 = build_from_lattice(ctx, p, st);
      if (rc0 != FEMX_OK) { femx_pattern_destroy(p); return rc0; }
      if (ctx->knobs.lattice == 0) p->lat_rows = 0;
      *out = p;
      return FEMX_OK;
    }
  }
  int *d_cnt = nullptr, *d_pair_ptr = nullptr, *d_row_ptr = nullptr, *d_flags = nullptr;
  int* d_pair_elem = nullptr;       // CSR-style incidence lists (temporary)
  unsigned* d_pair_code = nullptr;
  int st_code = FEMX_OK;
  auto cleanup = [&]() {
    free_async(d_cnt, st); free_async(d_pair_ptr, st); free_async(d_row_ptr, st); free_async(d_flags, st);
    free_async(d_pair_elem, st); free_async(d_pair_code, st);
  };
#define PB_TRY(x)                                   \
  do {                                              \
    st_code = (x);                                  \
    if (st_code != FEMX_OK) { cleanup(); femx_pattern_destroy(p); return st_code; } \
  } while (0)
#define PB_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      cleanup(); femx_pattern_destroy(p);                                              \
      return femx_fail(ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    }                                                                                  \
  } while (0)

  PB_TRY(tmp_alloc(ctx, &d_cnt, nr + 1, st));
  PB_TRY(tmp_alloc(ctx, &d_pair_ptr, nr + 1, st));
  PB_TRY(tmp_alloc(ctx, &d_row_ptr, nr + 1, st));
  PB_TRY(tmp_alloc(ctx, &d_flags, 4, st));  // [0] err, [1] max_row, [2] max tile nnz
  PB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (nr + 1), st));
  PB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, st));

  // 1-2: incidence histogram + scan
  if (total > 0)
    count_pairs<<<nblocks(total, 256), 256, 0, st>>>(d_conn, total, nn, (int)row_begin, (int)row_end,
                                                     (int)n_nodes, d_cnt, d_flags);
  long long n_pairs = 0;
  PB_TRY(exclusive_scan(ctx, d_cnt, nr, d_pair_ptr, &n_pairs, st));
  int h_flags[4];
  PB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  if (h_flags[0] & 1) {
    cleanup(); femx_pattern_destroy(p);
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_pattern_build: connectivity holds a node id outside [0,%lld)",
                     (long long)n_nodes);
  }
  if (h_flags[0] & 4) {
    cleanup(); femx_pattern_destroy(p);
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_pattern_build: an element lists the same node twice (degenerate element)");
  }
  p->n_pairs = n_pairs;
  PB_TRY(tmp_alloc(ctx, &d_pair_elem, n_pairs, st));
  PB_TRY(tmp_alloc(ctx, &d_pair_code, n_pairs, st));
  PB_TRY(dev_alloc(ctx, &p->d_rowinfo, nr + 1, &p->bytes, st));

  // 3: bucket fill + sort
  PB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (nr + 1), st));
  if (total > 0) {
    fill_pairs<<<nblocks(total, 256), 256, 0, st>>>(d_conn, total, (int)row_begin, (int)row_end,
                                                    d_pair_ptr, d_cnt, d_pair_elem);
    if (nr > 0) sort_pairs<<<nblocks(nr, 128), 128, 0, st>>>(d_pair_ptr, (int)nr, d_pair_elem);
  }
  // 4-5: row lengths + scan
  if (nr > 0) {
    if (nn == 3)
      row_lengths<3><<<nblocks(nr, 128), 128, 0, st>>>(d_conn, d_pair_ptr, d_pair_elem, (int)nr, d_cnt,
                                                       d_flags, d_flags + 1);
    else
      row_lengths<4><<<nblocks(nr, 128), 128, 0, st>>>(d_conn, d_pair_ptr, d_pair_elem, (int)nr, d_cnt,
                                                       d_flags, d_flags + 1);
  }
  long long nnz = 0;
  PB_TRY(exclusive_scan(ctx, d_cnt, nr, d_row_ptr, &nnz, st));
  PB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  if (h_flags[0] & 2) {
    cleanup(); femx_pattern_destroy(p);
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: a row has more than %d columns", FEMX_MAX_ROW);
  }
  if (nnz >= (1LL << 31) - 1) {
    cleanup(); femx_pattern_destroy(p);
    return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: %lld node-level nonzeros exceed 32-bit offsets", nnz);
  }
  p->nnz_node = nnz;
  p->max_row = h_flags[1];
  PB_TRY(dev_alloc(ctx, &p->d_col_idx, nnz + 8, &p->bytes, st));  // +8: 16-byte bulk copies may overrun the last row
  PB_CUDA(cudaMemsetAsync(p->d_col_idx + nnz, 0, sizeof(int) * 8, st));
  // 6: columns + scatter map
  if (nn == 3)
    row_fill<3><<<nblocks(nr + 1, 128), 128, 0, st>>>(d_conn, d_pair_ptr, d_pair_elem, d_row_ptr, (int)nr,
                                                      (int)row_begin, p->d_col_idx, d_pair_code, p->d_rowinfo);
  else
    row_fill<4><<<nblocks(nr + 1, 128), 128, 0, st>>>(d_conn, d_pair_ptr, d_pair_elem, d_row_ptr, (int)nr,
                                                      (int)row_begin, p->d_col_idx, d_pair_code, p->d_rowinfo);
  // SELL-32 transposition of the scatter map
  {
    const int64_t n_slices = (nr + 31) / 32;
    int* d_ssize = nullptr;
    PB_TRY(tmp_alloc(ctx, &d_ssize, n_slices + 1, st));
    st_code = dev_alloc(ctx, &p->d_slice_ptr, n_slices + 1, &p->bytes, st);
    long long n_sell = 0;
    if (st_code == FEMX_OK) {
      if (n_slices > 0) slice_sizes<<<nblocks(n_slices, 8), 256, 0, st>>>(d_pair_ptr, (int)nr, (int)n_slices, d_ssize);
      st_code = exclusive_scan(ctx, d_ssize, n_slices, p->d_slice_ptr, &n_sell, st);
    }
    free_async(d_ssize, st);
    if (st_code != FEMX_OK) { cleanup(); femx_pattern_destroy(p); return st_code; }
    if (n_sell >= (1LL << 31) - 1) {
      cleanup(); femx_pattern_destroy(p);
      return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_build: padded scatter map (%lld) exceeds 32-bit offsets", n_sell);
    }
    p->n_sell = n_sell;
    PB_TRY(dev_alloc(ctx, &p->d_sell_code, n_sell, &p->bytes, st));
    PB_TRY(dev_alloc(ctx, &p->d_sell_elem, n_sell, &p->bytes, st));
    if (n_sell > 0) {
      PB_CUDA(cudaMemsetAsync(p->d_sell_code, 0, sizeof(unsigned) * n_sell, st));
      PB_CUDA(cudaMemsetAsync(p->d_sell_elem, 0, sizeof(int) * n_sell, st));
      to_sell<<<nblocks(nr, 128), 128, 0, st>>>(d_pair_ptr, p->d_slice_ptr, (int)nr, d_pair_elem, d_pair_code,
                                                p->d_sell_elem, p->d_sell_code);
    }
  }
  if (nr > 0) {
    int64_t ntiles = (nr + p->tile_nodes - 1) / p->tile_nodes;
    tile_max<<<nblocks(ntiles, 128), 128, 0, st>>>(p->d_rowinfo, p->d_slice_ptr, (int)nr, p->tile_nodes, d_flags + 2);
  }
  // dominant stencil class (scalar problems; FEMX_SPEC=0 switches the detection off)
  if (nd == 1 && nr > 0 && n_pairs > 0 && ctx->knobs.spec != 0) {
    PB_TRY(detect_stencil_class(ctx, p, d_pair_ptr, d_pair_code, st));
    if (p->spec_np > 0 && ctx->knobs.lattice != 0) PB_TRY(detect_lattice(ctx, p, d_conn, st, true));
  }
  PB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  PB_CUDA(cudaGetLastError());
  p->max_tile_nnz = h_flags[2];
  p->max_tile_codes = h_flags[3];
  cleanup();
  *out = p;
  return FEMX_OK;
#undef PB_TRY
#undef PB_CUDA
}

void femx_pattern_destroy(femx_pattern* p) {
  if (!p) return;
  cudaFree(p->d_rowinfo);
  cudaFree(p->d_col_idx);
  cudaFree(p->d_slice_ptr);
  cudaFree(p->d_sell_code);
  cudaFree(p->d_sell_elem);
  cudaFree(p->d_other_rows);
  cudaFree(p->d_lat_tmpl);
  delete p;
}

int femx_pattern_info(const femx_pattern* p, int64_t* n_rows, int64_t* nnz, int64_t* max_row) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_info: pattern is NULL");
  if (n_rows) *n_rows = p->n_rows * p->nd;
  if (nnz) *nnz = p->nnz_node * p->nd * p->nd;
  if (max_row) *max_row = (int64_t)p->max_row * p->nd;
  return FEMX_OK;
}

int64_t femx_pattern_bytes(const femx_pattern* p) { return p ? p->bytes : 0; }

int femx_pattern_stencil(const femx_pattern* p, int* n_incid, int* row_len, int* self_pos, int64_t* rows,
                         uint32_t* h_codes, int32_t* h_offsets, int cap) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_stencil: pattern is NULL");
  if (n_incid) *n_incid = p->spec_np;
  if (row_len) *row_len = p->spec_rlen;
  if (self_pos) *self_pos = p->spec_self;
  if (rows) *rows = p->spec_rows;
  if (h_codes)
    for (int k = 0; k < p->spec_np && k < cap; ++k) h_codes[k] = p->spec_codes[k];
  if (h_offsets)
    for (int k = 0; k < p->spec_rlen && k < cap; ++k) h_offsets[k] = p->spec_off[k];
  return FEMX_OK;
}

int femx_pattern_lattice(const femx_pattern* p, int* n_per_cell, int64_t* h_cells, int64_t* h_strides, int64_t* node0,
                         int32_t* h_corners) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_lattice: pattern is NULL");
  const femx_lattice& L = p->lat;
  if (n_per_cell) *n_per_cell = L.ok ? L.P : 0;
  if (!L.ok) return FEMX_OK;
  for (int k = 0; k < 3; ++k) {
    if (h_cells) h_cells[k] = L.cn[k];
    if (h_strides) h_strides[k] = L.s[k];
  }
  if (node0) *node0 = L.node0;
  if (h_corners)
    for (int t = 0; t < L.P; ++t)
      for (int a = 0; a < p->nn; ++a) h_corners[t * p->nn + a] = L.corner[t][a];
  return FEMX_OK;
}

int femx_lattice_prefix(int dim, const int32_t* h_cells, const int64_t* h_strides, int64_t node0, const int32_t* h_weights,
                        int64_t n, const int64_t* h_nodes, int64_t* h_out) {
  if ((dim != 2 && dim != 3) || !h_cells || !h_strides || !h_weights || (n > 0 && (!h_nodes || !h_out)))
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_lattice_prefix: bad argument");
  lat_geom g = {};
  g.dim = dim;
  for (int d = 0; d < 3; ++d) { g.cn[d] = d < dim ? h_cells[d] : 1; g.s[d] = d < dim ? h_strides[d] : 0; }
  g.node0 = node0;
  if (g.cn[0] < 1 || g.cn[1] < 1 || g.cn[2] < 1 || g.s[0] != 1 || g.s[1] < g.cn[0] + 1 ||
      (dim == 3 && g.s[2] < (long long)g.cn[1] * g.s[1] + g.cn[0] + 1))
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_lattice_prefix: strides narrower than the lattice");
  const lat_pref P = lat_pref_make(g, h_weights);
  for (int64_t i = 0; i < n; ++i) {
    if (h_nodes[i] - node0 >= (1LL << 31)) return femx_fail(nullptr, FEMX_ERR_UNSUPPORTED, "femx_lattice_prefix: node ids beyond 32 bits");
    long long ijk[3];
    lat_locate(g, h_nodes[i] - node0, ijk);
    h_out[i] = lat_prefix(g, P, ijk);
  }
  return FEMX_OK;
}

int femx_pattern_export_csr(const femx_pattern* p, int64_t* d_rp64, int32_t* d_rp32, int32_t* d_col,
                            void* stream) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_export_csr: pattern is NULL");
  if (d_rp32 && p->nnz_node * p->nd * p->nd >= (1LL << 31) - 1)
    return femx_fail(p->ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_export_csr: nnz does not fit a 32-bit row_ptr");
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  int64_t n = p->n_rows * p->nd + 1;
  export_csr_k<<<nblocks(n, 128), 128, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)p->n_rows, p->nd,
                                                                  (int)p->col_base, nullptr, (long long*)d_rp64, d_rp32, d_col);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_pattern_export_csr_mapped(const femx_pattern* p, const int32_t* d_local_to_global, int64_t* d_rp64, int32_t* d_rp32,
                                   int32_t* d_col, void* stream) {
  if (!p || !d_local_to_global) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_export_csr_mapped: NULL argument");
  if (d_rp32 && p->nnz_node * p->nd * p->nd >= (1LL << 31) - 1)
    return femx_fail(p->ctx, FEMX_ERR_UNSUPPORTED, "femx_pattern_export_csr_mapped: nnz does not fit a 32-bit row_ptr");
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  int64_t n = p->n_rows * p->nd + 1;
  export_csr_k<<<nblocks(n, 128), 128, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)p->n_rows, p->nd, 0,
                                                                  d_local_to_global, (long long*)d_rp64, d_rp32, d_col);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_pattern_export_ell(const femx_pattern* p, int width, int32_t* d_len, int32_t* d_idx, void* stream) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_pattern_export_ell: pattern is NULL");
  if (p->nd != 1) return femx_fail(p->ctx, FEMX_ERR_INVALID, "femx_pattern_export_ell: nd must be 1");
  if (width < p->max_row)
    return femx_fail(p->ctx, FEMX_ERR_INVALID, "femx_pattern_export_ell: width %d < longest row %d", width, p->max_row);
  if (p->n_rows == 0) return FEMX_OK;
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  export_ell_k<<<nblocks(p->n_rows, 128), 128, 0, (cudaStream_t)stream>>>(p->d_rowinfo, p->d_col_idx, (int)p->n_rows,
                                                                          width, (int)p->col_base, d_len, d_idx);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

int femx_csr_to_ell(const femx_pattern* p, int dtype, int width, const void* d_values, void* d_ell, void* stream) {
  if (!p || !d_values || !d_ell) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_csr_to_ell: NULL argument");
  if (p->nd != 1) return femx_fail(p->ctx, FEMX_ERR_INVALID, "femx_csr_to_ell: nd must be 1");
  if (width < p->max_row)
    return femx_fail(p->ctx, FEMX_ERR_INVALID, "femx_csr_to_ell: width %d < longest row %d", width, p->max_row);
  if (p->n_rows == 0) return FEMX_OK;
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  if (dtype == FEMX_F64)
    csr_to_ell_k<double><<<nblocks(p->n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
        p->d_rowinfo, (int)p->n_rows, width, (const double*)d_values, (double*)d_ell);
  else
    csr_to_ell_k<float><<<nblocks(p->n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
        p->d_rowinfo, (int)p->n_rows, width, (const float*)d_values, (float*)d_ell);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

}  // extern "C"
