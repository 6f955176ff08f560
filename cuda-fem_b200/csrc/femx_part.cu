// Partitioning of a general (unstructured) mesh by owned CSR rows with duplicated ghost elements (SURVEY §8e:
// "contiguous node ranges after RCM" + ghost-element extraction; north_star (3)).  A rank owns the contiguous node range
// [node_lo, node_hi) of the caller's numbering; femx_partition_extract selects every element that touches an owned node
// (ascending, so every owned row keeps its ascending-element accumulation order => the same bits as the single-GPU matrix),
// renumbers the touched nodes by their rank in the sorted list of global ids (owned nodes stay contiguous and ordered, ghosts
// sit below / above them) and keeps the local -> global map for the column export.  Assembly of the sub-mesh then needs no
// communication: femx_pattern_build(row_begin, row_end) + femx_assemble_csr give the rank's rows of the global matrix.
// No reference counterpart (the reference is single-GPU).
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "femx_internal.h"

struct femx_part {
  femx_ctx* ctx = nullptr;
  int nn = 0;
  int64_t n_local_nodes = 0, n_local_elems = 0, row_begin = 0, row_end = 0;
  int32_t* d_conn = nullptr;   // [n_local_elems * nn] local node ids
  int32_t* d_l2g = nullptr;    // [n_local_nodes] ascending global ids
  int32_t* d_elem = nullptr;   // [n_local_elems] global element ids, ascending
};

namespace {

struct touches_range {
  const int32_t* conn;
  int nn;
  int lo, hi;
  __device__ bool operator()(int e) const {
    for (int a = 0; a < nn; ++a) {
      const int v = conn[(int64_t)e * nn + a];
      if (v >= lo && v < hi) return true;
    }
    return false;
  }
};

__global__ void gather_conn_k(const int32_t* __restrict__ conn, const int32_t* __restrict__ elem, int64_t total, int nn,
                              int32_t* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < total) out[k] = conn[(int64_t)elem[k / nn] * nn + k % nn];
}

template <class T>
__global__ void gather_vec_k(const T* __restrict__ g, const int32_t* __restrict__ l2g, int64_t n, T* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = g[l2g[k]];
}

inline unsigned nb(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

int femx_partition_extract(femx_ctx* ctx, int nn, int64_t n_nodes, int64_t n_elems, const int32_t* d_conn, int64_t node_lo,
                           int64_t node_hi, void* stream, femx_part** out) {
  if (!ctx || !out || (n_elems > 0 && !d_conn)) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_partition_extract: NULL argument");
  *out = nullptr;
  if ((nn != 3 && nn != 4) || node_lo < 0 || node_hi < node_lo || node_hi > n_nodes || n_nodes >= (1LL << 31) - 1 ||
      n_elems * nn >= (1LL << 31) - 1)
    return femx_fail(ctx, FEMX_ERR_INVALID, "femx_partition_extract: bad sizes (nn=%d, nodes [%lld,%lld) of %lld)", nn,
                     (long long)node_lo, (long long)node_hi, (long long)n_nodes);
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  femx_part* p = new femx_part();
  p->ctx = ctx; p->nn = nn;
  int32_t *d_sel = nullptr, *d_nodes = nullptr;
  auto fail = [&](int code) { cudaFree(d_sel); cudaFree(d_nodes); femx_part_destroy(p); return code; };
  try {
    auto pol = thrust::cuda::par.on(st);
    // 1. elements that touch an owned node, ascending
    if (cudaMalloc((void**)&d_sel, sizeof(int32_t) * (size_t)std::max<int64_t>(n_elems, 1)) != cudaSuccess)
      return fail(femx_fail(ctx, FEMX_ERR_NOMEM, "femx_partition_extract: out of memory"));
    touches_range pred = {d_conn, nn, (int)node_lo, (int)node_hi};
    int32_t* end = thrust::copy_if(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>((int)n_elems), d_sel, pred);
    const int64_t ne = end - d_sel;
    p->n_local_elems = ne;
    if (cudaMalloc((void**)&p->d_elem, sizeof(int32_t) * (size_t)std::max<int64_t>(ne, 1)) != cudaSuccess ||
        cudaMalloc((void**)&p->d_conn, sizeof(int32_t) * (size_t)std::max<int64_t>(ne * nn, 1)) != cudaSuccess)
      return fail(femx_fail(ctx, FEMX_ERR_NOMEM, "femx_partition_extract: out of memory"));
    cudaMemcpyAsync(p->d_elem, d_sel, sizeof(int32_t) * ne, cudaMemcpyDeviceToDevice, st);
    if (ne > 0) gather_conn_k<<<nb(ne * nn), 256, 0, st>>>(d_conn, p->d_elem, ne * nn, nn, p->d_conn);
    // 2. touched nodes + every owned node (an owned node without elements still owns an (empty) row), sorted, unique
    const int64_t n_own = node_hi - node_lo, n_all = ne * nn + n_own;
    if (cudaMalloc((void**)&d_nodes, sizeof(int32_t) * (size_t)std::max<int64_t>(n_all, 1)) != cudaSuccess)
      return fail(femx_fail(ctx, FEMX_ERR_NOMEM, "femx_partition_extract: out of memory"));
    cudaMemcpyAsync(d_nodes, p->d_conn, sizeof(int32_t) * ne * nn, cudaMemcpyDeviceToDevice, st);
    thrust::sequence(pol, d_nodes + ne * nn, d_nodes + n_all, (int32_t)node_lo);
    thrust::sort(pol, d_nodes, d_nodes + n_all);
    int32_t* uend = thrust::unique(pol, d_nodes, d_nodes + n_all);
    const int64_t nl = uend - d_nodes;
    p->n_local_nodes = nl;
    if (cudaMalloc((void**)&p->d_l2g, sizeof(int32_t) * (size_t)std::max<int64_t>(nl, 1)) != cudaSuccess)
      return fail(femx_fail(ctx, FEMX_ERR_NOMEM, "femx_partition_extract: out of memory"));
    cudaMemcpyAsync(p->d_l2g, d_nodes, sizeof(int32_t) * nl, cudaMemcpyDeviceToDevice, st);
    // 3. local connectivity: rank of each node in the sorted list
    if (ne > 0) thrust::lower_bound(pol, p->d_l2g, p->d_l2g + nl, p->d_conn, p->d_conn + ne * nn, p->d_conn);
    // 4. the owned range in local numbering (contiguous: every owned node is in the list)
    int32_t bounds[2] = {(int32_t)node_lo, (int32_t)node_hi};
    int32_t h_pos[2] = {0, 0};
    int32_t* d_q = nullptr;
    if (cudaMalloc((void**)&d_q, 4 * sizeof(int32_t)) != cudaSuccess) return fail(femx_fail(ctx, FEMX_ERR_NOMEM, "femx_partition_extract: out of memory"));
    cudaMemcpyAsync(d_q, bounds, sizeof bounds, cudaMemcpyHostToDevice, st);
    thrust::lower_bound(pol, p->d_l2g, p->d_l2g + nl, d_q, d_q + 2, d_q + 2);
    cudaMemcpyAsync(h_pos, d_q + 2, sizeof h_pos, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_q);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(femx_fail(ctx, FEMX_ERR_CUDA, "femx_partition_extract: %s", cudaGetErrorString(e)));
    p->row_begin = h_pos[0];
    p->row_end = h_pos[1];
    if (p->row_end - p->row_begin != n_own) return fail(femx_fail(ctx, FEMX_ERR_CUDA, "femx_partition_extract: owned range not contiguous"));
  } catch (const std::exception& ex) {
    (void)cudaGetLastError();
    return fail(femx_fail(ctx, FEMX_ERR_CUDA, "femx_partition_extract: %s", ex.what()));
  }
  cudaFree(d_sel);
  cudaFree(d_nodes);
  *out = p;
  return FEMX_OK;
}

void femx_part_destroy(femx_part* p) {
  if (!p) return;
  cudaFree(p->d_conn);
  cudaFree(p->d_l2g);
  cudaFree(p->d_elem);
  delete p;
}

int femx_part_info(const femx_part* p, int64_t* n_local_nodes, int64_t* n_local_elems, int64_t* row_begin, int64_t* row_end) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_part_info: NULL argument");
  if (n_local_nodes) *n_local_nodes = p->n_local_nodes;
  if (n_local_elems) *n_local_elems = p->n_local_elems;
  if (row_begin) *row_begin = p->row_begin;
  if (row_end) *row_end = p->row_end;
  return FEMX_OK;
}

int femx_part_arrays(const femx_part* p, const int32_t** d_conn_local, const int32_t** d_local_to_global,
                     const int32_t** d_elem_ids) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_part_arrays: NULL argument");
  if (d_conn_local) *d_conn_local = p->d_conn;
  if (d_local_to_global) *d_local_to_global = p->d_l2g;
  if (d_elem_ids) *d_elem_ids = p->d_elem;
  return FEMX_OK;
}

int femx_part_copy(const femx_part* p, int32_t* d_conn_local, int32_t* d_local_to_global, int32_t* d_elem_ids, void* stream) {
  if (!p) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_part_copy: NULL argument");
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (d_conn_local && p->n_local_elems)
    FEMX_CUDA_OK(p->ctx, cudaMemcpyAsync(d_conn_local, p->d_conn, sizeof(int32_t) * p->n_local_elems * p->nn, cudaMemcpyDeviceToDevice, st));
  if (d_local_to_global && p->n_local_nodes)
    FEMX_CUDA_OK(p->ctx, cudaMemcpyAsync(d_local_to_global, p->d_l2g, sizeof(int32_t) * p->n_local_nodes, cudaMemcpyDeviceToDevice, st));
  if (d_elem_ids && p->n_local_elems)
    FEMX_CUDA_OK(p->ctx, cudaMemcpyAsync(d_elem_ids, p->d_elem, sizeof(int32_t) * p->n_local_elems, cudaMemcpyDeviceToDevice, st));
  return FEMX_OK;
}

int femx_part_gather(const femx_part* p, int dtype, const void* d_global, void* d_local, void* stream) {
  if (!p || !d_global || !d_local) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_part_gather: NULL argument");
  if (p->n_local_nodes == 0) return FEMX_OK;
  FEMX_CUDA_OK(p->ctx, cudaSetDevice(p->ctx->device));
  if (dtype == FEMX_F64)
    gather_vec_k<double><<<nb(p->n_local_nodes), 256, 0, (cudaStream_t)stream>>>((const double*)d_global, p->d_l2g, p->n_local_nodes, (double*)d_local);
  else
    gather_vec_k<float><<<nb(p->n_local_nodes), 256, 0, (cudaStream_t)stream>>>((const float*)d_global, p->d_l2g, p->n_local_nodes, (float*)d_local);
  FEMX_CUDA_OK(p->ctx, cudaGetLastError());
  return FEMX_OK;
}

}  // extern "C"
