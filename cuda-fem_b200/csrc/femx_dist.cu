// Multi-GPU layer behind the C ABI (BASELINE configs 3-5): one femx_dist per rank
// (one process per GPU), NCCL over NVLink for the ONLY two exchanges the path has —
//   * the halo of the SpMV operand (grouped ncclSend/ncclRecv with the two slab neighbours), and
//   * one fused 2-double all-reduce per CG iteration (Chronopoulos-Gear single-reduction CG) —
// assembly itself needs none (owned rows + ghost elements).  The SpMV of the rows that read no ghost
// column runs on the compute stream while the halo travels on a second stream; one CG iteration is
// captured in a CUDA graph and replayed.  No reference counterpart (the reference is single-GPU:
// job.pbs:4,24 launches one rank).
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already mapped into the process — torch's —
// or the system's), so libfemx.so loads on a box without NCCL and never brings a second copy in.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "femx_internal.h"

int femx_spmv_range(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                    int64_t row_lo, int64_t row_hi, void* stream);
int femx_spmv_range2(const femx_pattern* p, int dtype, const void* d_values, const void* d_x, int64_t x_base, void* d_y,
                     int64_t row_lo, int64_t row_hi, int64_t row_lo2, int64_t row_hi2, void* stream);

namespace {

struct nccl_api {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string err;
};

const nccl_api* get_nccl() {
  static nccl_api api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
    struct { const char* name; void** slot; } syms[] = {
        {"ncclGetUniqueId", (void**)&api.GetUniqueId}, {"ncclCommInitRank", (void**)&api.CommInitRank},
        {"ncclCommDestroy", (void**)&api.CommDestroy}, {"ncclGroupStart", (void**)&api.GroupStart},
        {"ncclGroupEnd", (void**)&api.GroupEnd},       {"ncclSend", (void**)&api.Send},
        {"ncclRecv", (void**)&api.Recv},               {"ncclAllReduce", (void**)&api.AllReduce},
        {"ncclAllGather", (void**)&api.AllGather},
        {"ncclGetErrorString", (void**)&api.GetErrorString},
    };
    for (auto& s : syms) {
      *s.slot = dlsym(h, s.name);
      if (!*s.slot) { api.err = std::string("NCCL symbol missing: ") + s.name; return; }
    }
    api.ok = true;
  });
  return &api;
}

}  // namespace

struct femx_dist {
  femx_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t s_comm = nullptr;   // the halo travels here while interior rows are multiplied on the compute stream
  cudaStream_t s_main = nullptr;   // compute stream of femx_dist_cg (a capturable stream: the caller's may be the legacy stream)
  cudaEvent_t e_ready = nullptr, e_halo = nullptr, e_in = nullptr;
  // NVLink peer memory for the fused reduction of the CG (dot products -> all-reduce -> CG scalars in ONE kernel):
  // every rank owns 2 x world slots, peers[r] is rank r's buffer mapped through CUDA IPC (peers[rank] = own buffer)
  struct p2p_slot* p2p_mine = nullptr;
  struct p2p_slot** d_peers = nullptr;   // device array [world]
  std::vector<void*> p2p_opened;         // IPC mappings to close
  unsigned long long* d_seq = nullptr;   // reductions done so far (the same number on every rank)
  bool p2p = false;
};

struct p2p_slot { double v[2]; unsigned long long seq; unsigned long long pad; };

struct femx_dist_op {
  femx_dist* d = nullptr;
  const femx_pattern* pat = nullptr;
  int dtype = FEMX_F64;
  const void* vals = nullptr;
  int64_t n_owned = 0;               // dof rows owned
  int64_t ghost_lo = 0, ghost_hi = 0;  // operand entries below / above the owned range (local layout [lo | owned | hi])
  int64_t send_lo = 0, send_hi = 0;    // leading / trailing owned entries the neighbours need
  int64_t int_lo = 0, int_hi = 0;      // node rows [int_lo, int_hi) read no ghost column
  void* x_ext = nullptr;               // operand with ghost zones (SpMV entry point)
  // CG work vectors: r and w,p,s,x; r lives inside an extended buffer (it is the SpMV operand)
  void *r_ext = nullptr, *w = nullptr, *p = nullptr, *s = nullptr;
  double* d_sc = nullptr;     // [0] gamma [1] delta [2] gamma_old [3] alpha_old [4] alpha [5] beta
  double* d_hist = nullptr;   // residual history (gamma per iteration)
  int hist_cap = 0;
  int* d_it = nullptr;
  double* d_part = nullptr;   // partial sums of the dot products (private: no race with femx_dot2 users)
  cudaGraphExec_t graph = nullptr;
  const void* graph_x = nullptr;
  int use_graph = 1;
  // peer-memory halo of the CG: r_ext sits behind a header {flag_from_lo, flag_from_hi, cnt_lo, cnt_hi, err} in ONE allocation
  // that the two slab neighbours map through CUDA IPC; the update kernel stores its boundary entries of r into their ghost
  // zones and raises their flag, the boundary rows' SpMV waits for the own flags (no NCCL call, no second stream)
  char* halo_base = nullptr;
  void *peer_lo = nullptr, *peer_hi = nullptr;   // mapped allocations of rank - 1 / rank + 1
  void *push_lo_dst = nullptr, *push_hi_dst = nullptr;
  unsigned long long *push_lo_flag = nullptr, *push_hi_flag = nullptr;
  bool push = false;
};

#define FEMX_HALO_HDR 256

namespace {

#define ND_OK(d, call)                                                                                   \
  do {                                                                                                   \
    ncclResult_t r__ = (call);                                                                           \
    if (r__ != ncclSuccess)                                                                              \
      return femx_fail((d)->ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, get_nccl()->GetErrorString(r__)); \
  } while (0)

size_t esize(int dtype) { return dtype == FEMX_F64 ? 8 : 4; }
ncclDataType_t ntype(int dtype) { return dtype == FEMX_F64 ? ncclDouble : ncclFloat; }

// rows that read a ghost column: the sorted column list starts below row_begin or ends at / above row_end
__global__ void interior_range_k(const int2* __restrict__ rowinfo, const int* __restrict__ col, int n_rows, int row_begin,
                                 int row_end, int* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int lo = rowinfo[r].x, hi = rowinfo[r + 1].x;
  if (hi <= lo) return;
  if (col[lo] < row_begin) atomicMax(out, r + 1);   // interior starts after the last such row
  if (col[hi - 1] >= row_end) atomicMin(out + 1, r); // and ends at the first of these
}

template <class T>
__global__ void __launch_bounds__(256) dot2_part_k(int64_t n, const T* __restrict__ a, const T* __restrict__ b,
                                                   const T* __restrict__ c, double* __restrict__ part) {
  // part[blk] = sum a*a, part[1024 + blk] = sum b*c   (fixed grid, fixed tree: deterministic)
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double av = (double)a[i];
    s0 += av * av;
    s1 += (double)b[i] * (double)c[i];
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

__global__ void __launch_bounds__(256) dot2_fin_k(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
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

// Chronopoulos-Gear recurrences (device function): from the reduced gamma = (r,r), delta = (w,r); records gamma
__device__ __forceinline__ void cg_scalars(double* __restrict__ sc, double* __restrict__ hist, int* __restrict__ it) {
  const int k = *it;
  const double gamma = sc[0], delta = sc[1];
  double beta = 0.0, alpha;
  if (k == 0) {
    alpha = gamma / delta;
  } else {
    beta = gamma / sc[2];
    alpha = gamma / (delta - beta * gamma / sc[3]);
  }
  hist[k] = gamma;
  sc[2] = gamma; sc[3] = alpha; sc[4] = alpha; sc[5] = beta;
  *it = k + 1;
}

// ONE kernel (one CTA) for "finish the two dot products, sum them over the ranks, advance the CG scalars":
// the block reduces this rank's partial sums, thread r < world stores them (then a sequence number, behind a system-scope
// fence) into slot [parity][rank] of rank r's buffer over NVLink, polls slot [parity][r] of the own buffer, and thread 0 adds
// the world contributions in rank order — the same order on every rank, so every rank holds the same bits — and runs the
// recurrences.  Replaces dot2_fin_k + ncclAllReduce (2 doubles) + cg_scalars_k; slots alternate by parity so that a rank one
// reduction ahead never overwrites what a slower rank still reads.
__global__ void __launch_bounds__(256) cg_reduce_p2p_k(const double* __restrict__ part, int nblocks, double* __restrict__ sc,
                                                       double* __restrict__ hist, int* __restrict__ it,
                                                       p2p_slot* const* __restrict__ peers, int rank, int world,
                                                       unsigned long long* __restrict__ seq_counter) {
  __shared__ double sh0[256], sh1[256];
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { s0 += part[i]; s1 += part[FEMX_DOT_BLOCKS + i]; }
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh0[threadIdx.x] += sh0[threadIdx.x + o]; sh1[threadIdx.x] += sh1[threadIdx.x + o]; }
    __syncthreads();
  }
  const double m0 = sh0[0], m1 = sh1[0];
  __syncthreads();
  if (world > 1) {
    const unsigned long long seq = *seq_counter + 1;
    const int par = (int)(seq & 1);
    if ((int)threadIdx.x < world) {
      const int r = threadIdx.x;
      volatile p2p_slot* dst = peers[r] + par * world + rank;      // my contribution, in rank r's buffer
      dst->v[0] = m0;
      dst->v[1] = m1;
      __threadfence_system();
      dst->seq = seq;
      volatile p2p_slot* src = peers[rank] + par * world + r;      // rank r's contribution, in my buffer
      const long long t0 = clock64();
      bool lost = false;                                            // (bounded: a lost peer poisons the scalars instead of hanging the device)
      while (src->seq != seq) {
        if (clock64() - t0 > (1LL << 34)) { lost = true; break; }
      }
      __threadfence_system();
      sh0[r] = lost ? __longlong_as_double(0x7ff8000000000000LL) : src->v[0];
      sh1[r] = lost ? __longlong_as_double(0x7ff8000000000000LL) : src->v[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double g = 0.0, d = 0.0;
      for (int r = 0; r < world; ++r) { g += sh0[r]; d += sh1[r]; }
      sc[0] = g; sc[1] = d;
      *seq_counter = seq;
    }
  } else if (threadIdx.x == 0) {
    sc[0] = m0; sc[1] = m1;
  }
  if (threadIdx.x == 0) cg_scalars(sc, hist, it);
}

__global__ void cg_scalars_k(double* __restrict__ sc, double* __restrict__ hist, int* __restrict__ it) { cg_scalars(sc, hist, it); }

// p = r + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s
template <class T>
__global__ void cg_update_k(int64_t n, const double* __restrict__ sc, T* __restrict__ r, const T* __restrict__ w,
                            T* __restrict__ p, T* __restrict__ s, T* __restrict__ x) {
  const double alpha = sc[4], beta = sc[5];
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double pn = (double)r[i] + beta * (double)p[i];
  const double sn = (double)w[i] + beta * (double)s[i];
  p[i] = (T)pn;
  s[i] = (T)sn;
  x[i] = (T)((double)x[i] + alpha * pn);
  r[i] = (T)((double)r[i] - alpha * sn);
}

// The update kernel of the peer-memory halo: the same arithmetic; the threads of the first send_lo / last send_hi entries
// also store the new r into the neighbour's ghost zone (remote stores over NVLink), fence, and the LAST of the CTAs that touched
// a zone (local counter) raises the neighbour's flag to the sequence number of this iteration (= reductions done so far: the
// same number on every rank, monotone over solves).
template <class T>
__global__ void __launch_bounds__(256) cg_update_push_k(int64_t n, const double* __restrict__ sc, T* __restrict__ r, const T* __restrict__ w,
                                                        T* __restrict__ p, T* __restrict__ s, T* __restrict__ x, int64_t send_lo,
                                                        int64_t send_hi, T* __restrict__ dst_lo, T* __restrict__ dst_hi,
                                                        unsigned long long* flag_lo, unsigned long long* flag_hi,
                                                        unsigned* __restrict__ cnt, const unsigned long long* __restrict__ seq_counter,
                                                        unsigned n_cta_lo, unsigned n_cta_hi) {
  const double alpha = sc[4], beta = sc[5];
  const int64_t b0 = (int64_t)blockIdx.x * blockDim.x, i = b0 + threadIdx.x;
  bool remote = false;
  if (i < n) {
    const double pn = (double)r[i] + beta * (double)p[i];
    const double sn = (double)w[i] + beta * (double)s[i];
    p[i] = (T)pn;
    s[i] = (T)sn;
    x[i] = (T)((double)x[i] + alpha * pn);
    const T rn = (T)((double)r[i] - alpha * sn);
    r[i] = rn;
    if (i < send_lo) { dst_lo[i] = rn; remote = true; }
    if (i >= n - send_hi) { dst_hi[i - (n - send_hi)] = rn; remote = true; }
  }
  const int64_t b1 = min(b0 + (int64_t)blockDim.x, n) - 1;
  const bool cta_lo = b0 < send_lo, cta_hi = send_hi > 0 && b1 >= n - send_hi;
  if (cta_lo || cta_hi) {                      // (block-uniform)
    if (remote) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long seq = *seq_counter;
      if (cta_lo && atomicAdd(cnt, 1u) == n_cta_lo - 1) {
        *cnt = 0;
        __threadfence_system();
        *(volatile unsigned long long*)flag_lo = seq;
      }
      if (cta_hi && atomicAdd(cnt + 1, 1u) == n_cta_hi - 1) {
        cnt[1] = 0;
        __threadfence_system();
        *(volatile unsigned long long*)flag_hi = seq;
      }
    }
  }
}

// waits until both neighbours have delivered this iteration's ghost entries (bounded: a lost peer raises err instead of
// hanging the device)
__global__ void halo_wait_k(const unsigned long long* flags, int need_lo, int need_hi, const unsigned long long* __restrict__ seq_counter,
                            int* __restrict__ err) {
  const unsigned long long seq = *seq_counter;
  const int t = threadIdx.x;
  if ((t == 0 && need_lo) || (t == 1 && need_hi)) {
    const volatile unsigned long long* f = flags + t;
    const long long t0 = clock64();
    while (*f < seq) {
      if (clock64() - t0 > (1LL << 33)) { *err = 1; break; }
    }
  }
  __threadfence_system();
}

inline unsigned nb256(int64_t n) { return (unsigned)((n + 255) / 256); }

// ghost zones of `ext` (layout [ghost_lo | owned | ghost_hi]) from the neighbours, on the comm stream
int halo_exchange(femx_dist_op* op, void* ext) {
  femx_dist* d = op->d;
  const nccl_api* nc = get_nccl();
  const size_t es = esize(op->dtype);
  char* base = (char*)ext;
  char* own = base + op->ghost_lo * es;
  ND_OK(d, nc->GroupStart());
  if (op->ghost_lo > 0 && d->rank > 0) {
    ND_OK(d, nc->Send(own, (size_t)op->send_lo, ntype(op->dtype), d->rank - 1, d->comm, d->s_comm));
    ND_OK(d, nc->Recv(base, (size_t)op->ghost_lo, ntype(op->dtype), d->rank - 1, d->comm, d->s_comm));
  }
  if (op->ghost_hi > 0 && d->rank < d->world - 1) {
    ND_OK(d, nc->Send(own + (op->n_owned - op->send_hi) * es, (size_t)op->send_hi, ntype(op->dtype), d->rank + 1, d->comm, d->s_comm));
    ND_OK(d, nc->Recv(own + op->n_owned * es, (size_t)op->ghost_hi, ntype(op->dtype), d->rank + 1, d->comm, d->s_comm));
  }
  ND_OK(d, nc->GroupEnd());
  return FEMX_OK;
}

// y = A[owned rows] ext, the halo of `ext` exchanged on the way: interior rows overlap the exchange
int spmv_overlapped(femx_dist_op* op, void* ext, void* y, cudaStream_t st) {
  femx_dist* d = op->d;
  const femx_pattern* p = op->pat;
  const int64_t xb = (int64_t)p->col_base * p->nd;  // operand index = global dof - col_base*nd
  const bool comm = d->world > 1 && (op->ghost_lo > 0 || op->ghost_hi > 0);
  if (comm) {
    FEMX_CUDA_OK(d->ctx, cudaEventRecord(d->e_ready, st));          // the owned part of ext is final
    FEMX_CUDA_OK(d->ctx, cudaStreamWaitEvent(d->s_comm, d->e_ready, 0));
    int rc = halo_exchange(op, ext);
    if (rc != FEMX_OK) return rc;
    FEMX_CUDA_OK(d->ctx, cudaEventRecord(d->e_halo, d->s_comm));
  }
  int rc = FEMX_OK;
  if (op->int_hi > op->int_lo) rc = femx_spmv_range(p, op->dtype, op->vals, ext, xb, y, op->int_lo, op->int_hi, st);
  if (rc != FEMX_OK) return rc;
  if (comm) FEMX_CUDA_OK(d->ctx, cudaStreamWaitEvent(st, d->e_halo, 0));
  // the rows of the first and the last owned plane (they read the ghost zones): one launch
  return femx_spmv_range2(p, op->dtype, op->vals, ext, xb, y, 0, op->int_lo, op->int_hi, p->n_rows, st);
}

// (gamma, delta) = ((r,r), (w,r)) summed over the ranks, then the CG scalars for the next update — ONE reduction per iteration
int reduce2(femx_dist_op* op, const void* r, const void* w, cudaStream_t st) {
  femx_dist* d = op->d;
  const int64_t n = op->n_owned;
  const int blocks = (int)std::min<int64_t>(FEMX_DOT_BLOCKS, std::max<int64_t>(1, (n + 255) / 256));
  if (op->dtype == FEMX_F64)
    dot2_part_k<double><<<blocks, 256, 0, st>>>(n, (const double*)r, (const double*)w, (const double*)r, op->d_part);
  else
    dot2_part_k<float><<<blocks, 256, 0, st>>>(n, (const float*)r, (const float*)w, (const float*)r, op->d_part);
  if (d->world == 1 || d->p2p) {
    cg_reduce_p2p_k<<<1, 256, 0, st>>>(op->d_part, blocks, op->d_sc, op->d_hist, op->d_it, d->d_peers, d->rank, d->world, d->d_seq);
    FEMX_CUDA_OK(d->ctx, cudaGetLastError());
    return FEMX_OK;
  }
  dot2_fin_k<<<1, 256, 0, st>>>(op->d_part, blocks, op->d_sc);
  FEMX_CUDA_OK(d->ctx, cudaGetLastError());
  ND_OK(d, get_nccl()->AllReduce(op->d_sc, op->d_sc, 2, ncclDouble, ncclSum, d->comm, st));
  cg_scalars_k<<<1, 1, 0, st>>>(op->d_sc, op->d_hist, op->d_it);
  FEMX_CUDA_OK(d->ctx, cudaGetLastError());
  return FEMX_OK;
}

int cg_iteration(femx_dist_op* op, void* x, cudaStream_t st) {
  femx_dist* d = op->d;
  const int64_t n = op->n_owned;
  const size_t es = esize(op->dtype);
  void* r = (char*)op->r_ext + op->ghost_lo * es;
  if (op->push) {
    // peer-memory halo: update + push, interior rows, wait for the neighbours' flags, boundary rows — one stream, no NCCL
    const femx_pattern* p = op->pat;
    const int64_t xb = (int64_t)p->col_base * p->nd;
    unsigned long long* hdr = (unsigned long long*)op->halo_base;
    unsigned* cnt = (unsigned*)(hdr + 2);
    int* err = (int*)(hdr + 4);
    const unsigned nb = nb256(n);
    const unsigned c_lo = (unsigned)((op->send_lo + 255) / 256);
    const unsigned c_hi = op->send_hi > 0 ? nb - (unsigned)((n - op->send_hi) / 256) : 0;
    if (op->dtype == FEMX_F64)
      cg_update_push_k<double><<<nb, 256, 0, st>>>(n, op->d_sc, (double*)r, (const double*)op->w, (double*)op->p, (double*)op->s, (double*)x,
                                                   op->send_lo, op->send_hi, (double*)op->push_lo_dst, (double*)op->push_hi_dst,
                                                   op->push_lo_flag, op->push_hi_flag, cnt, d->d_seq, c_lo, c_hi);
    else
      cg_update_push_k<float><<<nb, 256, 0, st>>>(n, op->d_sc, (float*)r, (const float*)op->w, (float*)op->p, (float*)op->s, (float*)x,
                                                  op->send_lo, op->send_hi, (float*)op->push_lo_dst, (float*)op->push_hi_dst,
                                                  op->push_lo_flag, op->push_hi_flag, cnt, d->d_seq, c_lo, c_hi);
    FEMX_CUDA_OK(d->ctx, cudaGetLastError());
    int rc = FEMX_OK;
    if (op->int_hi > op->int_lo) rc = femx_spmv_range(p, op->dtype, op->vals, op->r_ext, xb, op->w, op->int_lo, op->int_hi, st);
    if (rc != FEMX_OK) return rc;
    halo_wait_k<<<1, 32, 0, st>>>(hdr, op->ghost_lo > 0 && d->rank > 0, op->ghost_hi > 0 && d->rank < d->world - 1, d->d_seq, err);
    FEMX_CUDA_OK(d->ctx, cudaGetLastError());
    rc = femx_spmv_range2(p, op->dtype, op->vals, op->r_ext, xb, op->w, 0, op->int_lo, op->int_hi, p->n_rows, st);
    if (rc != FEMX_OK) return rc;
    return reduce2(op, r, op->w, st);
  }
  if (op->dtype == FEMX_F64)
    cg_update_k<double><<<nb256(n), 256, 0, st>>>(n, op->d_sc, (double*)r, (const double*)op->w, (double*)op->p, (double*)op->s, (double*)x);
  else
    cg_update_k<float><<<nb256(n), 256, 0, st>>>(n, op->d_sc, (float*)r, (const float*)op->w, (float*)op->p, (float*)op->s, (float*)x);
  FEMX_CUDA_OK(d->ctx, cudaGetLastError());
  int rc = spmv_overlapped(op, op->r_ext, op->w, st);
  if (rc != FEMX_OK) return rc;
  return reduce2(op, r, op->w, st);
}

// Peer mapping of the CG's residual buffers (collective: every rank of the communicator calls it from femx_dist_op_create).
// One ncclAllGather carries {IPC handle, ghost_lo, n_owned, ghost_hi, ok} of every rank; a rank maps its two neighbours;
// an all-reduce (MIN) makes the outcome unanimous.
struct halo_record { cudaIpcMemHandle_t h; long long ghost_lo, n_owned, ghost_hi, ok; };

void dist_setup_push(femx_dist_op* op) {
  femx_dist* d = op->d;
  const nccl_api* nc = get_nccl();
  const int world = d->world;
  const size_t es = esize(op->dtype);
  bool ok = true;
  halo_record mine = {};
  mine.ghost_lo = op->ghost_lo; mine.n_owned = op->n_owned; mine.ghost_hi = op->ghost_hi;
  if (cudaIpcGetMemHandle(&mine.h, op->halo_base) != cudaSuccess) ok = false;
  mine.ok = ok ? 1 : 0;
  std::vector<halo_record> all(world);
  halo_record* d_all = nullptr;
  if (cudaMalloc((void**)&d_all, sizeof(halo_record) * world + 16) != cudaSuccess) { (void)cudaGetLastError(); d_all = nullptr; }
  // (from here on every rank makes the same collective calls whatever its own state)
  if (d_all) cudaMemcpy(d_all + d->rank, &mine, sizeof mine, cudaMemcpyHostToDevice);
  ncclResult_t r = d_all ? nc->AllGather(d_all + d->rank, d_all, sizeof mine, ncclUint8, d->comm, d->s_comm) : ncclSuccess;
  if (!d_all || r != ncclSuccess || cudaStreamSynchronize(d->s_comm) != cudaSuccess) ok = false;
  if (ok) cudaMemcpy(all.data(), d_all, sizeof(halo_record) * world, cudaMemcpyDeviceToHost);
  for (int q = 0; q < world && ok; ++q) ok = all[q].ok == 1;
  auto open = [&](int q, void** out) {
    if (cudaIpcOpenMemHandle(out, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { *out = nullptr; ok = false; }
  };
  if (ok && d->rank > 0 && op->send_lo > 0) {
    open(d->rank - 1, &op->peer_lo);
    if (ok && all[d->rank - 1].ghost_hi != op->send_lo) ok = false;
    if (ok) {
      op->push_lo_dst = (char*)op->peer_lo + FEMX_HALO_HDR + (all[d->rank - 1].ghost_lo + all[d->rank - 1].n_owned) * es;
      op->push_lo_flag = (unsigned long long*)op->peer_lo + 1;        // its flag_from_hi
    }
  }
  if (ok && d->rank < world - 1 && op->send_hi > 0) {
    open(d->rank + 1, &op->peer_hi);
    if (ok && all[d->rank + 1].ghost_lo != op->send_hi) ok = false;
    if (ok) {
      op->push_hi_dst = (char*)op->peer_hi + FEMX_HALO_HDR;
      op->push_hi_flag = (unsigned long long*)op->peer_hi;            // its flag_from_lo
    }
  }
  (void)cudaGetLastError();
  double flag = ok ? 1.0 : 0.0, *d_flag = (double*)(d_all + world);
  if (d_all) {
    cudaMemcpy(d_flag, &flag, sizeof flag, cudaMemcpyHostToDevice);
    r = nc->AllReduce(d_flag, d_flag, 1, ncclDouble, ncclMin, d->comm, d->s_comm);
    if (r == ncclSuccess && cudaStreamSynchronize(d->s_comm) == cudaSuccess) cudaMemcpy(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost);
    else flag = 0.0;
    cudaFree(d_all);
  } else {
    flag = 0.0;
  }
  op->push = flag == 1.0;
}

// Peer buffers of the fused reduction: every rank allocates 2 x world slots, the IPC handles travel over one ncclAllGather,
// every rank maps the others' buffers (NVLink peer access).  All ranks agree on success through an all-reduce (MIN), so either
// every rank uses the peer path or none does.
void dist_setup_p2p(femx_dist* d) {
  const nccl_api* nc = get_nccl();
  const int world = d->world;
  bool ok = true;
  unsigned char* d_h = nullptr;
  std::vector<unsigned char> h((size_t)world * sizeof(cudaIpcMemHandle_t));
  if (cudaMalloc((void**)&d->p2p_mine, sizeof(p2p_slot) * 2 * world) != cudaSuccess) ok = false;
  if (ok && cudaMemset(d->p2p_mine, 0, sizeof(p2p_slot) * 2 * world) != cudaSuccess) ok = false;
  cudaIpcMemHandle_t mine;
  if (ok && cudaIpcGetMemHandle(&mine, d->p2p_mine) != cudaSuccess) ok = false;
  if (cudaMalloc((void**)&d_h, h.size()) != cudaSuccess) { (void)cudaGetLastError(); return; }
  if (ok) cudaMemcpy(d_h + (size_t)d->rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice);
  // (collective calls are made by every rank whatever `ok` is, so that nobody hangs)
  ncclResult_t r = nc->AllGather(d_h + (size_t)d->rank * sizeof mine, d_h, sizeof mine, ncclUint8, d->comm, d->s_comm);
  if (r != ncclSuccess || cudaStreamSynchronize(d->s_comm) != cudaSuccess) ok = false;
  cudaMemcpy(h.data(), d_h, h.size(), cudaMemcpyDeviceToHost);
  std::vector<p2p_slot*> peers(world, nullptr);
  for (int q = 0; q < world && ok; ++q) {
    if (q == d->rank) { peers[q] = d->p2p_mine; continue; }
    cudaIpcMemHandle_t hq;
    memcpy(&hq, h.data() + (size_t)q * sizeof hq, sizeof hq);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    d->p2p_opened.push_back(p);
    peers[q] = (p2p_slot*)p;
  }
  if (ok && cudaMalloc((void**)&d->d_peers, sizeof(p2p_slot*) * world) != cudaSuccess) ok = false;
  if (ok) cudaMemcpy(d->d_peers, peers.data(), sizeof(p2p_slot*) * world, cudaMemcpyHostToDevice);
  (void)cudaGetLastError();
  // agreement (and a barrier: every buffer is zeroed and mapped before anyone's first reduction)
  double flag = ok ? 1.0 : 0.0, *d_flag = (double*)d_h;
  cudaMemcpy(d_flag, &flag, sizeof flag, cudaMemcpyHostToDevice);
  r = nc->AllReduce(d_flag, d_flag, 1, ncclDouble, ncclMin, d->comm, d->s_comm);
  if (r == ncclSuccess && cudaStreamSynchronize(d->s_comm) == cudaSuccess) cudaMemcpy(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost);
  else flag = 0.0;
  cudaFree(d_h);
  d->p2p = flag == 1.0;
}

}  // namespace

extern "C" {

int femx_dist_unique_id(void* h_id) {
  if (!h_id) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_unique_id: NULL argument");
  const nccl_api* nc = get_nccl();
  if (!nc->ok) return femx_fail(nullptr, FEMX_ERR_UNSUPPORTED, "femx_dist_unique_id: %s", nc->err.c_str());
  static_assert(sizeof(ncclUniqueId) == FEMX_DIST_ID_BYTES, "ncclUniqueId size");
  ncclResult_t r = nc->GetUniqueId((ncclUniqueId*)h_id);
  if (r != ncclSuccess) return femx_fail(nullptr, FEMX_ERR_CUDA, "ncclGetUniqueId: %s", nc->GetErrorString(r));
  return FEMX_OK;
}

int femx_dist_create(femx_ctx* ctx, int rank, int world, const void* h_id, femx_dist** out) {
  if (!ctx || !out) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_dist_create: NULL argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_dist_create: rank %d of %d", rank, world);
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  femx_dist* d = new femx_dist();
  d->ctx = ctx; d->rank = rank; d->world = world;
  if (world > 1) {
    const nccl_api* nc = get_nccl();
    if (!nc->ok || !h_id) {
      delete d;
      return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_dist_create: %s", !h_id ? "unique id is NULL" : nc->err.c_str());
    }
    ncclUniqueId id;
    memcpy(&id, h_id, sizeof id);
    ncclResult_t r = nc->CommInitRank(&d->comm, world, id, rank);
    if (r != ncclSuccess) {
      delete d;
      return femx_fail(ctx, FEMX_ERR_CUDA, "ncclCommInitRank: %s", nc->GetErrorString(r));
    }
  }
  cudaError_t e = cudaStreamCreateWithFlags(&d->s_comm, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->s_main, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->e_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->e_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->e_halo, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    femx_dist_destroy(d);
    return femx_fail(ctx, FEMX_ERR_CUDA, "femx_dist_create: %s", cudaGetErrorString(e));
  }
  // sequence counter of the fused reduction (also used at world == 1) and, for world > 1, the peer buffers
  e = cudaMalloc((void**)&d->d_seq, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(d->d_seq, 0, sizeof(unsigned long long));
  if (e != cudaSuccess) {
    femx_dist_destroy(d);
    return femx_fail(ctx, FEMX_ERR_NOMEM, "femx_dist_create: %s", cudaGetErrorString(e));
  }
  if (world > 1 && ctx->knobs.dist_p2p != 0) dist_setup_p2p(d);   // failure is not an error: the NCCL all-reduce stays
  *out = d;
  return FEMX_OK;
}

void femx_dist_destroy(femx_dist* d) {
  if (!d) return;
  for (void* p : d->p2p_opened) cudaIpcCloseMemHandle(p);
  cudaFree(d->p2p_mine);
  cudaFree(d->d_peers);
  cudaFree(d->d_seq);
  if (d->comm) get_nccl()->CommDestroy(d->comm);
  if (d->s_comm) cudaStreamDestroy(d->s_comm);
  if (d->s_main) cudaStreamDestroy(d->s_main);
  if (d->e_in) cudaEventDestroy(d->e_in);
  if (d->e_ready) cudaEventDestroy(d->e_ready);
  if (d->e_halo) cudaEventDestroy(d->e_halo);
  delete d;
}

int femx_dist_slab(int64_t n_planes, int world, int rank, int64_t* r0, int64_t* r1, int64_t* lo, int64_t* hi) {
  if (world < 1 || rank < 0 || rank >= world || n_planes < world)
    return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_slab: cannot split %lld node planes over %d ranks (rank %d)",
                     (long long)n_planes, world, rank);
  const int64_t a = (rank * n_planes) / world, b = ((rank + 1) * n_planes) / world;
  if (r0) *r0 = a;
  if (r1) *r1 = b;
  if (lo) *lo = std::max<int64_t>(a - 1, 0);
  if (hi) *hi = std::min<int64_t>(b, n_planes - 1);
  return FEMX_OK;
}

int femx_dist_allreduce(femx_dist* d, double* d_buf, int n, int op_max, void* stream) {
  if (!d || !d_buf) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_allreduce: NULL argument");
  if (d->world == 1) return FEMX_OK;
  ND_OK(d, get_nccl()->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, op_max ? ncclMax : ncclSum, d->comm, (cudaStream_t)stream));
  return FEMX_OK;
}

int femx_dist_op_create(femx_dist* d, const femx_pattern* pat, int dtype, const void* d_values, femx_dist_op** out) {
  if (!d || !pat || !d_values || !out) return femx_fail(d ? d->ctx : nullptr, FEMX_ERR_INVALID, "femx_dist_op_create: NULL argument");
  *out = nullptr;
  femx_ctx* ctx = d->ctx;
  if (pat->ctx != ctx) return femx_fail(ctx, FEMX_ERR_INVALID, "femx_dist_op_create: pattern belongs to another context");
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  femx_dist_op* op = new femx_dist_op();
  op->d = d; op->pat = pat; op->dtype = dtype; op->vals = d_values;
  const int nd = pat->nd;
  op->n_owned = pat->n_rows * nd;
  op->ghost_lo = pat->row_begin * nd;
  op->ghost_hi = (pat->n_nodes - pat->row_end) * nd;
  op->use_graph = ctx->knobs.dist_graph;
  const size_t es = esize(dtype);
  const int64_t n_ext = op->ghost_lo + op->n_owned + op->ghost_hi;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes ? bytes : 8); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes ? bytes : 8); } };
  alloc(&op->x_ext, n_ext * es);
  alloc((void**)&op->halo_base, FEMX_HALO_HDR + n_ext * es);
  if (e == cudaSuccess) op->r_ext = op->halo_base + FEMX_HALO_HDR;
  alloc(&op->w, op->n_owned * es);
  alloc(&op->p, op->n_owned * es);
  alloc(&op->s, op->n_owned * es);
  alloc((void**)&op->d_sc, 8 * sizeof(double));
  alloc((void**)&op->d_it, 4 * sizeof(int));
  alloc((void**)&op->d_part, 2 * FEMX_DOT_BLOCKS * sizeof(double));
  if (e != cudaSuccess) {
    femx_dist_op_destroy(op);
    return femx_fail(ctx, FEMX_ERR_NOMEM, "femx_dist_op_create: %s", cudaGetErrorString(e));
  }
  // rows that read no ghost column: [int_lo, int_hi)
  {
    int h[2] = {0, (int)pat->n_rows};
    cudaMemcpy(op->d_it + 2, h, sizeof h, cudaMemcpyHostToDevice);
    if (pat->n_rows > 0)
      interior_range_k<<<nb256(pat->n_rows), 256>>>(pat->d_rowinfo, pat->d_col_idx, (int)pat->n_rows, (int)pat->row_begin,
                                                    (int)pat->row_end, op->d_it + 2);
    e = cudaMemcpy(h, op->d_it + 2, sizeof h, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      femx_dist_op_destroy(op);
      return femx_fail(ctx, FEMX_ERR_CUDA, "femx_dist_op_create: %s", cudaGetErrorString(e));
    }
    op->int_lo = h[0]; op->int_hi = std::max(h[0], h[1]);
  }
  // what the neighbours need from this rank: their ghost zone sizes (one int64 each way, through NCCL)
  op->send_lo = op->send_hi = 0;
  if (d->world > 1) {
    long long* d_cnt = nullptr;
    e = cudaMalloc((void**)&d_cnt, 4 * sizeof(long long));
    long long h[4] = {op->ghost_lo, op->ghost_hi, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(d_cnt, h, sizeof h, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { femx_dist_op_destroy(op); return femx_fail(ctx, FEMX_ERR_CUDA, "femx_dist_op_create: %s", cudaGetErrorString(e)); }
    const nccl_api* nc = get_nccl();
    ncclResult_t r = nc->GroupStart();
    if (d->rank > 0) {  // lower neighbour: tell it my ghost_lo (= what it must send up), learn its ghost_hi
      if (r == ncclSuccess) r = nc->Send(d_cnt + 0, 1, ncclInt64, d->rank - 1, d->comm, d->s_comm);
      if (r == ncclSuccess) r = nc->Recv(d_cnt + 2, 1, ncclInt64, d->rank - 1, d->comm, d->s_comm);
    }
    if (d->rank < d->world - 1) {
      if (r == ncclSuccess) r = nc->Send(d_cnt + 1, 1, ncclInt64, d->rank + 1, d->comm, d->s_comm);
      if (r == ncclSuccess) r = nc->Recv(d_cnt + 3, 1, ncclInt64, d->rank + 1, d->comm, d->s_comm);
    }
    if (r == ncclSuccess) r = nc->GroupEnd();
    if (r == ncclSuccess) {
      e = cudaStreamSynchronize(d->s_comm);
      if (e == cudaSuccess) e = cudaMemcpy(h, d_cnt, sizeof h, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_cnt);
    if (r != ncclSuccess || e != cudaSuccess) {
      femx_dist_op_destroy(op);
      return femx_fail(ctx, FEMX_ERR_CUDA, "femx_dist_op_create: halo size exchange failed: %s",
                       r != ncclSuccess ? nc->GetErrorString(r) : cudaGetErrorString(e));
    }
    op->send_lo = d->rank > 0 ? h[2] : 0;                // lower neighbour's ghost_hi = my leading owned entries
    op->send_hi = d->rank < d->world - 1 ? h[3] : 0;     // upper neighbour's ghost_lo = my trailing owned entries
    if (op->send_lo > op->n_owned || op->send_hi > op->n_owned) {
      femx_dist_op_destroy(op);
      return femx_fail(ctx, FEMX_ERR_UNSUPPORTED, "femx_dist_op_create: a neighbour's ghost zone is larger than this rank's owned range");
    }
    // (d->p2p and the option are the same on every rank: either all ranks make the collective calls below or none)
    if (d->p2p && ctx->knobs.dist_push != 0) dist_setup_push(op);
  }
  *out = op;
  return FEMX_OK;
}

void femx_dist_op_destroy(femx_dist_op* op) {
  if (!op) return;
  if (op->graph) cudaGraphExecDestroy(op->graph);
  if (op->peer_lo) cudaIpcCloseMemHandle(op->peer_lo);
  if (op->peer_hi) cudaIpcCloseMemHandle(op->peer_hi);
  cudaFree(op->x_ext); cudaFree(op->halo_base); cudaFree(op->w); cudaFree(op->p); cudaFree(op->s);
  cudaFree(op->d_sc); cudaFree(op->d_it); cudaFree(op->d_part); cudaFree(op->d_hist);
  delete op;
}

int femx_dist_info(const femx_dist* d, int* rank, int* world, int* p2p_reduction) {
  if (!d) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_info: NULL argument");
  if (rank) *rank = d->rank;
  if (world) *world = d->world;
  if (p2p_reduction) *p2p_reduction = d->p2p ? 1 : 0;
  return FEMX_OK;
}

int femx_dist_op_info(const femx_dist_op* op, int64_t* n_owned, int64_t* ghost_lo, int64_t* ghost_hi, int64_t* interior_lo,
                      int64_t* interior_hi) {
  if (!op) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_op_info: NULL argument");
  if (n_owned) *n_owned = op->n_owned;
  if (ghost_lo) *ghost_lo = op->ghost_lo;
  if (ghost_hi) *ghost_hi = op->ghost_hi;
  if (interior_lo) *interior_lo = op->int_lo;
  if (interior_hi) *interior_hi = op->int_hi;
  return FEMX_OK;
}

int femx_dist_op_peer_halo(const femx_dist_op* op, int* peer_halo) {
  if (!op || !peer_halo) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_op_peer_halo: NULL argument");
  *peer_halo = op->push ? 1 : 0;
  return FEMX_OK;
}

int femx_dist_spmv(femx_dist_op* op, const void* d_x_owned, void* d_y_owned, void* stream) {
  if (!op || !d_x_owned || !d_y_owned) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_spmv: NULL argument");
  femx_ctx* ctx = op->d->ctx;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = esize(op->dtype);
  FEMX_CUDA_OK(ctx, cudaMemcpyAsync((char*)op->x_ext + op->ghost_lo * es, d_x_owned, op->n_owned * es, cudaMemcpyDeviceToDevice, st));
  return spmv_overlapped(op, op->x_ext, d_y_owned, st);
}

int femx_dist_cg(femx_dist_op* op, const void* d_b_owned, void* d_x_owned, int iters, double* h_residuals, float* h_ms,
                 void* stream) {
  if (!op || !d_b_owned || !d_x_owned || iters < 0) return femx_fail(nullptr, FEMX_ERR_INVALID, "femx_dist_cg: bad argument");
  femx_dist* d = op->d;
  femx_ctx* ctx = d->ctx;
  FEMX_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // the solve runs on the layer's own stream, ordered after the caller's: the iteration is stream-captured, and the
  // caller's stream may be the legacy default stream, which cannot be captured
  FEMX_CUDA_OK(ctx, cudaEventRecord(d->e_in, (cudaStream_t)stream));
  FEMX_CUDA_OK(ctx, cudaStreamWaitEvent(d->s_main, d->e_in, 0));
  cudaStream_t st = d->s_main;
  const size_t es = esize(op->dtype);
  const int64_t n = op->n_owned;
  if (op->hist_cap < iters + 1) {
    cudaFree(op->d_hist);
    op->d_hist = nullptr;
    FEMX_CUDA_OK(ctx, cudaMalloc((void**)&op->d_hist, sizeof(double) * (iters + 1)));
    op->hist_cap = iters + 1;
  }
  void* r = (char*)op->r_ext + op->ghost_lo * es;
  cudaEvent_t t0, t1;
  FEMX_CUDA_OK(ctx, cudaEventCreate(&t0));
  FEMX_CUDA_OK(ctx, cudaEventCreate(&t1));
  auto fail = [&](int code) { cudaEventDestroy(t0); cudaEventDestroy(t1); return code; };
#define CG_CUDA(call)                                                                                         \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess) return fail(femx_fail(ctx, FEMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__))); \
  } while (0)
  // x0 = 0, r = b, w = A r, (gamma, delta) = ((r,r), (w,r))
  CG_CUDA(cudaMemsetAsync(d_x_owned, 0, n * es, st));
  CG_CUDA(cudaMemsetAsync(op->p, 0, n * es, st));
  CG_CUDA(cudaMemsetAsync(op->s, 0, n * es, st));
  CG_CUDA(cudaMemsetAsync(op->d_it, 0, sizeof(int), st));
  CG_CUDA(cudaMemcpyAsync(r, d_b_owned, n * es, cudaMemcpyDeviceToDevice, st));
  CG_CUDA(cudaEventRecord(t0, st));
  int rc = spmv_overlapped(op, op->r_ext, op->w, st);
  if (rc == FEMX_OK) rc = reduce2(op, r, op->w, st);
  if (rc != FEMX_OK) return fail(rc);
  if (iters > 0 && op->use_graph && (!op->graph || op->graph_x != d_x_owned)) {
    // one iteration, two streams (the halo travels on s_comm), captured once and replayed; the captured
    // iteration writes the x it was captured with
    if (op->graph) { cudaGraphExecDestroy(op->graph); op->graph = nullptr; }
    cudaGraph_t g = nullptr;
    CG_CUDA(cudaStreamSynchronize(st));
    CG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = cg_iteration(op, d_x_owned, st);
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc == FEMX_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&op->graph, g, 0) == cudaSuccess) {
      op->graph_x = d_x_owned;
    } else {  // e.g. an NCCL build that cannot be captured: plain launches
      (void)cudaGetLastError();
      op->graph = nullptr;
      op->use_graph = 0;
    }
    if (g) cudaGraphDestroy(g);
  }
  for (int it = 0; it < iters; ++it) {
    if (op->graph) {
      CG_CUDA(cudaGraphLaunch(op->graph, st));
    } else {
      rc = cg_iteration(op, d_x_owned, st);
      if (rc != FEMX_OK) return fail(rc);
    }
  }
  CG_CUDA(cudaEventRecord(t1, st));
  std::vector<double> hist(iters + 1);
  int halo_err = 0;
  CG_CUDA(cudaMemcpyAsync(hist.data(), op->d_hist, sizeof(double) * (iters + 1), cudaMemcpyDeviceToHost, st));
  if (op->push) CG_CUDA(cudaMemcpyAsync(&halo_err, op->halo_base + 32, sizeof(int), cudaMemcpyDeviceToHost, st));
  CG_CUDA(cudaStreamSynchronize(st));
  if (halo_err) {
    cudaMemsetAsync(op->halo_base + 32, 0, sizeof(int), st);
    return fail(femx_fail(ctx, FEMX_ERR_CUDA, "femx_dist_cg: a neighbour's ghost entries did not arrive (peer-memory halo timed out)"));
  }
  if (h_residuals)
    for (int k = 0; k <= iters; ++k) h_residuals[k] = std::sqrt(std::max(hist[k], 0.0));
  if (h_ms) cudaEventElapsedTime(h_ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return FEMX_OK;
#undef CG_CUDA
}

}  // extern "C"
