// femx_sparse2 — the reference's fea_test_sm_sym_sparse2 / fea_symbolic_nvrtc_sparse2 main()
// re-written against the femx C ABI (plain C++, CUDA runtime only: no Python, no torch).
//
//   build:  g++ -O2 -std=c++14 examples/femx_sparse2.cpp -Iinclude -I/usr/local/cuda/include \
//               -Lcuda-fem_b200 -lfemx -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuda-fem_b200 -o femx_sparse2
//   run:    ./femx_sparse2 [MESH_W MESH_H]          (default 1000 100, the reference's configuration)
//
// Flow (reference lines: fea_test_sm_sym_sparse2.cu):
//   RectangleMesh(-3,3,-3,3,MESH_W,MESH_H)            :302   → femx_mesh_rectangle (device)
//   mesh.getNeighborNodesList(...)                    :331   → femx_pattern_build + femx_pattern_export_ell
//   fea_kernel<<<...>>>(dA, dNbrNodeLen, ...)         :368-372 → femx_assemble_csr + femx_csr_to_ell
//   print the first 16 rows as "(i,col) val"          :389-394 → same format, so outputs can be diffed
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "femx.h"

#define CHECK(call)                                                                        \
  do {                                                                                     \
    int st_ = (call);                                                                      \
    if (st_) { fprintf(stderr, "%s failed (%d): %s\n", #call, st_, femx_last_error(ctx)); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  const long MESH_W = argc > 2 ? atol(argv[1]) : 1000, MESH_H = argc > 2 ? atol(argv[2]) : 100;
  const int fp32 = argc > 3 ? atoi(argv[3]) : 1;  // the reference computes in float
  const long M = (MESH_W + 1) * (MESH_H + 1), NE = 2 * MESH_W * MESH_H;
  const int MAX_NEIGHBOR = 7;
  const size_t rs = fp32 ? 4 : 8;
  femx_ctx* ctx = nullptr;
  CHECK(femx_ctx_create(0, &ctx));

  void *dX, *dY, *dVals, *dA;
  int32_t *dConn, *dLen, *dIdx;
  cudaMalloc(&dX, M * rs); cudaMalloc(&dY, M * rs);
  cudaMalloc(&dConn, NE * 3 * sizeof(int32_t));
  CHECK(femx_mesh_rectangle(ctx, -3.0, 3.0, -3.0, 3.0, MESH_W, MESH_H, 0, MESH_W, fp32 ? FEMX_F32 : FEMX_F64,
                            dX, dY, nullptr, dConn, nullptr));

  femx_form_desc d = {};
  d.dim = 2; d.nn = 3; d.nd = 1; d.dtype = fp32 ? FEMX_F32 : FEMX_F64;
  d.builtin = FEMX_FORM_POISSON; d.fmad = 1;
  femx_form* form = nullptr;
  CHECK(femx_form_compile(ctx, &d, &form));

  femx_pattern* pat = nullptr;
  CHECK(femx_pattern_build(ctx, 3, 1, M, NE, dConn, 0, M, 0, nullptr, &pat));
  int64_t n_rows, nnz, max_row;
  CHECK(femx_pattern_info(pat, &n_rows, &nnz, &max_row));
  cudaMalloc(&dLen, M * sizeof(int32_t)); cudaMalloc(&dIdx, M * MAX_NEIGHBOR * sizeof(int32_t));
  CHECK(femx_pattern_export_ell(pat, MAX_NEIGHBOR, dLen, dIdx, nullptr));

  femx_mesh_view m = {};
  m.dim = 2; m.nn = 3; m.n_nodes = M; m.n_elems = NE; m.d_conn = dConn;
  m.d_node_xyz[0] = dX; m.d_node_xyz[1] = dY; m.node_stride = 1;
  cudaMalloc(&dVals, nnz * rs); cudaMalloc(&dA, M * MAX_NEIGHBOR * rs);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CHECK(femx_assemble_csr(form, pat, &m, dVals, nullptr));  // warm-up (JIT)
  cudaEventRecord(e0);
  CHECK(femx_assemble_csr(form, pat, &m, dVals, nullptr));
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  CHECK(femx_csr_to_ell(pat, fp32 ? FEMX_F32 : FEMX_F64, MAX_NEIGHBOR, dVals, dA, nullptr));

  std::vector<int32_t> len(M), idx(M * MAX_NEIGHBOR);
  std::vector<float> Af(fp32 ? M * MAX_NEIGHBOR : 0);
  std::vector<double> Ad(fp32 ? 0 : M * MAX_NEIGHBOR);
  cudaMemcpy(len.data(), dLen, M * sizeof(int32_t), cudaMemcpyDeviceToHost);
  cudaMemcpy(idx.data(), dIdx, M * MAX_NEIGHBOR * sizeof(int32_t), cudaMemcpyDeviceToHost);
  cudaMemcpy(fp32 ? (void*)Af.data() : (void*)Ad.data(), dA, M * MAX_NEIGHBOR * rs, cudaMemcpyDeviceToHost);

  printf("mesh = %ld*%ld\n", MESH_W, MESH_H);
  printf("Number of nodes = %ld, nnz = %lld, longest row = %lld\n", M, (long long)nnz, (long long)max_row);
  printf("GPU Time: %gms\n", ms);
  for (long i = 0; i < 16 && i < M; i++) {
    for (int j = 0; j < len[i]; j++)
      printf("(%ld,%d) %g    ", i, idx[i * MAX_NEIGHBOR + j], fp32 ? (double)Af[i * MAX_NEIGHBOR + j] : Ad[i * MAX_NEIGHBOR + j]);
    printf("\n");
  }
  femx_form_destroy(form); femx_pattern_destroy(pat); femx_ctx_destroy(ctx);
  return 0;
}
