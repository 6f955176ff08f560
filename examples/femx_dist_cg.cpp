// femx_dist_cg — the multi-GPU layer from plain C++ (no Python, no torch, no MPI): one process per GPU,
//     ./femx_dist_cg <rank> <world> <id file> [n = 64] [iterations = 100]
// started once per rank with the same <id file> (rank 0 writes the NCCL unique id there, the others wait for it), e.g.
//     for r in 0 1 2 3; do ./femx_dist_cg $r 4 /tmp/femx.id 128 & done; wait
// Every rank assembles its z-slab of the n^3 Kuhn cube (grad u . grad v + u v) — rows of the GLOBAL matrix, no
// communication —, then solves A x = A 1 with the layer's CG: halo and the one reduction per iteration go through NVLink peer
// memory when the ranks can map each other's buffers, through NCCL otherwise.  Rank 0 prints the residual history's ends.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "femx.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int st_ = (call);                                                                                 \
    if (st_) { fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, #call, st_, femx_last_error(ctx)); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <rank> <world> <id file> [n] [iterations]\n", argv[0]); return 2; }
  const int rank = atoi(argv[1]), world = atoi(argv[2]);
  const char* idfile = argv[3];
  const int64_t n = argc > 4 ? atoll(argv[4]) : 64;
  const int iters = argc > 5 ? atoi(argv[5]) : 100;
  femx_ctx* ctx = nullptr;
  CHECK(femx_ctx_create(rank, &ctx));   // device = rank: one process per GPU of the node

  // the rank's slab: owned node planes [r0, r1), mesh planes [lo, hi] (one ghost cell layer per side)
  int64_t r0, r1, lo, hi;
  CHECK(femx_dist_slab(n + 1, world, rank, &r0, &r1, &lo, &hi));
  const int64_t plane = (n + 1) * (n + 1), n_nodes = (hi - lo + 1) * plane, n_elems = 6 * n * n * (hi - lo);
  double *dX, *dY, *dZ;
  int32_t* dConn;
  cudaMalloc(&dX, n_nodes * 8); cudaMalloc(&dY, n_nodes * 8); cudaMalloc(&dZ, n_nodes * 8);
  cudaMalloc(&dConn, n_elems * 4 * sizeof(int32_t));
  CHECK(femx_mesh_box(ctx, 0, 1, 0, 1, 0, 1, n, n, n, lo, hi, FEMX_F64, dX, dY, dZ, dConn, nullptr));

  // symbolic pass (rows of the global matrix: local rows [row_begin, row_end), columns + col_base) and numeric pass
  femx_pattern* pat = nullptr;
  CHECK(femx_pattern_build(ctx, 4, 1, n_nodes, n_elems, dConn, (r0 - lo) * plane, (r1 - lo) * plane, lo * plane, nullptr, &pat));
  int64_t n_rows, nnz;
  CHECK(femx_pattern_info(pat, &n_rows, &nnz, nullptr));
  femx_form_desc desc = {};
  desc.dim = 3; desc.nn = 4; desc.nd = 1; desc.dtype = FEMX_F64; desc.builtin = FEMX_FORM_POISSON_MASS; desc.params[0] = 1.0;
  desc.fmad = 1;
  femx_form* form = nullptr;
  CHECK(femx_form_compile(ctx, &desc, &form));
  femx_mesh_view mesh = {};
  mesh.dim = 3; mesh.nn = 4; mesh.n_nodes = n_nodes; mesh.n_elems = n_elems; mesh.d_conn = dConn;
  mesh.d_node_xyz[0] = dX; mesh.d_node_xyz[1] = dY; mesh.d_node_xyz[2] = dZ;
  double* dVals;
  cudaMalloc(&dVals, nnz * 8);
  CHECK(femx_assemble_csr(form, pat, &mesh, dVals, nullptr));

  // the communicator: rank 0 publishes the unique id through the file
  unsigned char id[FEMX_DIST_ID_BYTES] = {0};
  if (world > 1) {
    if (rank == 0) {
      CHECK(femx_dist_unique_id(id));
      char tmp[4096];
      snprintf(tmp, sizeof tmp, "%s.tmp", idfile);
      FILE* f = fopen(tmp, "wb");
      if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "cannot write %s\n", tmp); return 1; }
      fclose(f);
      rename(tmp, idfile);   // (atomic: the readers never see a partial file)
    } else {
      FILE* f = nullptr;
      for (int tries = 0; tries < 600 && !(f = fopen(idfile, "rb")); ++tries) usleep(100000);
      if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "rank %d: no id in %s\n", rank, idfile); return 1; }
      fclose(f);
    }
  }
  femx_dist* dd = nullptr;
  CHECK(femx_dist_create(ctx, rank, world, world > 1 ? id : nullptr, &dd));
  femx_dist_op* op = nullptr;
  CHECK(femx_dist_op_create(dd, pat, FEMX_F64, dVals, &op));
  int p2p = 0, peer_halo = 0;
  CHECK(femx_dist_info(dd, nullptr, nullptr, &p2p));
  CHECK(femx_dist_op_peer_halo(op, &peer_halo));

  // b = A 1 (so that the exact solution is 1), then CG from x0 = 0
  std::vector<double> ones(n_rows, 1.0), res(iters + 1);
  double *dOnes, *dB, *dSol;
  cudaMalloc(&dOnes, n_rows * 8); cudaMalloc(&dB, n_rows * 8); cudaMalloc(&dSol, n_rows * 8);
  cudaMemcpy(dOnes, ones.data(), n_rows * 8, cudaMemcpyHostToDevice);
  CHECK(femx_dist_spmv(op, dOnes, dB, nullptr));
  float ms = 0;
  CHECK(femx_dist_cg(op, dB, dSol, iters, res.data(), &ms, nullptr));
  if (rank == 0)
    printf("world %d, n %lld: %lld owned rows on rank 0, ||r_0|| = %.6e, ||r_%d|| = %.6e, %.3f ms (%.0f iterations/s); "
           "reduction: %s, halo: %s\n", world, (long long)n, (long long)n_rows, res[0], iters, res[iters], ms,
           iters / (ms * 1e-3), p2p ? "NVLink peer memory" : (world > 1 ? "ncclAllReduce" : "local"),
           peer_halo ? "NVLink peer memory" : (world > 1 ? "ncclSend/ncclRecv" : "none"));
  femx_dist_op_destroy(op);
  femx_dist_destroy(dd);
  femx_form_destroy(form);
  femx_pattern_destroy(pat);
  femx_ctx_destroy(ctx);
  if (rank == 0 && world > 1) unlink(idfile);
  return std::isfinite(res[iters]) && res[iters] < res[0] ? 0 : 1;
}
