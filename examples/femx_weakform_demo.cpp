// femx_weakform_demo — the reference's main() (fea_symbolic_nvrtc_sparse.cpp:484-503) with the C++ weak-form front end of
// include/femx_weakform.hpp: the weak form is written as two lambdas, compiled through NVRTC, assembled on the GPU, and
// compared with the built-in emitter (matrix) on a 2-D and a 3-D mesh.  Prints "ok" lines; exit code 0 = all equal to 1e-12.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "femx.h"
#include "femx_weakform.hpp"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int st_ = (call);                                                                                 \
    if (st_) { fprintf(stderr, "%s failed (%d): %s\n", #call, st_, femx_last_error(ctx)); return 1; } \
  } while (0)

using namespace femx::wf;

static double rel_diff(const std::vector<double>& a, const std::vector<double>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
  return std::sqrt(num / den);
}

int main() {
  femx_ctx* ctx = nullptr;
  CHECK(femx_ctx_create(0, &ctx));
  int bad = 0;
  for (int dim = 2; dim <= 3; ++dim) {
    const int nn = dim + 1;
    const long nx = 24, ny = 17, nz = 9;
    const long M = dim == 2 ? (nx + 1) * (ny + 1) : (nx + 1) * (ny + 1) * (nz + 1);
    const long NE = dim == 2 ? 2 * nx * ny : 6 * nx * ny * nz;
    double *dX, *dY, *dZ = nullptr;
    int32_t* dConn;
    cudaMalloc(&dX, M * 8); cudaMalloc(&dY, M * 8); cudaMalloc(&dZ, M * 8);
    cudaMalloc(&dConn, NE * nn * sizeof(int32_t));
    if (dim == 2) CHECK(femx_mesh_rectangle(ctx, -3, 3, -3, 3, ny, nx, 0, ny, FEMX_F64, dX, dY, nullptr, dConn, nullptr));
    else CHECK(femx_mesh_box(ctx, 0, 1, 0, 2, 0, 1, nx, ny, nz, 0, nz, FEMX_F64, dX, dY, dZ, dConn, nullptr));

    // ---- the weak form, as the reference writes it
    FunctionSpace fs(dim);
    Ex f = -2.0 * (fs.x * fs.x + fs.y * fs.y) + 36.0;
    WeakForm wf(fs);
    wf.build([&](Fn u, Fn v) { return dot(grad(u), grad(v)) + 2.5 * u * v; }, [&](Fn v) { return f * v; });
    femx_form_desc d = wf.desc(FEMX_F64);
    femx_form *form = nullptr, *ref = nullptr;
    CHECK(femx_form_compile(ctx, &d, &form));
    femx_form_desc b = femx_form_desc();
    b.dim = dim; b.nn = nn; b.nd = 1; b.dtype = FEMX_F64; b.builtin = FEMX_FORM_POISSON_MASS; b.params[0] = 2.5; b.fmad = 1;
    CHECK(femx_form_compile(ctx, &b, &ref));

    femx_pattern* pat = nullptr;
    CHECK(femx_pattern_build(ctx, nn, 1, M, NE, dConn, 0, M, 0, nullptr, &pat));
    int64_t n_rows, nnz, max_row;
    CHECK(femx_pattern_info(pat, &n_rows, &nnz, &max_row));
    femx_mesh_view m = femx_mesh_view();
    m.dim = dim; m.nn = nn; m.n_nodes = M; m.n_elems = NE; m.d_conn = dConn;
    m.d_node_xyz[0] = dX; m.d_node_xyz[1] = dY; m.d_node_xyz[2] = dim == 3 ? dZ : nullptr; m.node_stride = 1;
    double *dV, *dW, *dB;
    cudaMalloc(&dV, nnz * 8); cudaMalloc(&dW, nnz * 8); cudaMalloc(&dB, M * 8);
    CHECK(femx_assemble_csr(form, pat, &m, dV, nullptr));
    CHECK(femx_assemble_csr(ref, pat, &m, dW, nullptr));
    CHECK(femx_assemble_rhs(form, pat, &m, dB, nullptr));
    std::vector<double> v(nnz), w(nnz), rhs(M);
    cudaMemcpy(v.data(), dV, nnz * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(w.data(), dW, nnz * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(rhs.data(), dB, M * 8, cudaMemcpyDeviceToHost);
    const double err = rel_diff(v, w);
    double load = 0;
    for (double t : rhs) load += t;   // = integral of f over the domain
    // 2-D: int_{[-3,3]^2} (-2(x^2+y^2)+36) = 36*36 - 2*2*(6*18) = 864;  3-D box [0,1]x[0,2]x[0,1]: 72 - 2*(2/3 + 8/3) = 196/3
    const double want = dim == 2 ? 864.0 : 196.0 / 3.0;
    const bool ok = err <= 1e-12 && std::fabs(load - want) <= 1e-6 * std::fabs(want);   // (the 2-D default rule has 8-digit weights)
    printf("%s dim=%d nnz=%lld relF(lambda form vs built-in)=%.2e  sum(b)=%.9f (exact %.9f)\n", ok ? "ok" : "MISMATCH", dim,
           (long long)nnz, err, load, want);
    bad += !ok;
    femx_pattern_destroy(pat); femx_form_destroy(form); femx_form_destroy(ref);
    cudaFree(dX); cudaFree(dY); cudaFree(dZ); cudaFree(dConn); cudaFree(dV); cudaFree(dW); cudaFree(dB);
  }
  femx_ctx_destroy(ctx);
  return bad;
}
