// What HBM sustains for write-only, read-only and copy traffic on this part (the roofline denominator of bench.py is the
// COPY figure: bytes read + bytes written).  The assembly kernels are write-heavy (cfg3: 1.94 GB written, 0.57 GB read).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_rw hbm_rw.cu      Run: ./hbm_rw [GiB=2]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void write_k(double2* __restrict__ p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2(v, v);
}
__global__ void read_k(const double2* __restrict__ p, size_t n, double* out) {
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = p[i];
    s += v.x + v.y;
  }
  if (s == 123.456) *out = s;
}
__global__ void copy_k(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// 3 parts written for 1 part read: the mix of the cfg3 numeric pass
__global__ void mix_k(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = a[i];
    b[3 * i] = v; b[3 * i + 1] = v; b[3 * i + 2] = v;
  }
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 2.0;
  const size_t bytes = (size_t)(gib * (1ull << 30)), n = bytes / sizeof(double2);
  double2 *a, *b;
  double* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, 3 * bytes); cudaMalloc(&out, 8);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, 3 * bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 16, reps = 20;
  auto time = [&](const char* name, double moved, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %8.1f GB/s  (%.3f ms per pass, %.2f GB moved)\n", name, moved / (ms / reps * 1e-3) / 1e9, ms / reps, moved / 1e9);
  };
  time("write only (16 B stores)", (double)bytes, [&] { write_k<<<grid, 512>>>(a, n, 1.0); });
  time("cudaMemsetAsync", (double)bytes, [&] { cudaMemsetAsync(a, 1, bytes); });
  time("read only (16 B loads)", (double)bytes, [&] { read_k<<<grid, 512>>>(a, n, out); });
  time("copy (read + write)", 2.0 * bytes, [&] { copy_k<<<grid, 512>>>(a, b, n); });
  time("1 read : 3 written", 4.0 * bytes, [&] { mix_k<<<grid, 512>>>(a, b, n); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
