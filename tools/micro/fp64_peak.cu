// Practical fp64 issue rate of one B200 SM: independent DFMA / DMUL / DADD chains per thread,
// for a few (warps per SM, chains per thread) combinations.  Prints thread-instructions per clock per SM
// (the data sheet figure is 64).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIX>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MIX == 0) x[i] = fma(x[i], a, b);
      else if (MIX == 2) { x[i] = fma(x[(i + 1) % ILP], x[(i + 3) % ILP], x[i]); }   // three distinct register operands
      else if (MIX == 3) { x[i] = fma(x[(i + 1) % ILP], x[(i + 3) % ILP], x[i]); x[i] = __dmul_rn(x[(i + 2) % ILP], x[(i + 5) % ILP]); x[i] = __dadd_rn(x[(i + 4) % ILP], x[i]); }
      else if (MIX == 4) { x[i] = fma(x[(i + 1) % ILP], a, x[i]); }                    // two distinct registers + one shared by consecutive DFMAs (.reuse)
      else if (MIX == 5) { x[i] = __dmul_rn(x[(i + 1) % ILP], x[(i + 3) % ILP]); }       // DMUL, two distinct registers
      else if (MIX == 1) { x[i] = fma(x[i], a, b); x[i] = __dmul_rn(x[i], a); x[i] = fma(x[i], a, b); x[i] = __dadd_rn(x[i], b); x[i] = __dmul_rn(x[i], b); x[i] = fma(x[i], b, a); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, int MIX>
void run(int warps_per_sm, int sms, double clk_ghz) {
  int threads = 128, blocks_per_sm = warps_per_sm / 4;
  int blocks = blocks_per_sm * sms;
  double* d;
  cudaMalloc(&d, sizeof(double) * blocks * threads);
  int iters = 20000;
  k<ILP, MIX><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, MIX><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double per_thread = (double)iters * ILP * (MIX == 0 || MIX == 2 || MIX == 4 || MIX == 5 ? 1 : (MIX == 3 ? 3 : 6));
  double total = per_thread * threads * blocks;
  double per_clk_sm = total / (ms * 1e-3) / (clk_ghz * 1e9) / sms;
  printf("mix=%d warps/SM=%2d ILP=%2d : %.2f ms, %.1f DP thread-inst/clk/SM (at %.3f GHz), %.2f T inst/s\n", MIX, warps_per_sm, ILP, ms,
         per_clk_sm, clk_ghz, total / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ghz = clk_khz * 1e-6;
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, %.3f GHz\n", p.name, sms, ghz);
  run<1, 0>(16, sms, ghz); run<2, 0>(16, sms, ghz); run<4, 0>(16, sms, ghz); run<8, 0>(16, sms, ghz);
  run<4, 0>(8, sms, ghz); run<4, 0>(32, sms, ghz); run<8, 0>(64, sms, ghz);
  run<1, 1>(16, sms, ghz); run<2, 1>(16, sms, ghz); run<4, 1>(16, sms, ghz); run<8, 1>(16, sms, ghz); run<4, 1>(32, sms, ghz);
  run<8, 2>(16, sms, ghz); run<16, 2>(16, sms, ghz); run<8, 2>(20, sms, ghz); run<8, 3>(16, sms, ghz); run<16, 3>(16, sms, ghz); run<16, 3>(20, sms, ghz);
  run<8, 4>(16, sms, ghz); run<16, 4>(16, sms, ghz); run<8, 5>(16, sms, ghz); run<16, 5>(16, sms, ghz);
  return 0;
}
