#!/usr/bin/env python
"""Builds the REFERENCE ITSELF (its two compilable nvcc twins) for sm_100 into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (the authoring
container); the built .so files travel to the GPU box, the sources do not.  No reference
source is copied into the repository: this script reads the files where they lie, applies
the minimal fixes listed below IN MEMORY, appends a small extern "C" shim, and hands the text
to nvcc.  The generated translation units are written under oracle/_ref/gen/ (git-ignored
build products) so that nvcc errors can be read.

Sources compiled
  fea_test_sm_sym_sparse.cu   kernel K4: COO triplets            (operator surface #1)
  fea_test_sm_sym_sparse2.cu  kernel K5: ELL(7) + global atomicAdd, and the HOST pattern
                              builder Mesh::getNeighborNodesList   (operator surface #2)

Minimal fixes — nothing else is touched:
  size   MESH_W / MESH_H become run-time (__managed__) values instead of #defines, so that one
         build serves every mesh (the reference bakes the mesh size into the binary)
  Q2     COO kernel: localFlatMatrix is never zeroed            → thread x==0 zeroes its slot + barrier
  Q13    ELL kernel: every thread zeroes the whole array, no barrier → same fix as Q2
  Q3     node indices staged through `__shared__ float`         → int
  Q8     staging loads not guarded by gEleIdx < NE              → guarded
  Q4     chunked launches restart at element 0                  → the shim launches ONE 1-D grid
  main   renamed (the shim replaces assembleWithCuda's driver role)
Each build exists as written (fp32) and mechanically retyped to fp64 (float→double, powf→pow,
literal suffixes dropped).
"""
import os
import re
import subprocess
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sub_once(text, old, new, what):
    assert text.count(old) == 1, f"patch '{what}': expected exactly one match, found {text.count(old)}"
    return text.replace(old, new)


def patch_common(t, mesh_w_line, mesh_h_line):
    t = sub_once(t, mesh_w_line, "__device__ __managed__ long ref_mesh_w = 2;\n#define MESH_W ref_mesh_w\n", "MESH_W")
    t = sub_once(t, mesh_h_line, "__device__ __managed__ long ref_mesh_h = 2;\n#define MESH_H ref_mesh_h\n", "MESH_H")
    t = sub_once(t, "__shared__ float sGIdx[BLOCK_Z*NNODE];", "__shared__ int sGIdx[BLOCK_Z*NNODE];", "Q3")
    t = sub_once(t, "if(threadIdx.x==0 && threadIdx.y==0)\n", "if(threadIdx.x==0 && threadIdx.y==0 && gEleIdx < NE)\n", "Q8")
    zero = ("int lfmIdx = threadIdx.z*BLOCK_Y + threadIdx.y; //local flat matrix index of the integrand of threadIdx.y\n"
            "\tif(threadIdx.x == 0) localFlatMatrix[lfmIdx] = 0.0f; __syncthreads(); /* Q2/Q13 fix */\n")
    t = sub_once(t, "int lfmIdx = threadIdx.z*BLOCK_Y + threadIdx.y; //local flat matrix index of the integrand of threadIdx.y\n",
                 zero, "Q2/Q13")
    t = sub_once(t, "int main()", "int ref_main_unused()", "main")
    return t


def retype_fp64(t):
    t = t.replace("powf", "pow")
    t = re.sub(r"\bfloat\b", "double", t)
    t = re.sub(r"(\d+\.\d*)f\b", r"\1", t)
    t = t.replace("double elapsed = 0;", "float elapsed = 0;")  # host timer of the unused driver stays float
    return t


SHIM_COMMON = r'''
// ---------------------------------------------------------------- shim (not reference code)
#include <cstdint>
extern "C" void ref_set_mesh(long w, long h) { ref_mesh_w = w; ref_mesh_h = h; cudaDeviceSynchronize(); }
extern "C" void ref_set_mesh_host_only(long w, long h) { (void)w; (void)h; }
// The reference's own host mesh (RectangleMesh::generate) and flattening loop
// (assembleWithCuda: X[NNODE*i+k] = e->nodes[k]->x ...), returned to the caller.
extern "C" int ref_host_mesh(double x0, double x1, double y0, double y1, int nRow, int nCol,
                             double* nodeX, double* nodeY, int* flag, REAL* X, REAL* Y, int* gIdx) {
  RectangleMesh mesh(x0, x1, y0, y1, nRow, nCol);
  for (size_t i = 0; i < mesh.nodes.size(); i++) {
    if (nodeX) nodeX[i] = mesh.nodes[i]->x;
    if (nodeY) nodeY[i] = mesh.nodes[i]->y;
    if (flag) flag[i] = mesh.nodes[i]->flag;
  }
  for (size_t i = 0; i < mesh.elements.size(); i++) {
    Element* e = mesh.elements[i];
    for (int k = 0; k < 3; k++) {
      if (X) X[3 * i + k] = e->nodes[k]->x;
      if (Y) Y[3 * i + k] = e->nodes[k]->y;
      if (gIdx) gIdx[3 * i + k] = e->nodes[k]->index;
    }
  }
  return (int)mesh.elements.size();
}
static float ref_time_launches(int iters, cudaEvent_t a, cudaEvent_t b) {
  float ms = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); return ms / (iters > 0 ? iters : 1);
}
'''

SHIM_COO = r'''
// One 1-D grid (Q4), block (7,9,16) as the reference (fea_test_sm_sym_sparse.cu:263-265).
extern "C" int ref_assemble_coo(long ne, REAL* dA, int* dRow, int* dCol, REAL* dX, REAL* dY, int* dGIdx,
                                int iters, float* ms_per_launch) {
  dim3 blk(BLOCK_X, BLOCK_Y, BLOCK_Z);
  unsigned grid = (unsigned)((ne + BLOCK_Z - 1) / BLOCK_Z);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < (iters > 0 ? iters : 1); i++) fea_kernel<<<grid, blk>>>(dA, dRow, dCol, dX, dY, dGIdx);
  cudaEventRecord(b);
  float ms = ref_time_launches(iters, a, b);
  if (ms_per_launch) *ms_per_launch = ms;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return (int)cudaGetLastError();
}
'''

SHIM_ELL = r'''
// The reference's HOST pattern builder, untouched (Mesh::getNeighborNodesList).
extern "C" int ref_neighbor_list(int nRow, int nCol, int* len, int maxLen, int* idx) {
  RectangleMesh mesh(-3.0, 3.0, -3.0, 3.0, nRow, nCol);
  mesh.getNeighborNodesList(len, maxLen, idx);
  return (int)mesh.nodes.size();
}
// dA must be zeroed by the caller before EACH launch (the reference copies a zeroed host A:
// fea_test_sm_sym_sparse2.cu:314,359); with iters > 1 only the timing is meaningful.
extern "C" int ref_assemble_ell(long ne, REAL* dA, int* dLen, int* dIdx, REAL* dX, REAL* dY, int* dGIdx,
                                int iters, float* ms_per_launch) {
  dim3 blk(BLOCK_X, BLOCK_Y, BLOCK_Z);
  unsigned grid = (unsigned)((ne + BLOCK_Z - 1) / BLOCK_Z);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < (iters > 0 ? iters : 1); i++) fea_kernel<<<grid, blk>>>(dA, dLen, dIdx, dX, dY, dGIdx);
  cudaEventRecord(b);
  float ms = ref_time_launches(iters, a, b);
  if (ms_per_launch) *ms_per_launch = ms;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return (int)cudaGetLastError();
}
'''


def build_one(src_name, out_stem, shim, mesh_w_line, mesh_h_line):
    text = open(os.path.join(REF, src_name)).read()
    text = patch_common(text, mesh_w_line, mesh_h_line)
    if "sparse2" in src_name:
        text = sub_once(text, "  for(int i=0; i<BLOCK_Y*BLOCK_Z; i++) localFlatMatrix[i] = 0.0f;\n", "", "Q13 (racy zero loop)")
    for prec in ("f32", "f64"):
        t = text if prec == "f32" else retype_fp64(text)
        real = "float" if prec == "f32" else "double"
        unit = t + f"\n#define REAL {real}\n" + SHIM_COMMON + shim
        gen = os.path.join(GEN, f"{out_stem}_{prec}.cu")
        open(gen, "w").write(unit)
        so = os.path.join(OUT, f"lib{out_stem}_{prec}.so")
        cmd = [NVCC, "-gencode", "arch=compute_100,code=sm_100", "-O2", "-std=c++14", "-w", "-shared",
               "-Xcompiler", "-fPIC", "-cudart", "static", "-o", so, gen]
        subprocess.check_call(cmd)
        print("built", os.path.relpath(so, HERE))


SHIM_ATOMIC = r'''
// ---- shim (not reference code): times the reference's three accumulation variants on n values ----------
// (atomicadd.cu:73-129: naive global atomicAdd, shared-memory staged block sum + one atomicAdd per block, CAS-loop double add;
//  its main() launches only the first on SIZE = 50)
// n_cas: the CAS-loop variant serialises completely (every resident thread retries until it wins: work ~ n x resident threads),
// it is timed on a much smaller n
extern "C" int ref_atomic_variants(long n, long n_cas, int iters, float* ms3, double* results3) {
  float* dIn = 0; double* dInD = 0; float* dRes = 0; double* dResD = 0;
  cudaMalloc(&dIn, n * sizeof(float)); cudaMalloc(&dInD, n * sizeof(double));
  cudaMalloc(&dRes, sizeof(float)); cudaMalloc(&dResD, sizeof(double));
  float* h = (float*)malloc(n * sizeof(float)); double* hd = (double*)malloc(n * sizeof(double));
  for (long i = 0; i < n; i++) { h[i] = 1.0f; hd[i] = 1.0; }
  cudaMemcpy(dIn, h, n * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(dInD, hd, n * sizeof(double), cudaMemcpyHostToDevice);
  dim3 blk(BLOCK_X_NAIVE, BLOCK_Y_NAIVE, 1);
  dim3 grd(BLOCK_COUNT_X, (unsigned)((n + (long)BLOCK_X_NAIVE * BLOCK_Y_NAIVE * BLOCK_COUNT_X - 1) / ((long)BLOCK_X_NAIVE * BLOCK_Y_NAIVE * BLOCK_COUNT_X)), 1);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int v = 0; v < 3; v++) {
    float best = 1e30f;
    ref_size = v == 2 ? n_cas : n;
    if (v == 2) grd.y = (unsigned)((n_cas + (long)BLOCK_X_NAIVE * BLOCK_Y_NAIVE * BLOCK_COUNT_X - 1) / ((long)BLOCK_X_NAIVE * BLOCK_Y_NAIVE * BLOCK_COUNT_X));
    for (int it = 0; it < iters; it++) {
      cudaMemset(dRes, 0, sizeof(float)); cudaMemset(dResD, 0, sizeof(double));
      cudaEventRecord(a);
      if (v == 0) reductionKernel<<<grd, blk>>>(dRes, dIn);
      if (v == 1) reductionKernel2<<<grd, blk>>>(dRes, dIn);
      if (v == 2) reductionKernel3<<<grd, blk>>>(dResD, dInD);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    ms3[v] = best;
    if (v < 2) { float r; cudaMemcpy(&r, dRes, sizeof r, cudaMemcpyDeviceToHost); results3[v] = r; }
    else cudaMemcpy(&results3[v], dResD, sizeof(double), cudaMemcpyDeviceToHost);
  }
  int err = (int)cudaGetLastError();
  cudaFree(dIn); cudaFree(dInD); cudaFree(dRes); cudaFree(dResD); free(h); free(hd);
  return err;
}
'''


def build_atomicadd():
    """atomicadd.cu: the three accumulation variants as a timed side baseline.  In-memory edits: SIZE becomes a
    run-time value, main is renamed, the per-thread debug printf of reductionKernel2 is removed."""
    t = open(os.path.join(REF, "atomicadd.cu")).read()
    t = sub_once(t, "#define SIZE 50\n", "__device__ __managed__ long ref_size = 50;\n#define SIZE ref_size\n", "SIZE")
    t = sub_once(t, "int main()", "int ref_main_unused()", "main")
    t = sub_once(t, '    printf("(%d %d) (%d %d) (%d %d)\\n",blockDim.x, blockDim.y, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);\n', "", "debug printf")
    gen = os.path.join(GEN, "ref_atomicadd.cu")
    open(gen, "w").write(t + SHIM_ATOMIC)
    so = os.path.join(OUT, "libref_atomicadd.so")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100,code=sm_100", "-O2", "-std=c++14", "-w", "-shared",
                           "-Xcompiler", "-fPIC", "-cudart", "static", "-o", so, gen])
    print("built", os.path.relpath(so, HERE))


def main():
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: /root/reference not present; keeping prebuilt oracle/_ref as is")
        return 0
    os.makedirs(GEN, exist_ok=True)
    build_one("fea_test_sm_sym_sparse.cu", "ref_coo", SHIM_COO, "#define MESH_W 10000\n", "#define MESH_H 1000\n")
    build_one("fea_test_sm_sym_sparse2.cu", "ref_ell", SHIM_ELL, "#define MESH_W 1000L\n", "#define MESH_H 100L\n")
    # The reference's own PROGRAM, completely unmodified (its main(), its 1000x100 mesh, its racy zero
    # loop Q13 and all): its stdout is diffed against examples/femx_sparse2 on the GPU box.
    exe = os.path.join(OUT, "fea_test_sm_sym_sparse2")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100,code=sm_100", "-O2", "-std=c++14", "-w", "-cudart", "static",
                           "-o", exe, os.path.join(REF, "fea_test_sm_sym_sparse2.cu")])
    print("built", os.path.relpath(exe, HERE), "(unmodified reference program)")
    build_atomicadd()
    return 0


if __name__ == "__main__":
    sys.exit(main())
