"""CPU tier: the element-once lattice pass (femx_jit_src.h: kFemxJitLattice + the macros femx_lattice.cpp
generates), compiled for the HOST and run by a small CTA emulator.

The generated CUDA source is compiled as is with g++ (-ffp-contract=off; FEMX_HOST_EMU swaps the handful of
device-only pieces: thread / block ids, __syncthreads, the bulk store).  The emulator runs every CTA with one OS
thread per CUDA thread and a std::barrier for __syncthreads, shared memory is one buffer per CTA — so the
inter-thread exchange (fields, parity buffers, halo columns, run detection of the image stores) is executed
exactly as written.  Checked against the oracle on jittered Kuhn boxes: every interior (class) row within
1e-12, bitwise symmetric, and untouched outside the class rows."""
import os
import subprocess

import numpy as np
import pytest

import femx
from oracle import oracle as orc
from tools.lattice_offline import kuhn_corners, kuhn_offsets

HARNESS_PRE = r'''
#define FEMX_HOST_EMU 1
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <thread>
#include <vector>
using std::max;
using std::min;
#define __device__
#define __forceinline__ inline
#define __restrict__
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double femx_mul(double a, double b) { return a * b; }   // -ffp-contract=off: one rounding
static inline float femx_mul(float a, float b) { return a * b; }
static inline double femx_rcp(double a) { return 1.0 / a; }
static inline float femx_rcp(float a) { return 1.0f / a; }
static thread_local int emu_tid = 0;
static int emu_bid = 0;
static unsigned char* emu_smem = nullptr;
static std::barrier<>* emu_bar = nullptr;              // the CTA's mbarrier: split arrive / wait
static std::vector<std::barrier<>*> emu_warp_bar;       // __syncwarp
static std::vector<int> emu_flags;                      // ballot scratch
static thread_local std::optional<std::barrier<>::arrival_token> emu_tok;
#define FEMX_TID emu_tid
#define FEMX_BID emu_bid
#define FEMX_LT_KERNEL static void
static inline unsigned char* femx_emu_smem() { return emu_smem; }
static inline void femx_lt_fence() {}
static inline void femx_lt_bulk_wait() {}
static inline void femx_lt_prefetch(const void*) {}
static inline void femx_lt_bar_init(void*, int) {}
static inline void femx_lt_bar_arrive(void*) { emu_tok.emplace(emu_bar->arrive()); }
static inline void femx_lt_bar_wait(void*, int) { emu_bar->wait(std::move(*emu_tok)); emu_tok.reset(); }
static inline void femx_lt_syncwarp() { emu_warp_bar[emu_tid >> 5]->arrive_and_wait(); }
static inline unsigned femx_lt_ballot(int pred) {
  emu_flags[emu_tid] = pred ? 1 : 0;
  femx_lt_syncwarp();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)emu_flags[(emu_tid & ~31) + l] << l;
  femx_lt_syncwarp();
  return m;
}
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
'''

HARNESS_POST = r'''
static inline void femx_lt_bulk_store(real* dst, const real* src, unsigned bytes) {
  if (((size_t)dst | (size_t)src | bytes) & 15) { fprintf(stderr, "misaligned bulk store\n"); abort(); }
  memcpy(dst, src, bytes);
}
'''

HARNESS_MAIN = r'''
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int hdr[16];
  if (fread(hdr, 4, 16, f) != 16) return 2;
  const int n_nodes = hdr[0], n_rows = hdr[1], row0 = hdr[2], kc = hdr[9], smem = hdr[10], nnz = hdr[11];
  femx_lat lat;
  lat.cnx = hdr[3]; lat.cny = hdr[4]; lat.cnz = hdr[5]; lat.sy = hdr[6]; lat.sz = hdr[7]; lat.node0 = hdr[8];
  lat.klo = hdr[12]; lat.khi = hdr[13];
  lat.ntx = (lat.cnx - 1 + FEMX_LT_TX - 2) / (FEMX_LT_TX - 1);
  lat.nty = (lat.cny - 1 + FEMX_LT_TY - 2) / (FEMX_LT_TY - 1);
  lat.kc = kc;
  std::vector<double> tmp(n_nodes);
  std::vector<real> C[3];
  for (int c = 0; c < 3; ++c) {
    if (fread(tmp.data(), 8, n_nodes, f) != (size_t)n_nodes) return 2;
    C[c].assign(tmp.begin(), tmp.end());
  }
  std::vector<int2> rowinfo(n_rows + 1);
  if (fread(rowinfo.data(), 8, n_rows + 1, f) != (size_t)n_rows + 1) return 2;
  fclose(f);
  // one slack element in front so that the values start 8 bytes off a 16-byte boundary when hdr[14] says so
  std::vector<real> store(nnz + 8, real(-777));
  size_t base = 0;
  while (((size_t)(store.data() + base)) & 15) ++base;
  real* vals = store.data() + base;
  const int ntz = lat.khi >= lat.klo ? (lat.khi - lat.klo + 1 + kc - 1) / kc : 0;
  const int blocks = lat.ntx * lat.nty * ntz;
  std::vector<unsigned char> sm(smem + 256);
  emu_smem = sm.data();
  while ((size_t)emu_smem & 127) ++emu_smem;
  for (int b = 0; b < blocks; ++b) {
    emu_bid = b;
    std::barrier<> bar(LT_NT);
    emu_bar = &bar;
    emu_flags.assign(LT_NT, 0);
    for (auto* w : emu_warp_bar) delete w;
    emu_warp_bar.clear();
    for (int w = 0; w < LT_NT / 32; ++w) emu_warp_bar.push_back(new std::barrier<>(32));
    std::vector<std::thread> th;
    for (int t = 0; t < LT_NT; ++t)
      th.emplace_back([&, t] {
        emu_tid = t;
        femx_csr(rowinfo.data(), nullptr, nullptr, nullptr, nullptr, C[0].data(), C[1].data(), C[2].data(), 1, vals,
                 n_rows, row0, lat);
      });
    for (auto& x : th) x.join();
  }
  FILE* out = fopen(argv[2], "wb");
  std::vector<double> o(vals, vals + nnz);
  fwrite(o.data(), 8, nnz, out);
  fclose(out);
  return 0;
}
'''


def jitter_box(nx, ny, nz, seed=12345, amp=0.2):
    X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
    rng = np.random.default_rng(seed)
    h = 1.0 / max(nx, ny, nz)
    i = np.arange(len(X)) % (nx + 1)
    j = (np.arange(len(X)) // (nx + 1)) % (ny + 1)
    k = np.arange(len(X)) // ((nx + 1) * (ny + 1))
    inner = (i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz)
    for C_ in (X, Y, Z):
        C_ += np.where(inner, rng.uniform(-amp * h, amp * h, len(X)), 0.0)
    return X, Y, Z, conn, inner


def run_emu(tmp_path, builtin, dims, tile, kc, dtype=femx.F64, row_range=None):
    nx, ny, nz = dims
    X, Y, Z, conn, inner = jitter_box(nx, ny, nz)
    sy, sz = nx + 1, (nx + 1) * (ny + 1)
    offs, self_pos = kuhn_offsets(sy, sz)
    os.environ["FEMX_LT_TX"], os.environ["FEMX_LT_TY"] = str(tile[0]), str(tile[1])
    try:
        form = femx.Form(None, 3, getattr(femx, builtin), dtype=dtype, params=(1.5,), offline=True)
        _, info = form.cubin_lattice(kuhn_corners(), sy, sz, offs, self_pos)
        src = form.source
        form.close()
    finally:
        os.environ.pop("FEMX_LT_TX"), os.environ.pop("FEMX_LT_TY")
    assert (info["tx"], info["ty"]) == tuple(tile)
    cpp = tmp_path / "emu.cpp"
    # the harness pieces go around the generated source: prelude | source up to the kernel | bulk store | kernel | main
    cut = src.index("FEMX_LT_KERNEL\nfemx_csr")
    cpp.write_text(HARNESS_PRE + src[:cut] + HARNESS_POST + src[cut:] + HARNESS_MAIN)
    exe = tmp_path / "emu"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-o", str(exe), str(cpp)])
    n_nodes = len(X)
    rp, ci = orc.pattern(conn, n_nodes)
    r_lo, r_hi = row_range or (0, n_nodes)
    n_rows = r_hi - r_lo
    rowinfo = np.zeros((n_rows + 1, 2), np.int32)
    rowinfo[:, 0] = rp[r_lo:r_hi + 1] - rp[r_lo]
    rowinfo[:n_rows, 1] = np.where(inner[r_lo:r_hi], 24 | (1 << 23), 1)
    nnz = int(rp[r_hi] - rp[r_lo])
    klo = max(1, r_lo // sz)
    khi = min(nz - 1, (r_hi - 1) // sz)
    hdr = np.array([n_nodes, n_rows, r_lo, nx, ny, nz, sy, sz, 0, kc, info["smem"], nnz, klo, khi, 0, 0], np.int32)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as fh:
        fh.write(hdr.tobytes())
        for C_ in (X, Y, Z):
            fh.write(np.ascontiguousarray(C_, np.float64).tobytes())
        fh.write(rowinfo.tobytes())
    outp = tmp_path / "out.bin"
    subprocess.check_call([str(exe), str(inp), str(outp)])
    got = np.fromfile(outp, np.float64)
    ref = orc.assemble_csr(getattr(orc, builtin), 3, 1, conn, X, Y, Z, rp, ci, params=(1.5,))
    return got, ref, rp, ci, inner, (r_lo, r_hi)


@pytest.mark.parametrize("builtin", ["POISSON_MASS", "POISSON", "MASS"])
@pytest.mark.parametrize("dims,tile,kc", [((6, 5, 4), (4, 3), 2), ((9, 7, 5), (8, 5), 8), ((5, 5, 6), (16, 16), 3), ((40, 3, 3), (32, 3), 1)])
def test_lattice_pass_on_host_equals_oracle(tmp_path, builtin, dims, tile, kc):
    got, ref, rp, ci, inner, _ = run_emu(tmp_path, builtin, dims, tile, kc)
    rows = np.flatnonzero(inner)
    assert len(rows) == (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    mask = np.zeros(len(ref), bool)
    for r in rows:
        assert rp[r + 1] - rp[r] == 15
        mask[rp[r]:rp[r + 1]] = True
    err = np.linalg.norm(got[mask] - ref[mask]) / np.linalg.norm(ref[mask])
    assert err <= 1e-12, err
    assert np.all(got[~mask] == -777.0)          # nothing outside the class rows is touched
    # A(p,q) and A(q,p) are the same sum of the same fields: bitwise symmetric among class rows
    pos = {}
    for r in rows:
        for k in range(rp[r], rp[r + 1]):
            pos[(r, ci[k])] = k
    for (r, c), k in pos.items():
        if (c, r) in pos:
            assert got[k] == got[pos[(c, r)]]


def test_lattice_pass_on_host_slab_rows(tmp_path):
    """Owned rows [row_begin, row_end) in the middle of the mesh: same bits as the whole-mesh run."""
    dims, tile, kc = (5, 4, 7), (4, 4), 2
    whole, ref, rp, ci, inner, _ = run_emu(tmp_path, "POISSON_MASS", dims, tile, kc)
    sz = (dims[0] + 1) * (dims[1] + 1)
    lo, hi = 2 * sz, 5 * sz
    part, _, _, _, _, _ = run_emu(tmp_path, "POISSON_MASS", dims, tile, kc, row_range=(lo, hi))
    assert np.array_equal(part, whole[rp[lo]:rp[hi]])


def test_lattice_pass_on_host_fp32(tmp_path):
    got, ref, rp, ci, inner, _ = run_emu(tmp_path, "POISSON_MASS", (6, 6, 4), (8, 4), 4, dtype=femx.F32)
    rows = np.flatnonzero(inner)
    mask = np.zeros(len(ref), bool)
    for r in rows:
        mask[rp[r]:rp[r + 1]] = True
    err = np.linalg.norm(got[mask] - ref[mask]) / np.linalg.norm(ref[mask])
    assert err <= 1e-5, err
