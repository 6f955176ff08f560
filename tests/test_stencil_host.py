"""CPU tier: the generated numeric-pass code, compiled for the HOST.

The emitter's output for a stencil class — FEMX_PROLOGUE, the row macros, FEMX_CSR_CASES (one incidence of
the generic loop) and FEMX_SPEC_LOAD / FEMX_SPEC_BODY (the specialised straight-line body) — is plain C over
fma / femx_mul.  This test compiles that very text with g++ (-ffp-contract=off, so every rounding is the one
the text spells out) around a small harness and checks, for every interior row of a jittered structured mesh,
that the specialised body produces the SAME BITS as the generic incidence loop, and that both agree with
the oracle's CSR values.  It pins the codegen (positions, first-touch flags, shared face cross products and
their signs, accumulate form, streaming order) without a GPU."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import femx
from oracle import oracle as orc
from tools.stencil_offline import row_codes

HARNESS = r'''
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef long long i64;
%(defines)s
typedef FEMX_REAL real;
using std::max;
using std::min;
static inline real femx_mul(real a, real b) { return a * b; }   // -ffp-contract=off: exactly one rounding
static inline real femx_rcp(real a) { return real(1) / a; }     // (the same function on both paths)
template <class T> static inline T femx_pow(T a, double e) { return (T)std::pow((double)a, e); }
#define pow femx_pow
#define powf femx_pow
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T femx_ldg_pinned(const T* p) { return *p; }
#define NDOF (NN * ND)
#define FEMX_CS 1
#define FEMX_FIRST(J) ((code >> (21 + (J))) & 1u)
#define FEMX_GATHER_NEXT
struct femx_soff { int v[24]; };

int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int hdr[6];  // n_nodes, n_rows, rlen, np, self, unused
  if (fread(hdr, 4, 6, f) != 6) return 2;
  const int n_nodes = hdr[0], n_rows = hdr[1], rlen = hdr[2], np = hdr[3], self = hdr[4];
  std::vector<real> Xv(n_nodes), Yv(n_nodes), Zv(n_nodes);
  std::vector<double> tmp(n_nodes);
  real* coord[3] = {Xv.data(), Yv.data(), Zv.data()};
  for (int c = 0; c < 3; ++c) {
    if (fread(tmp.data(), 8, n_nodes, f) != (size_t)n_nodes) return 2;
    for (int i = 0; i < n_nodes; ++i) coord[c][i] = (real)tmp[i];
  }
  const real *X = Xv.data(), *Y = Yv.data(), *Z = Zv.data();
  (void)Z;
  std::vector<unsigned> codes(np);
  if (fread(codes.data(), 4, np, f) != (size_t)np) return 2;
  std::vector<int> rows(n_rows), cols((size_t)n_rows * rlen);
  if (fread(rows.data(), 4, n_rows, f) != (size_t)n_rows) return 2;
  if (fread(cols.data(), 4, cols.size(), f) != cols.size()) return 2;
  fclose(f);
  FILE* out = fopen(argv[2], "wb");
  for (int r = 0; r < n_rows; ++r) {
    const int* scol = cols.data() + (size_t)r * rlen;
    std::vector<real> g(rlen, real(NAN)), s(rlen, real(NAN));
    {  // ---- the generic incidence loop (skeleton of femx_generic_row; the per-incidence code is the product's)
      real* srow = g.data();
      const int rstride = rlen * ND;
      (void)rstride;
      const real sx = X[scol[self]], sy = Y[scol[self]], sz = DIM == 3 ? Z[scol[self]] : real(0);
      (void)sz;
      real dacc[ND * ND] = {real(0)};
      for (int it = 0; it < np; ++it) {
        const unsigned code = codes[it];
        real ox[NN - 1], oy[NN - 1], oz[NN - 1];
        int po[NN - 1];
        for (int j = 0; j < NN - 1; ++j) {
          const int p = (code >> (7 * j)) & 127;
          ox[j] = X[scol[p]]; oy[j] = Y[scol[p]]; oz[j] = DIM == 3 ? Z[scol[p]] : real(0);
          po[j] = p * ND;
        }
        (void)oz;
        switch (FEMX_ROTINV ? 0 : (int)((code >> 28) & 3)) {
          FEMX_CSR_CASES
        }
      }
#if FEMX_ROWSUM
      {  // (femx_generic_row's row-sum diagonal)
        real S_ = real(0);
        for (int k = 0; k < rlen; ++k)
          if (k != self) S_ += srow[k];
        dacc[0] = fma(FEMX_CJ, dacc[0], -S_);
      }
#endif
      srow[self] = dacc[0];
    }
    {  // ---- the specialised body, exactly as the kernel runs it
      real* srow = s.data();
      const int node_ = rows[r], node_max = n_nodes - 1;
      const bool mine = true;
      femx_soff soff;
      for (int k = 0; k < rlen; ++k) soff.v[k] = scol[k] - rows[r];
      FEMX_SPEC_LOAD
      FEMX_SPEC_BODY
    }
    for (int k = 0; k < rlen; ++k) { double v = (double)g[k]; fwrite(&v, 8, 1, out); }
    for (int k = 0; k < rlen; ++k) { double v = (double)s[k]; fwrite(&v, 8, 1, out); }
  }
  fclose(out);
  return 0;
}
'''


def _run_case(tmp_path, dim, form, dtype="f64", env=None):
    rng = np.random.RandomState(4)
    if dim == 2:
        n = 6
        X, Y, _, conn = orc.rect_mesh(0, 1, 0, 2, n, n)
        Z = np.zeros_like(X)
        m = n + 1
        interior = [i * m + j for i in range(1, n) for j in range(1, n)]
        nn = 3
    else:
        n = 5
        X, Y, Z, conn = orc.box_mesh(n, n, n)
        m = n + 1
        interior = [(k * m + j) * m + i for k in range(1, n) for j in range(1, n) for i in range(1, n)]
        nn = 4
    X = X + rng.uniform(-0.02, 0.02, X.shape)
    Y = Y + rng.uniform(-0.02, 0.02, Y.shape)
    if dim == 3:
        Z = Z + rng.uniform(-0.02, 0.02, Z.shape)
    rp, ci = orc.pattern(conn, len(X))
    codes, rlen, self_pos = row_codes(conn, nn, interior[len(interior) // 2])
    for r in interior:   # every interior row is in the class
        assert row_codes(conn, nn, r) == (codes, rlen, self_pos)
    cols = np.array([ci[rp[r]:rp[r + 1]] for r in interior], dtype=np.int32)
    assert cols.shape == (len(interior), rlen)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        form.cubin_stencil(codes, rlen, self_pos)
        src = form.source
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    defines = src[:src.index("typedef FEMX_REAL real;")]
    if dtype == "f32":
        assert "#define FEMX_REAL float" in defines
    cpp = tmp_path / "harness.cpp"
    cpp.write_text(HARNESS % {"defines": defines})
    exe = tmp_path / "harness"
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-w", "-o", str(exe), str(cpp)])
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("6i", len(X), len(interior), rlen, len(codes), self_pos, 0))
        for a in (X, Y, Z):
            f.write(np.asarray(a, np.float64).tobytes())
        f.write(np.asarray(codes, np.uint32).tobytes())
        f.write(np.asarray(interior, np.int32).tobytes())
        f.write(cols.tobytes())
    outp = tmp_path / "out.bin"
    subprocess.check_call([str(exe), str(inp), str(outp)])
    res = np.fromfile(outp, np.float64).reshape(len(interior), 2, rlen)
    gen, spec = res[:, 0, :], res[:, 1, :]
    assert not np.isnan(gen).any() and not np.isnan(spec).any()      # every slot written on both paths
    assert np.array_equal(gen.view(np.uint64), spec.view(np.uint64)), "specialised body and generic loop differ"
    return X, Y, Z, conn, rp, ci, interior, gen


@pytest.mark.parametrize("dim,builtin", [(2, "POISSON"), (2, "POISSON_MASS"), (2, "MASS"),
                                         (3, "POISSON"), (3, "POISSON_MASS"), (3, "MASS")])
@pytest.mark.parametrize("env", [None, {"FEMX_SPEC_AHEAD": "99"}, {"FEMX_SHAREDFACES": "0", "FEMX_ACCF": "0"},
                                 {"FEMX_SPEC_AHEAD": "0"}, {"FEMX_ROWSUM": "1", "FEMX_RCP3": "1"},
                                 {"FEMX_CHAINORDER": "1"}, {"FEMX_CHAINORDER": "1", "FEMX_ROWSUM": "1"}])
def test_generated_specialised_body_equals_generic_loop_on_the_host(tmp_path, dim, builtin, env):
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})          # some knobs are read when the form is created
    try:
        form = femx.Form(None, dim, getattr(femx, builtin), params=(1.5,), offline=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if env and env.get("FEMX_ROWSUM") == "1" and dim == 3 and builtin != "MASS":
        form.cubin_stencil(*__import__("tools.stencil_offline", fromlist=["x"]).interior_class(3))
        assert "#define FEMX_ROWSUM 1" in form.source
    X, Y, Z, conn, rp, ci, interior, gen = _run_case(tmp_path, dim, form, env=env)
    form.close()
    # ... and both are the operator the oracle assembles (host reciprocal = IEEE division here)
    ov = orc.assemble_csr(getattr(orc, builtin), dim, 1, conn, X, Y, Z if dim == 3 else None, rp, ci, params=(1.5,))
    ref = np.array([ov[rp[r]:rp[r + 1]] for r in interior])
    assert np.linalg.norm(gen - ref) / np.linalg.norm(ref) <= 1e-12


def test_generated_body_fp32(tmp_path):
    form = femx.Form(None, 3, femx.POISSON_MASS, params=(1.0,), dtype=femx.F32, offline=True)
    _run_case(tmp_path, 3, form, dtype="f32")
    form.close()


def test_generated_body_reference_strings_fmad_off(tmp_path, golden_dir):
    """The reference's GiNaC strings (quadrature loop, one case per local vertex) through the same harness."""
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    form = femx.Form(None, 2, entries=j["integrand"], fmad=False, offline=True)
    X, Y, Z, conn, rp, ci, interior, gen = _run_case(tmp_path, 2, form)
    form.close()
    ov = orc.assemble_csr(orc.POISSON, 2, 1, conn, X, Y, None, rp, ci)
    ref = np.array([ov[rp[r]:rp[r + 1]] for r in interior])
    assert np.linalg.norm(gen - ref) / np.linalg.norm(ref) <= 1e-6     # the reference's float literals / 8-digit rule
