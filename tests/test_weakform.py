"""The sympy weak-form front-end (femx.weakform) — the reference's GiNaC workflow
(FunctionSpace / WeakForm::build, fea_symbolic_nvrtc_sparse.cpp:226-362) — against the
reference's recorded GiNaC output and, on the GPU, against the built-in emitter / the oracle."""
import json
import os

import numpy as np
import pytest

import femx
from femx.weakform import FunctionSpace, VectorFunctionSpace, WeakForm, div, dot, grad, inner, sym
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _eval(prologue, expr, env):
    env = dict(env)
    env["real"] = float
    for line in prologue.replace("\n  ", "\n").strip().split("\n"):
        if line:
            name, e = line[len("const real "):].rstrip(";").split(" = ", 1)
            env[name] = eval(e, {"__builtins__": {}}, env)
    return eval(expr, {"__builtins__": {}}, env)


@pytest.fixture(scope="module")
def poisson2d():
    fs = FunctionSpace(2)
    return WeakForm(fs).build(lambda u, v: dot(grad(u), grad(v)),
                              lambda v: (-2 * (fs.x ** 2 + fs.y ** 2) + 36) * v)


def test_generated_strings_equal_reference_ginac_output(poisson2d):
    """Same weak form, same f as the reference's main() (:494-503): LHS and RHS strings agree in value
    with the GiNaC strings the reference recorded (string order is not stable)."""
    j = json.load(open(os.path.join(GOLDEN, "ref_integrand_strings.json")))
    rng = np.random.RandomState(3)
    for _ in range(5):
        c = rng.uniform(-2, 2, 6)
        env = dict(x1=c[0], x2=c[1], x3=c[2], y1=c[3], y2=c[4], y3=c[5], r=0.2, s=0.3, t=0.5, pow=pow)
        for k in range(9):
            a = _eval(poisson2d.prologue, poisson2d.entries[k // 3][k % 3], env)
            b = eval(j["integrand"][k].replace("powf", "pow"), {"__builtins__": {}}, env)
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b))
        for k in range(3):
            a = _eval(poisson2d.prologue, poisson2d.rhs[k], env)
            b = eval(j["rhs"][k], {"__builtins__": {}}, env)
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b))


def test_prologue_is_free_of_the_quadrature_point_and_compiles(poisson2d):
    assert " r" not in poisson2d.prologue.replace("real", "") and "*s" not in poisson2d.prologue
    f = poisson2d.compile(None, offline=True)
    assert f.cubin("csr")[:4] == b"\x7fELF" and f.cubin("rhs")[:4] == b"\x7fELF"
    f.close()


@pytest.mark.gpu
def test_dsl_poisson_matches_golden_on_gpu(ctx, poisson2d):
    import torch
    g = np.load(os.path.join(GOLDEN, "ref_poisson2d_jitter12.npz"))
    mesh = femx.Mesh(2, torch.from_numpy(g["conn"]).cuda(), (torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["Y"]).cuda()))
    form = poisson2d.compile(ctx)
    A, r, c = form.assemble_coo(mesh)
    assert np.linalg.norm(A.cpu().numpy() - g["A"]) <= 1e-12 * np.linalg.norm(g["A"])
    pat = femx.Pattern(ctx, mesh)
    b = form.assemble_rhs(pat, mesh)
    assert np.linalg.norm(b.cpu().numpy() - g["rhs"]) <= 1e-12 * np.linalg.norm(g["rhs"])
    form.close(); pat.close()


@pytest.mark.gpu
def test_dsl_tets_and_elasticity_match_oracle_on_gpu(ctx):
    import torch
    X, Y, Z, conn = orc.box_mesh(4, 3, 5)
    X = X + np.random.RandomState(2).uniform(-0.02, 0.02, X.shape)
    mesh = femx.Mesh(3, torch.from_numpy(conn).cuda(), tuple(torch.from_numpy(a).cuda() for a in (X, Y, Z)))
    rp, ci = orc.pattern(conn, len(X))
    fs = FunctionSpace(3)
    form = WeakForm(fs).build(lambda u, v: dot(grad(u), grad(v)) + 2.5 * u * v).compile(ctx)
    pat = femx.Pattern(ctx, mesh)
    v = form.assemble_csr(pat, mesh)
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci, params=(2.5,))
    assert np.linalg.norm(v.cpu().numpy() - ov) <= 1e-12 * np.linalg.norm(ov)
    form.close(); pat.close()
    # 2-D elasticity written as a weak form
    X, Y, _, conn = orc.rect_mesh(0, 2, 0, 1, 7, 9)
    mesh = femx.Mesh(2, torch.from_numpy(conn).cuda(), (torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()))
    lam, mu = 1.2, 0.7
    vs = VectorFunctionSpace(2)
    form = WeakForm(vs).build(lambda u, v: lam * div(u) * div(v) + 2 * mu * inner(sym(grad(u)), sym(grad(v)))).compile(ctx)
    pat = femx.Pattern(ctx, mesh, nd=2)
    v = form.assemble_csr(pat, mesh)
    rp, ci = orc.pattern(conn, len(X))
    drp, dci = orc.expand_pattern(2, rp, ci)
    ov = orc.assemble_csr(orc.ELASTICITY, 2, 2, conn, X, Y, None, drp, dci, params=(lam, mu))
    assert np.linalg.norm(v.cpu().numpy() - ov) <= 1e-12 * np.linalg.norm(ov)
    form.close(); pat.close()


# ---------------------------------------------------------------- C++ front end (include/femx_weakform.hpp) ---
CPP_PROBE = r'''
#include <cstdio>
#include "femx_weakform.hpp"
using namespace femx::wf;
int main() {
  FunctionSpace fs(2);
  Ex f = -2.0 * (fs.x * fs.x + fs.y * fs.y) + 36.0;
  WeakForm wf(fs);
  wf.build([&](Fn u, Fn v) { return dot(grad(u), grad(v)); }, [&](Fn v) { return f * v; });
  femx_form_desc d = wf.desc(FEMX_F64, 0);
  printf("%s@@\n", d.prologue);
  for (int k = 0; k < 9; ++k) printf("%s@@\n", d.entries[k]);
  for (int k = 0; k < 3; ++k) printf("%s@@\n", d.rhs_entries[k]);
  return 0;
}
'''


def test_cpp_weakform_front_end_generates_the_reference_entries(tmp_path):
    """The C++ lambda front end (the reference's UX, fea_symbolic_nvrtc_sparse.cpp:494-503) on the host: its nine entry strings for
    dot(grad u, grad v) and three load strings for f v, compiled through the OFFLINE NVRTC path and evaluated numerically,
    equal the reference's recorded GiNaC output (golden fixture) on a jittered triangle."""
    import ctypes
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "probe.cpp"
    src.write_text(CPP_PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["g++", "-std=c++11", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().split("@@\n")
    prologue, entries, rhs = out[0], out[1:10], out[10:13]
    assert len(entries) == 9 and all(entries) and len(rhs) == 3
    # numerical evaluation of the generated C text: compile it into a tiny host function
    body = "typedef double real;\nextern \"C\" void ev(const double* p, double* o) {\n"
    body += "  const double x1=p[0],x2=p[1],x3=p[2],y1=p[3],y2=p[4],y3=p[5],r=p[6],s=p[7],t=p[8]; (void)r;(void)s;(void)t;\n"
    body += "  " + prologue.replace("\n", "\n  ") + "\n"
    for k, e in enumerate(entries + rhs):
        body += f"  o[{k}] = {e};\n"
    body += "}\n"
    (tmp_path / "ev.cpp").write_text(body)
    so = tmp_path / "ev.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", str(so), str(tmp_path / "ev.cpp")])
    lib = ctypes.CDLL(str(so))
    golden = json.load(open(os.path.join(root, "tests", "golden", "ref_integrand_strings.json")))
    X = np.array([0.1, 1.3, 0.4]); Y = np.array([-0.2, 0.3, 1.1])
    for (r, s) in ((0.2, 0.3), (0.6, 0.1)):
        t = 1 - r - s
        p = np.array([*X, *Y, r, s, t])
        o = np.zeros(12)
        lib.ev(p.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p))
        ns = dict(x1=X[0], x2=X[1], x3=X[2], y1=Y[0], y2=Y[1], y3=Y[2], r=r, s=s, t=t, pow=pow)
        want = [eval(e.replace("powf", "pow").replace(".0f", ".0").replace("f)", ")"), dict(ns)) for e in golden["integrand"]]
        assert np.allclose(o[:9], want, rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
def test_cpp_weakform_demo_program():
    """examples/femx_weakform_demo: the reference's main() with the lambda front end, plain C++ against the C ABI (2-D and 3-D,
    matrix against the built-in emitter to 1e-12, load vector against the exact integral of f)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "femx_weakform_demo")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "examples")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok dim=") == 2, out.stdout
