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
