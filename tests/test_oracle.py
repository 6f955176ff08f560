"""CPU tier: the oracle against the reference's golden vectors and known answers."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

CASES = ["ref_poisson2d_2x2.npz", "ref_poisson2d_4x4.npz", "ref_poisson2d_10x7.npz",
         "ref_poisson2d_jitter12.npz"]


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_kat_2x2_triplets():
    # SURVEY §4 known answers: RectangleMesh(-3,3,-3,3,2,2), elements 0 and 1
    X, Y, flag, conn = orc.rect_mesh(-3, 3, -3, 3, 2, 2)
    assert conn[:4].tolist() == [[0, 1, 3], [1, 4, 3], [1, 2, 4], [2, 5, 4]]
    assert flag.tolist() == [1, 1, 1, 1, 0, 1, 1, 1, 1]
    A, r, c = orc.assemble_coo(orc.POISSON, 2, 1, conn, X, Y)
    s = 1.00000002  # sum of the 8-digit weights * 2
    exp0 = [(0, 0, 1), (0, 1, -.5), (0, 3, -.5), (1, 0, -.5), (1, 1, .5), (1, 3, 0), (3, 0, -.5), (3, 1, 0), (3, 3, .5)]
    exp1 = [(1, 1, .5), (1, 4, -.5), (1, 3, 0), (4, 1, -.5), (4, 4, 1), (4, 3, -.5), (3, 1, 0), (3, 4, -.5), (3, 3, .5)]
    for k, (i, j, v) in enumerate(exp0 + exp1):
        assert (r[k], c[k]) == (i, j)
        assert abs(A[k] - v * s) < 1e-14


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_strings(golden_dir, case):
    """Element values: oracle vs the reference's GiNaC strings evaluated verbatim."""
    g = np.load(os.path.join(golden_dir, case))
    A, r, c = orc.assemble_coo(orc.POISSON, 2, 1, g["conn"], g["X"], g["Y"])
    assert np.array_equal(r, g["rowA"]) and np.array_equal(c, g["colA"])
    assert relF(A, g["A"]) <= 1e-13
    assert np.max(np.abs(A - g["A"])) <= 1e-13 * np.max(np.abs(g["A"]))


@pytest.mark.parametrize("case", CASES)
def test_oracle_pattern_and_ell(golden_dir, case):
    """Pattern bit-exact vs getNeighborNodesList (python-set restatement), ELL values."""
    g = np.load(os.path.join(golden_dir, case))
    n = len(g["X"])
    rp, ci = orc.pattern(g["conn"], n)
    ln, idx = orc.ell_pattern(rp, ci, 7)
    assert np.array_equal(ln, g["ell_len"]) and np.array_equal(idx, g["ell_idx"])
    vals = orc.assemble_csr(orc.POISSON, 2, 1, g["conn"], g["X"], g["Y"], None, rp, ci)
    ell = np.zeros((n, 7))
    for i in range(n):
        ell[i, : ln[i]] = vals[rp[i]:rp[i + 1]]
    assert relF(ell, g["ell_val"]) <= 1e-13


def test_rect_mesh_matches_golden_mesh(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_poisson2d_10x7.npz"))
    X, Y, _, conn = orc.rect_mesh(-3, 3, -3, 3, 10, 7)
    assert np.array_equal(conn, g["conn"])
    assert np.array_equal(X, g["X"]) and np.array_equal(Y, g["Y"])


def test_pattern_sizes():
    # 64x64 → 29,057 nnz; 1000x100 → 703,301; closed forms
    for (nr, nc) in [(64, 64), (1000, 100)]:
        _, _, _, conn = orc.rect_mesh(0, 1, 0, 1, nr, nc)
        rp, ci = orc.pattern(conn, (nr + 1) * (nc + 1))
        nnz2d = (nr + 1) * (nc + 1) + 2 * (nr * (nc + 1) + nc * (nr + 1) + nr * nc)
        assert rp[-1] == nnz2d and np.diff(rp).max() == 7
    assert 29057 == 65 * 65 + 2 * (2 * 64 * 65 + 64 * 64)
    for n in (4, 8):
        _, _, _, conn = orc.box_mesh(n, n, n)
        rp, ci = orc.pattern(conn, (n + 1) ** 3)
        assert rp[-1] == (n + 1) ** 3 + 2 * (3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3)
        assert np.diff(rp).max() == 15


def test_box_mesh_orientation_and_volume():
    X, Y, Z, conn = orc.box_mesh(3, 4, 5, hi=(1.0, 2.0, 3.0))
    P = np.stack([X, Y, Z], 1)[conn]  # [ne,4,3]
    J = np.transpose(P[:, :3, :] - P[:, 3:4, :], (0, 2, 1))
    det = np.linalg.det(J)
    assert np.all(det > 0)
    assert abs(det.sum() / 6 - 6.0) < 1e-12


def test_invariants_poisson_mass_elasticity():
    # row sums of grad.grad = 0; 1^T M 1 = |Omega|; rigid translations in the elasticity nullspace
    X, Y, Z, conn = orc.box_mesh(3, 3, 3)
    n = len(X)
    rp, ci = orc.pattern(conn, n)
    v = orc.assemble_csr(orc.POISSON, 3, 1, conn, X, Y, Z, rp, ci)
    assert np.abs(orc.spmv(rp, ci, v, np.ones(n))).max() < 1e-12
    m = orc.assemble_csr(orc.MASS, 3, 1, conn, X, Y, Z, rp, ci)
    assert abs(m.sum() - 1.0) < 1e-12
    pm = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci)
    assert np.allclose(pm, v + m, atol=1e-14)
    drp, dci = orc.expand_pattern(3, rp, ci)
    ev = orc.assemble_csr(orc.ELASTICITY, 3, 3, conn, X, Y, Z, drp, dci, params=(0.5769, 0.3846))
    for c in range(3):
        t = np.zeros(3 * n)
        t[c::3] = 1.0
        assert np.abs(orc.spmv(drp, dci, ev, t)).max() < 1e-12
    # infinitesimal rotation about z: u = (-y, x, 0)
    rot = np.zeros(3 * n)
    rot[0::3] = -Y
    rot[1::3] = X
    assert np.abs(orc.spmv(drp, dci, ev, rot)).max() < 1e-12
    # symmetry
    import scipy.sparse as sp
    Ae = sp.csr_matrix((ev, dci, drp))
    assert abs(Ae - Ae.T).max() < 1e-13


def test_mass_2d_exact_for_quadratics():
    # reference 7-point rule is degree 5 → P1 mass matrix exact up to the 8-digit literals
    Ae = orc.element_matrix(orc.MASS, 2, 1, [0, 1, 0], [0, 0, 1])
    exact = np.array([[2, 1, 1], [1, 2, 1], [1, 1, 2]]) / 24.0
    assert np.abs(Ae - exact).max() < 5e-9


def test_cg_converges_on_spd():
    X, Y, Z, conn = orc.box_mesh(4, 4, 4)
    n = len(X)
    rp, ci = orc.pattern(conn, n)
    v = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci)
    b = orc.spmv(rp, ci, v, np.ones(n))
    x, res = orc.cg(rp, ci, v, b, 100)
    assert res[-1] / res[0] < 1e-10
    assert np.abs(x - 1).max() < 1e-8


def test_golden_strings_fixture(golden_dir):
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    assert len(j["integrand"]) == 9 and len(j["triW"]) == 7
    # funIdx 0 ≡ ((x2-x3)^2+(y2-y3)^2)/jac
    x1, x2, x3, y1, y2, y3 = 0.3, 1.7, 0.2, -0.4, 0.1, 1.9
    val = eval(j["integrand"][0].replace("powf", "pow"), {"__builtins__": {}},
               dict(x1=x1, x2=x2, x3=x3, y1=y1, y2=y2, y3=y3, pow=pow))
    jac = (x1 - x3) * (y2 - y3) - (y1 - y3) * (x2 - x3)
    assert abs(val - ((x2 - x3) ** 2 + (y2 - y3) ** 2) / jac) < 1e-13


@pytest.mark.parametrize("case", CASES)
def test_oracle_rhs_matches_reference_strings(golden_dir, case):
    """Load vector: oracle vs the reference's (generated-then-discarded) RHS strings, fea_symbolic.cu:335,339,343."""
    g = np.load(os.path.join(golden_dir, case))
    b, be = orc.assemble_rhs(1, 2, 1, g["conn"], g["X"], g["Y"])
    assert relF(be, g["rhs_elem"]) <= 1e-13
    assert relF(b, g["rhs"]) <= 1e-13


def test_rhs_constant_source_integrates_to_volume():
    X, Y, Z, conn = orc.box_mesh(3, 4, 2, hi=(1.0, 2.0, 0.5))
    b, _ = orc.assemble_rhs(0, 3, 1, conn, X, Y, Z, fvec=(2.0,))
    assert abs(b.sum() - 2.0) < 1e-12          # f * |Omega| = 2 * 1
    X, Y, _, conn = orc.rect_mesh(0, 2, 0, 3, 5, 4)
    b, _ = orc.assemble_rhs(0, 2, 2, conn, X, Y, None, fvec=(1.0, -3.0))
    assert abs(b[0::2].sum() - 6.00000012) < 1e-9 and abs(b[1::2].sum() + 18.00000036) < 1e-9  # 8-digit weights (Q9)
