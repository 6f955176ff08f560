"""The reference's own code, recompiled (oracle/_ref, built by oracle/build_ref.py where
/root/reference exists): host mesh + host pattern builder pin the oracle on the CPU tier;
its CUDA kernels (fixes Q2/Q3/Q4/Q8/Q13 only) pin the engine on the GPU tier."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refimpl

needs_ref = pytest.mark.skipif(not refimpl.available(), reason="oracle/_ref not built (no /root/reference at build time)")


@needs_ref
@pytest.mark.parametrize("shape", [(2, 2), (4, 4), (10, 7), (64, 64), (100, 13)])
def test_oracle_mesh_and_pattern_match_reference_host_code(shape):
    nR, nC = shape
    nx, ny, fl, X, Y, g = refimpl.host_mesh(-3.0, 3.0, -3.0, 3.0, nR, nC)
    oX, oY, ofl, oconn = orc.rect_mesh(-3.0, 3.0, -3.0, 3.0, nR, nC)
    assert np.array_equal(nx, oX) and np.array_equal(ny, oY) and np.array_equal(fl, ofl)
    assert np.array_equal(g, oconn)
    assert np.array_equal(X, oX[oconn].ravel()) and np.array_equal(Y, oY[oconn].ravel())
    ln, idx = refimpl.neighbor_list(nR, nC)           # Mesh::getNeighborNodesList, untouched
    rp, ci = orc.pattern(oconn, len(oX))
    oln, oidx = orc.ell_pattern(rp, ci, 7)
    assert np.array_equal(ln, oln) and np.array_equal(idx, oidx)


def _relF(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 2), (10, 7), (64, 64), (1000, 100)])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_engine_matches_reference_kernels(ctx, shape, prec):
    """Same mesh, same integrand: reference K4/K5 vs femx.  Pattern bit-exact; values within
    1e-12 (fp64 retype) / 1e-5 (fp32 as written) relative Frobenius norm."""
    import torch
    import femx
    nR, nC = shape
    tol = 1e-12 if prec == "f64" else 1e-5
    dt = femx.F64 if prec == "f64" else femx.F32
    mesh = ctx.rectangle_mesh(-3.0, 3.0, -3.0, 3.0, nR, nC, dtype=dt)
    exp = mesh.expanded(ctx)
    gIdx = mesh.conn.reshape(-1).contiguous()
    form = femx.Form(ctx, 2, femx.POISSON, dtype=dt)
    # ---- K4: COO triplets
    rA, rrow, rcol, _ = refimpl.assemble_coo(prec, nR, nC, exp.elem_xyz[0], exp.elem_xyz[1], gIdx)
    A, row, col = form.assemble_coo(exp)
    assert torch.equal(row, rrow) and torch.equal(col, rcol)
    assert _relF(A.cpu().numpy(), rA.double().cpu().numpy()) <= tol
    # ---- K5: pattern + numeric.  Host pattern from the reference's own getNeighborNodesList
    ln, idx = refimpl.neighbor_list(nR, nC)
    pat = femx.Pattern(ctx, mesh)
    eln, eidx = pat.ell(7)
    assert np.array_equal(eln.cpu().numpy(), ln) and np.array_equal(eidx.cpu().numpy(), idx)
    rell, _ = refimpl.assemble_ell(prec, nR, nC, exp.elem_xyz[0], exp.elem_xyz[1], gIdx,
                                   torch.from_numpy(ln).cuda(), torch.from_numpy(idx).cuda())
    vals = form.assemble_csr(pat, mesh)
    ell = pat.values_to_ell(vals, 7).reshape(-1)
    assert _relF(ell.cpu().numpy(), rell.double().cpu().numpy()) <= tol
    form.close(); pat.close()


def _parse_rows(text):
    """'(i,col) val    (i,col) val ...' lines → {(i,col): val}"""
    import re
    out = {}
    for m in re.finditer(r"\((\d+),(\d+)\)\s+(\S+)", text):
        out[(int(m.group(1)), int(m.group(2)))] = float(m.group(3))
    return out


@needs_ref
@pytest.mark.gpu
def test_cpp_client_output_equals_unmodified_reference_program():
    """The reference's own binary (fea_test_sm_sym_sparse2.cu compiled UNMODIFIED for sm_100, its main(),
    its 1000x100 mesh) and the plain C++ client of the femx C ABI (examples/femx_sparse2) print the
    same first 16 matrix rows: same (row, col) keys — the pattern — and values within 1e-5 (fp32)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_exe = os.path.join(root, "oracle", "_ref", "fea_test_sm_sym_sparse2")
    my_exe = os.path.join(root, "examples", "femx_sparse2")
    if not (os.path.exists(ref_exe) and os.path.exists(my_exe)):
        pytest.skip("binaries not built")
    ref = _parse_rows(subprocess.check_output([ref_exe], timeout=300).decode())
    mine = _parse_rows(subprocess.check_output([my_exe, "1000", "100", "1"], timeout=300).decode())
    assert len(ref) >= 60 and set(ref) == set(mine)
    scale = max(abs(v) for v in ref.values())
    for k in ref:
        assert abs(ref[k] - mine[k]) <= 1e-5 * scale, (k, ref[k], mine[k])
