"""Multi-GPU layer behind the C ABI (femx_dist_*), one-rank tier: the operator, the overlapped SpMV (interior rows /
rows that read ghost columns) and the single-reduction CG against the oracle.  The N-rank tier runs under torchrun:
tools/dist_check.py (gpurun --gpus 2) and bench.py --gpus N."""
import numpy as np
import pytest

import femx
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_dist_operator_world1_spmv_and_cg_against_oracle(ctx):
    import torch
    mesh = ctx.box_mesh(12, 10, 9)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    vals = form.assemble_csr(pat, mesh)
    d = femx.Dist(ctx)
    op = d.operator(pat, vals)
    assert (op.n_owned, op.ghost_lo, op.ghost_hi) == (mesh.n_nodes, 0, 0)
    assert (op.interior_lo, op.interior_hi) == (0, mesh.n_nodes)
    x = torch.from_numpy(np.random.RandomState(3).uniform(-1, 1, mesh.n_nodes)).cuda()
    assert torch.equal(op.spmv(x), pat.spmv(vals, x))
    X, Y, Z, conn = orc.box_mesh(12, 10, 9)
    rp, ci = orc.pattern(conn, len(X))
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci, params=(1.0,))
    b = orc.spmv(rp, ci, ov, np.ones(len(X)))
    xs, res, ms = op.cg(torch.from_numpy(b).cuda(), 60)
    _, ores = orc.cg(rp, ci, ov, b, 60)
    # same Krylov iterates (Chronopoulos-Gear recurrences vs the textbook loop): histories agree while the
    # residual is above rounding level
    assert len(ores) == 61 and np.allclose(res, ores, rtol=1e-6)
    assert res[-1] < 1e-4 * res[0]
    assert float((xs - 1.0).abs().max()) < 1e-3
    # a second solve replays the captured graph with the same bits
    xs2, res2, _ = op.cg(torch.from_numpy(b).cuda(), 60, xs.clone())
    assert np.array_equal(res, res2)
    op.close(); d.close(); form.close(); pat.close()


def test_dist_operator_on_a_slab_splits_interior_and_ghost_rows(ctx):
    """A middle slab built on one GPU: rows of the first / last owned plane read ghost columns."""
    import torch
    nx, ny, nz = 6, 5, 12
    plane = (nx + 1) * (ny + 1)
    r0, r1, lo, hi = femx.dist_slab(nz + 1, 3, 1)
    mesh = ctx.box_mesh(nx, ny, nz, k_lo=lo, k_hi=hi)
    pat = femx.Pattern(ctx, mesh, row_begin=(r0 - lo) * plane, row_end=(r1 - lo) * plane, col_base=lo * plane)
    vals = femx.Form(ctx, 3, femx.POISSON_MASS).assemble_csr(pat, mesh)
    d = femx.Dist(ctx)
    op = d.operator(pat, vals)
    assert (op.n_owned, op.ghost_lo, op.ghost_hi) == ((r1 - r0) * plane, plane, plane)
    assert (op.interior_lo, op.interior_hi) == (plane, (r1 - r0 - 1) * plane)
    # femx_spmv_rows pieces = the whole product
    x = torch.from_numpy(np.random.RandomState(1).uniform(-1, 1, mesh.n_nodes)).cuda()
    y = pat.spmv(vals, x, x_base=lo * plane)
    y2 = torch.zeros_like(y)
    import ctypes as C
    for a, b in ((0, 37), (37, op.interior_hi), (op.interior_hi, pat.n_rows)):
        ctx.check(femx.lib().femx_spmv_rows(pat.h, femx.F64, C.c_void_p(vals.data_ptr()), C.c_void_p(x.data_ptr()),
                                            C.c_int64(lo * plane), C.c_void_p(y2.data_ptr()), C.c_int64(a), C.c_int64(b),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(y, y2)
    op.close(); d.close(); pat.close()
