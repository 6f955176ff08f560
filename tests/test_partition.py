"""Partition of an UNSTRUCTURED mesh by owned CSR rows with duplicated ghost elements (femx_partition_extract): every rank's
sub-mesh, assembled alone and without communication, yields its rows of the global matrix — pattern and values bit for bit."""
import numpy as np
import pytest

import femx

pytestmark = pytest.mark.gpu


def _delaunay(n_pts, seed):
    from scipy.spatial import Delaunay
    pts = np.random.RandomState(seed).uniform(0, 1, (n_pts, 2))
    # a locality-preserving numbering (sort along x then y in strips), the stand-in for RCM
    key = np.lexsort((pts[:, 1], np.floor(pts[:, 0] * 12)))
    pts = pts[key]
    tri = Delaunay(pts).simplices.astype(np.int32)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    flip = ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return pts, tri


@pytest.mark.parametrize("world", [1, 3, 4])
def test_unstructured_partition_rows_equal_global_rows(ctx, world):
    import torch
    pts, tri = _delaunay(4000, 7)
    mesh = femx.Mesh(2, torch.from_numpy(tri).cuda(), (torch.from_numpy(pts[:, 0].copy()).cuda(), torch.from_numpy(pts[:, 1].copy()).cuda()))
    form = femx.Form(ctx, 2, femx.POISSON_MASS)
    pat = femx.Pattern(ctx, mesh)
    vals = form.assemble_csr(pat, mesh)
    rp, ci = pat.csr("int64")
    n = mesh.n_nodes
    got_v, got_c, got_len = [], [], []
    for r in range(world):
        lo, hi = (r * n) // world, ((r + 1) * n) // world
        part = femx.Partition(ctx, mesh, lo, hi)
        assert part.row_end - part.row_begin == hi - lo
        assert torch.equal(part.l2g[part.row_begin:part.row_end].long(), torch.arange(lo, hi, device="cuda"))
        assert bool((part.l2g[1:] > part.l2g[:-1]).all())
        assert part.n_elems < mesh.n_elems or world == 1
        sp = part.pattern()
        sv = form.assemble_csr(sp, part.mesh)
        srp, sci = part.global_csr(sp)
        got_v.append(sv); got_c.append(sci); got_len.append(srp[1:] - srp[:-1])
        sp.close(); part.close()
    assert torch.equal(torch.cat(got_len), rp[1:] - rp[:-1])
    assert torch.equal(torch.cat(got_c), ci)
    assert torch.equal(torch.cat(got_v), vals)
    form.close(); pat.close()


def test_partition_of_tets_and_isolated_nodes(ctx):
    """3-D, vector pattern, and an owned node that no element touches (it keeps an empty row)."""
    import torch
    from oracle import oracle as orc
    X, Y, Z, conn = orc.box_mesh(5, 4, 6)
    rng = np.random.RandomState(1)
    perm = rng.permutation(len(conn))
    conn = conn[perm]                                       # no lattice: the general passes
    n = len(X) + 2                                          # two isolated nodes at the end
    coords = [np.concatenate([c, [9.0, 9.5]]) for c in (X, Y, Z)]
    mesh = femx.Mesh(3, torch.from_numpy(np.ascontiguousarray(conn)).cuda(), tuple(torch.from_numpy(c).cuda() for c in coords))
    form = femx.Form(ctx, 3, femx.ELASTICITY, nd=3, params=(0.6, 0.4))
    pat = femx.Pattern(ctx, mesh, nd=3)
    vals = form.assemble_csr(pat, mesh)
    rp, ci = pat.csr("int64")
    parts_v, parts_c = [], []
    for lo, hi in ((0, 100), (100, 101), (101, n)):
        part = femx.Partition(ctx, mesh, lo, hi)
        sp = part.pattern(nd=3)
        parts_v.append(form.assemble_csr(sp, part.mesh))
        # dof-level columns: global node id * nd + component
        srp, sci = part.global_csr(sp)
        parts_c.append(sci)
        sp.close(); part.close()
    assert torch.equal(torch.cat(parts_v), vals)
    assert torch.equal(torch.cat(parts_c), ci)
    form.close(); pat.close()
