"""Host-side I/O: Gmsh MSH 2.2 reader, Matrix Market writer (no device needed)."""
import numpy as np
import pytest

import femx
from oracle import oracle as orc


def _write_msh(path, X, Y, Z, conn, extra_boundary=False, tag_offset=1):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(X))
        for i in range(len(X)):
            f.write("%d %.17g %.17g %.17g\n" % (i + tag_offset, X[i], Y[i], Z[i]))
        f.write("$EndNodes\n$Elements\n")
        lines = []
        if extra_boundary:   # a point, a line and (for volume meshes) a boundary triangle: must be ignored
            lines.append("15 2 0 1 %d" % tag_offset)
            lines.append("1 2 0 1 %d %d" % (tag_offset, tag_offset + 1))
            if conn.shape[1] == 4:
                lines.append("2 2 0 1 %d %d %d" % tuple(conn[0, :3] + tag_offset))
        etype = 2 if conn.shape[1] == 3 else 4
        for e in conn:
            lines.append(("%d 2 7 1 " % etype) + " ".join(str(v + tag_offset) for v in e))
        f.write("%d\n" % len(lines))
        for k, l in enumerate(lines):
            f.write("%d %s\n" % (k + 1, l))
        f.write("$EndElements\n")


@pytest.mark.parametrize("dim", [2, 3])
def test_gmsh_round_trip(tmp_path, dim):
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(-3, 3, -3, 3, 6, 5)
        Z = np.zeros_like(X)
    else:
        X, Y, Z, conn = orc.box_mesh(3, 2, 4)
    p = tmp_path / "mesh.msh"
    _write_msh(p, X, Y, Z, conn, extra_boundary=True, tag_offset=101)   # non-contiguous-from-1 tags
    d, rX, rY, rZ, rconn = femx.read_gmsh(p)
    assert d == dim
    assert np.array_equal(rX, X) and np.array_equal(rY, Y) and np.array_equal(rZ, Z)
    assert np.array_equal(rconn, conn)


def test_gmsh_errors(tmp_path):
    with pytest.raises(femx.FemxError):
        femx.read_gmsh(tmp_path / "missing.msh")
    p = tmp_path / "v4.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(femx.FemxError) as ei:
        femx.read_gmsh(p)
    assert ei.value.status == 4
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 2 0 1 2 3\n$EndElements\n")
    with pytest.raises(femx.FemxError) as ei:
        femx.read_gmsh(p)
    assert "unknown node" in str(ei.value)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_matrix_market_matches_scipy(tmp_path):
    import scipy.io
    import scipy.sparse as sp
    X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, 5, 4)
    rp, ci = orc.pattern(conn, len(X))
    v = orc.assemble_csr(orc.POISSON_MASS, 2, 1, conn, X, Y, None, rp, ci)
    p = tmp_path / "A.mtx"
    femx.write_matrix_market(p, rp, ci, v)
    A = scipy.io.mmread(str(p)).tocsr()
    B = sp.csr_matrix((v, ci, rp), shape=(len(X), len(X)))
    assert (A != B).nnz == 0 or abs(A - B).max() == 0.0


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_assemble_from_gmsh_file_and_export(ctx, tmp_path):
    """File in → assembled operator → file out, cross-checked with scipy."""
    import scipy.io
    import torch
    X, Y, Z, conn = orc.box_mesh(4, 3, 3)
    p = tmp_path / "box.msh"
    _write_msh(p, X, Y, Z, conn)
    d, rX, rY, rZ, rconn = femx.read_gmsh(p)
    mesh = femx.Mesh(d, torch.from_numpy(rconn).cuda(), tuple(torch.from_numpy(a).cuda() for a in (rX, rY, rZ)))
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    v = form.assemble_csr(pat, mesh)
    rp, ci = pat.csr("int64")
    out = tmp_path / "A.mtx"
    femx.write_matrix_market(out, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy())
    A = scipy.io.mmread(str(out)).tocsr()
    orp, oci = orc.pattern(conn, len(X))
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, orp, oci)
    import scipy.sparse as sp
    B = sp.csr_matrix((ov, oci, orp), shape=A.shape)
    assert abs(A - B).max() <= 1e-12 * abs(B).max()
    form.close(); pat.close()
