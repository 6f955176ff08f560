"""Element-once lattice numeric pass (femx_pattern_lattice / kFemxJitLattice) on the GPU.

The symbolic pass must recognise what femx_mesh_box / RectangleMesh-style generators produce; the lattice
pass must give the oracle's values (1e-12), be bitwise reproducible run to run, bitwise symmetric, bitwise
independent of the tile shape / k-chunk / slab partition, and agree with the stencil-class pass to rounding."""
import numpy as np
import pytest

import femx
from oracle import oracle as orc
from tools.lattice_offline import kuhn_corners

pytestmark = pytest.mark.gpu

OID = {femx.POISSON: orc.POISSON, femx.POISSON_MASS: orc.POISSON_MASS, femx.MASS: orc.MASS}


def _jittered_box(ctx, nx, ny, nz, amp=0.2, seed=12345, dtype=None):
    import torch
    X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
    rng = np.random.default_rng(seed)
    h = 1.0 / max(nx, ny, nz)
    coords = [c + rng.uniform(-amp * h, amp * h, c.shape) for c in (X, Y, Z)]
    tdt = dtype or torch.float64
    mesh = femx.Mesh(3, torch.from_numpy(conn).cuda(), tuple(torch.from_numpy(c).to(tdt).cuda() for c in coords))
    return mesh, conn, coords


def test_symbolic_pass_finds_the_kuhn_lattice(ctx):
    mesh = ctx.box_mesh(7, 6, 5)
    pat = femx.Pattern(ctx, mesh)
    lat = pat.lattice()
    assert lat is not None
    assert lat["n_per_cell"] == 6 and lat["cells"] == [7, 6, 5] and lat["strides"] == [1, 8, 56] and lat["node0"] == 0
    assert lat["corners"] == kuhn_corners()
    pat.close()
    # 2-D: RectangleMesh::generate is a lattice too (2 triangles per cell); no element-once pass there yet
    mesh = ctx.rectangle_mesh(0, 1, 0, 1, 9, 11)
    pat = femx.Pattern(ctx, mesh)
    lat = pat.lattice()
    assert lat["n_per_cell"] == 2 and lat["cells"][:2] == [11, 9] and lat["strides"][:2] == [1, 12]
    assert lat["corners"] == [[0, 1, 2], [1, 3, 2]]
    pat.close()


def test_no_lattice_when_elements_are_permuted_or_perturbed(ctx):
    import torch
    X, Y, Z, conn = orc.box_mesh(6, 5, 4)
    coords = tuple(torch.from_numpy(c).cuda() for c in (X, Y, Z))
    c2 = conn.copy()
    c2[[100, 101]] = c2[[101, 100]]          # two elements swapped: same matrix, not cell-major any more
    pat = femx.Pattern(ctx, femx.Mesh(3, torch.from_numpy(c2).cuda(), coords))
    assert pat.lattice() is None and pat.stencil()["rows"] > 0
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    v = form.assemble_csr(pat, femx.Mesh(3, torch.from_numpy(c2).cuda(), coords))   # stencil-class pass
    rp, ci = orc.pattern(conn, len(X))
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci, params=(1.0,))
    assert np.linalg.norm(v.cpu().numpy() - ov) / np.linalg.norm(ov) <= 1e-12
    assert "FEMX_LATTICE" not in form.source
    form.close(); pat.close()


@pytest.mark.parametrize("builtin", [femx.POISSON_MASS, femx.POISSON, femx.MASS])
@pytest.mark.parametrize("dims", [(13, 11, 9), (37, 5, 3), (6, 6, 6), (3, 40, 35)])
def test_lattice_pass_equals_oracle(ctx, builtin, dims):
    import torch
    mesh, conn, coords = _jittered_box(ctx, *dims)
    pat = femx.Pattern(ctx, mesh)
    assert pat.lattice() is not None
    form = femx.Form(ctx, 3, builtin, params=(1.5,))
    v = form.assemble_csr(pat, mesh)
    assert "#define FEMX_LATTICE 1" in form.source
    v2 = form.assemble_csr(pat, mesh)
    assert torch.equal(v, v2)                                  # run to run: same bits
    rp, ci = orc.pattern(conn, mesh.n_nodes)
    ov = orc.assemble_csr(OID[builtin], 3, 1, conn, *coords, rp, ci, params=(1.5,))
    assert np.linalg.norm(v.cpu().numpy() - ov) / np.linalg.norm(ov) <= 1e-12
    # against the stencil-class pass (owner-computes row loop): equal to rounding
    ctx.set_option("lattice", 0)
    try:
        vs = form.assemble_csr(pat, mesh)
    finally:
        ctx.set_option("lattice", 1)
    assert "FEMX_LATTICE" not in form.source
    assert torch.linalg.norm(v - vs) / torch.linalg.norm(vs) <= 1e-14
    # symmetric to the bit among interior rows
    import scipy.sparse as sp
    A = sp.csr_matrix((v.cpu().numpy(), ci, rp), shape=(mesh.n_nodes,) * 2)
    nx, ny, nz = dims
    idx = np.arange(mesh.n_nodes)
    i, j, k = idx % (nx + 1), (idx // (nx + 1)) % (ny + 1), idx // ((nx + 1) * (ny + 1))
    inner = np.flatnonzero((i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz))
    if len(inner):
        S = A[inner][:, inner]
        assert (S != S.T).nnz == 0
    form.close(); pat.close()


def test_lattice_pass_is_independent_of_tile_shape_and_chunk(ctx):
    import torch
    mesh, conn, coords = _jittered_box(ctx, 21, 17, 12)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    ref = form.assemble_csr(pat, mesh).clone()
    try:
        for tx, ty, kc, pf in ((4, 4, 1, 1), (32, 8, 5, 0), (8, 30, 100, 1), (16, 16, 3, 0), (2, 9, 4, 0)):
            for name, val in (("lt_tx", tx), ("lt_ty", ty), ("lt_kc", kc), ("lt_pf", pf)):
                ctx.set_option(name, val)
            v = form.assemble_csr(pat, mesh)
            assert torch.equal(v, ref), (tx, ty, kc, pf)
    finally:
        for name, val in (("lt_tx", 0), ("lt_ty", 0), ("lt_kc", 0), ("lt_pf", 0)):
            ctx.set_option(name, val)
    form.close(); pat.close()


def test_lattice_pass_slab_rows_concatenate(ctx):
    """z-slabs with ghost layers (the multi-GPU layout): owned rows concatenate to the single-device matrix, bit for bit."""
    import torch
    nx, ny, nz = 9, 8, 14
    mesh, conn, coords = _jittered_box(ctx, nx, ny, nz)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    whole = form.assemble_csr(pat, mesh)
    plane = (nx + 1) * (ny + 1)
    parts = []
    for r0, r1 in ((0, 4), (4, 9), (9, 15)):
        lo, hi = max(r0 - 1, 0), min(r1, nz)
        sel = slice(lo * plane, (hi + 1) * plane)
        e0, e1 = 6 * nx * ny * lo, 6 * nx * ny * hi
        sub = femx.Mesh(3, (mesh.conn[e0:e1] - lo * plane).contiguous(), tuple(c[sel].contiguous() for c in mesh.node_xyz))
        sp = femx.Pattern(ctx, sub, row_begin=(r0 - lo) * plane, row_end=(r1 - lo) * plane, col_base=lo * plane)
        assert sp.lattice() is not None
        parts.append(form.assemble_csr(sp, sub))
        assert "#define FEMX_LATTICE 1" in form.source
        sp.close()
    assert torch.equal(torch.cat(parts), whole)
    form.close(); pat.close()


def test_lattice_pass_fp32(ctx):
    import torch
    mesh, conn, coords = _jittered_box(ctx, 12, 9, 7, dtype=torch.float32)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS, dtype=femx.F32)
    v = form.assemble_csr(pat, mesh)
    assert "#define FEMX_LATTICE 1" in form.source
    rp, ci = orc.pattern(conn, mesh.n_nodes)
    c64 = [c.astype(np.float32).astype(np.float64) for c in coords]
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, *c64, rp, ci, params=(1.0,))
    assert np.linalg.norm(v.double().cpu().numpy() - ov) / np.linalg.norm(ov) <= 1e-5
    form.close(); pat.close()


def test_unaligned_value_buffer_is_rejected(ctx):
    import torch
    mesh = ctx.box_mesh(4, 4, 4)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    buf = torch.empty(pat.nnz + 1, dtype=torch.float64, device="cuda")
    with pytest.raises(femx.FemxError) as e:
        form.assemble_csr(pat, mesh, buf[1:])
    assert e.value.status == 1 and "16-byte" in str(e.value)
    form.close(); pat.close()


def _gapped(ctx, kind, dims):
    """The same cells in a node numbering with holes: lattice node (i, j, k) is node node0 + i + j sy + k sz with strides
    wider than the lattice, unused nodes in front of it, between its lines and planes, and behind it (rows of length 0)."""
    import torch
    rng = np.random.default_rng(5)
    if kind == "rect":
        nx, ny = dims                      # nx cells along the fast node index (RectangleMesh: columns)
        X, Y, _, conn = orc.rect_mesh(0.0, 1.0, 0.0, 2.0, ny, nx)
        Z, nz = None, 0
    else:
        nx, ny, nz = dims
        X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
    node0, sy = 7, nx + 1 + 2
    sz = (ny + 1) * sy + 5
    old = np.arange(len(X))
    i, j, k = old % (nx + 1), (old // (nx + 1)) % (ny + 1), old // ((nx + 1) * (ny + 1))
    new = node0 + i + j * sy + k * sz
    n_nodes = int(new.max()) + 1 + 4
    coords = []
    for c in (X, Y, Z):
        if c is None:
            continue
        full = rng.uniform(0.0, 1.0, n_nodes)
        full[new] = c + rng.uniform(-0.02, 0.02, len(c))
        coords.append(torch.from_numpy(full).cuda())
    conn2 = torch.from_numpy(new[conn].astype(np.int32)).cuda()
    return femx.Mesh(2 if kind == "rect" else 3, conn2, tuple(coords))


@pytest.mark.parametrize("case", ["rect 9x11", "rect 1x40", "box 7x6x5", "box 37x5x1", "box 1x1x1", "box 2x2x9 slab", "box 6x5x4 nd3",
                                  "rect 6x5 gaps", "box 5x4x3 gaps", "box 4x3x6 gaps slab"])
def test_lattice_templated_symbolic_pass_equals_general_pass(ctx, case):
    """On lattice meshes the symbolic pass writes the rows from (at most 3^dim) templates; everything it produces —
    CSR, class, scatter map (exercised through the generic numeric pass and the load vector) — must equal the general
    histogram / sort / merge pipeline bit for bit."""
    import torch
    kind, dims = case.split()[0], [int(v) for v in case.split()[1].split("x")]
    nd = 3 if "nd3" in case else 1
    if "gaps" in case:
        mesh = _gapped(ctx, kind, dims)
        rows = dict()
        if "slab" in case:   # a row range that starts and ends inside the holes / inside a line
            rows = dict(row_begin=61, row_end=mesh.n_nodes - 40, col_base=0)
    elif kind == "rect":
        mesh = ctx.rectangle_mesh(0, 1, 0, 2, dims[0], dims[1])
        rows = dict()
    else:
        mesh = ctx.box_mesh(*dims)
        rows = dict()
        if "slab" in case:
            plane = (dims[0] + 1) * (dims[1] + 1)
            rows = dict(row_begin=2 * plane, row_end=7 * plane + 3, col_base=11)
    out = []
    for flag in (1, 0):
        ctx.set_option("lattice_pattern", flag)
        try:
            pat = femx.Pattern(ctx, mesh, nd=nd, **rows)
        finally:
            ctx.set_option("lattice_pattern", 1)
        rp, ci = pat.csr("int64")
        st = pat.stencil()
        info = (pat.n_rows, pat.nnz, pat.max_row)
        dim = mesh.dim
        form = femx.Form(ctx, dim, femx.ELASTICITY if nd == 3 else femx.POISSON_MASS, nd=nd, params=(0.6, 0.4) if nd == 3 else (1.0,))
        ctx.set_option("spec", 0)
        try:
            v = form.assemble_csr(pat, mesh)      # generic pass: reads the whole scatter map
        finally:
            ctx.set_option("spec", 1)
        b = form.assemble_rhs(pat, mesh)
        v2 = form.assemble_csr(pat, mesh)         # default path (class / lattice pass where available)
        lat = pat.lattice()
        out.append((rp.cpu(), ci.cpu(), st, info, v.cpu(), b.cpu(), v2.cpu(), lat))
        form.close(); pat.close()
    a, g = out
    if "1x1x1" not in case:
        assert a[7] is not None                                       # the templated pass ran (a lattice was found)
    assert torch.equal(a[0], g[0]) and torch.equal(a[1], g[1])
    assert a[3] == g[3], (a[3], g[3])
    # (with holes in the numbering most sample rows of the general pass are empty and it may decline a class that the
    # lattice pass, which counts its rows exactly, keeps: then only the class-independent results are compared bit for bit)
    same_class = "gaps" not in case or g[2]["rows"] > 0
    if same_class:
        for k in ("n_incid", "row_len", "self_pos", "rows", "codes", "offsets"):
            assert a[2][k] == g[2][k], (k, a[2], g[2])
    assert torch.equal(a[4], g[4]) and torch.equal(a[5], g[5])
    if same_class:
        assert torch.equal(a[6], g[6])
    else:
        assert torch.allclose(a[6], g[6], rtol=1e-12, atol=1e-14)
