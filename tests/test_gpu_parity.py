"""GPU tier: the CUDA path through the C ABI against the CPU oracle and the golden
fixtures.  Pattern / index work must be bit-exact; values within the north-star
tolerance  ||A - A_ref||_F / ||A_ref||_F <= 1e-12 (fp64), 1e-5 (fp32)."""
import json
import os

import numpy as np
import pytest

import femx
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL64 = 1e-12
TOL32 = 1e-5
FORM_IDS = {femx.POISSON: orc.POISSON, femx.POISSON_MASS: orc.POISSON_MASS, femx.MASS: orc.MASS,
            femx.ELASTICITY: orc.ELASTICITY}


def relF(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def to_dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host_mesh_to_dev(dim, conn, coords, dtype=femx.F64):
    import torch
    tdt = torch.float64 if dtype == femx.F64 else torch.float32
    return femx.Mesh(dim, to_dev(conn), tuple(to_dev(c, tdt) for c in coords))


# ------------------------------------------------------------------ meshes ---
def test_device_rectangle_mesh_matches_reference_semantics(ctx):
    m = ctx.rectangle_mesh(-3.0, 3.0, -3.0, 3.0, 10, 7, flags=True)
    X, Y, flag, conn = orc.rect_mesh(-3.0, 3.0, -3.0, 3.0, 10, 7)
    assert np.array_equal(m.conn.cpu().numpy(), conn)
    assert np.array_equal(m.node_xyz[0].cpu().numpy(), X)
    assert np.array_equal(m.node_xyz[1].cpu().numpy(), Y)
    assert np.array_equal(m.flag.cpu().numpy(), flag)


def test_device_box_mesh_matches_oracle(ctx):
    m = ctx.box_mesh(5, 4, 3, hi=(1.0, 2.0, 3.0))
    X, Y, Z, conn = orc.box_mesh(5, 4, 3, hi=(1.0, 2.0, 3.0))
    assert np.array_equal(m.conn.cpu().numpy(), conn)
    for a, b in zip(m.node_xyz, (X, Y, Z)):
        assert np.array_equal(a.cpu().numpy(), b)


# ------------------------------------------------- kernel ABI #1: COO triplets ---
@pytest.mark.parametrize("case", ["ref_poisson2d_2x2.npz", "ref_poisson2d_4x4.npz", "ref_poisson2d_10x7.npz",
                                  "ref_poisson2d_jitter12.npz"])
def test_coo_against_reference_golden(ctx, golden_dir, case):
    g = np.load(os.path.join(golden_dir, case))
    mesh = host_mesh_to_dev(2, g["conn"], (g["X"], g["Y"]))
    form = femx.Form(ctx, 2, femx.POISSON)
    A, r, c = form.assemble_coo(mesh)
    assert np.array_equal(r.cpu().numpy(), g["rowA"]) and np.array_equal(c.cpu().numpy(), g["colA"])
    assert relF(A.cpu().numpy(), g["A"]) <= TOL64
    # element-expanded coordinates (the reference's X[3e+k] layout)
    A2, _, _ = form.assemble_coo(mesh.expanded(ctx))
    assert np.array_equal(A2.cpu().numpy(), A.cpu().numpy())
    form.close()


def test_coo_reference_strings_through_nvrtc(ctx, golden_dir):
    """The reference's own GiNaC output strings, compiled at run time (operator surface #1)."""
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    g = np.load(os.path.join(golden_dir, "ref_poisson2d_jitter12.npz"))
    mesh = host_mesh_to_dev(2, g["conn"], (g["X"], g["Y"]))
    for fmad in (True, False):  # the reference compiles with --fmad=false
        form = femx.Form(ctx, 2, entries=j["integrand"], fmad=fmad)
        A, r, c = form.assemble_coo(mesh)
        assert relF(A.cpu().numpy(), g["A"]) <= TOL64
        assert np.array_equal(r.cpu().numpy(), g["rowA"])
        form.close()
    # fp32, as the reference runs it: tolerance 1e-5
    mesh32 = host_mesh_to_dev(2, g["conn"], (g["X"], g["Y"]), femx.F32)
    form = femx.Form(ctx, 2, entries=j["integrand"], dtype=femx.F32, fmad=False)
    A, _, _ = form.assemble_coo(mesh32)
    assert relF(A.cpu().numpy(), g["A"]) <= TOL32
    form.close()


@pytest.mark.parametrize("dim,builtin,nd", [(2, femx.POISSON_MASS, 1), (2, femx.MASS, 1), (3, femx.POISSON, 1),
                                           (3, femx.POISSON_MASS, 1), (2, femx.ELASTICITY, 2),
                                           (3, femx.ELASTICITY, 3)])
def test_coo_builtin_forms(ctx, dim, builtin, nd):
    rng = np.random.RandomState(12345)
    params = (0.5769, 0.3846) if builtin == femx.ELASTICITY else (2.5,)
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(0, 1, 0, 2, 9, 6)
        X = X + rng.uniform(-0.03, 0.03, X.shape)
        coords = (X, Y)
        oc = (X, Y, None)
    else:
        X, Y, Z, conn = orc.box_mesh(4, 3, 5)
        Z = Z + rng.uniform(-0.02, 0.02, Z.shape)
        coords = (X, Y, Z)
        oc = coords
    mesh = host_mesh_to_dev(dim, conn, coords)
    form = femx.Form(ctx, dim, builtin, nd=nd, params=params)
    A, r, c = form.assemble_coo(mesh)
    oA, orow, ocol = orc.assemble_coo(FORM_IDS[builtin], dim, nd, conn, *oc, params=params)
    assert np.array_equal(r.cpu().numpy(), orow) and np.array_equal(c.cpu().numpy(), ocol)
    assert relF(A.cpu().numpy(), oA) <= TOL64
    form.close()


# ------------------------------------------- kernel ABI #2: pattern + numeric ---
@pytest.mark.parametrize("case", ["ref_poisson2d_2x2.npz", "ref_poisson2d_4x4.npz", "ref_poisson2d_10x7.npz",
                                  "ref_poisson2d_jitter12.npz"])
def test_pattern_and_ell_against_reference_golden(ctx, golden_dir, case):
    g = np.load(os.path.join(golden_dir, case))
    mesh = host_mesh_to_dev(2, g["conn"], (g["X"], g["Y"]))
    pat = femx.Pattern(ctx, mesh)
    ln, idx = pat.ell(7)
    assert np.array_equal(ln.cpu().numpy(), g["ell_len"])
    assert np.array_equal(idx.cpu().numpy(), g["ell_idx"])
    form = femx.Form(ctx, 2, femx.POISSON)
    vals = form.assemble_csr(pat, mesh)
    ell = pat.values_to_ell(vals, 7).cpu().numpy()
    assert relF(ell, g["ell_val"]) <= TOL64
    with pytest.raises(femx.FemxError):
        pat.ell(pat.max_row - 1)
    form.close(); pat.close()


def _csr_case(ctx, dim, builtin, nd, conn, coords, params=(), dtype=femx.F64, tol=TOL64):
    n = len(coords[0])
    mesh = host_mesh_to_dev(dim, conn, coords, dtype)
    pat = femx.Pattern(ctx, mesh, nd=nd)
    rp, ci = pat.csr("int64")
    orp, oci = orc.pattern(conn, n)
    drp, dci = orc.expand_pattern(nd, orp, oci) if nd > 1 else (orp, oci)
    assert pat.nnz == drp[-1] and pat.n_rows == n * nd
    assert np.array_equal(rp.cpu().numpy(), drp), "row_ptr not bit-exact"
    assert np.array_equal(ci.cpu().numpy(), dci), "col_idx not bit-exact"
    rp32, _ = pat.csr("int32")
    assert np.array_equal(rp32.cpu().numpy().astype(np.int64), drp)
    form = femx.Form(ctx, dim, builtin, nd=nd, params=params, dtype=dtype)
    import torch
    vals = torch.full((pat.nnz,), float("nan"), dtype=torch.float64 if dtype == femx.F64 else torch.float32, device="cuda")
    form.assemble_csr(pat, mesh, vals)          # every value must be overwritten (NaN sentinel)
    assert not torch.isnan(vals).any()
    oc = coords if dim == 3 else (coords[0], coords[1], None)
    ov = orc.assemble_csr(FORM_IDS[builtin], dim, nd, conn, *oc, drp, dci, params=params)
    assert relF(vals.cpu().numpy(), ov) <= tol
    # bitwise run-to-run determinism
    vals2 = form.assemble_csr(pat, mesh)
    assert np.array_equal(vals.cpu().numpy(), vals2.cpu().numpy())
    # element-expanded coordinate layout (the reference's X[nn*e+k]): the same bits as the node layout on the
    # owner-computes passes; the element-once lattice pass (3-D structured meshes, node layout only) sums in a
    # different order and agrees to rounding
    vals3 = form.assemble_csr(pat, mesh.expanded(ctx))
    ctx.set_option("lattice", 0)
    try:
        vals4 = form.assemble_csr(pat, mesh)
    finally:
        ctx.set_option("lattice", 1)
    assert np.array_equal(vals4.cpu().numpy(), vals3.cpu().numpy())
    assert relF(vals.cpu().numpy(), vals3.cpu().numpy()) <= (1e-14 if dtype == femx.F64 else 1e-6)
    # SpMV against the oracle
    x = np.random.RandomState(7).uniform(-1, 1, n * nd)
    y = pat.spmv(vals, to_dev(x, vals.dtype)).cpu().numpy()
    assert relF(y, orc.spmv(drp, dci, ov, x)) <= max(tol, 1e-13) * 10
    form.close(); pat.close()


def test_csr_config1_poisson_64x64(ctx):
    """BASELINE config 1: 2-D P1 Poisson, 64x64 unit square, 8,192 triangles, 29,057 nnz."""
    X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, 64, 64)
    _csr_case(ctx, 2, femx.POISSON, 1, conn, (X, Y))


def test_csr_reference_config_1000x100(ctx):
    """The reference's configured sparse2 run (fea_test_sm_sym_sparse2.cu:16-17, 302): 703,301 nnz."""
    X, Y, _, conn = orc.rect_mesh(-3, 3, -3, 3, 1000, 100)
    _csr_case(ctx, 2, femx.POISSON, 1, conn, (X, Y))


def test_csr_fp32(ctx):
    X, Y, _, conn = orc.rect_mesh(-3, 3, -3, 3, 40, 30)
    _csr_case(ctx, 2, femx.POISSON, 1, conn, (X, Y), dtype=femx.F32, tol=TOL32)


@pytest.mark.parametrize("builtin", [femx.POISSON, femx.POISSON_MASS])
def test_csr_tets(ctx, builtin):
    X, Y, Z, conn = orc.box_mesh(9, 7, 8)
    rng = np.random.RandomState(3)
    X = X + rng.uniform(-0.01, 0.01, X.shape)
    _csr_case(ctx, 3, builtin, 1, conn, (X, Y, Z), params=(1.0,))


def test_csr_elasticity_3d(ctx):
    X, Y, Z, conn = orc.box_mesh(6, 5, 4)
    _csr_case(ctx, 3, femx.ELASTICITY, 3, conn, (X, Y, Z), params=(0.5769, 0.3846))


def test_csr_elasticity_2d(ctx):
    X, Y, _, conn = orc.rect_mesh(0, 2, 0, 1, 12, 17)
    _csr_case(ctx, 2, femx.ELASTICITY, 2, conn, (X, Y), params=(1.2, 0.7))


def test_csr_unstructured_shuffled_numbering(ctx):
    """Random node renumbering + random element order + random local rotations:
    nothing in the engine may rely on the structured numbering."""
    rng = np.random.RandomState(12345)
    X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, 23, 31)
    n = len(X)
    perm = rng.permutation(n)            # new id of old node
    X2 = np.empty(n); Y2 = np.empty(n)
    X2[perm] = X; Y2[perm] = Y
    conn2 = perm[conn].astype(np.int32)
    conn2 = conn2[rng.permutation(len(conn2))]
    rot = rng.randint(0, 3, len(conn2))
    conn2 = np.stack([np.roll(c, k) for c, k in zip(conn2, rot)]).astype(np.int32)
    # flip a few elements to clockwise: the reference's signed jac gives negated entries
    conn2[::7] = conn2[::7][:, [0, 2, 1]]
    _csr_case(ctx, 2, femx.POISSON_MASS, 1, conn2, (X2, Y2))


def test_csr_high_valence_fan(ctx):
    """A fan of 40 triangles around one node: row length 41 (> the structured 7)."""
    k = 40
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    X = np.concatenate([[0.0], np.cos(ang)]); Y = np.concatenate([[0.0], np.sin(ang)])
    conn = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], np.int32)
    _csr_case(ctx, 2, femx.POISSON, 1, conn, (X, Y))


def test_pattern_rejects_bad_connectivity(ctx):
    conn = np.array([[0, 1, 5]], np.int32)
    mesh = host_mesh_to_dev(2, conn, (np.zeros(3), np.zeros(3)))
    with pytest.raises(femx.FemxError) as ei:
        femx.Pattern(ctx, mesh)
    assert ei.value.status == 1


def test_empty_mesh(ctx):
    import torch
    mesh = femx.Mesh(2, torch.empty((0, 3), dtype=torch.int32, device="cuda"),
                     (torch.zeros(4, dtype=torch.float64, device="cuda"),) * 2)
    pat = femx.Pattern(ctx, mesh)
    assert pat.nnz == 0 and pat.n_rows == 4
    rp, ci = pat.csr("int64")
    assert rp.cpu().tolist() == [0] * 5
    form = femx.Form(ctx, 2, femx.POISSON)
    A, r, c = form.assemble_coo(mesh)
    assert A.numel() == 0
    v = form.assemble_csr(pat, mesh)
    assert v.numel() == 0
    form.close(); pat.close()


def test_form_pattern_mismatch_is_an_error(ctx):
    m = ctx.rectangle_mesh(0, 1, 0, 1, 4, 4)
    pat = femx.Pattern(ctx, m)
    f3 = femx.Form(ctx, 3, femx.POISSON)
    with pytest.raises(femx.FemxError):
        f3.assemble_csr(pat, m)
    f3.close(); pat.close()


# --------------------------------------------------- slabs (multi-GPU layout) ---
@pytest.mark.parametrize("parts", [2, 3])
def test_slab_rows_concatenate_to_global_matrix_2d(ctx, parts):
    """Owned-row slabs with ghost elements: concatenation == single-device matrix, bitwise."""
    import torch
    nR, nC = 24, 13
    whole = ctx.rectangle_mesh(0, 1, 0, 1, nR, nC)
    form = femx.Form(ctx, 2, femx.POISSON_MASS)
    pw = femx.Pattern(ctx, whole)
    vw = form.assemble_csr(pw, whole)
    rpw, ciw = pw.csr("int64")
    vals, cols, lens = [], [], []
    bounds = [round(p * (nR + 1) / parts) for p in range(parts + 1)]  # owned node rows
    for p in range(parts):
        r0, r1 = bounds[p], bounds[p + 1]          # owned grid rows [r0, r1)
        lo, hi = max(r0 - 1, 0), min(r1, nR)       # slab incl. ghost rows: node rows [lo, hi]
        slab = ctx.rectangle_mesh(0, 1, 0, 1, nR, nC, row_lo=lo, row_hi=hi)
        base = lo * (nC + 1)
        pat = femx.Pattern(ctx, slab, row_begin=(r0 - lo) * (nC + 1), row_end=(r1 - lo) * (nC + 1), col_base=base)
        v = form.assemble_csr(pat, slab)
        rp, ci = pat.csr("int64")
        vals.append(v); cols.append(ci); lens.append(torch.diff(rp))
        pat.close()
    assert torch.equal(torch.cat(lens), torch.diff(rpw))
    assert torch.equal(torch.cat(cols), ciw)
    assert torch.equal(torch.cat(vals), vw)
    form.close(); pw.close()


def test_slab_rows_concatenate_to_global_matrix_3d(ctx):
    import torch
    n = 6
    whole = ctx.box_mesh(n, n, n)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    pw = femx.Pattern(ctx, whole)
    vw = form.assemble_csr(pw, whole)
    rpw, ciw = pw.csr("int64")
    plane = (n + 1) * (n + 1)
    vals, cols = [], []
    for (k0, k1) in [(0, 3), (3, 7)]:              # owned node planes [k0, k1)
        lo, hi = max(k0 - 1, 0), min(k1, n)
        slab = ctx.box_mesh(n, n, n, k_lo=lo, k_hi=hi)
        pat = femx.Pattern(ctx, slab, row_begin=(k0 - lo) * plane, row_end=(k1 - lo) * plane, col_base=lo * plane)
        vals.append(form.assemble_csr(pat, slab)); cols.append(pat.csr("int64")[1])
        pat.close()
    assert torch.equal(torch.cat(cols), ciw)
    assert torch.equal(torch.cat(vals), vw)
    form.close(); pw.close()


# ------------------------------------------- size-independent properties ---
def test_large_mesh_properties(ctx):
    """1024x1024 (2.1M elements): closed-form nnz, zero row sums, symmetry via x^T A y = y^T A x,
    COO and CSR agree (sum of triplets == sum of CSR values), determinism."""
    import torch
    n = 1024
    mesh = ctx.rectangle_mesh(0, 1, 0, 1, n, n)
    pat = femx.Pattern(ctx, mesh)
    assert pat.nnz == (n + 1) ** 2 + 2 * (2 * n * (n + 1) + n * n) and pat.max_row == 7
    form = femx.Form(ctx, 2, femx.POISSON)
    v = form.assemble_csr(pat, mesh)
    ones = torch.ones(pat.n_rows, dtype=torch.float64, device="cuda")
    assert pat.spmv(v, ones).abs().max().item() < 1e-10
    g = torch.Generator(device="cuda"); g.manual_seed(12345)
    x = torch.rand(pat.n_rows, dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand(pat.n_rows, dtype=torch.float64, device="cuda", generator=g)
    a = torch.dot(x, pat.spmv(v, y)).item(); b = torch.dot(y, pat.spmv(v, x)).item()
    assert abs(a - b) <= 1e-11 * abs(a)
    A, r, c = form.assemble_coo(mesh)
    # scatter the triplets with index_add (order-dependent, so compare with tolerance)
    rp, ci = pat.csr("int64")
    dense_rows = torch.zeros(pat.n_rows, dtype=torch.float64, device="cuda")
    dense_rows.index_add_(0, r.long(), A * x[c.long()])
    assert (dense_rows - pat.spmv(v, x)).norm().item() <= 1e-12 * dense_rows.norm().item()
    assert torch.equal(v, form.assemble_csr(pat, mesh))
    form.close(); pat.close()


# ------------------------------------------------ validation layer: SpMV + CG ---
def test_cg_matches_oracle_history(ctx):
    """cfg5 in miniature: assembled 3-D operator, b = A 1, CG from 0 — residual history vs the oracle's CG."""
    import torch
    from femx.dist import SlabOperator, make_slab
    n = 10
    mesh = ctx.box_mesh(n, n, n)
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    vals = form.assemble_csr(pat, mesh)
    slab = make_slab(0, 1, n, (n + 1) ** 2)
    op = SlabOperator(ctx, pat, vals, slab)
    ones = torch.ones(pat.n_rows, dtype=torch.float64, device="cuda")
    b = op.matvec(ones).clone()
    x, hist = op.cg(b, 40)
    X, Y, Z, conn = orc.box_mesh(n, n, n)
    rp, ci = orc.pattern(conn, len(X))
    ov = orc.assemble_csr(orc.POISSON_MASS, 3, 1, conn, X, Y, Z, rp, ci)
    ob = orc.spmv(rp, ci, ov, np.ones(len(X)))
    ox, ores = orc.cg(rp, ci, ov, ob, 40)
    h = hist.cpu().numpy()
    assert len(h) == len(ores) == 41
    assert np.all(np.abs(h - ores) <= 1e-8 * ores[0] + 1e-6 * ores)
    assert h[-1] < 1e-2 * h.max()
    assert np.abs(x.cpu().numpy() - ox).max() < 1e-8
    form.close(); pat.close()


# ------------------------------------------------------- load vector ---
@pytest.mark.parametrize("case", ["ref_poisson2d_2x2.npz", "ref_poisson2d_10x7.npz", "ref_poisson2d_jitter12.npz"])
def test_rhs_reference_strings(ctx, golden_dir, case):
    """b = sum_e sum_q w_q rhs_li with the RHS strings the reference generates and discards."""
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    g = np.load(os.path.join(golden_dir, case))
    mesh = host_mesh_to_dev(2, g["conn"], (g["X"], g["Y"]))
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 2, entries=j["integrand"], rhs=j["rhs"])
    b = form.assemble_rhs(pat, mesh)
    assert relF(b.cpu().numpy(), g["rhs"]) <= TOL64
    b2 = form.assemble_rhs(pat, mesh.expanded(ctx))
    assert np.array_equal(b.cpu().numpy(), b2.cpu().numpy())
    form.close(); pat.close()


@pytest.mark.parametrize("dim,builtin,nd,fvec", [(2, femx.POISSON, 1, (2.5,)), (3, femx.POISSON_MASS, 1, (1.0,)),
                                                (3, femx.ELASTICITY, 3, (0.1, -0.2, -9.81))])
def test_rhs_constant_source(ctx, dim, builtin, nd, fvec):
    rng = np.random.RandomState(4)
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(0, 1, 0, 2, 14, 9)
        X = X + rng.uniform(-0.02, 0.02, X.shape)
        coords, oc = (X, Y), (X, Y, None)
    else:
        X, Y, Z, conn = orc.box_mesh(5, 6, 4)
        Y = Y + rng.uniform(-0.02, 0.02, Y.shape)
        coords = oc = (X, Y, Z)
    mesh = host_mesh_to_dev(dim, conn, coords)
    pat = femx.Pattern(ctx, mesh, nd=nd)
    form = femx.Form(ctx, dim, builtin, nd=nd, params=(0.6, 0.4), rhs_vec=fvec)
    b = form.assemble_rhs(pat, mesh)
    ob, _ = orc.assemble_rhs(0, dim, nd, conn, *oc, fvec=fvec)
    assert relF(b.cpu().numpy(), ob) <= TOL64
    assert np.array_equal(b.cpu().numpy(), form.assemble_rhs(pat, mesh).cpu().numpy())
    form.close(); pat.close()


def test_poisson_solve_manufactured(ctx):
    """A x = b end to end: -lap(u) + u = f with the assembled operator and load vector; CG recovers
    the discrete solution the oracle's CG finds (the reference's f, its mesh, its quadrature)."""
    import torch
    from femx.dist import SlabOperator, make_slab
    j = json.load(open(os.path.join(GOLDEN_DIR, "ref_integrand_strings.json")))
    nR = nC = 24
    mesh = ctx.rectangle_mesh(-3.0, 3.0, -3.0, 3.0, nR, nC)
    pat = femx.Pattern(ctx, mesh)
    fA = femx.Form(ctx, 2, femx.POISSON_MASS)
    fb = femx.Form(ctx, 2, entries=j["integrand"], rhs=j["rhs"])
    vals = fA.assemble_csr(pat, mesh)
    b = fb.assemble_rhs(pat, mesh)
    op = SlabOperator(ctx, pat, vals, make_slab(0, 1, nR, nC + 1))
    x, hist = op.cg(b, 200)
    X, Y, _, conn = orc.rect_mesh(-3.0, 3.0, -3.0, 3.0, nR, nC)
    rp, ci = orc.pattern(conn, len(X))
    ov = orc.assemble_csr(orc.POISSON_MASS, 2, 1, conn, X, Y, None, rp, ci)
    ob, _ = orc.assemble_rhs(1, 2, 1, conn, X, Y)
    ox, ores = orc.cg(rp, ci, ov, ob, 200)
    assert hist[-1].item() < 1e-8 * hist[0].item()
    assert np.abs(x.cpu().numpy() - ox).max() <= 1e-8 * np.abs(ox).max()
    fA.close(); fb.close(); pat.close()


# --------------------------------------------------- Dirichlet conditions ---
@pytest.mark.parametrize("nd", [1, 2])
def test_dirichlet_elimination_matches_oracle(ctx, nd):
    import torch
    nR, nC = 13, 9
    mesh = ctx.rectangle_mesh(0.0, 1.0, 0.0, 1.0, nR, nC, flags=True)
    builtin = femx.POISSON_MASS if nd == 1 else femx.ELASTICITY
    form = femx.Form(ctx, 2, builtin, nd=nd, params=(1.0, 0.5) if nd > 1 else (1.0,), rhs_vec=(1.0, -2.0))
    pat = femx.Pattern(ctx, mesh, nd=nd)
    vals = form.assemble_csr(pat, mesh)
    b = form.assemble_rhs(pat, mesh)
    flag = mesh.flag.repeat_interleave(nd).contiguous()           # the reference's boundary flag, per dof
    g = torch.linspace(-1.0, 1.0, pat.n_rows, dtype=torch.float64, device="cuda")
    X, Y, oflag, conn = orc.rect_mesh(0.0, 1.0, 0.0, 1.0, nR, nC)
    rp, ci = orc.pattern(conn, len(X))
    drp, dci = orc.expand_pattern(nd, rp, ci) if nd > 1 else (rp, ci)
    ov = vals.cpu().numpy().copy(); ob = b.cpu().numpy().copy()
    orc.apply_dirichlet(drp, dci, np.repeat(oflag, nd), g.cpu().numpy(), ov, ob)
    pat.apply_dirichlet(flag, g, vals, b)
    assert np.array_equal(vals.cpu().numpy(), ov)
    assert np.allclose(b.cpu().numpy(), ob, rtol=1e-14, atol=1e-15)
    # the constrained system is symmetric and solving it reproduces g on the boundary
    import scipy.sparse as sp
    import scipy.sparse.linalg
    A = sp.csr_matrix((vals.cpu().numpy(), dci, drp))
    assert abs(A - A.T).max() < 1e-13
    x = sp.linalg.spsolve(A.tocsc(), b.cpu().numpy())
    fl = np.repeat(oflag, nd).astype(bool)
    assert np.abs(x[fl] - g.cpu().numpy()[fl]).max() < 1e-12
    form.close(); pat.close()


# ---------------------------------------------------------------- edge cases ---
def test_row_longer_than_127_columns_is_unsupported(ctx):
    k = 130
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    X = np.concatenate([[0.0], np.cos(ang)]); Y = np.concatenate([[0.0], np.sin(ang)])
    conn = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], np.int32)
    mesh = host_mesh_to_dev(2, conn, (X, Y))
    with pytest.raises(femx.FemxError) as ei:
        femx.Pattern(ctx, mesh)
    assert ei.value.status == 4 and "128" in str(ei.value)


def test_oversized_mesh_is_rejected_before_any_work(ctx):
    import ctypes as C
    import torch
    conn = torch.zeros(3, dtype=torch.int32, device="cuda")
    h = C.c_void_p()
    st = femx.lib().femx_pattern_build(ctx.h, 3, 1, C.c_int64(10), C.c_int64(800_000_000), C.c_void_p(conn.data_ptr()),
                                       C.c_int64(0), C.c_int64(10), C.c_int64(0), None, C.byref(h))
    assert st == 4 and b"32-bit" in femx.lib().femx_last_error(ctx.h)


def test_isolated_nodes_and_duplicate_elements(ctx):
    """Nodes that belong to no element give empty rows; an element listed twice contributes twice."""
    X = np.array([0.0, 1.0, 0.0, 5.0, 1.0, 7.0]); Y = np.array([0.0, 0.0, 1.0, 5.0, 1.0, 7.0])
    conn = np.array([[0, 1, 2], [1, 4, 2], [0, 1, 2]], np.int32)          # nodes 3 and 5 are isolated
    mesh = host_mesh_to_dev(2, conn, (X, Y))
    pat = femx.Pattern(ctx, mesh)
    rp, ci = pat.csr("int64")
    orp, oci = orc.pattern(conn, 6)
    assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(ci.cpu().numpy(), oci)
    assert orp[4] - orp[3] == 0 and orp[6] - orp[5] == 0
    form = femx.Form(ctx, 2, femx.POISSON_MASS)
    v = form.assemble_csr(pat, mesh)
    ov = orc.assemble_csr(orc.POISSON_MASS, 2, 1, conn, X, Y, None, orp, oci)
    assert relF(v.cpu().numpy(), ov) <= TOL64
    form.close(); pat.close()


def test_csr_tets_fp32(ctx):
    X, Y, Z, conn = orc.box_mesh(7, 5, 6)
    _csr_case(ctx, 3, femx.POISSON_MASS, 1, conn, (X, Y, Z), params=(1.0,), dtype=femx.F32, tol=TOL32)


def test_degenerate_element_is_rejected(ctx):
    conn = np.array([[0, 1, 2], [1, 1, 2]], np.int32)
    mesh = host_mesh_to_dev(2, conn, (np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0])))
    with pytest.raises(femx.FemxError) as ei:
        femx.Pattern(ctx, mesh)
    assert ei.value.status == 1 and "twice" in str(ei.value)


# ------------------------------------------------------------ unstructured meshes ---
def _delaunay(dim, npts, seed):
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(seed)
    P = rng.uniform(0, 1, (npts, dim))
    tri = Delaunay(P)
    conn = tri.simplices.astype(np.int32)
    # orient positively (the engine keeps the reference's signed jac; mixed signs are legal but make
    # the operator indefinite) and drop slivers the quadrature cannot resolve
    V = P[conn]
    J = np.transpose(V[:, :dim, :] - V[:, dim:dim + 1, :], (0, 2, 1))
    det = np.linalg.det(J)
    keep = np.abs(det) > 1e-9
    conn, det = conn[keep], det[keep]
    neg = det < 0
    conn[neg, 0], conn[neg, 1] = conn[neg, 1].copy(), conn[neg, 0].copy()
    return P, conn


@pytest.mark.parametrize("dim,npts", [(2, 3000), (3, 1500)])
def test_csr_delaunay_mesh(ctx, dim, npts):
    """Random Delaunay triangulation / tetrahedralisation: irregular valence (rows of 3..30+ columns),
    random numbering — nothing structured."""
    P, conn = _delaunay(dim, npts, 12345)
    coords = tuple(np.ascontiguousarray(P[:, k]) for k in range(dim))
    _csr_case(ctx, dim, femx.POISSON_MASS, 1, conn, coords, params=(1.0,))


def test_elasticity_delaunay_3d(ctx):
    P, conn = _delaunay(3, 600, 7)
    coords = tuple(np.ascontiguousarray(P[:, k]) for k in range(3))
    _csr_case(ctx, 3, femx.ELASTICITY, 3, conn, coords, params=(0.5769, 0.3846))
