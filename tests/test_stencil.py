"""Stencil-class specialisation of the numeric pass (femx_pattern_stencil / FEMX_SPEC_BODY).

CPU tier: the specialised kernel compiles offline for the interior stencils of the structured
meshes.  GPU tier: the symbolic pass finds exactly the class a CPU restatement of its rules
predicts, and the specialised pass produces the SAME BITS as the generic incidence loop."""
import os

import numpy as np
import pytest

import femx
from oracle import oracle as orc
from tools.stencil_offline import interior_class, row_codes


@pytest.mark.parametrize("dim,form,np_,rlen,self_pos", [(2, "POISSON", 6, 7, 3), (2, "POISSON_MASS", 6, 7, 3),
                                                        (3, "POISSON", 24, 15, 7), (3, "POISSON_MASS", 24, 15, 7)])
def test_specialised_kernel_compiles_offline(dim, form, np_, rlen, self_pos):
    codes, r, s = interior_class(dim)
    assert (len(codes), r, s) == (np_, rlen, self_pos)
    f = femx.Form(None, dim, getattr(femx, form), offline=True)
    cubin = f.cubin_stencil(codes, r, s)
    assert len(cubin) > 1000
    src = f.source
    assert "#define FEMX_SPEC 1" in src and "FEMX_SPEC_BODY" in src
    # every column of the class is stored exactly once
    body = src[src.index("#define FEMX_SPEC_BODY"):]
    body = body[:body.index("\n\n") if "\n\n" in body else len(body)]
    for k in range(r):
        assert body.count(f"srow[{k}] =") == 1
    f.close()


def test_specialised_kernel_rejects_bad_class():
    f = femx.Form(None, 2, femx.POISSON, offline=True)
    codes, r, s = interior_class(2)
    with pytest.raises(femx.FemxError):
        f.cubin_stencil(codes, 3, 5)                 # own position outside the row
    with pytest.raises(femx.FemxError):
        f.cubin_stencil([c | 127 for c in codes], r, s)  # a code naming a column beyond the row
    f.close()
    g = femx.Form(None, 3, femx.ELASTICITY, nd=3, params=(0.5, 0.4), offline=True)
    with pytest.raises(femx.FemxError):
        g.cubin_stencil(interior_class(3)[0], 15, 7)  # vector forms have no specialised pass
    g.close()


# ------------------------------------------------------------------ GPU tier ---
def _dev_mesh(dim, conn, coords):
    import torch
    return femx.Mesh(dim, torch.from_numpy(np.ascontiguousarray(conn)).cuda(),
                     tuple(torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in coords))


@pytest.mark.gpu
def test_dominant_class_2d(ctx):
    nR, nC = 16, 12
    X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, nR, nC)
    pat = femx.Pattern(ctx, _dev_mesh(2, conn, (X, Y)))
    st = pat.stencil()
    row = (nR // 2) * (nC + 1) + nC // 2
    codes, rlen, self_pos = row_codes(conn, 3, row)
    assert (st["n_incid"], st["row_len"], st["self_pos"]) == (6, 7, 3) == (len(codes), rlen, self_pos)
    assert st["codes"] == codes
    assert st["rows"] == (nR - 1) * (nC - 1)         # every interior node, nothing else
    assert st["offsets"] == [-(nC + 1), -nC, -1, 0, 1, nC, nC + 1]
    pat.close()


@pytest.mark.gpu
def test_dominant_class_sampling_does_not_alias_with_the_mesh_lines(ctx):
    """rows / 256 a multiple of the line length: evenly spaced samples would all sit in one mesh column
    (the boundary column for the 2-GPU slabs of cfg2) — the samples are jittered inside their stride."""
    nR, nC = 511, 16            # 512 x 17 nodes, 8704 / 256 = 34 = 2 lines
    mesh = ctx.rectangle_mesh(0, 1, 0, 1, nR, nC)
    pat = femx.Pattern(ctx, mesh)
    st = pat.stencil()
    assert (st["n_incid"], st["row_len"]) == (6, 7) and st["rows"] == (nR - 1) * (nC - 1)
    pat.close()
    # a slab as bench.py builds it for rank 0 of 2: owned rows = whole lines
    nR, nC = 64, 255            # 256 nodes per line
    slab = ctx.rectangle_mesh(0, 1, 0, 1, nR, nC, row_lo=0, row_hi=33)
    pat = femx.Pattern(ctx, slab, row_begin=0, row_end=32 * (nC + 1), col_base=0)
    st = pat.stencil()
    assert (st["n_incid"], st["row_len"]) == (6, 7) and st["rows"] == 31 * (nC - 1)
    pat.close()


@pytest.mark.gpu
def test_dominant_class_3d(ctx):
    nx, ny, nz = 7, 6, 5
    X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
    pat = femx.Pattern(ctx, _dev_mesh(3, conn, (X, Y, Z)))
    st = pat.stencil()
    row = ((nz // 2) * (ny + 1) + ny // 2) * (nx + 1) + nx // 2
    codes, rlen, self_pos = row_codes(conn, 4, row)
    assert (st["n_incid"], st["row_len"], st["self_pos"]) == (24, 15, 7)
    assert st["codes"] == codes
    assert st["rows"] == (nx - 1) * (ny - 1) * (nz - 1)
    px, pxy = nx + 1, (nx + 1) * (ny + 1)
    half = [1, px, px + 1, pxy, pxy + 1, pxy + px, pxy + px + 1]     # Kuhn: the 7 neighbours "ahead"
    assert st["offsets"] == sorted([-o for o in half] + [0] + half)
    pat.close()


@pytest.mark.gpu
def test_no_class_on_unstructured_or_vector_patterns(ctx):
    from scipy.spatial import Delaunay
    pts = np.random.RandomState(5).uniform(0, 1, (1500, 2))
    tri = Delaunay(pts).simplices.astype(np.int32)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    flip = ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    pat = femx.Pattern(ctx, _dev_mesh(2, tri, (pts[:, 0].copy(), pts[:, 1].copy())))
    assert pat.stencil()["rows"] * 4 < pat.n_rows     # at most a stray minority
    pat.close()
    X, Y, Z, conn = orc.box_mesh(5, 5, 5)
    pat = femx.Pattern(ctx, _dev_mesh(3, conn, (X, Y, Z)), nd=3)
    assert pat.stencil()["rows"] == 0
    pat.close()


def _both_paths(ctx, form, pat, mesh):
    """assemble_csr with the specialised body and with the stencil-class pass switched off (generic loop for
    every row).  The element-once lattice pass (tests/test_lattice.py) is switched off for both: it is a
    different summation order, equal to these to rounding, not to the bit."""
    import torch
    ctx.set_option("lattice", 0)
    try:
        v_spec = form.assemble_csr(pat, mesh)
        n_variants = form.source.count("#define FEMX_SPEC 1")
        ctx.set_option("spec", 0)
        v_gen = form.assemble_csr(pat, mesh)
        assert "#define FEMX_SPEC 0" in form.source
    finally:
        ctx.set_option("spec", 1)
        ctx.set_option("lattice", 1)
    torch.cuda.synchronize()
    assert n_variants == 1, "the specialised kernel was not the one launched"
    return v_spec, v_gen


@pytest.mark.gpu
@pytest.mark.parametrize("dim,builtin,dtype", [(2, femx.POISSON, femx.F64), (2, femx.POISSON_MASS, femx.F64),
                                               (2, femx.MASS, femx.F64), (3, femx.POISSON, femx.F64),
                                               (3, femx.POISSON_MASS, femx.F64), (3, femx.MASS, femx.F64),
                                               (2, femx.POISSON, femx.F32), (3, femx.POISSON_MASS, femx.F32)])
def test_specialised_pass_equals_generic_bitwise(ctx, dim, builtin, dtype):
    """Jittered structured meshes (every element has its own geometry): same bits from both paths,
    and the oracle's values within tolerance."""
    import torch
    rng = np.random.RandomState(11)
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(-1, 2, 0, 1, 37, 29)
        coords = [X + rng.uniform(-0.01, 0.01, X.shape), Y + rng.uniform(-0.01, 0.01, Y.shape)]
    else:
        X, Y, Z, conn = orc.box_mesh(13, 11, 9)
        coords = [c + rng.uniform(-0.01, 0.01, c.shape) for c in (X, Y, Z)]
    tdt = torch.float64 if dtype == femx.F64 else torch.float32
    mesh = femx.Mesh(dim, torch.from_numpy(conn).cuda(), tuple(torch.from_numpy(c).to(tdt).cuda() for c in coords))
    pat = femx.Pattern(ctx, mesh)
    assert pat.stencil()["rows"] * 2 >= pat.n_rows
    form = femx.Form(ctx, dim, builtin, params=(1.5,), dtype=dtype)
    v_spec, v_gen = _both_paths(ctx, form, pat, mesh)
    assert torch.equal(v_spec, v_gen)
    if dtype == femx.F64:
        orp, oci = orc.pattern(conn, len(X))
        oid = {femx.POISSON: orc.POISSON, femx.POISSON_MASS: orc.POISSON_MASS, femx.MASS: orc.MASS}[builtin]
        oc = coords if dim == 3 else (coords[0], coords[1], None)
        ov = orc.assemble_csr(oid, dim, 1, conn, *oc, orp, oci, params=(1.5,))
        assert np.linalg.norm(v_spec.cpu().numpy() - ov) / np.linalg.norm(ov) <= 1e-12
    form.close(); pat.close()


@pytest.mark.gpu
def test_specialised_pass_custom_strings_fmad_off(ctx, golden_dir):
    """The reference's own integrand strings (compiled with --fmad=false as the reference does) take the
    specialised pass too; with contraction on they stay on the generic loop."""
    import json
    import torch
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    X, Y, _, conn = orc.rect_mesh(-3, 3, -3, 3, 20, 15)
    rng = np.random.RandomState(2)
    X = X + rng.uniform(-0.02, 0.02, X.shape)
    mesh = _dev_mesh(2, conn, (X, Y))
    pat = femx.Pattern(ctx, mesh)
    form = femx.Form(ctx, 2, entries=j["integrand"], fmad=False)
    v_spec, v_gen = _both_paths(ctx, form, pat, mesh)
    assert torch.equal(v_spec, v_gen)
    form.close()
    form = femx.Form(ctx, 2, entries=j["integrand"], fmad=True)
    form.assemble_csr(pat, mesh)
    assert "#define FEMX_SPEC 0" in form.source
    form.close(); pat.close()


@pytest.mark.gpu
def test_specialised_pass_partial_and_mixed_tiles(ctx):
    """Mesh sizes that leave ragged tiles and mixed (boundary + interior) tiles everywhere."""
    import torch
    for nR, nC in ((3, 3), (5, 200), (129, 2), (64, 127)):
        mesh = ctx.rectangle_mesh(0, 1, 0, 2, nR, nC)
        pat = femx.Pattern(ctx, mesh)
        form = femx.Form(ctx, 2, femx.POISSON_MASS)
        v1 = form.assemble_csr(pat, mesh)
        ctx.set_option("spec", 0)
        try:
            v0 = form.assemble_csr(pat, mesh)
        finally:
            ctx.set_option("spec", 1)
        assert torch.equal(v0, v1), (nR, nC)
        form.close(); pat.close()
