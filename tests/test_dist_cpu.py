"""CPU tier, world_size 2 over gloo: the host logic of the multi-GPU layer — slab bounds, ghost
layers and the halo exchange — checked with the oracle's matrices (no GPU compute here)."""
import os
import socket

import numpy as np
import pytest

from femx.dist import HaloExchange, make_slab


def test_slabs_cover_all_rows_once():
    for world in (1, 2, 3, 4, 8):
        for cells in (8, 64, 257):
            slabs = [make_slab(r, world, cells, 10) for r in range(world)]
            assert slabs[0].r0 == 0 and slabs[-1].r1 == cells + 1
            for a, b in zip(slabs, slabs[1:]):
                assert a.r1 == b.r0
            for s in slabs:
                assert s.lo == max(s.r0 - 1, 0) and s.hi == min(s.r1, cells)
                assert s.row_end - s.row_begin == s.n_owned
    with pytest.raises(ValueError):
        make_slab(0, 9, 7, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nR, nC, out):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plane = nC + 1
        s = make_slab(rank, world, nR, plane)
        # this rank's slab of the global mesh, local numbering, exactly as femx_mesh_rectangle lays it out
        X, Y, _, conn = orc.rect_mesh(0.0, 1.0, 0.0, 1.0, nR, nC)
        cells = conn.reshape(nR, nC, 2, 3)[s.cells_lo:s.cells_hi].reshape(-1, 3) - s.col_base
        Xl, Yl = X[s.col_base:s.col_base + s.n_local], Y[s.col_base:s.col_base + s.n_local]
        rp, ci = orc.pattern(cells.astype(np.int32), s.n_local)
        v = orc.assemble_csr(orc.POISSON_MASS, 2, 1, cells.astype(np.int32), Xl, Yl, None, rp, ci)
        # owned rows only (rows of ghost planes are incomplete by construction)
        lo_r, hi_r = s.row_begin, s.row_end
        xg = np.random.RandomState(5).uniform(-1, 1, (nR + 1) * plane)  # same global x on every rank
        x_ext = torch.zeros(s.n_local, dtype=torch.float64)
        halo = HaloExchange(s)
        halo.owned_view(x_ext).copy_(torch.from_numpy(xg[s.r0 * plane:s.r1 * plane]))
        halo.exchange(x_ext)
        assert np.array_equal(x_ext.numpy(), xg[s.col_base:s.col_base + s.n_local]), "halo exchange"
        y_ext = orc.spmv(rp, ci, v, x_ext.numpy())
        y_owned = torch.from_numpy(y_ext[lo_r:hi_r].copy())
        ys = [torch.zeros((make_slab(r, world, nR, plane).n_owned,), dtype=torch.float64) for r in range(world)]
        dist.all_gather(ys, y_owned) if len({t.numel() for t in ys}) == 1 else None
        if len({t.numel() for t in ys}) != 1:   # ragged: gather through object lists
            objs = [None] * world
            dist.all_gather_object(objs, y_owned.numpy())
            ys = [torch.from_numpy(o) for o in objs]
        if rank == 0:
            grp, gci = orc.pattern(conn, len(X))
            gv = orc.assemble_csr(orc.POISSON_MASS, 2, 1, conn, X, Y, None, grp, gci)
            yg = orc.spmv(grp, gci, gv, xg)
            got = np.concatenate([t.numpy() for t in ys])
            out.put(float(np.abs(got - yg).max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nR", [(2, 9), (3, 10)])
def test_halo_exchange_and_owned_rows_gloo(world, nR):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nR, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-12


def test_c_abi_slab_equals_python_slab():
    """femx_dist_slab (C++ multi-GPU layer) and femx.dist.make_slab (the gloo-tested host logic) agree."""
    import femx
    for world in (1, 2, 3, 4, 8):
        for cells in (8, 64, 256, 257):
            for r in range(world):
                s = make_slab(r, world, cells, 10)
                assert femx.dist_slab(cells + 1, world, r) == (s.r0, s.r1, s.lo, s.hi)
    with pytest.raises(femx.FemxError):
        femx.dist_slab(7, 9, 0)


def test_dist_layer_needs_a_device():
    """No CPU fallback: the communicator cannot be created without a context, and a context needs a GPU."""
    import torch
    import femx
    if torch.cuda.is_available():
        pytest.skip("CPU-tier check")
    with pytest.raises(femx.FemxError):
        femx.Context(0)
