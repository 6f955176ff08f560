"""Multi-GPU layer: owned-row slabs with ghost elements (assembly needs NO communication),
plus the validation solver — SpMV / CG whose halo exchange and global reductions go through
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).

No reference counterpart (the reference is single-GPU: job.pbs:4,24 launches one rank).
Partition: 1-D slabs along the slowest mesh axis (node rows in 2-D, node planes in 3-D).
Rank p owns node planes [r0, r1); its slab holds planes [lo, hi] = [max(r0-1,0), min(r1, n-1)]
so that every element touching an owned row is present (ghost elements duplicated on both
sides).  Local node id = global id - lo*plane; CSR columns are global.
"""
from dataclasses import dataclass


@dataclass
class Slab:
    rank: int
    world: int
    n_planes: int   # node planes along the sharded axis (cells + 1)
    plane: int      # nodes per plane
    r0: int         # owned planes [r0, r1)
    r1: int
    lo: int         # slab planes [lo, hi]
    hi: int

    @property
    def row_begin(self):      # local node range of the owned rows
        return (self.r0 - self.lo) * self.plane

    @property
    def row_end(self):
        return (self.r1 - self.lo) * self.plane

    @property
    def col_base(self):
        return self.lo * self.plane

    @property
    def n_owned(self):
        return (self.r1 - self.r0) * self.plane

    @property
    def n_local(self):
        return (self.hi - self.lo + 1) * self.plane

    @property
    def cells_lo(self):       # cell layers [lo, hi)
        return self.lo

    @property
    def cells_hi(self):
        return self.hi


def make_slab(rank, world, n_cells_axis, plane):
    """Even split of the n_cells_axis+1 node planes over `world` ranks."""
    n_planes = n_cells_axis + 1
    if world > n_planes:
        raise ValueError(f"cannot split {n_planes} node planes over {world} ranks")
    r0 = (rank * n_planes) // world
    r1 = ((rank + 1) * n_planes) // world
    lo = max(r0 - 1, 0)
    hi = min(r1, n_planes - 1)
    return Slab(rank, world, n_planes, plane, r0, r1, lo, hi)


class HaloExchange:
    """x_ext covers the slab's node planes [lo, hi] (times nd dofs); the owned part is filled by the
    caller, exchange() fills the ghost plane(s) from the neighbouring ranks."""

    def __init__(self, slab, nd=1, group=None):
        self.s = slab
        self.nd = nd
        self.group = group

    def owned_view(self, x_ext):
        s = self.s
        return x_ext[s.row_begin * self.nd: s.row_end * self.nd]

    def exchange(self, x_ext):
        import torch.distributed as dist
        s, nd = self.s, self.nd
        if s.world == 1:
            return
        w = s.plane * nd
        ops = []
        own = self.owned_view(x_ext)
        if s.rank > 0:       # lower neighbour: send my first owned plane, receive its last owned plane
            ops.append(dist.P2POp(dist.isend, own[:w].contiguous(), s.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, x_ext[:w], s.rank - 1, self.group))
        if s.rank < s.world - 1:
            ops.append(dist.P2POp(dist.isend, own[-w:].contiguous(), s.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, x_ext[(s.hi - s.lo) * w:], s.rank + 1, self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class SlabOperator:
    """The assembled rows of one rank: y_owned = A[rows] @ x (device), and an unpreconditioned CG
    across ranks.  Everything stays on the device; alpha/beta are never read back in the loop."""

    def __init__(self, ctx, pattern, values, slab, nd=1, group=None):
        import torch
        self.ctx, self.pat, self.vals, self.slab, self.nd = ctx, pattern, values, slab, nd
        self.halo = HaloExchange(slab, nd, group)
        self.group = group
        self.dev = values.device
        self.x_ext = torch.zeros(slab.n_local * nd, dtype=values.dtype, device=self.dev)

    def _allreduce(self, t):
        import torch.distributed as dist
        if self.slab.world > 1:
            dist.all_reduce(t, group=self.group)

    def matvec(self, x_owned, y=None):
        """y = A[owned rows] @ x, x given by its owned part on every rank."""
        self.halo.owned_view(self.x_ext).copy_(x_owned)
        self.halo.exchange(self.x_ext)
        return self.pat.spmv(self.vals, self.x_ext, x_base=self.slab.col_base * self.nd, y=y)

    def cg(self, b, iters):
        """iters steps of CG from x0 = 0.  Returns x (owned part) and the residual-norm history
        (iters+1 values, device tensor)."""
        import ctypes as C

        import torch

        from . import F32, F64, _dbl, _i64, _stream, _vp, lib
        L, ctx = lib(), self.ctx
        dt = F64 if b.dtype == torch.float64 else F32
        n = b.numel()
        x = torch.zeros_like(b)
        r = b.clone()
        p = b.clone()
        Ap = torch.empty_like(b)
        hist = torch.zeros(iters + 1, dtype=torch.float64, device=self.dev)
        rr = torch.zeros(2, dtype=torch.float64, device=self.dev)      # [r.r, unused]
        pap = torch.zeros(2, dtype=torch.float64, device=self.dev)     # [p.Ap, unused]
        rr_new = torch.zeros(2, dtype=torch.float64, device=self.dev)
        st = _stream(None)
        ctx.check(L.femx_dot2(ctx.h, dt, _i64(n), _vp(r), _vp(r), _vp(None), _vp(None), _vp(rr), st))
        self._allreduce(rr)
        hist[0] = rr[0]
        for it in range(iters):
            self.matvec(p, Ap)
            ctx.check(L.femx_dot2(ctx.h, dt, _i64(n), _vp(p), _vp(Ap), _vp(None), _vp(None), _vp(pap), st))
            self._allreduce(pap)
            # x += (rr/pAp) p ; r -= (rr/pAp) Ap
            ctx.check(L.femx_axpy_ratio(ctx.h, dt, _i64(n), _vp(rr), _vp(pap), _dbl(1.0), _vp(p), _vp(x), st))
            ctx.check(L.femx_axpy_ratio(ctx.h, dt, _i64(n), _vp(rr), _vp(pap), _dbl(-1.0), _vp(Ap), _vp(r), st))
            ctx.check(L.femx_dot2(ctx.h, dt, _i64(n), _vp(r), _vp(r), _vp(None), _vp(None), _vp(rr_new), st))
            self._allreduce(rr_new)
            hist[it + 1] = rr_new[0]
            # p = r + (rr_new/rr) p
            ctx.check(L.femx_xpby_ratio(ctx.h, dt, _i64(n), _vp(rr_new), _vp(rr), _vp(r), _vp(p), st))
            rr, rr_new = rr_new, rr
        return x, hist.sqrt()
