#!/usr/bin/env python
"""N-rank check of the multi-GPU layer (run under torchrun, e.g. gpurun --gpus 2):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [n=48]
Every rank assembles its z-slab of an n^3 Kuhn cube (jittered) and ALSO the whole mesh; checks
  * owned rows of the slab == the same rows of the whole-mesh matrix, bit for bit (assembly needs no communication),
  * femx_dist_spmv == the whole-mesh SpMV restricted to the owned rows (halo exchange correct),
  * the N-rank CG residual history == the 1-rank history to 1e-10 (relative, above rounding level)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import femx
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = femx.Context(lr)
    plane = (n + 1) ** 2
    whole = ctx.box_mesh(n, n, n)
    g = torch.Generator(device="cpu").manual_seed(12345)
    jit = [(torch.rand(whole.n_nodes, generator=g, dtype=torch.float64) - 0.5) * (0.3 / n) for _ in range(3)]
    whole.node_xyz = tuple(c + j.cuda() for c, j in zip(whole.node_xyz, jit))
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    pw = femx.Pattern(ctx, whole)
    vw = form.assemble_csr(pw, whole)
    rpw, _ = pw.csr("int64")
    r0, r1, lo, hi = femx.dist_slab(n + 1, world, rank)
    sel = slice(lo * plane, (hi + 1) * plane)
    e0, e1 = 6 * n * n * lo, 6 * n * n * hi
    slab = femx.Mesh(3, (whole.conn[e0:e1] - lo * plane).contiguous(), tuple(c[sel].contiguous() for c in whole.node_xyz))
    ps = femx.Pattern(ctx, slab, row_begin=(r0 - lo) * plane, row_end=(r1 - lo) * plane, col_base=lo * plane)
    vs = form.assemble_csr(ps, slab)
    a, b = int(rpw[r0 * plane]), int(rpw[r1 * plane])
    ok_rows = bool(torch.equal(vs, vw[a:b]))
    dd = femx.Dist.from_torch(ctx)
    op = dd.operator(ps, vs)
    x = torch.from_numpy(np.random.RandomState(7).uniform(-1, 1, whole.n_nodes)).cuda()
    yw = pw.spmv(vw, x)
    ys = op.spmv(x[r0 * plane:r1 * plane].contiguous())
    ok_spmv = bool(torch.equal(ys, yw[r0 * plane:r1 * plane]))
    bw = pw.spmv(vw, torch.ones_like(x))
    d1 = femx.Dist(ctx)
    o1 = d1.operator(pw, vw)
    _, res1, _ = o1.cg(bw, 100)
    xs, resn, ms = op.cg(bw[r0 * plane:r1 * plane].contiguous(), 100)
    k = np.flatnonzero(res1 > 1e-10 * res1[0])
    rel = float(np.max(np.abs(resn[k] - res1[k]) / res1[k]))
    ok = torch.tensor([float(ok_rows), float(ok_spmv), float(rel <= 1e-10)], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "n": n, "slab_rows_bitwise": bool(ok[0]), "dist_spmv_bitwise": bool(ok[1]),
                          "cg_history_rel_diff_vs_1_rank": rel, "cg_history_ok": bool(ok[2]),
                          "residual@0": float(resn[0]), "residual@100": float(resn[100]), "cg_ms": ms,
                          "lattice": ps.lattice() is not None, "p2p_reduction": dd.p2p_reduction, "peer_halo": op.peer_halo}))
    op.close(); dd.close(); o1.close(); d1.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if bool(ok.min() > 0) else 1)


if __name__ == "__main__":
    main()
