#!/usr/bin/env python
"""BASELINE config 5: assemble the 3-D operator (grad.grad + u v on a Kuhn cube) on N GPUs as
owned-row slabs, then validate it with SpMV + CG (halo exchange + all-reduce over NCCL).

  python tools/cg_validate.py --size 256 --cg-iters 100                       (1 GPU)
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/cg_validate.py --size 256

Prints one JSON line: assembly time, SpMV time, CG time/iteration, residual history checkpoints,
and the checks  A 1 = M 1 (row sums = lumped mass → sum = volume)  and  ||x - 1||_inf after CG on b = A 1.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=256)
    ap.add_argument("--cg-iters", dest="iters", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import femx
    from femx.dist import SlabOperator, make_slab
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = femx.Context(lr)
    n = args.n
    slab = make_slab(rank, world, n, (n + 1) ** 2)
    mesh = ctx.box_mesh(n, n, n, k_lo=slab.cells_lo, k_hi=slab.cells_hi)
    pat = femx.Pattern(ctx, mesh, row_begin=slab.row_begin, row_end=slab.row_end, col_base=slab.col_base)
    form = femx.Form(ctx, 3, femx.POISSON_MASS)
    vals = form.assemble_csr(pat, mesh)

    def timed(fn, reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    asm_ms = timed(lambda: form.assemble_csr(pat, mesh, vals), 10)
    op = SlabOperator(ctx, pat, vals, slab)
    ones = torch.ones(slab.n_owned, dtype=torch.float64, device="cuda")
    y = torch.empty_like(ones)
    op.matvec(ones, y)                       # warm-up: first NCCL point-to-point sets the channels up
    spmv_ms = timed(lambda: op.matvec(ones, y), 10)
    vol = y.sum().reshape(1).clone()
    if world > 1:
        dist.all_reduce(vol)
    b = y.clone()
    x, hist = op.cg(b, 3)  # warm-up
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    x, hist = op.cg(b, args.iters)
    torch.cuda.synchronize()
    cg_s = time.perf_counter() - t0
    err = (x - 1.0).abs().max().reshape(1)
    if world > 1:
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    h = hist.cpu().tolist()
    if rank == 0:
        print(json.dumps({
            "config": f"3-D P1 tets {n}^3 Kuhn cube, grad.grad + u v, fp64, {world} GPU(s), owned-row slabs",
            "elements": 6 * n ** 3, "rows": (n + 1) ** 3, "assemble_ms": asm_ms,
            "elements_per_s": 6 * n ** 3 / (asm_ms * 1e-3), "spmv_ms": spmv_ms,
            "cg_iters": args.iters, "cg_ms_per_iter": 1e3 * cg_s / args.iters,
            "sum_A1": float(vol.item()), "volume": 1.0,
            "residual": {"r0": h[0], "r10": h[min(10, len(h) - 1)], "r50": h[min(50, len(h) - 1)], "r_last": h[-1]},
            "x_minus_1_inf": float(err.item()),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
