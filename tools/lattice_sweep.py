"""Times the numeric pass of a 3-D workload for a list of lattice-pass configurations (one process, one
pattern).  python tools/lattice_sweep.py [n=256] [steps=20]  — prints one line per configuration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import femx  # noqa: E402

CONFIGS = [
    dict(lattice=0),
    dict(),
    dict(lt_side=0),
    dict(lt_unroll=2),
    dict(lt_unroll=2, lt_side=0),
    dict(lt_kc=64),
    dict(lt_kc=32),
    dict(lt_kc=128),
    dict(lt_unroll=2, rcp3=1),
    dict(lt_tx=8, lt_ty=16),
    dict(lt_tx=8, lt_ty=16, lt_unroll=2),
    dict(lt_tx=32, lt_ty=4),
    dict(lt_tx=16, lt_ty=8, lt_regs=152),
    dict(lt_tx=16, lt_ty=8, lt_minb=4, lt_unroll=2),
    dict(lt_tx=16, lt_ty=6, lt_minb=4),
    dict(lt_tx=16, lt_ty=12, lt_minb=2)]
DEFAULTS = dict(lattice=1, lt_tx=0, lt_ty=0, lt_minb=0, lt_kc=0, lt_pf=0, lt_regs=0, carveout=-1, rcp3=0, lt_unroll=2, lt_side=1)


def main():
    global CONFIGS
    if os.environ.get("LATTICE_CFGS"):      # e.g. LATTICE_CFGS='[{"lt_tx":8,"lt_ty":16,"lt_minb":3}]' (ncu runs: one configuration)
        CONFIGS = json.loads(os.environ["LATTICE_CFGS"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ctx = femx.Context(0)
    if os.environ.get("LATTICE_SLAB"):     # "rank/world": time one rank's z-slab of the n^3 cube (strong-scaling tuning)
        rank, world = (int(v) for v in os.environ["LATTICE_SLAB"].split("/"))
        r0, r1, lo, hi = femx.dist_slab(n + 1, world, rank)
        plane = (n + 1) ** 2
        mesh = ctx.box_mesh(n, n, n, k_lo=lo, k_hi=hi)
        pat = femx.Pattern(ctx, mesh, row_begin=(r0 - lo) * plane, row_end=(r1 - lo) * plane, col_base=lo * plane)
    else:
        mesh = ctx.box_mesh(n, n, n)
        pat = femx.Pattern(ctx, mesh)
    b_alg = mesh.n_elems * 16 + mesh.n_nodes * 24 + pat.nnz * 8
    print(json.dumps(dict(n=n, elems=mesh.n_elems, nnz=pat.nnz, lattice=pat.lattice() is not None)), flush=True)
    ref = None
    for cfg in CONFIGS:
        opts = dict(DEFAULTS)
        opts.update(cfg)
        for k, v in opts.items():
            ctx.set_option(k, v)
        form = femx.Form(ctx, 3, femx.POISSON_MASS)
        vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
        try:
            for _ in range(3):
                form.assemble_csr(pat, mesh, vals)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                form.assemble_csr(pat, mesh, vals)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if ref is None:
                ref = vals.clone()
                err = 0.0
            else:
                err = float(torch.linalg.norm(vals - ref) / torch.linalg.norm(ref))
            print(json.dumps(dict(cfg=cfg, ms=round(ms, 4), frac=round(b_alg / (ms * 1e-3) / 6.544e12, 4), relerr_vs_first=err)), flush=True)
        except Exception as e:  # a configuration that does not fit must not stop the sweep
            print(json.dumps(dict(cfg=cfg, error=str(e)[:200])), flush=True)
        form.close()
    pat.close(); ctx.close()


if __name__ == "__main__":
    main()
