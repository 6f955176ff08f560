#!/usr/bin/env python
"""One workload, a few numeric passes (ncu target): python tools/measure_one.py cfg3|cfg2|cfg4 [steps] [coo|cg]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import femx  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
extra = sys.argv[3] if len(sys.argv) > 3 else ""
wl = bench.WORKLOADS[name]
ctx = femx.Context(0)
mesh, slab = bench.build_problem(ctx, femx, wl, 0, 1)
pat = femx.Pattern(ctx, mesh, nd=wl["nd"])
form = femx.Form(ctx, wl["dim"], getattr(femx, wl["form"]), nd=wl["nd"], params=wl["params"])
vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
for _ in range(steps):
    form.assemble_csr(pat, mesh, vals)
torch.cuda.synchronize()
if extra == "coo":
    for _ in range(steps):
        A, r, c = form.assemble_coo(mesh)
    torch.cuda.synchronize()
if extra == "cg":
    d = femx.Dist(ctx)
    op = d.operator(pat, vals)
    b = op.spmv(torch.ones(op.n_owned, dtype=torch.float64, device="cuda"))
    op.cg(b, 8)
    op.close(); d.close()
print("ok", name, pat.nnz)
