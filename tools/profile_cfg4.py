#!/usr/bin/env python
"""One warm launch of femx_csr on cfg4 (3-D elasticity, 192^3) for `ncu --set full`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import torch, femx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
ctx = femx.Context(0)
mesh = ctx.box_mesh(n, n, n)
pat = femx.Pattern(ctx, mesh, nd=3)
form = femx.Form(ctx, 3, femx.ELASTICITY, nd=3, params=(0.5769, 0.3846))
vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
for _ in range(3):
    form.assemble_csr(pat, mesh, vals)
torch.cuda.synchronize()
print("ok", pat.nnz)
