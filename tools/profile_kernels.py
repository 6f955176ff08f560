#!/usr/bin/env python
"""Runs each kernel of the path once on cfg2 (symbolic pass, COO, CSR, RHS) — the command that is
profiled with `ncu --set full` to commit per-kernel counters (profiles/r01_kernels_cfg2.txt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

import femx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = femx.Context(0)
mesh = ctx.rectangle_mesh(0, 1, 0, 1, n, n)
form = femx.Form(ctx, 2, femx.POISSON)
for rep in range(2):          # second repetition = warm
    pat = femx.Pattern(ctx, mesh)
    vals = form.assemble_csr(pat, mesh)
    A, r, c = form.assemble_coo(mesh)
    b = form.assemble_rhs(pat, mesh)
    torch.cuda.synchronize()
    if rep == 0:
        pat.close()
        del vals, A, r, c, b
print("ok", pat.nnz)
