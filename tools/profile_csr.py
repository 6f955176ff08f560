#!/usr/bin/env python
"""Three launches of femx_csr on one BASELINE config (cfg2 | cfg3 | cfg4) — the command profiled with
  ncu --set full --clock-control none --import-source on -k regex:femx_csr --launch-skip 2 --launch-count 1
(FEMX_JIT_DUMP=<dir> keeps the generated source for the source page)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import torch

import femx
from measure_all import CFG

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
c = CFG[name]
ctx = femx.Context(0)
n = c["n"]
mesh = ctx.rectangle_mesh(0, 1, 0, 1, n[0], n[1]) if c["dim"] == 2 else ctx.box_mesh(*n)
pat = femx.Pattern(ctx, mesh, nd=c["nd"])
form = femx.Form(ctx, c["dim"], getattr(femx, c["form"]), nd=c["nd"], params=(0.5769, 0.3846) if c["nd"] > 1 else (1.0,))
vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
for _ in range(3):
    form.assemble_csr(pat, mesh, vals)
torch.cuda.synchronize()
print("ok", name, pat.nnz, pat.stencil()["rows"])
