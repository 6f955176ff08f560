"""Small lattice-pass run for compute-sanitizer: python tools/debug_lattice.py [FORM] [nx ny nz]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import torch
import femx
form_name = sys.argv[1] if len(sys.argv) > 1 else "POISSON_MASS"
dims = [int(a) for a in sys.argv[2:5]] if len(sys.argv) > 4 else [9, 7, 8]
ctx = femx.Context(0)
mesh = ctx.box_mesh(*dims)
pat = femx.Pattern(ctx, mesh)
print("lattice", pat.lattice(), "stencil rows", pat.stencil()["rows"], flush=True)
form = femx.Form(ctx, 3, getattr(femx, form_name), params=(1.0,))
for it in range(2):
    v = form.assemble_csr(pat, mesh)
    torch.cuda.synchronize()
    print("launch", it, "ok, sum", float(v.sum()), flush=True)
ctx.set_option("lattice", 0)
v0 = form.assemble_csr(pat, mesh)
print("rel diff vs stencil-class pass", float(torch.linalg.norm(v - v0) / torch.linalg.norm(v0)))
