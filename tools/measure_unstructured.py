#!/usr/bin/env python
"""Numeric pass on an UNSTRUCTURED mesh: random Delaunay triangulation (scipy), nodes renumbered
along a Hilbert-like sort (x-major bins) as a mesh generator / RCM would, vs the same mesh with
random numbering.  Shows what the structured benchmark does not: irregular valence, SELL padding,
gather locality.  One JSON line per case."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import femx
from scipy.spatial import Delaunay

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.RandomState(12345)
P = rng.uniform(0, 1, (npts, 2))
t0 = time.time()
tri = Delaunay(P)
conn = tri.simplices.astype(np.int32)
V = P[conn]
det = (V[:, 0, 0] - V[:, 2, 0]) * (V[:, 1, 1] - V[:, 2, 1]) - (V[:, 0, 1] - V[:, 2, 1]) * (V[:, 1, 0] - V[:, 2, 0])
neg = det < 0
conn[neg, 0], conn[neg, 1] = conn[neg, 1].copy(), conn[neg, 0].copy()
gen_s = time.time() - t0
ctx = femx.Context(0)
form = femx.Form(ctx, 2, femx.POISSON)
PEAK = 6544.0
for name in ("spatially sorted numbering", "random numbering"):
    if name.startswith("spatial"):
        nb = int(np.sqrt(npts) / 4)
        key = (np.floor(P[:, 1] * nb).astype(np.int64) * (1 << 32)) + (P[:, 0] * (1 << 31)).astype(np.int64)
        order = np.argsort(key)
    else:
        order = rng.permutation(npts)
    new_id = np.empty(npts, np.int64); new_id[order] = np.arange(npts)
    c2 = new_id[conn].astype(np.int32)
    c2 = c2[np.argsort(c2.min(1), kind="stable")]            # elements in node order, as generators emit them
    X = torch.from_numpy(np.ascontiguousarray(P[order, 0])).cuda(); Y = torch.from_numpy(np.ascontiguousarray(P[order, 1])).cuda()
    mesh = femx.Mesh(2, torch.from_numpy(c2).cuda(), (X, Y))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pat = femx.Pattern(ctx, mesh)
    torch.cuda.synchronize(); pat_ms = 1e3 * (time.perf_counter() - t0)
    vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
    for _ in range(3):
        form.assemble_csr(pat, mesh, vals)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for k in range(10):
        form.assemble_csr(pat, mesh, vals); ev[k + 1].record()
    torch.cuda.synchronize()
    ms = sum(ev[k].elapsed_time(ev[k + 1]) for k in range(10)) / 10
    b_alg = mesh.n_elems * 12 + npts * 16 + pat.nnz * 8
    ones = torch.ones(npts, dtype=torch.float64, device="cuda")
    print(json.dumps({"mesh": f"2-D Delaunay, {npts} random points, {mesh.n_elems} triangles, {name}", "nnz": pat.nnz,
                      "max_row": pat.max_row, "pattern_ms": pat_ms, "csr_ms": ms, "elements_per_s": mesh.n_elems / (ms * 1e-3),
                      "roofline_frac": b_alg / (ms * 1e-3) / 1e9 / PEAK,
                      "max_abs_row_sum": pat.spmv(vals, ones).abs().max().item()}), flush=True)
    pat.close()
