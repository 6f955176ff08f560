#!/usr/bin/env python
"""Every BASELINE config at full size on one B200: symbolic pass, numeric CSR pass, COO pass,
with the size-independent parity properties checked on the results (closed-form nnz, zero row
sums / volume, symmetry via x'Ay = y'Ax, COO vs CSR agreement, bitwise determinism).
One JSON line per config → profiles/r01_measure_all.jsonl.

  python tools/measure_all.py [cfg1 cfg2 cfg3 cfg4] [--coo]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

PEAK = 6544.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])

CFG = {
    "cfg1": dict(dim=2, n=(64, 64), form="POISSON", nd=1),
    "cfg2": dict(dim=2, n=(4096, 4096), form="POISSON", nd=1),
    "cfg3": dict(dim=3, n=(256, 256, 256), form="POISSON_MASS", nd=1),
    "cfg4": dict(dim=3, n=(192, 192, 192), form="ELASTICITY", nd=3),
}


def timeit(fn, reps, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    return sum(t) / len(t), t[0]


def main():
    import torch
    import femx
    names = [a for a in sys.argv[1:] if a in CFG] or list(CFG)
    do_coo = "--coo" in sys.argv
    fp32 = "--fp32" in sys.argv      # the reference computes in float; fp64 is the headline
    ctx = femx.Context(0)
    for name in names:
        c = CFG[name]
        dim, nd = c["dim"], c["nd"]
        if dim == 2:
            mesh = ctx.rectangle_mesh(0, 1, 0, 1, c["n"][0], c["n"][1], dtype=femx.F32 if fp32 else femx.F64)
            n = c["n"][0]
            nnz_closed = (c["n"][0] + 1) * (c["n"][1] + 1) + 2 * (c["n"][0] * (c["n"][1] + 1) + c["n"][1] * (c["n"][0] + 1) + c["n"][0] * c["n"][1])
        else:
            n = c["n"][0]
            mesh = ctx.box_mesh(n, n, n, dtype=femx.F32 if fp32 else femx.F64)
            nnz_closed = (n + 1) ** 3 + 2 * (3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pat = femx.Pattern(ctx, mesh, nd=nd)
        torch.cuda.synchronize()
        pat_ms = 1e3 * (time.perf_counter() - t0)
        pat.close()
        t0 = time.perf_counter()
        pat = femx.Pattern(ctx, mesh, nd=nd)
        torch.cuda.synchronize()
        pat_ms = min(pat_ms, 1e3 * (time.perf_counter() - t0))
        form = femx.Form(ctx, dim, getattr(femx, c["form"]), nd=nd, params=(0.5769, 0.3846) if nd > 1 else (1.0,),
                         dtype=femx.F32 if fp32 else femx.F64)
        tdt = torch.float32 if fp32 else torch.float64
        rs = 4 if fp32 else 8
        vals = torch.empty(pat.nnz, dtype=tdt, device="cuda")
        ms, ms_min = timeit(lambda: form.assemble_csr(pat, mesh, vals), 10 if name != "cfg1" else 50)
        ne, nn_ = mesh.n_elems, mesh.nn
        b_alg = ne * nn_ * 4 + mesh.n_nodes * dim * rs + pat.nnz * rs
        out = {"config": name + (" fp32" if fp32 else ""), "elements": ne, "nodes": mesh.n_nodes, "nnz": pat.nnz, "nnz_closed_form_ok": pat.nnz == nnz_closed * nd * nd,
               "pattern_build_ms": pat_ms, "pattern_nnz_per_s": pat.nnz / (pat_ms * 1e-3), "pattern_bytes": pat.bytes,
               "csr_ms": ms, "csr_ms_min": ms_min, "elements_per_s": ne / (ms * 1e-3), "nnz_per_s": pat.nnz / (ms * 1e-3),
               "algorithmic_bytes": b_alg, "bytes_per_element": b_alg / ne, "achieved_GBs": b_alg / (ms * 1e-3) / 1e9,
               "roofline_frac": b_alg / (ms * 1e-3) / 1e9 / PEAK, "peak_GBs": PEAK,
               "stencil_rows": pat.stencil()["rows"], "stencil_pass": os.environ.get("FEMX_SPEC", "1") != "0"}
        # ---- size-independent parity properties
        v2 = torch.empty_like(vals)
        form.assemble_csr(pat, mesh, v2)
        out["bitwise_deterministic"] = bool(torch.equal(vals, v2))
        del v2
        nr = pat.n_rows
        g = torch.Generator(device="cuda"); g.manual_seed(12345)
        x = torch.rand(nr, dtype=tdt, device="cuda", generator=g)
        y = torch.rand(nr, dtype=tdt, device="cuda", generator=g)
        Ax, Ay = pat.spmv(vals, x), pat.spmv(vals, y)
        a, b = torch.dot(y, Ax).item(), torch.dot(x, Ay).item()
        out["symmetry_rel"] = abs(a - b) / abs(a)
        ones = torch.ones(nr, dtype=tdt, device="cuda")
        A1 = pat.spmv(vals, ones)
        if c["form"] == "POISSON":
            out["max_abs_row_sum"] = A1.abs().max().item()            # grad.grad: constants in the null space
        elif c["form"] == "POISSON_MASS":
            out["sum_A1_minus_volume"] = A1.sum().item() - 1.0       # 1'A1 = |Omega|
        else:
            t = torch.zeros(nr, dtype=tdt, device="cuda"); t[0::3] = 1.0
            out["max_abs_A_translation"] = pat.spmv(vals, t).abs().max().item()   # rigid translation
        del Ax, Ay, A1
        if do_coo and nd == 1:
            n2 = ne * nn_ * nn_
            A = torch.empty(n2, dtype=tdt, device="cuda")
            r = torch.empty(n2, dtype=torch.int32, device="cuda")
            cc = torch.empty(n2, dtype=torch.int32, device="cuda")
            cms, cms_min = timeit(lambda: form.assemble_coo(mesh, A, r, cc), 5)
            b_coo = ne * nn_ * 4 + mesh.n_nodes * dim * rs + n2 * (8 + rs)
            out.update({"coo_ms": cms, "coo_elements_per_s": ne / (cms * 1e-3), "coo_algorithmic_bytes": b_coo,
                        "coo_achieved_GBs": b_coo / (cms * 1e-3) / 1e9, "coo_roofline_frac": b_coo / (cms * 1e-3) / 1e9 / PEAK})
            # COO and CSR describe the same operator: scatter the triplets against x and compare with A x
            yy = torch.zeros(nr, dtype=tdt, device="cuda")
            yy.index_add_(0, r.long(), A * x[cc.long()])
            ref = pat.spmv(vals, x)
            out["coo_vs_csr_rel"] = ((yy - ref).norm() / ref.norm()).item()
            del A, r, cc, yy, ref
        print(json.dumps(out), flush=True)
        form.close(); pat.close()
        del vals, mesh
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
