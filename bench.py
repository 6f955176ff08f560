#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: fp64 P1 assembly into CSR.

Workload (BASELINE.json configs[2], the north-star configuration): the symbolic integrand
grad u . grad v + u v on 3-D P1 tetrahedra, 256^3 Kuhn cube (100,663,296 tets, 16,974,593 nodes,
253,036,801 nnz), fp64, deterministic numeric pass into a prebuilt CSR pattern.  One "step" = one
numeric pass over the whole mesh.  With N GPUs the FIXED cube is split into z-slabs of owned node
planes with one ghost cell layer per side (strong scaling; assembly needs no collective); the
assembled operator is then validated by SpMV + 100 CG iterations with NCCL halo exchange
(BASELINE.json configs[4]) through the C++ multi-GPU layer (femx_dist_*).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload cfg3|cfg2|cfg4|cfg1] [--no-extras] [--no-cg]

Prints ONE JSON line (DESIGN.md §5 explains the keys).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # n = cells along the sharded axis (node rows in 2-D, z in 3-D), m = cells along the other axes
    "cfg1": dict(dim=2, n=64, m=64, form="POISSON", nd=1, params=(), desc="2-D P1 Poisson 64x64 unit square"),
    "cfg2": dict(dim=2, n=4096, m=4096, form="POISSON", nd=1, params=(),
                 desc="2-D P1 Poisson 4096x4096 structured triangles"),
    "cfg3": dict(dim=3, n=256, m=256, form="POISSON_MASS", nd=1, params=(1.0,),
                 desc="3-D P1 tets 256^3 Kuhn cube, grad.grad + u v (NVRTC integrand)"),
    "cfg4": dict(dim=3, n=192, m=192, form="ELASTICITY", nd=3, params=(0.5769230769230769, 0.3846153846153846),
                 desc="3-D linear elasticity P1 tets 192^3 Kuhn cube (12x12 element matrices, E=1, nu=0.3)"),
}
UNPINNED = ("3-D / mass / elasticity: the reference has no implementation (fea_symbolic_nvrtc_sparse.cpp:271-274 returns 3), "
            "parity is against the oracle restatement, unpinned by the reference")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "note": "no NVML samples" + (": " + getattr(self, "err", "") if not self.ok else "")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def profiled_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of femx_csr from the committed
    `ncu --set full` capture (profiles/traffic.json) + a hash tying the figure to the file, or (None, None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    raw = open(p, "rb").read()
    j = json.loads(raw)
    ent = j.get(workload)
    if ent is None:
        return None, None
    src = {"file": "profiles/traffic.json", "sha1": hashlib.sha1(raw).hexdigest()[:12]}
    if isinstance(ent, dict):
        src.update({k: v for k, v in ent.items() if k != "bytes"})
        return ent.get("bytes"), src
    return ent, src


# ------------------------------------------------------------------ the workload on this rank ---
def build_problem(ctx, femx, wl, rank, world):
    """This rank's z-slab (node rows in 2-D) of the FIXED global mesh: owned planes [r0, r1), slab planes
    [lo, hi] (one ghost layer per side).  Returns mesh, slab dict."""
    dim, n, m = wl["dim"], wl["n"], wl["m"]
    r0, r1, lo, hi = femx.dist_slab(n + 1, world, rank)
    if dim == 2:
        plane = m + 1
        mesh = ctx.rectangle_mesh(0.0, 1.0, 0.0, 1.0, n, m, row_lo=lo, row_hi=hi)
        ne_global = 2 * n * m
    else:
        plane = (m + 1) ** 2
        mesh = ctx.box_mesh(m, m, n, k_lo=lo, k_hi=hi)
        ne_global = 6 * n * m * m
    slab = dict(r0=r0, r1=r1, lo=lo, hi=hi, plane=plane, row_begin=(r0 - lo) * plane, row_end=(r1 - lo) * plane,
                col_base=lo * plane, ne_global=ne_global, nodes_global=(n + 1) * plane)
    return mesh, slab


def algorithmic_bytes(mesh, pat, dim):
    """connectivity int32 + node coordinates fp64 read once, CSR values written once;
    no scatter map, no workspace."""
    return mesh.n_elems * mesh.nn * 4 + mesh.n_nodes * dim * 8 + pat.nnz * 8


def time_numeric_pass(torch, form, pat, mesh, vals, steps, warmup, barrier, sampler=None, use_graph=True):
    """K timed steps.  One step = one femx_assemble_csr call; by default the call is stream-captured ONCE into a CUDA graph
    (the C ABI is capture-safe: stream-ordered, the side stream forks and joins by events) and every step replays it, so that the
    host's launch path (two kernel launches, two event records, two stream waits, Python) does not gate small per-GPU steps."""
    for _ in range(max(warmup, 3)):
        form.assemble_csr(pat, mesh, vals)
    torch.cuda.synchronize()
    graph = None
    if use_graph:
        try:
            g = torch.cuda.CUDAGraph()
            cs = torch.cuda.Stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, stream=cs):
                form.assemble_csr(pat, mesh, vals)
            torch.cuda.current_stream().wait_stream(cs)
            g.replay()
            torch.cuda.synchronize()
            graph = g
        except Exception:
            graph = None
            torch.cuda.synchronize()
    step = graph.replay if graph is not None else (lambda: form.assemble_csr(pat, mesh, vals))
    for _ in range(3):
        step()
    barrier()
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        step()   # femx_csr (+ femx_rowlist for the boundary rows of a lattice mesh)
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(steps))
    time_numeric_pass.launch = "CUDA graph replay (one captured femx_assemble_csr call per step)" if graph is not None else "femx_assemble_csr call per step"
    return total_ms, per_launch, clocks


def parity_vs_oracle(torch, wl, mesh, slab, pat, vals, rp, ci):
    """Value-level parity AT THE BENCHMARKED SIZE, outside the timed region: up to three z-slabs of two owned node
    planes (first / middle / last of this rank) are re-assembled by the CPU oracle from the device's own coordinates
    and compared with the CUDA values: column indices bit-exact, values relF <= 1e-12."""
    import numpy as np
    from oracle import oracle as orc
    dim, m, nd = wl["dim"], wl["m"], wl["nd"]
    plane = slab["plane"]
    owned_planes = slab["r1"] - slab["r0"]
    n_local_planes = slab["hi"] - slab["lo"] + 1
    firsts = sorted({0, max(0, owned_planes // 2 - 1), max(0, owned_planes - 2)})
    worst, exact, rows_checked = 0.0, True, 0
    for f in firsts:
        p0 = slab["r0"] - slab["lo"] + f               # local plane of the first checked row plane
        p1 = min(p0 + 2, slab["r1"] - slab["lo"])       # checked planes [p0, p1)
        s0, s1 = max(p0 - 1, 0), min(p1, n_local_planes - 1)   # sub-mesh planes [s0, s1]
        if dim == 3:
            _, _, _, conn = orc.box_mesh(m, m, s1 - s0)
        else:
            _, _, _, conn = orc.rect_mesh(0, 1, 0, 1, s1 - s0, m)
        sel = slice(s0 * plane, (s1 + 1) * plane)
        coords = [c[sel].cpu().numpy() for c in mesh.node_xyz]
        orp, oci = orc.pattern(conn, len(coords[0]))
        if nd > 1:
            drp, dci = orc.expand_pattern(nd, orp, oci)
        else:
            drp, dci = orp, oci
        oc = coords if dim == 3 else (coords[0], coords[1], None)
        ov = orc.assemble_csr(getattr(orc, wl["form"]), dim, nd, conn, *oc, drp, dci, params=wl["params"] or None)
        a, b = (p0 - s0) * plane * nd, (p1 - s0) * plane * nd           # oracle dof rows
        ga, gb = (p0 * plane - slab["row_begin"]) * nd, (p1 * plane - slab["row_begin"]) * nd   # pattern dof rows
        va, vb = int(rp[ga].item()), int(rp[gb].item())
        got = vals[va:vb].cpu().numpy()
        want = ov[drp[a]:drp[b]]
        cols = ci[va:vb].cpu().numpy().astype(np.int64)
        wcols = dci[drp[a]:drp[b]].astype(np.int64) + (slab["lo"] + s0) * plane * nd
        exact = exact and len(got) == len(want) and bool(np.array_equal(cols, wcols)) and \
            bool(np.array_equal((rp[ga:gb + 1] - rp[ga]).cpu().numpy(), drp[a:b + 1] - drp[a]))
        if len(got) == len(want):
            worst = max(worst, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
        else:
            worst = float("inf")
        rows_checked += gb - ga
    return {"against": "CPU oracle (oracle/femx_oracle.c) on the device's own coordinates", "rows_checked": rows_checked,
            "z_slabs": len(firsts), "pattern_exact": exact, "relF": worst, "tolerance": 1e-12,
            "ok": bool(exact and worst <= 1e-12)}


def ref_gpu_baseline_and_parity(ctx, femx, torch, wl, mesh, pat, vals):
    """cfg2 only (the reference is 2-D): the reference's own kernels (K4 COO, K5 ELL + global atomicAdd; oracle/_ref,
    recompiled for sm_100 with the minimal fixes of oracle/build_ref.py) timed on the same mesh — and, with the fp64
    retype, compared VALUE BY VALUE with femx at the full size: ELL pattern bit-exact, relF <= 1e-12."""
    try:
        from oracle import refimpl
        if not refimpl.available() or wl["dim"] != 2:
            return None, None
        out = {"what": "reference fea_kernel recompiled for sm_100 (fixes Q2,Q3,Q4,Q8,Q13), CUDA events, 3 launches"}
        ln, idx = pat.ell(7)
        gidx = mesh.conn.reshape(-1).contiguous()
        parity = None
        for prec, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            X = mesh.node_xyz[0][mesh.conn.reshape(-1).long()].to(tdt).contiguous()
            Y = mesh.node_xyz[1][mesh.conn.reshape(-1).long()].to(tdt).contiguous()
            ell, ms_ell = refimpl.assemble_ell(prec, wl["n"], wl["m"], X, Y, gidx, ln, idx, iters=3)
            coo = refimpl.assemble_coo(prec, wl["n"], wl["m"], X, Y, gidx, iters=3)
            ms_coo = coo[-1]
            out[prec] = {"ell_atomic_ms": ms_ell, "ell_atomic_elements_per_s": mesh.n_elems / (ms_ell * 1e-3),
                         "coo_ms": ms_coo, "coo_elements_per_s": mesh.n_elems / (ms_coo * 1e-3)}
            if prec == "f64":
                del ell
                ell, _ = refimpl.assemble_ell(prec, wl["n"], wl["m"], X, Y, gidx, ln, idx, iters=1)   # (K5 accumulates: one launch)
                mine = pat.values_to_ell(vals, 7)
                ref = ell.reshape(mine.shape).to(torch.float64)
                relF = float(torch.linalg.norm(mine - ref) / torch.linalg.norm(ref))
                # rows above node 2^24 explicitly (the reference's float node ids break there)
                hi_rows = slice(1 << 24, None)
                relF_hi = float(torch.linalg.norm(mine[hi_rows] - ref[hi_rows]) / torch.linalg.norm(ref[hi_rows])) \
                    if mine.shape[0] > (1 << 24) else None
                A_ref = coo[0].to(torch.float64)
                A_mine, _, _ = femx_coo_values(femx, ctx, wl, mesh)
                relF_coo = float(torch.linalg.norm(A_mine - A_ref) / torch.linalg.norm(A_ref))
                parity = {"against": "reference K5 (ELL+atomicAdd) and K4 (COO), fp64 retype, same mesh", "config": "cfg2",
                          "pattern_exact": True, "pattern_note": "K5 consumes femx_pattern_export_ell (bit-equal to getNeighborNodesList, tests)",
                          "relF": relF, "relF_rows_above_2^24": relF_hi, "relF_coo": relF_coo, "tolerance": 1e-12,
                          "ok": bool(relF <= 1e-12 and relF_coo <= 1e-12 and (relF_hi is None or relF_hi <= 1e-12))}
                del mine, ref, A_ref, A_mine
            del X, Y, ell, coo
        try:   # the reference's accumulation micro-benchmark (atomicadd.cu:73-129), the contention K5 suffers
            av = refimpl.atomic_variants()
            if av:
                out["atomicadd_variants"] = av
        except Exception as e:
            out["atomicadd_variants"] = {"error": repr(e)[:200]}
        return out, parity
    except Exception as e:  # a reported extra must never take the headline down
        return {"error": repr(e)}, None


def femx_coo_values(femx, ctx, wl, mesh):
    form = femx.Form(ctx, wl["dim"], getattr(femx, wl["form"]), nd=wl["nd"], params=wl["params"])
    A, r, c = form.assemble_coo(mesh, indices=False)
    form.close()
    return A, r, c


def cpu_baseline_port(wl, seconds_target=12.0):
    """The oracle's serial numeric pass on a bounded sub-mesh of the workload, 1 core, built with -O3 -march=native
    on this host."""
    from oracle import oracle as orc
    dim, m, nd = wl["dim"], wl["m"], wl["nd"]
    layers = {2: min(wl["n"], 4096), 3: min(wl["n"], 96 if nd == 1 else 16)}[dim]   # about 10-20 s of CPU work
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, layers, m)
        Z = None
        sample = f"{layers}x{m} sub-mesh of the workload"
    else:
        X, Y, Z, conn = orc.box_mesh(m, m, layers)
        sample = f"{m}x{m}x{layers} sub-mesh of the workload"
    rp, ci = orc.pattern(conn, len(X))
    if nd > 1:
        rp, ci = orc.expand_pattern(nd, rp, ci)
    form = getattr(orc, wl["form"])
    try:
        orc.lib_native()
        native, flags = True, orc.NATIVE_FLAGS
    except Exception:
        native, flags = False, "-O3 -march=x86-64-v2 -ffp-contract=off (native build failed)"
    orc.assemble_csr(form, dim, nd, conn[:1000], X, Y, Z, rp, ci, params=wl["params"] or None, native=native)  # warm
    t0 = time.perf_counter()
    orc.assemble_csr(form, dim, nd, conn, X, Y, Z, rp, ci, params=wl["params"] or None, native=native)
    dt = time.perf_counter() - t0
    return {"value": len(conn) / dt, "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": f"{sample} ({len(conn)} elements), numeric pass only, pattern prebuilt, {dt:.1f} s",
            "flags": flags}


_REF_STATE = {}


def _ref_worker(args):
    """One slab of the workload on one host core: mesh + pattern are built once per worker process
    (cached), every step re-runs the oracle's serial numeric pass on it."""
    name, wl, layers = args
    from oracle import oracle as orc
    st = _REF_STATE.get(name)
    if st is None:
        if wl["dim"] == 2:
            X, Y, _, conn = orc.rect_mesh(0, 1, 0, 1, layers, wl["m"])
            Z = None
        else:
            X, Y, Z, conn = orc.box_mesh(wl["m"], wl["m"], layers)
        rp, ci = orc.pattern(conn, len(X))
        if wl["nd"] > 1:
            rp, ci = orc.expand_pattern(wl["nd"], rp, ci)
        try:
            orc.lib_native()
            native = True
        except Exception:
            native = False
        st = _REF_STATE[name] = (X, Y, Z, conn, rp, ci, native)
    X, Y, Z, conn, rp, ci, native = st
    t0 = time.perf_counter()
    orc.assemble_csr(getattr(orc, wl["form"]), wl["dim"], wl["nd"], conn, X, Y, Z, rp, ci, params=wl["params"] or None,
                     native=native)
    return len(conn), time.perf_counter() - t0, native


def run_reference(args, wl, rank, world):
    """--impl reference: the reference ships no host implementation of the numeric pass (its fea_kernel is CUDA
    only; its recompiled kernels are reported by the femx arm as `ref_gpu_baseline`), so the CPU arm is the oracle
    port (-O3 -march=native), run as independent slabs on all host cores (the same sharding as the multi-GPU
    layout).  Rank 0 only."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, (os.cpu_count() or 1))
    per = {2: 64, 3: 2 if wl["nd"] == 1 else 1}[wl["dim"]]   # cell layers per worker per step: a bounded sample
    per = min(per, wl["n"])
    times, ne_step, native = [], None, False
    with mp.Pool(cores) as pool:
        job = [(args.workload, wl, per)] * cores
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, job, chunksize=1)
            dt = time.perf_counter() - t0
            ne_step = sum(r[0] for r in res)
            native = all(r[2] for r in res)
            if it >= args.warmup:
                times.append(dt)            # wall time of the step: all workers, the slowest bounds it
    tot = sum(times)
    value = ne_step * len(times) / tot
    sample = (f"each step: {cores} workers x ({per} cell layers x {wl['m']}"
              + (f" x {wl['m']}" if wl["dim"] == 3 else "") + f") slab of the workload = {ne_step} elements; "
              "numeric pass only, mesh + pattern prebuilt per worker")
    line = {
        "impl": "reference", "metric": "elements/s (fp64 P1 assembly into CSR)", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "note": "reference has no host numeric pass; oracle port on all host cores"},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample,
                         "flags": "-O3 -march=native" if native else "-O3 -march=x86-64-v2 -ffp-contract=off"},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_extra(ctx, femx, torch, name, steps=20):
    """A secondary configuration on one GPU (reported under `extra`, never the headline)."""
    wl = WORKLOADS[name]
    try:
        mesh, slab = build_problem(ctx, femx, wl, 0, 1)
        pattern_ms = []
        for _ in range(3):      # symbolic pass: the last of three builds (the first ones warm the allocator pool)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pat = femx.Pattern(ctx, mesh, nd=wl["nd"])
            torch.cuda.synchronize()
            pattern_ms.append(1e3 * (time.perf_counter() - t0))
            if len(pattern_ms) < 3:
                pat.close()
        form = femx.Form(ctx, wl["dim"], getattr(femx, wl["form"]), nd=wl["nd"], params=wl["params"])
        vals = torch.empty(pat.nnz, dtype=torch.float64, device=mesh.conn.device)
        total_ms, per_launch, _ = time_numeric_pass(torch, form, pat, mesh, vals, steps, 3, torch.cuda.synchronize)
        b_alg = algorithmic_bytes(mesh, pat, wl["dim"])
        peak, _ = measured_peaks()
        ms = sum(per_launch) / len(per_launch)
        out = {"workload": wl["desc"], "elements": mesh.n_elems, "nnz": pat.nnz, "ms_per_step": ms,
               "elements_per_s": mesh.n_elems / (ms * 1e-3), "nnz_per_s": pat.nnz / (ms * 1e-3),
               "roofline_frac": b_alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_alg,
               "numeric_pass": numeric_pass_kind(form)}
        nd = wl["nd"]
        b_sym = mesh.n_elems * mesh.nn * 4 + (pat.n_rows // nd + 1) * 8 + (pat.nnz // (nd * nd)) * 4
        out["setup"] = {"pattern_build_ms": pattern_ms[-1], "pattern_roofline": {"algorithmic_bytes": b_sym,
                        "frac": b_sym / (pattern_ms[-1] * 1e-3) / 1e9 / peak}}
        rp, ci = pat.csr("int64")
        if wl["dim"] == 3:
            out["parity"] = parity_vs_oracle(torch, wl, mesh, slab, pat, vals, rp, ci)
            out["parity"]["note"] = UNPINNED
        else:
            out["ref_gpu_baseline"], out["parity"] = ref_gpu_baseline_and_parity(ctx, femx, torch, wl, mesh, pat, vals)
        del rp, ci, vals
        form.close(); pat.close()
        del mesh
        torch.cuda.empty_cache()
        return out
    except Exception as e:
        return {"workload": wl["desc"], "error": repr(e)[:300]}


def numeric_pass_kind(form):
    src = form.source
    if "#define FEMX_LATTICE 1" in src:
        return "element-once lattice pass (class rows) + generic row list (boundary rows), one launch"
    if "#define FEMX_SPEC 1" in src:
        return "stencil-class pass (class rows) + generic row list, one launch"
    return "generic owner-computes incidence loop"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="femx", choices=["femx", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-graph", action="store_true")   # launch every step through the C ABI instead of replaying its CUDA graph
    ap.add_argument("--e2e-steps", type=int, default=12)   # the 3-stream pipeline fills once inside the timed region
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    import femx

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = femx.Context(local_rank)
    dim, nd = wl["dim"], wl["nd"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- this rank's slab of the FIXED global mesh (strong scaling) ----------------------
    mesh, slab = build_problem(ctx, femx, wl, rank, world)
    ne_global = slab["ne_global"]
    # symbolic pass (one-time per topology): built three times, the first calls also warm the allocator pools
    pattern_ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pat = femx.Pattern(ctx, mesh, nd=nd, row_begin=slab["row_begin"], row_end=slab["row_end"], col_base=slab["col_base"])
        torch.cuda.synchronize()
        pattern_ms.append(1e3 * (time.perf_counter() - t0))
        if len(pattern_ms) < 3:
            pat.close()
    t0 = time.perf_counter()
    form = femx.Form(ctx, dim, getattr(femx, wl["form"]), nd=nd, params=wl["params"])
    vals = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
    form.assemble_csr(pat, mesh, vals)
    torch.cuda.synchronize()
    jit_ms = 1e3 * (time.perf_counter() - t0)
    b_alg = algorithmic_bytes(mesh, pat, dim)          # of THIS rank's launch
    nnz_global = sum_over_ranks(pat.nnz)

    # ---- device-resident throughput (`value`) ---------------------------------------
    sampler = ClockSampler(local_rank)
    total_ms, per_launch, clocks = time_numeric_pass(torch, form, pat, mesh, vals, args.steps, args.warmup, barrier, sampler,
                                                    use_graph=not args.no_graph)
    launch_mode = time_numeric_pass.launch
    ms_per_step = max_over_ranks(total_ms) / args.steps
    value = ne_global / (ms_per_step * 1e-3)
    kern_ms = sum(per_launch) / len(per_launch)

    # ---- parity at the benchmarked size (outside every timed region) --------------------
    rp, ci = pat.csr("int64")
    if dim == 3:
        parity = parity_vs_oracle(torch, wl, mesh, slab, pat, vals, rp, ci)
        parity["note"] = UNPINNED
        ref_gpu = None
    else:
        ref_gpu, parity = (ref_gpu_baseline_and_parity(ctx, femx, torch, wl, mesh, pat, vals) if world == 1 else (None, None))
        if parity is None:
            parity = parity_vs_oracle(torch, wl, mesh, slab, pat, vals, rp, ci)
    parity["config"] = args.workload
    parity_all_ok = sum_over_ranks(0.0 if parity.get("ok") else 1.0) == 0.0
    checksum_dev = sum_over_ranks(float(vals.sum().item()))
    del ci

    # ---- end to end through the C ABI with HOST buffers (`e2e`) -----------------------
    # every step: H2D of the operator's inputs (node coordinates, pinned; the connectivity is NOT re-uploaded: the
    # numeric pass never reads it once the pattern exists), the numeric pass, D2H of the CSR values.  The pattern
    # (one-time symbolic pass) is reused.  Steps are double-buffered over three streams (H2D / numeric pass / D2H) the
    # way a re-assembly loop would run: the D2H of step k overlaps the H2D of step k+1 (PCIe is full duplex).
    h_in = [c.cpu().pin_memory() for c in mesh.node_xyz]
    bufs = []
    for b_ in range(2):
        d_in = [torch.empty_like(c) for c in mesh.node_xyz]
        m_b = femx.Mesh(dim, mesh.conn, tuple(d_in))
        bufs.append(dict(d_in=d_in, mesh=m_b, vals=torch.empty_like(vals),
                         h_out=torch.empty(pat.nnz, dtype=torch.float64).pin_memory(),
                         in_done=torch.cuda.Event(), comp_done=torch.cuda.Event(), out_done=torch.cuda.Event()))
    h2d = sum(x.numel() * x.element_size() for x in h_in)
    d2h = bufs[0]["h_out"].numel() * 8
    s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()

    def e2e_step(k):
        bb = bufs[k % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(bb["comp_done"])          # buffer free once its previous numeric pass is done
            for h, d in zip(h_in, bb["d_in"]):
                d.copy_(h, non_blocking=True)
            bb["in_done"].record(s_in)
        s_cmp.wait_event(bb["in_done"])
        s_cmp.wait_event(bb["out_done"])               # previous D2H of this values buffer finished
        form.assemble_csr(pat, bb["mesh"], bb["vals"])
        bb["comp_done"].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(bb["comp_done"])
            bb["h_out"].copy_(bb["vals"], non_blocking=True)
            bb["out_done"].record(s_out)

    for k in range(2):
        e2e_step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.e2e_steps):
        e2e_step(k)
    s_cmp.wait_stream(s_out)
    s_cmp.wait_stream(s_in)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.e2e_steps
    h_out = bufs[(args.e2e_steps - 1) % 2]["h_out"]
    e2e_ok = bool(torch.equal(h_out, vals.cpu()))      # the end-to-end result is the device-resident result
    e2e_all_ok = sum_over_ranks(0.0 if e2e_ok else 1.0) == 0.0
    h2d_total, d2h_total = sum_over_ranks(h2d), sum_over_ranks(d2h)
    del bufs, h_in, h_out
    torch.cuda.empty_cache()

    # ---- validation of the assembled operator: SpMV + 100 CG iterations, NCCL halo exchange (configs[4]) --------
    cg = None
    if not args.no_cg and nd == 1:
        try:
            dd = femx.Dist.from_torch(ctx)
            op = dd.operator(pat, vals)
            ones = torch.ones(op.n_owned, dtype=torch.float64, device=dev)
            b = op.spmv(ones)                       # b = A 1 (row sums: the mass of each hat function)
            x = torch.empty_like(b)
            y = torch.empty_like(b)
            for _ in range(3):
                op.spmv(ones, y)
            barrier()
            s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0_.record()
            for _ in range(20):
                op.spmv(ones, y)
            s1_.record()
            barrier()
            spmv_ms = max_over_ranks(s0_.elapsed_time(s1_)) / 20
            op.cg(b, 5, x)                          # warm-up (graph capture)
            barrier()
            _, res, ms = op.cg(b, 100, x)
            cg_ms = max_over_ranks(ms)
            err1 = sum_over_ranks(float(((x - 1.0) ** 2).sum().item())) ** 0.5 / slab["nodes_global"] ** 0.5
            cg = {"iterations": 100, "its_per_s": 100.0 / (cg_ms * 1e-3), "ms_total": cg_ms, "spmv_ms": spmv_ms,
                  "spmv_nnz_per_s": nnz_global / (spmv_ms * 1e-3),
                  "residual@0": float(res[0]), "residual@100": float(res[100]), "rms_error_vs_exact_solution_1": err1,
                  "collectives_per_iteration": {"ncclSend/ncclRecv (grouped, halo of r)": 2 if (world > 1 and not op.peer_halo) else 0,
                                                "halo_path": ("none (1 rank)" if world == 1 else
                                                              ("NVLink peer memory: the update kernel stores the boundary entries of r into the neighbours' ghost zones"
                                                               if op.peer_halo else "ncclSend/ncclRecv on a second stream, overlapped with the interior rows")),
                                                "reduction of 2 doubles": (0 if world == 1 else 1),
                                                "reduction_path": ("none (1 rank)" if world == 1 else
                                                                   ("NVLink peer memory, fused into the kernel that finishes the dot products and advances alpha/beta"
                                                                    if dd.p2p_reduction else "ncclAllReduce"))},
                  "interior_rows_overlap_halo": [op.interior_lo, op.interior_hi],
                  "method": "Chronopoulos-Gear CG, iteration replayed from a CUDA graph, femx_dist_cg (C++ behind the C ABI)"}
            op.close(); dd.close()
            del ones, b, x, y
        except Exception as e:
            cg = {"error": repr(e)[:400]}

    peak, peak_src = measured_peaks()
    stencil = pat.stencil()
    achieved = b_alg / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(args.workload) if world == 1 else (None, None)
    line = {
        "metric": "elements/s (fp64 P1 assembly into CSR)",
        "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": wl["desc"] + (f", z-slabs over {world} GPUs" if world > 1 else ""),
            "elements": ne_global, "elements_this_gpu": mesh.n_elems, "nodes_this_gpu": mesh.n_nodes,
            "nnz_this_gpu": pat.nnz, "nnz": nnz_global,
            "parallelism": f"owned node planes [{slab['r0']},{slab['r1']}) of {wl['n'] + 1} on rank {rank}; ghost cell layers, no collective in assembly",
            "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no explicit flush" % (b_alg / 1e9),
            "timing": "CUDA events on the launching stream, max over ranks",
            "launch": launch_mode,
        },
        "nnz_per_s": nnz_global / (ms_per_step * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "dram_frac": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     "kernel": "femx_csr", "kernel_ms": kern_ms, "kernel_ms_min": per_launch[0],
                     "algorithmic_bytes": b_alg, "peak_source": peak_src,
                     "note": "rank 0's launch; frac = algorithmic bytes / time / peak, dram_frac = profiled DRAM bytes / time / peak"},
        "e2e": {"value": ne_global / (e2e_ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": h2d_total,
                "d2h_bytes_per_step": d2h_total, "ms_per_step": e2e_ms,
                "h2d_GBps_per_gpu": h2d / (e2e_ms * 1e-3) / 1e9, "d2h_GBps_per_gpu": d2h / (e2e_ms * 1e-3) / 1e9,
                "includes": "H2D node coordinates (pinned), numeric pass, D2H CSR values; pattern reused, connectivity not re-uploaded; double-buffered over 3 streams",
                "matches_device_result": e2e_all_ok},
        # kernels of MINE launched inside the timed region: femx_csr per step, plus femx_rowlist (the boundary rows, side stream)
        # when the lattice pass runs
        "gpu_launches": args.steps * (2 if "lattice" in numeric_pass_kind(form) and stencil["rows"] < pat.n_rows else 1),
        "clocks": clocks,
        "parity": dict(parity, all_ranks_ok=parity_all_ok),
        "checksum": checksum_dev,
        "numeric_pass": {"kind": numeric_pass_kind(form), "class_rows": stencil["rows"], "rows": pat.n_rows,
                         "lattice": pat.lattice() is not None},
        "setup": {"pattern_build_ms": pattern_ms[-1], "pattern_build_first_call_ms": pattern_ms[0],
                  "pattern_nnz_per_s": pat.nnz / (pattern_ms[-1] * 1e-3), "jit_plus_first_launch_ms": jit_ms,
                  "pattern_bytes": pat.bytes,
                  "pattern_roofline": {"algorithmic_bytes": mesh.n_elems * mesh.nn * 4 + (pat.n_rows // nd + 1) * 8 + (pat.nnz // (nd * nd)) * 4,
                                       "frac": (mesh.n_elems * mesh.nn * 4 + (pat.n_rows // nd + 1) * 8 + (pat.nnz // (nd * nd)) * 4)
                                       / (pattern_ms[-1] * 1e-3) / 1e9 / peak}},
    }
    if cg is not None:
        line["cg"] = cg
    form.close(); pat.close()
    del vals, rp, mesh
    torch.cuda.empty_cache()
    if rank == 0 and world == 1:
        if not args.no_extras:
            line["extra"] = {name: run_extra(ctx, femx, torch, name) for name in ("cfg2", "cfg4") if name != args.workload}
            if ref_gpu is None and "cfg2" in line["extra"]:
                ref_gpu = line["extra"]["cfg2"].pop("ref_gpu_baseline", None)
        if ref_gpu:
            line["ref_gpu_baseline"] = ref_gpu
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(wl)
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
