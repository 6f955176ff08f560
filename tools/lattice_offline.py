"""Offline look at the element-once lattice pass (no GPU): JIT-compiles it for the Kuhn lattice cell
(what femx_mesh_box produces) and prints ptxas-level facts and a SASS opcode histogram.

  python tools/lattice_offline.py [FORM] [outdir]      (env FEMX_LT_TX / FEMX_LT_TY / FEMX_LT_MINB / FEMX_LT_REGS apply)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import femx  # noqa: E402

PERM = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
ODD = [0, 1, 1, 0, 0, 1]


def kuhn_corners():
    """corner codes (dx | dy << 1 | dz << 2) of the 6 Kuhn tets in femx_mesh_box's vertex order"""
    out = []
    for p, odd in zip(PERM, ODD):
        c = [0, 0, 0]
        v = [0]
        for a in p:
            c[a] += 1
            v.append(c[0] | c[1] << 1 | c[2] << 2)
        if not odd:
            v[1], v[2] = v[2], v[1]
        out.append(v)
    return out


def kuhn_offsets(sy, sz):
    offs = sorted({0} | {s * (dx + dy * sy + dz * sz) for s in (1, -1)
                         for dx, dy, dz in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))})
    return offs, offs.index(0)


def sass_histogram(cubin_path):
    sass = subprocess.run(["cuobjdump", "-sass", cubin_path], capture_output=True, text=True).stdout
    ops = {}
    for line in sass.splitlines():
        for tok in line.split():
            if tok[:1].isupper() and tok.split(".")[0] in ("DFMA", "DMUL", "DADD", "LDG", "LDS", "STS", "STG", "MUFU", "STL", "LDL", "BAR",
                                                           "UBLKCP", "SYNCS", "UTMACMDFLUSH", "DSETP", "SHFL", "MOV", "IMAD"):
                k = tok.split(".")[0]
                ops[k] = ops.get(k, 0) + 1
                break
    return dict(sorted(ops.items()))


def main():
    form_name = sys.argv[1] if len(sys.argv) > 1 else "POISSON_MASS"
    out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/lattice"
    os.makedirs(out, exist_ok=True)
    os.environ["FEMX_JIT_DUMP"] = out
    f = femx.Form(None, 3, getattr(femx, form_name), offline=True)
    sy, sz = 257, 257 * 257
    offs, self_pos = kuhn_offsets(sy, sz)
    import time
    t0 = time.time()
    cubin, info = f.cubin_lattice(kuhn_corners(), sy, sz, offs, self_pos)
    print(f"NVRTC: {time.time() - t0:.2f} s, cubin {len(cubin)} B, plan {info}")
    path = os.path.join(out, "femx_csr_lattice.cubin")
    open(path, "wb").write(cubin)
    print(subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout.strip())
    print("SASS op counts (static, whole kernel):", sass_histogram(path))
    f.close()


if __name__ == "__main__":
    main()
