"""Offline look at the stencil-specialised numeric pass (no GPU): computes the scatter codes of an
interior row of a structured mesh on the CPU (same rules as femx_pattern.cu: row_fill), JIT-compiles
the specialised femx_csr for it and prints ptxas-level facts from the cubin.

  python tools/stencil_offline.py [2|3] [FORM] [outdir]
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import femx  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def row_codes(conn, nn, row):
    """(codes, row_len, self_pos) of node `row`: incidences in ascending e*nn+li order, 7-bit positions of
    the other vertices (order femx_oth), first-touch flags at bits 21+j, li at bits 28-29."""
    conn = np.asarray(conn).reshape(-1, nn)
    inc = np.argwhere(conn == row)
    inc = sorted((int(e), int(li)) for e, li in inc)
    cols = sorted(set(int(v) for e, _ in inc for v in conn[e]))
    pos = {c: k for k, c in enumerate(cols)}
    seen, codes = set(), []
    for e, li in inc:
        code = li << 28
        # the device loop visits local vertices a = 0..nn-1 in order; j = slot of a among the others
        for a in range(nn):
            if a == li:
                continue
            j = (a ^ li) - 1 if nn == 4 else (a - li - 1 + 3) % 3
            p = pos[int(conn[e, a])]
            code |= p << (7 * j)
            if p not in seen:
                code |= 1 << (21 + j)
                seen.add(p)
        codes.append(code)
    return codes, len(cols), pos[row]


def interior_class(dim, n=6):
    if dim == 2:
        X, Y, _, conn = orc.rect_mesh(0.0, 1.0, 0.0, 1.0, n, n)
        row = (n // 2) * (n + 1) + n // 2
        return row_codes(conn, 3, row)
    X, Y, Z, conn = orc.box_mesh(n, n, n)
    m = n + 1
    row = ((n // 2) * m + n // 2) * m + n // 2
    return row_codes(conn, 4, row)


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    form_name = sys.argv[2] if len(sys.argv) > 2 else ("POISSON_MASS" if dim == 3 else "POISSON")
    out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/spec"
    os.makedirs(out, exist_ok=True)
    codes, rlen, self_pos = interior_class(dim)
    print(f"class: {len(codes)} incidences, {rlen} columns, own position {self_pos}")
    os.environ["FEMX_JIT_DUMP"] = out
    f = femx.Form(None, dim, getattr(femx, form_name), offline=True)
    import time
    t0 = time.time()
    cubin = f.cubin_stencil(codes, rlen, self_pos)
    print(f"NVRTC: {time.time() - t0:.2f} s, cubin {len(cubin)} B")
    path = os.path.join(out, f"femx_csr_spec_{dim}d.cubin")
    open(path, "wb").write(cubin)
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    print(res.strip())
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ops = {}
    for line in sass.splitlines():
        parts = line.split()
        for tok in parts:
            if tok[:1].isupper() and tok.split(".")[0] in ("DFMA", "DMUL", "DADD", "LDG", "LDS", "STS", "MUFU", "STL", "LDL", "BAR", "UBLKCP", "DSETP"):
                k = tok.split(".")[0]
                ops[k] = ops.get(k, 0) + 1
                break
    print("SASS op counts (static, whole kernel):", dict(sorted(ops.items())))
    f.close()


if __name__ == "__main__":
    main()
