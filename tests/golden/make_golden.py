"""Generates tests/golden/*.npz|json from the REFERENCE ITSELF (run in the authoring
container, where /root/reference exists; the fixtures travel, the reference does not).

What is pinned
  ref_integrand_strings.json  the 9 GiNaC-emitted LHS expressions the reference pasted
                              into its kernels (fea_test_sm_sym_sparse2.cu:188-205;
                              identical text at fea_test_sm_sym_sparse.cu:153-170) and the
                              quadrature literals (:30-33).
  (+ "rhs")                   the 3 RHS strings rhs[j] = f*phi_j*jac the reference generated and
                              discarded (fea_symbolic.cu:335,339,343)
  ref_poisson2d_*.npz         those strings evaluated VERBATIM (python eval, float64,
                              powf→pow) on RectangleMesh inputs built with the reference's
                              mesh semantics (fea_test_sm_sym_sparse2.cu:119-165), summed
                              over the 7 quadrature points as the kernel does (:260-264):
                              COO triplets in slot order e*9+li*3+lj, and the reference
                              host pattern (getNeighborNodesList, :72-100, python sets)
                              with the ELL(7) accumulation of the kernel (:274-282).
Usage:  python tests/golden/make_golden.py
"""
import json
import os
import re

import numpy as np

REF = "/root/reference/fea_test_sm_sym_sparse2.cu"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_reference():
    src = open(REF).read().splitlines()
    exprs = {}
    for i, line in enumerate(src):
        m = re.match(r"\s*if\(funIdx == (\d)\)", line)
        if m:
            body = src[i + 1].strip()
            assert body.startswith("return ") and body.endswith(";")
            exprs[int(m.group(1))] = body[len("return "):-1]
    assert sorted(exprs) == list(range(9))
    # RHS strings rhs[j] = f*phi_j*jac, f = -2(x^2+y^2)+36: generated and then DISCARDED by the
    # reference (fea_symbolic_nvrtc_sparse.cpp:346-351); its recorded output: fea_symbolic.cu:335,339,343
    sym = open("/root/reference/fea_symbolic.cu").read().splitlines()
    rhs = [sym[k - 1].strip() for k in (335, 339, 343)]
    assert all("18.0" in r and "std::" not in r for r in rhs)
    lit = {}
    for name in ("triW", "triR", "triS", "triT"):
        line = next(l for l in src if f"float {name}[7]" in l)
        vals = re.search(r"\{(.*)\}", line).group(1)
        lit[name] = [v.strip().rstrip("f") for v in vals.split(",")]
    lit["rhs"] = rhs
    return [exprs[k] for k in range(9)], lit


def rect_mesh(x0, x1, y0, y1, n_row, n_col):
    """RectangleMesh::generate restated line by line (fea_test_sm_sym_sparse2.cu:119-165)."""
    stepx = (x1 - x0) / n_col
    stepy = (y1 - y0) / n_row
    X, Y = [], []
    for i in range(n_row + 1):
        y = y0 + i * stepy
        for j in range(n_col + 1):
            X.append(x0 + j * stepx)
            Y.append(y)
    conn = []
    for i in range(n_row):
        for j in range(n_col):
            n1 = i * (n_col + 1) + j
            n2 = n1 + 1
            n3 = (i + 1) * (n_col + 1) + j
            conn.append((n1, n2, n3))
            n1 = i * (n_col + 1) + j + 1
            n2 = (i + 1) * (n_col + 1) + j + 1
            n3 = n2 - 1
            conn.append((n1, n2, n3))
    return np.array(X), np.array(Y), np.array(conn, np.int32)


def neighbor_list(conn, n_nodes):
    """getNeighborNodesList (fea_test_sm_sym_sparse2.cu:72-100) with python sets."""
    nb = [set() for _ in range(n_nodes)]
    for e in conn:
        for j in e:
            for jj in e:
                nb[j].add(int(jj))
    return [sorted(s) for s in nb]


def evaluate(exprs, lit, X, Y, conn):
    w = [float(v) for v in lit["triW"]]
    r_, s_, t_ = ([float(v) for v in lit[k]] for k in ("triR", "triS", "triT"))
    codes = [compile(e.replace("powf", "pow"), f"<integrand{k}>", "eval") for k, e in enumerate(exprs)]
    A = np.zeros((len(conn), 9))
    for e, (a, b, c) in enumerate(conn):
        env = dict(x1=X[a], x2=X[b], x3=X[c], y1=Y[a], y2=Y[b], y3=Y[c], pow=pow)
        for k in range(9):
            acc = 0.0
            for q in range(7):
                env.update(r=r_[q], s=s_[q], t=t_[q])
                acc += w[q] * eval(codes[k], {"__builtins__": {}}, env)
            A[e, k] = acc
    return A


def evaluate_rhs(lit, X, Y, conn):
    w = [float(v) for v in lit["triW"]]
    r_, s_, t_ = ([float(v) for v in lit[k]] for k in ("triR", "triS", "triT"))
    codes = [compile(e, f"<rhs{k}>", "eval") for k, e in enumerate(lit["rhs"])]
    B = np.zeros((len(conn), 3))
    for e, (a, b, c) in enumerate(conn):
        env = dict(x1=X[a], x2=X[b], x3=X[c], y1=Y[a], y2=Y[b], y3=Y[c], pow=pow)
        for k in range(3):
            acc = 0.0
            for q in range(7):
                env.update(r=r_[q], s=s_[q], t=t_[q])
                acc += w[q] * eval(codes[k], {"__builtins__": {}}, env)
            B[e, k] = acc
    return B


def make_case(name, exprs, lit, X, Y, conn):
    A = evaluate(exprs, lit, X, Y, conn)
    B = evaluate_rhs(lit, X, Y, conn)
    bvec = np.zeros(len(X))
    for e in range(len(conn)):          # element order (the loop the author sketched: fea_kernal.cu:193-214)
        for k in range(3):
            bvec[conn[e, k]] += B[e, k]
    ne = len(conn)
    row = np.empty((ne, 9), np.int32)
    col = np.empty((ne, 9), np.int32)
    for k in range(9):
        row[:, k] = conn[:, k // 3]
        col[:, k] = conn[:, k % 3]
    nb = neighbor_list(conn, len(X))
    width = 7
    ell_len = np.array([len(s) for s in nb], np.int32)
    assert ell_len.max() <= width
    ell_idx = np.zeros((len(X), width), np.int32)
    for i, s in enumerate(nb):
        ell_idx[i, : len(s)] = s
    ell_val = np.zeros((len(X), width))
    for e in range(ne):  # element order; the reference's atomicAdd order is unspecified
        for k in range(9):
            gi, gj = row[e, k], col[e, k]
            ell_val[gi, nb[gi].index(gj)] += A[e, k]
    np.savez_compressed(os.path.join(HERE, name), X=X, Y=Y, conn=conn, A=A.ravel(), rowA=row.ravel(),
                        colA=col.ravel(), ell_len=ell_len, ell_idx=ell_idx, ell_val=ell_val, rhs_elem=B, rhs=bvec)
    print(name, "elements", ne, "nnz", int(ell_len.sum()))


def main():
    exprs, lit = parse_reference()
    json.dump({"source": "fea_test_sm_sym_sparse2.cu:188-205 and :30-33 (GiNaC csrc_float output as pasted by the reference)",
               "integrand": exprs, **lit}, open(os.path.join(HERE, "ref_integrand_strings.json"), "w"), indent=1)
    # the reference's own run configurations: 2x2 (fea_symbolic_nvrtc_sparse.cpp:366-367,491),
    # 4x4 (fea_test_sm_sym.cu), both on [-3,3]^2; plus a non-square and a jittered mesh
    X, Y, conn = rect_mesh(-3.0, 3.0, -3.0, 3.0, 2, 2)
    make_case("ref_poisson2d_2x2.npz", exprs, lit, X, Y, conn)
    X, Y, conn = rect_mesh(-3.0, 3.0, -3.0, 3.0, 4, 4)
    make_case("ref_poisson2d_4x4.npz", exprs, lit, X, Y, conn)
    X, Y, conn = rect_mesh(-3.0, 3.0, -3.0, 3.0, 10, 7)  # MESH_W → nRow, MESH_H → nCol
    make_case("ref_poisson2d_10x7.npz", exprs, lit, X, Y, conn)
    X, Y, conn = rect_mesh(0.0, 1.0, 0.0, 1.0, 12, 12)
    rng = np.random.RandomState(12345)
    inner = (X > 0) & (X < 1) & (Y > 0) & (Y < 1)
    X = X + inner * rng.uniform(-0.2, 0.2, X.shape) / 12
    Y = Y + inner * rng.uniform(-0.2, 0.2, Y.shape) / 12
    make_case("ref_poisson2d_jitter12.npz", exprs, lit, X, Y, conn)


if __name__ == "__main__":
    main()
