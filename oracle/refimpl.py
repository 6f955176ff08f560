"""ctypes front-end of oracle/_ref/*.so — the REFERENCE's own code recompiled for sm_100
(see oracle/build_ref.py).  TEST / BASELINE INFRASTRUCTURE ONLY; the product never imports it.

Host parts (RectangleMesh, getNeighborNodesList) run anywhere; the kernels need a GPU and take
raw device pointers (torch tensors' data_ptr()).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_libs = {}


def available():
    return all(os.path.exists(os.path.join(_REF, f"libref_{k}_{p}.so")) for k in ("coo", "ell") for p in ("f32", "f64"))


def lib(kind, prec):
    key = (kind, prec)
    if key not in _libs:
        _libs[key] = C.CDLL(os.path.join(_REF, f"libref_{kind}_{prec}.so"))
    return _libs[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_mesh(x0, x1, y0, y1, n_row, n_col, prec="f64"):
    """The reference's RectangleMesh + flattening loop: node coords, flags, X[3e+k], Y[3e+k], gIdx[3e+k]."""
    n = (n_row + 1) * (n_col + 1)
    ne = 2 * n_row * n_col
    rt = np.float64 if prec == "f64" else np.float32
    nx, ny, fl = np.empty(n), np.empty(n), np.empty(n, np.int32)
    X, Y, g = np.empty(3 * ne, rt), np.empty(3 * ne, rt), np.empty(3 * ne, np.int32)
    lib("ell", prec).ref_host_mesh(C.c_double(x0), C.c_double(x1), C.c_double(y0), C.c_double(y1), n_row, n_col,
                                   _p(nx), _p(ny), _p(fl), _p(X), _p(Y), _p(g))
    return nx, ny, fl, X, Y, g.reshape(-1, 3)


def neighbor_list(n_row, n_col, width=7):
    """Mesh::getNeighborNodesList of the reference, untouched: len[i], idx[i*width+j]."""
    n = (n_row + 1) * (n_col + 1)
    ln = np.zeros(n, np.int32)
    idx = np.zeros((n, width), np.int32)
    lib("ell", "f64").ref_neighbor_list(n_row, n_col, _p(ln), width, _p(idx))
    return ln, idx


def _vp(t):
    return C.c_void_p(t.data_ptr())


def assemble_coo(prec, n_row, n_col, X, Y, gIdx, iters=1):
    """Reference kernel K4 on device tensors X,Y [3*NE], gIdx [3*NE] → A,rowA,colA [9*NE], ms/launch."""
    import torch
    L = lib("coo", prec)
    L.ref_set_mesh(C.c_long(n_row), C.c_long(n_col))
    ne = gIdx.numel() // 3
    A = torch.zeros(9 * ne, dtype=X.dtype, device=X.device)
    row = torch.zeros(9 * ne, dtype=torch.int32, device=X.device)
    col = torch.zeros(9 * ne, dtype=torch.int32, device=X.device)
    ms = C.c_float()
    torch.cuda.synchronize()
    err = L.ref_assemble_coo(C.c_long(ne), _vp(A), _vp(row), _vp(col), _vp(X), _vp(Y), _vp(gIdx), iters, C.byref(ms))
    torch.cuda.synchronize()
    assert err == 0, f"reference COO kernel: CUDA error {err}"
    return A, row, col, ms.value


def assemble_ell(prec, n_row, n_col, X, Y, gIdx, ell_len, ell_idx, iters=1):
    """Reference kernel K5 (linear search + global atomicAdd) → A [n_nodes*7], ms/launch."""
    import torch
    L = lib("ell", prec)
    L.ref_set_mesh(C.c_long(n_row), C.c_long(n_col))
    ne = gIdx.numel() // 3
    A = torch.zeros(ell_idx.numel(), dtype=X.dtype, device=X.device)
    ms = C.c_float()
    torch.cuda.synchronize()
    err = L.ref_assemble_ell(C.c_long(ne), _vp(A), _vp(ell_len), _vp(ell_idx), _vp(X), _vp(Y), _vp(gIdx), iters,
                             C.byref(ms))
    torch.cuda.synchronize()
    assert err == 0, f"reference ELL kernel: CUDA error {err}"
    return A, ms.value


def atomic_variants(n=10_240_000, n_cas=20_480, iters=3):   # n: a multiple of the 32x32x100 grid row (the staged variant reads no unset shared memory)
    """The reference's three accumulation variants (atomicadd.cu:73-129) on n ones: ms per launch and the sums
    (fp32 naive / fp32 shared-memory staged / fp64 CAS loop).  The contention K5 suffers, as a micro-benchmark."""
    path = os.path.join(_REF, "libref_atomicadd.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    ms = (C.c_float * 3)()
    res = (C.c_double * 3)()
    err = L.ref_atomic_variants(C.c_long(n), C.c_long(n_cas), int(iters), ms, res)
    names = ("naive_global_atomicAdd_f32", "smem_staged_block_sum_f32", "cas_loop_f64")
    sizes = (n, n, n_cas)
    return {"cuda_error": err, **{k: {"n": sizes[i], "ms": ms[i], "sum": res[i], "values_per_s": sizes[i] / (ms[i] * 1e-3)}
                                  for i, k in enumerate(names)}}
