"""ctypes/numpy front-end of oracle/femx_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POISSON, POISSON_MASS, MASS, ELASTICITY = 1, 2, 3, 4


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfemx_oracle.so")
        src = os.path.join(_HERE, "femx_oracle.c")
        if not os.path.exists(path) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)
        ):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_pattern.restype = C.c_int64
        _LIB.orc_cg.restype = C.c_int
    return _LIB


_NATIVE = None
NATIVE_FLAGS = "-O3 -march=native"


def lib_native():
    """The same source built with -O3 -march=native on THIS host (bench.py's CPU baseline legs only)."""
    global _NATIVE
    if _NATIVE is None:
        # -march=native code must not travel between hosts: one build per CPU model
        import hashlib
        try:
            model = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
        except OSError:
            model = []
        tag = hashlib.sha1("".join(model).encode()).hexdigest()[:12]
        out = os.path.join(_HERE, "_native", tag)
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libfemx_oracle_native.so")
        src = os.path.join(_HERE, "femx_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["gcc"] + NATIVE_FLAGS.split() + ["-fPIC", "-std=c99", "-shared", "-o", so, src, "-lm"])
        _NATIVE = C.CDLL(so)
        _NATIVE.orc_pattern.restype = C.c_int64
    return _NATIVE


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(v):
    return C.c_double(float(v))


def _i64(v):
    return C.c_int64(int(v))


def rect_mesh(x0, x1, y0, y1, n_row, n_col):
    """RectangleMesh(x0,x1,y0,y1,nRow,nCol) → X, Y, flag, conn[NE,3]."""
    nn = (n_row + 1) * (n_col + 1)
    ne = 2 * n_row * n_col
    X = np.empty(nn)
    Y = np.empty(nn)
    flag = np.empty(nn, np.int32)
    conn = np.empty((ne, 3), np.int32)
    lib().orc_rect_mesh(_d(x0), _d(x1), _d(y0), _d(y1), _i64(n_row), _i64(n_col),
                        _p(X), _p(Y), _p(flag), _p(conn))
    return X, Y, flag, conn


def box_mesh(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)):
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    ne = 6 * nx * ny * nz
    X, Y, Z = np.empty(nn), np.empty(nn), np.empty(nn)
    conn = np.empty((ne, 4), np.int32)
    lib().orc_box_mesh(_d(lo[0]), _d(hi[0]), _d(lo[1]), _d(hi[1]), _d(lo[2]), _d(hi[2]),
                       _i64(nx), _i64(ny), _i64(nz), _p(X), _p(Y), _p(Z), _p(conn))
    return X, Y, Z, conn


def _params(params):
    p = np.zeros(4)
    if params is not None:
        p[: len(params)] = params
    return p


def element_matrix(form, dim, nd, x, y, z=None, params=None, rule=None):
    nn = dim + 1
    n = nn * nd
    Ae = np.empty((n, n))
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    z = np.zeros(4) if z is None else np.ascontiguousarray(z, np.float64)
    p = _params(params)
    if rule is None:
        lib().orc_element_matrix(form, dim, nd, _p(p), _p(x), _p(y), _p(z), 0,
                                 None, None, None, None, _p(Ae))
    else:
        w, r, s = (np.ascontiguousarray(a, np.float64) for a in rule[:3])
        t = np.ascontiguousarray(rule[3], np.float64) if len(rule) > 3 else None
        lib().orc_element_matrix(form, dim, nd, _p(p), _p(x), _p(y), _p(z), len(w),
                                 _p(w), _p(r), _p(s), _p(t), _p(Ae))
    return Ae


def assemble_coo(form, dim, nd, conn, X, Y, Z=None, params=None):
    conn = np.ascontiguousarray(conn, np.int32)
    ne = conn.shape[0]
    n = (dim + 1) * nd
    A = np.empty(ne * n * n)
    row = np.empty(ne * n * n, np.int32)
    col = np.empty(ne * n * n, np.int32)
    p = _params(params)
    lib().orc_assemble_coo(form, dim, nd, _p(p), _i64(ne), _p(conn), _p(X), _p(Y),
                           _p(Z), _p(A), _p(row), _p(col))
    return A, row, col


def assemble_rhs(kind, dim, nd, conn, X, Y, Z=None, fvec=(1.0, 0.0, 0.0)):
    """Load vector: kind 0 constant source fvec, kind 1 the reference's f = -2(x^2+y^2)+36."""
    conn = np.ascontiguousarray(conn, np.int32)
    ne, nn = conn.shape
    f = np.zeros(3)
    f[: len(fvec)] = fvec
    b = np.empty(len(X) * nd)
    be = np.empty(ne * nn * nd)
    lib().orc_assemble_rhs(kind, dim, nd, _p(f), _i64(ne), _p(conn), _p(X), _p(Y), _p(Z), _i64(len(b)), _p(b), _p(be))
    return b, be.reshape(ne, nn * nd)


def pattern(conn, n_nodes):
    """Node-level CSR pattern (row_ptr int64, col_idx int32) of getNeighborNodesList."""
    conn = np.ascontiguousarray(conn, np.int32)
    ne, nn = conn.shape
    row_ptr = np.empty(n_nodes + 1, np.int64)
    nnz = lib().orc_pattern(nn, _i64(n_nodes), _i64(ne), _p(conn), _p(row_ptr), None)
    col = np.empty(nnz, np.int32)
    lib().orc_pattern(nn, _i64(n_nodes), _i64(ne), _p(conn), _p(row_ptr), _p(col))
    return row_ptr, col


def ell_pattern(row_ptr, col_idx, width):
    n = len(row_ptr) - 1
    ln = np.empty(n, np.int32)
    idx = np.empty((n, width), np.int32)
    lib().orc_csr_to_ell_pattern(_i64(n), _p(row_ptr), _p(col_idx), width, _p(ln), _p(idx))
    return ln, idx


def expand_pattern(nd, row_ptr, col_idx):
    n = len(row_ptr) - 1
    nnz = int(row_ptr[-1]) * nd * nd
    drp = np.empty(n * nd + 1, np.int64)
    dci = np.empty(nnz, np.int32)
    lib().orc_expand_pattern(nd, _i64(n), _p(row_ptr), _p(col_idx), _p(drp), _p(dci))
    return drp, dci


def assemble_csr(form, dim, nd, conn, X, Y, Z, drow_ptr, dcol_idx, params=None, native=False):
    conn = np.ascontiguousarray(conn, np.int32)
    vals = np.empty(int(drow_ptr[-1]))
    p = _params(params)
    (lib_native() if native else lib()).orc_assemble_csr(form, dim, nd, _p(p), _i64(conn.shape[0]), _p(conn), _p(X),
                           _p(Y), _p(Z), _p(drow_ptr), _p(dcol_idx),
                           _i64(len(drow_ptr) - 1), _p(vals))
    return vals


def apply_dirichlet(row_ptr, col_idx, flag, g, vals, rhs):
    """In place: symmetric elimination of the dofs with flag != 0."""
    flag = np.ascontiguousarray(flag, np.int32)
    g = np.ascontiguousarray(g, np.float64)
    lib().orc_apply_dirichlet(_i64(len(row_ptr) - 1), _p(row_ptr), _p(col_idx), _p(flag), _p(g), _p(vals), _p(rhs))


def spmv(row_ptr, col_idx, vals, x):
    y = np.empty(len(row_ptr) - 1)
    lib().orc_spmv(_i64(len(y)), _p(row_ptr), _p(col_idx), _p(vals), _p(x), _p(y))
    return y


def cg(row_ptr, col_idx, vals, b, iters):
    n = len(row_ptr) - 1
    x = np.empty(n)
    res = np.zeros(iters + 1)
    it = lib().orc_cg(_i64(n), _p(row_ptr), _p(col_idx), _p(vals), _p(b), _p(x), iters, _p(res))
    return x, res[: it + 1]
