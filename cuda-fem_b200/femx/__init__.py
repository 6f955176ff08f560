"""Python host-side mirror of the femx C ABI (include/femx.h), via ctypes.

This is plumbing for tests and bench.py: torch supplies device memory and
streams, every compute call goes straight through libfemx.so.  There is no CPU
fallback — if the library is missing or no device is present the calls raise.

Names follow the reference's vocabulary (fea_symbolic_nvrtc_sparse*.cpp):
RectangleMesh, WeakForm (→ Form), getNeighborNodesList (→ Pattern), rowA/colA/A.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libfemx.so")

F64, F32 = 0, 1
CUSTOM, POISSON, POISSON_MASS, MASS, ELASTICITY = 0, 1, 2, 3, 4

_STATUS = {1: "INVALID", 2: "CUDA", 3: "NVRTC", 4: "UNSUPPORTED", 5: "NOMEM"}


class FemxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"femx error {status} ({_STATUS.get(status, '?')}): {msg}")
        self.status = status


class _FormDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("nn", C.c_int), ("nd", C.c_int), ("dtype", C.c_int),
        ("builtin", C.c_int), ("params", C.c_double * 4),
        ("entries", C.POINTER(C.c_char_p)), ("prologue", C.c_char_p),
        ("nq", C.c_int), ("qw", C.POINTER(C.c_double)), ("qr", C.POINTER(C.c_double)),
        ("qs", C.POINTER(C.c_double)), ("qt", C.POINTER(C.c_double)),
        ("qu", C.POINTER(C.c_double)), ("fmad", C.c_int), ("integrated", C.c_int),
        ("rhs_entries", C.POINTER(C.c_char_p)), ("rhs_vec", C.c_double * 3),
    ]


class _MeshView(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("nn", C.c_int), ("n_nodes", C.c_int64), ("n_elems", C.c_int64),
        ("d_conn", C.c_void_p), ("d_node_xyz", C.c_void_p * 3), ("node_stride", C.c_int64),
        ("d_elem_xyz", C.c_void_p * 3),
    ]


# every symbol include/femx.h declares (tests check the library exports them all)
SYMBOLS = [
    "femx_ctx_create", "femx_ctx_destroy", "femx_last_error", "femx_version", "femx_ctx_set_option",
    "femx_pattern_lattice", "femx_form_cubin_lattice",
    "femx_dist_unique_id", "femx_dist_create", "femx_dist_info",
    "femx_partition_extract", "femx_part_destroy", "femx_part_info", "femx_part_arrays", "femx_part_copy", "femx_part_gather",
    "femx_pattern_export_csr_mapped", "femx_dist_destroy", "femx_dist_slab", "femx_dist_allreduce",
    "femx_dist_op_create", "femx_dist_op_destroy", "femx_dist_op_info", "femx_dist_op_peer_halo", "femx_lattice_prefix", "femx_dist_spmv", "femx_dist_cg", "femx_spmv_rows",
    "femx_form_compile", "femx_form_compile_offline", "femx_form_destroy", "femx_form_source",
    "femx_form_log", "femx_form_entry", "femx_form_prologue", "femx_form_cubin",
    "femx_mesh_rectangle", "femx_mesh_expand", "femx_mesh_box",
    "femx_assemble_coo", "femx_pattern_build", "femx_pattern_destroy", "femx_pattern_info",
    "femx_pattern_bytes", "femx_pattern_export_csr", "femx_pattern_export_ell",
    "femx_pattern_stencil", "femx_form_cubin_stencil",
    "femx_assemble_csr", "femx_assemble_rhs", "femx_apply_dirichlet", "femx_csr_to_ell", "femx_spmv", "femx_dot2", "femx_axpy_ratio",
    "femx_xpby_ratio", "femx_io_read_gmsh", "femx_io_free", "femx_io_write_matrix_market",
]

_lib = None


def lib():
    """Load libfemx.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FemxError(2, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.femx_last_error.restype = C.c_char_p
        L.femx_version.restype = C.c_char_p
        L.femx_form_source.restype = C.c_char_p
        L.femx_form_log.restype = C.c_char_p
        L.femx_form_entry.restype = C.c_char_p
        L.femx_form_prologue.restype = C.c_char_p
        L.femx_pattern_bytes.restype = C.c_int64
        for n in ("femx_form_source", "femx_form_log", "femx_form_prologue", "femx_last_error",
                  "femx_form_destroy", "femx_pattern_destroy", "femx_ctx_destroy", "femx_pattern_bytes"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.femx_form_entry.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _vp(x):
    """device pointer of a torch tensor / int / None → c_void_p"""
    if x is None:
        return C.c_void_p(None)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _stream(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def _i64(v):
    return C.c_int64(int(v))


def _dbl(v):
    return C.c_double(float(v))


class Context:
    """One per device (replaces cuInit/cuCtxCreate, fea_symbolic_nvrtc_sparse.cpp:557-559)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        st = lib().femx_ctx_create(int(device), C.byref(self.h))
        if st:
            raise FemxError(st, lib().femx_last_error(None).decode())
        self.device = device

    def check(self, st):
        if st:
            raise FemxError(st, lib().femx_last_error(self.h).decode())

    def set_option(self, name, value):
        """Tuning option (femx_ctx_set_option): "spec", "lattice", "lt_tx", ... — experiments and tests."""
        self.check(lib().femx_ctx_set_option(self.h, name.encode(), int(value)))

    def close(self):
        if self.h:
            lib().femx_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---- structured generators (RectangleMesh::generate on the device)
    def rectangle_mesh(self, x0, x1, y0, y1, n_row, n_col, row_lo=0, row_hi=None, dtype=F64,
                       flags=False, stream=None):
        import torch
        row_hi = n_row if row_hi is None else row_hi
        nl = (row_hi - row_lo + 1) * (n_col + 1)
        ne = 2 * (row_hi - row_lo) * n_col
        dev = torch.device("cuda", self.device)
        tdt = torch.float64 if dtype == F64 else torch.float32
        X = torch.empty(nl, dtype=tdt, device=dev)
        Y = torch.empty(nl, dtype=tdt, device=dev)
        flag = torch.empty(nl, dtype=torch.int32, device=dev) if flags else None
        conn = torch.empty((ne, 3), dtype=torch.int32, device=dev)
        self.check(lib().femx_mesh_rectangle(self.h, _dbl(x0), _dbl(x1), _dbl(y0), _dbl(y1), _i64(n_row),
                                             _i64(n_col), _i64(row_lo), _i64(row_hi), dtype, _vp(X), _vp(Y),
                                             _vp(flag), _vp(conn), _stream(stream)))
        m = Mesh(2, conn, (X, Y))
        m.flag = flag
        return m

    def box_mesh(self, nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), k_lo=0, k_hi=None, dtype=F64,
                 stream=None):
        import torch
        k_hi = nz if k_hi is None else k_hi
        nl = (k_hi - k_lo + 1) * (nx + 1) * (ny + 1)
        ne = 6 * (k_hi - k_lo) * nx * ny
        dev = torch.device("cuda", self.device)
        tdt = torch.float64 if dtype == F64 else torch.float32
        X, Y, Z = (torch.empty(nl, dtype=tdt, device=dev) for _ in range(3))
        conn = torch.empty((ne, 4), dtype=torch.int32, device=dev)
        self.check(lib().femx_mesh_box(self.h, _dbl(lo[0]), _dbl(hi[0]), _dbl(lo[1]), _dbl(hi[1]), _dbl(lo[2]),
                                       _dbl(hi[2]), _i64(nx), _i64(ny), _i64(nz), _i64(k_lo), _i64(k_hi), dtype,
                                       _vp(X), _vp(Y), _vp(Z), _vp(conn), _stream(stream)))
        return Mesh(3, conn, (X, Y, Z))


class Mesh:
    """Device mesh: conn[NE, nn] int32 (gIdx layout) + node coordinates (SoA) and/or
    element-expanded coordinates X[nn*e+k] (the reference's layout)."""

    def __init__(self, dim, conn, node_xyz=None, elem_xyz=None, n_nodes=None):
        self.dim = dim
        self.nn = dim + 1
        self.conn = conn
        self.node_xyz = node_xyz
        self.elem_xyz = elem_xyz
        self.n_elems = 0 if conn is None else int(conn.shape[0])
        if conn is None and elem_xyz is not None:
            self.n_elems = int(elem_xyz[0].numel()) // self.nn
        self.n_nodes = int(n_nodes) if n_nodes is not None else (
            int(node_xyz[0].numel()) if node_xyz is not None else 0)
        self.flag = None

    def expanded(self, ctx, stream=None):
        """Element-expanded copy (replaces the host loop fea_symbolic_nvrtc_sparse.cpp:571-583)."""
        import torch
        out = []
        for c in self.node_xyz:
            e = torch.empty(self.n_elems * self.nn, dtype=c.dtype, device=c.device)
            dt = F64 if c.dtype == torch.float64 else F32
            ctx.check(lib().femx_mesh_expand(ctx.h, dt, self.nn, _i64(self.n_elems), _vp(self.conn), _vp(c),
                                             _vp(e), _stream(stream)))
            out.append(e)
        return Mesh(self.dim, self.conn, None, tuple(out), n_nodes=self.n_nodes)

    def view(self):
        v = _MeshView()
        v.dim, v.nn = self.dim, self.nn
        v.n_nodes, v.n_elems = self.n_nodes, self.n_elems
        v.d_conn = None if self.conn is None else self.conn.data_ptr()
        v.node_stride = 1
        for k in range(3):
            v.d_node_xyz[k] = None
            v.d_elem_xyz[k] = None
        if self.elem_xyz is not None:
            for k, c in enumerate(self.elem_xyz):
                v.d_elem_xyz[k] = c.data_ptr()
        elif self.node_xyz is not None:
            for k, c in enumerate(self.node_xyz):
                v.d_node_xyz[k] = c.data_ptr()
        return v


class Form:
    """A JIT-compiled element integrand (WeakForm::build + NVRTC + module load,
    fea_symbolic_nvrtc_sparse.cpp:307-356, 506-561)."""

    def __init__(self, ctx, dim, builtin=POISSON, nd=1, dtype=F64, params=(), entries=None, prologue=None,
                 rule=None, fmad=True, offline=False, integrated=False, rhs=None, rhs_vec=(1.0, 0.0, 0.0)):
        self.ctx = ctx
        self.dim, self.nn, self.nd, self.dtype = dim, dim + 1, nd, dtype
        self.n = self.nn * nd
        d = _FormDesc()
        d.dim, d.nn, d.nd, d.dtype, d.builtin = dim, dim + 1, nd, dtype, builtin
        for i in range(4):
            d.params[i] = float(params[i]) if i < len(params) else 0.0
        keep = []
        if entries is not None:
            flat = [e for row in entries for e in row] if isinstance(entries[0], (list, tuple)) else list(entries)
            arr = (C.c_char_p * len(flat))(*[s.encode() for s in flat])
            keep.append(arr)
            d.entries = C.cast(arr, C.POINTER(C.c_char_p))
            d.builtin = CUSTOM
        d.prologue = prologue.encode() if prologue else None
        if rhs is not None:
            rarr = (C.c_char_p * len(rhs))(*[s_.encode() for s_ in rhs])
            keep.append(rarr)
            d.rhs_entries = C.cast(rarr, C.POINTER(C.c_char_p))
        for i in range(3):
            d.rhs_vec[i] = float(rhs_vec[i]) if i < len(rhs_vec) else 0.0
        if rule is not None:
            cols = [list(map(float, c)) if c is not None else None for c in rule]
            while len(cols) < 5:
                cols.append(None)
            d.nq = len(cols[0])
            ptrs = []
            for c in cols:
                if c is None:
                    ptrs.append(None)
                else:
                    a = (C.c_double * len(c))(*c)
                    keep.append(a)
                    ptrs.append(C.cast(a, C.POINTER(C.c_double)))
            d.qw, d.qr, d.qs, d.qt, d.qu = ptrs
        d.fmad = 1 if fmad else 0
        d.integrated = 1 if integrated else 0
        self.h = C.c_void_p()
        if offline:
            st = lib().femx_form_compile_offline(C.byref(d), C.byref(self.h))
            if st:
                raise FemxError(st, lib().femx_last_error(None).decode())
        else:
            ctx.check(lib().femx_form_compile(ctx.h, C.byref(d), C.byref(self.h)))

    def _check(self, st):
        if st:
            raise FemxError(st, lib().femx_last_error(self.ctx.h if self.ctx else None).decode())

    @property
    def source(self):
        return lib().femx_form_source(self.h).decode()

    @property
    def log(self):
        return lib().femx_form_log(self.h).decode()

    @property
    def prologue(self):
        return lib().femx_form_prologue(self.h).decode()

    def entry(self, li, lj):
        s = lib().femx_form_entry(self.h, li, lj)
        return None if s is None else s.decode()

    def cubin(self, kernel="csr"):
        p = C.c_void_p()
        n = C.c_size_t()
        self._check(lib().femx_form_cubin(self.h, kernel.encode(), C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value)

    def cubin_stencil(self, codes, row_len, self_pos):
        """Numeric-pass kernel specialised for an explicit stencil class (diagnostic; works offline)."""
        arr = (C.c_uint32 * len(codes))(*[int(c) for c in codes])
        p = C.c_void_p()
        n = C.c_size_t()
        self._check(lib().femx_form_cubin_stencil(self.h, len(codes), int(row_len), int(self_pos), arr,
                                                  C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value)

    def cubin_lattice(self, corners, stride_y, stride_z, offsets, self_pos):
        """Element-once lattice pass for an explicit lattice cell (diagnostic; works offline).
        Returns (cubin, info dict)."""
        flat = [int(c) for t in corners for c in t]
        carr = (C.c_int32 * len(flat))(*flat)
        oarr = (C.c_int32 * len(offsets))(*[int(v) for v in offsets])
        info = (C.c_int * 6)()
        p = C.c_void_p()
        n = C.c_size_t()
        self._check(lib().femx_form_cubin_lattice(self.h, len(corners), carr, _i64(stride_y), _i64(stride_z),
                                                  len(offsets), int(self_pos), oarr, C.byref(p), C.byref(n), info))
        return C.string_at(p.value, n.value), dict(tx=info[0], ty=info[1], threads=info[2], nslot=info[3],
                                                   smem=info[4], minb=info[5])

    def _tdtype(self):
        import torch
        return torch.float64 if self.dtype == F64 else torch.float32

    def assemble_coo(self, mesh, A=None, rowA=None, colA=None, stream=None, indices=True):
        """COO triplets, slot e*n*n + li*n + lj (kernel ABI #1)."""
        import torch
        n2 = mesh.n_elems * self.n * self.n
        dev = torch.device("cuda", self.ctx.device)
        if A is None:
            A = torch.empty(n2, dtype=self._tdtype(), device=dev)
        if indices and rowA is None:
            rowA = torch.empty(n2, dtype=torch.int32, device=dev)
        if indices and colA is None:
            colA = torch.empty(n2, dtype=torch.int32, device=dev)
        v = mesh.view()
        self._check(lib().femx_assemble_coo(self.h, C.byref(v), _vp(A), _vp(rowA), _vp(colA), _stream(stream)))
        return A, rowA, colA

    def assemble_csr(self, pattern, mesh, values=None, stream=None):
        """Deterministic numeric pass into the pattern's CSR (kernel ABI #2)."""
        import torch
        if values is None:
            values = torch.empty(pattern.nnz, dtype=self._tdtype(), device=torch.device("cuda", self.ctx.device))
        v = mesh.view()
        self._check(lib().femx_assemble_csr(self.h, pattern.h, C.byref(v), _vp(values), _stream(stream)))
        return values

    def assemble_rhs(self, pattern, mesh, out=None, stream=None):
        """Load vector b (one value per dof row of the pattern), deterministic."""
        import torch
        if out is None:
            out = torch.empty(pattern.n_rows, dtype=self._tdtype(), device=torch.device("cuda", self.ctx.device))
        v = mesh.view()
        self._check(lib().femx_assemble_rhs(self.h, pattern.h, C.byref(v), _vp(out), _stream(stream)))
        return out

    def close(self):
        if self.h:
            lib().femx_form_destroy(self.h)
            self.h = C.c_void_p()


class Pattern:
    """CSR pattern + scatter map (the symbolic pass; replaces getNeighborNodesList,
    fea_symbolic_nvrtc_sparse2.cpp:181-210)."""

    def __init__(self, ctx, mesh, nd=1, row_begin=0, row_end=None, col_base=0, stream=None):
        self.ctx = ctx
        self.nd = nd
        self.nn = mesh.nn
        row_end = mesh.n_nodes if row_end is None else row_end
        self.h = C.c_void_p()
        ctx.check(lib().femx_pattern_build(ctx.h, mesh.nn, nd, _i64(mesh.n_nodes), _i64(mesh.n_elems),
                                           _vp(mesh.conn), _i64(row_begin), _i64(row_end), _i64(col_base),
                                           _stream(stream), C.byref(self.h)))
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        ctx.check(lib().femx_pattern_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_rows, self.nnz, self.max_row = a.value, b.value, c.value

    @property
    def bytes(self):
        return lib().femx_pattern_bytes(self.h)

    def stencil(self):
        """Dominant stencil class found by the symbolic pass: dict(n_incid, row_len, self_pos, rows,
        codes, offsets); rows == 0 when there is none (unstructured mesh, nd > 1, FEMX_SPEC=0)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        r = C.c_int64()
        codes = (C.c_uint32 * 32)()
        offs = (C.c_int32 * 32)()
        self.ctx.check(lib().femx_pattern_stencil(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(r), codes, offs, 32))
        return dict(n_incid=a.value, row_len=b.value, self_pos=c.value, rows=r.value,
                    codes=[int(codes[k]) for k in range(a.value)], offsets=[int(offs[k]) for k in range(b.value)])

    def lattice(self):
        """Lattice structure found by the symbolic pass (femx_pattern_lattice), or None."""
        P = C.c_int()
        cells, strides = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        node0 = C.c_int64()
        corners = (C.c_int32 * 32)()
        self.ctx.check(lib().femx_pattern_lattice(self.h, C.byref(P), cells, strides, C.byref(node0), corners))
        if P.value == 0:
            return None
        nn = self.nn
        return dict(n_per_cell=P.value, cells=list(cells), strides=list(strides), node0=node0.value,
                    corners=[[int(corners[t * nn + a]) for a in range(nn)] for t in range(P.value)])

    def csr(self, index_dtype="int32", stream=None):
        import torch
        dev = torch.device("cuda", self.ctx.device)
        col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        if index_dtype == "int64":
            rp = torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev)
            self.ctx.check(lib().femx_pattern_export_csr(self.h, _vp(rp), _vp(None), _vp(col), _stream(stream)))
        else:
            rp = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
            self.ctx.check(lib().femx_pattern_export_csr(self.h, _vp(None), _vp(rp), _vp(col), _stream(stream)))
        return rp, col

    def ell(self, width, stream=None):
        """gNbrNodeLen / gNbrNodeIdx in the reference's padded layout."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        ln = torch.empty(self.n_rows, dtype=torch.int32, device=dev)
        idx = torch.empty((self.n_rows, width), dtype=torch.int32, device=dev)
        self.ctx.check(lib().femx_pattern_export_ell(self.h, int(width), _vp(ln), _vp(idx), _stream(stream)))
        return ln, idx

    def values_to_ell(self, values, width, stream=None):
        import torch
        out = torch.empty((self.n_rows, width), dtype=values.dtype, device=values.device)
        dt = F64 if values.dtype == torch.float64 else F32
        self.ctx.check(lib().femx_csr_to_ell(self.h, dt, int(width), _vp(values), _vp(out), _stream(stream)))
        return out

    def apply_dirichlet(self, flag, g, values, rhs=None, stream=None):
        """Symmetric elimination of the dofs with flag != 0 (values g); in place."""
        import torch
        dt = F64 if values.dtype == torch.float64 else F32
        self.ctx.check(lib().femx_apply_dirichlet(self.h, dt, _vp(flag), _vp(g), _vp(values), _vp(rhs), _stream(stream)))

    def spmv(self, values, x, x_base=0, y=None, stream=None):
        import torch
        if y is None:
            y = torch.empty(self.n_rows, dtype=values.dtype, device=values.device)
        dt = F64 if values.dtype == torch.float64 else F32
        self.ctx.check(lib().femx_spmv(self.h, dt, _vp(values), _vp(x), _i64(x_base), _vp(y), _stream(stream)))
        return y

    def close(self):
        if self.h:
            lib().femx_pattern_destroy(self.h)
            self.h = C.c_void_p()


class Partition:
    """Ghost-element sub-mesh of the rank that owns global nodes [node_lo, node_hi) of an unstructured mesh
    (femx_partition_extract): local connectivity, local -> global node map, owned local rows [row_begin, row_end)."""

    def __init__(self, ctx, mesh, node_lo, node_hi, stream=None):
        import torch
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx.check(lib().femx_partition_extract(ctx.h, mesh.nn, _i64(mesh.n_nodes), _i64(mesh.n_elems), _vp(mesh.conn),
                                               _i64(node_lo), _i64(node_hi), _stream(stream), C.byref(self.h)))
        v = [C.c_int64() for _ in range(4)]
        ctx.check(lib().femx_part_info(self.h, *[C.byref(x) for x in v]))
        self.n_nodes, self.n_elems, self.row_begin, self.row_end = (x.value for x in v)
        dev = mesh.conn.device
        self.l2g = torch.empty(self.n_nodes, dtype=torch.int32, device=dev)
        self.conn = torch.empty((self.n_elems, mesh.nn), dtype=torch.int32, device=dev)
        self.elem_ids = torch.empty(self.n_elems, dtype=torch.int32, device=dev)
        ctx.check(lib().femx_part_copy(self.h, _vp(self.conn), _vp(self.l2g), _vp(self.elem_ids), _stream(stream)))
        # node coordinates of the sub-mesh
        xyz = []
        for c in mesh.node_xyz:
            o = torch.empty(self.n_nodes, dtype=c.dtype, device=dev)
            dt = F64 if c.dtype == torch.float64 else F32
            ctx.check(lib().femx_part_gather(self.h, dt, _vp(c), _vp(o), _stream(stream)))
            xyz.append(o)
        self.mesh = Mesh(mesh.dim, self.conn, tuple(xyz))

    def pattern(self, nd=1, stream=None):
        return Pattern(self.ctx, self.mesh, nd=nd, row_begin=self.row_begin, row_end=self.row_end, stream=stream)

    def global_csr(self, pattern, stream=None):
        """row_ptr (int64, relative to the rank's first row) and GLOBAL column ids of the rank's rows."""
        import torch
        dev = self.conn.device
        rp = torch.empty(pattern.n_rows + 1, dtype=torch.int64, device=dev)
        col = torch.empty(pattern.nnz, dtype=torch.int32, device=dev)
        self.ctx.check(lib().femx_pattern_export_csr_mapped(pattern.h, _vp(self.l2g), _vp(rp), _vp(None), _vp(col), _stream(stream)))
        return rp, col

    def close(self):
        if self.h:
            lib().femx_part_destroy.argtypes = [C.c_void_p]
            lib().femx_part_destroy(self.h)
            self.h = C.c_void_p()


DIST_ID_BYTES = 128


def dist_slab(n_planes, world, rank):
    """Owned node planes [r0, r1) and slab planes [lo, hi] of `rank` (femx_dist_slab)."""
    a, b, c, d = (C.c_int64() for _ in range(4))
    st = lib().femx_dist_slab(_i64(n_planes), int(world), int(rank), C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    if st:
        raise FemxError(st, lib().femx_last_error(None).decode())
    return a.value, b.value, c.value, d.value


def dist_unique_id():
    """NCCL unique id (bytes) — rank 0 creates it, the launcher's channel carries it to the other ranks."""
    buf = (C.c_ubyte * DIST_ID_BYTES)()
    st = lib().femx_dist_unique_id(buf)
    if st:
        raise FemxError(st, lib().femx_last_error(None).decode())
    return bytes(buf)


class Dist:
    """One rank of the multi-GPU layer (femx_dist_*): NCCL communicator + halo stream, in C++ behind the ABI."""

    def __init__(self, ctx, rank=0, world=1, unique_id=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = C.c_void_p()
        idbuf = None
        if unique_id is not None:
            idbuf = (C.c_ubyte * DIST_ID_BYTES).from_buffer_copy(bytes(unique_id))
        ctx.check(lib().femx_dist_create(ctx.h, int(rank), int(world), idbuf, C.byref(self.h)))

    @staticmethod
    def from_torch(ctx):
        """Bootstraps from an initialised torch.distributed group: rank 0's id is broadcast over it."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return Dist(ctx)
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", ctx.device)
        t = torch.zeros(DIST_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return Dist(ctx, rank, world, bytes(t.cpu().numpy().tobytes()))

    @property
    def p2p_reduction(self):
        """True when the CG reduction runs over NVLink peer memory (femx_dist_info)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.ctx.check(lib().femx_dist_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return bool(c.value)

    def allreduce(self, t, op="sum", stream=None):
        """in-place on a float64 device tensor"""
        self.ctx.check(lib().femx_dist_allreduce(self.h, _vp(t), int(t.numel()), 1 if op == "max" else 0, _stream(stream)))
        return t

    def operator(self, pattern, values):
        return DistOp(self, pattern, values)

    def close(self):
        if self.h:
            lib().femx_dist_destroy.argtypes = [C.c_void_p]
            lib().femx_dist_destroy(self.h)
            self.h = C.c_void_p()


class DistOp:
    """This rank's rows of the assembled operator: halo-exchanging SpMV and CG (femx_dist_spmv / femx_dist_cg)."""

    def __init__(self, dist, pattern, values):
        import torch
        self.dist, self.pat, self.vals = dist, pattern, values
        self.dt = F64 if values.dtype == torch.float64 else F32
        self.h = C.c_void_p()
        dist.ctx.check(lib().femx_dist_op_create(dist.h, pattern.h, self.dt, _vp(values), C.byref(self.h)))
        v = [C.c_int64() for _ in range(5)]
        dist.ctx.check(lib().femx_dist_op_info(self.h, *[C.byref(x) for x in v]))
        self.n_owned, self.ghost_lo, self.ghost_hi, self.interior_lo, self.interior_hi = (x.value for x in v)
        ph = C.c_int()
        dist.ctx.check(lib().femx_dist_op_peer_halo(self.h, C.byref(ph)))
        self.peer_halo = bool(ph.value)   # the CG's halo goes through NVLink peer memory (no NCCL call in the iteration)

    def spmv(self, x, y=None, stream=None):
        import torch
        if y is None:
            y = torch.empty_like(x)
        self.dist.ctx.check(lib().femx_dist_spmv(self.h, _vp(x), _vp(y), _stream(stream)))
        return y

    def cg(self, b, iters, x=None, stream=None):
        """x (owned part), residual norms (numpy, iters+1), device milliseconds of the solve"""
        import numpy as np
        import torch
        if x is None:
            x = torch.empty_like(b)
        res = (C.c_double * (iters + 1))()
        ms = C.c_float()
        self.dist.ctx.check(lib().femx_dist_cg(self.h, _vp(b), _vp(x), int(iters), res, C.byref(ms), _stream(stream)))
        return x, np.array(res[:]), ms.value

    def close(self):
        if self.h:
            lib().femx_dist_op_destroy.argtypes = [C.c_void_p]
            lib().femx_dist_op_destroy(self.h)
            self.h = C.c_void_p()


def lattice_prefix(cells, strides, node0, weights, nodes):
    """Closed-form prefix sums over the nodes of a lattice (femx_lattice_prefix; host only): numpy int64 array."""
    import numpy as np
    dim = len(cells)
    c = (C.c_int32 * dim)(*[int(v) for v in cells])
    s_ = (C.c_int64 * dim)(*[int(v) for v in strides])
    w = (C.c_int32 * 27)(*[int(v) for v in weights])
    q = np.ascontiguousarray(nodes, np.int64)
    out = np.empty(len(q), np.int64)
    st = lib().femx_lattice_prefix(dim, c, s_, _i64(node0), w, _i64(len(q)), q.ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.c_void_p))
    if st:
        raise FemxError(st, lib().femx_last_error(None).decode())
    return out


def read_gmsh(path):
    """Gmsh MSH 2.x ASCII → (dim, X, Y, Z, conn[NE, nn]) as numpy arrays (host)."""
    import numpy as np
    L = lib()
    dim, nn_, ne = C.c_int(), C.c_int64(), C.c_int64()
    px, py, pz = (C.POINTER(C.c_double)() for _ in range(3))
    pc = C.POINTER(C.c_int32)()
    st = L.femx_io_read_gmsh(str(path).encode(), C.byref(dim), C.byref(nn_), C.byref(ne), C.byref(px), C.byref(py),
                             C.byref(pz), C.byref(pc))
    if st:
        raise FemxError(st, L.femx_last_error(None).decode())
    n, e, k = nn_.value, ne.value, dim.value + 1
    try:
        X, Y, Z = (np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy() for p in (px, py, pz))
        conn = np.ctypeslib.as_array(pc, shape=(max(e * k, 1),))[: e * k].copy().reshape(e, k)
    finally:
        L.femx_io_free.argtypes = [C.c_void_p]
        for p in (px, py, pz, pc):
            L.femx_io_free(C.cast(p, C.c_void_p))
    return dim.value, X, Y, Z, conn


def write_matrix_market(path, row_ptr, col_idx, values, n_cols=None):
    """Host CSR (numpy / CPU tensors) → Matrix Market coordinate file."""
    import numpy as np
    rp = np.ascontiguousarray(np.asarray(row_ptr), np.int64)
    ci = np.ascontiguousarray(np.asarray(col_idx), np.int32)
    v = np.ascontiguousarray(np.asarray(values), np.float64)
    n = len(rp) - 1
    st = lib().femx_io_write_matrix_market(str(path).encode(), C.c_int64(n), C.c_int64(n if n_cols is None else n_cols),
                                           rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                           v.ctypes.data_as(C.c_void_p))
    if st:
        raise FemxError(st, lib().femx_last_error(None).decode())
