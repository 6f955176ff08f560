"""Symbolic weak-form front-end: the reference's GiNaC workflow with sympy.

The reference writes the weak form as lambdas over symbolic shape functions and lets GiNaC
differentiate them (fea_symbolic_nvrtc_sparse.cpp:50-115 sfR/sfS with registered derivatives,
:226-290 FunctionSpace, :293-362 WeakForm::build), prints every entry with csrc_float and pastes
the strings into the kernel template.  This module does the same with sympy and hands the strings
to femx.Form (NVRTC, sm_100a):

    fs = FunctionSpace(2)                                  # FunctionSpace(mesh, lst(x,y), "Lagrange", 1)
    wf = WeakForm(fs)
    wf.build(lambda u, v: dot(grad(u), grad(v)),           # wf.build([&](ex u, ex v){...},
             lambda v: (-2*(fs.x**2 + fs.y**2) + 36) * v)  #          [&](ex v){ return f*v; });
    form = wf.compile(ctx)                                 # nvrtcCompileProgram … cuModuleGetFunction

Differences from the reference: (1) common sub-expressions are pulled into a per-element
prologue by sympy.cse instead of being re-expanded in every entry (the reason the reference
kernel is instruction-bound); (2) 3-D tetrahedra and vector-valued spaces
(VectorFunctionSpace) exist; (3) the RHS strings are kept (the reference drops them, :346-351).
Entry (li, lj) = a(u = phi_lj, v = phi_li) * jac, as the reference.

sympy is needed only here; the engine itself does not depend on it.
"""
import sympy as sp

from . import CUSTOM, F64, Form

__all__ = ["FunctionSpace", "VectorFunctionSpace", "WeakForm", "grad", "dot", "div", "sym", "inner"]


def _make_ref_coord(name, index, space):
    """A reference coordinate r(x,y[,z]) as an undefined sympy function whose derivative with
    respect to the physical coordinate c is (J^-1)[index, c] — the sympy counterpart of
    REGISTER_FUNCTION(sfR, ... derivative_func(sfR_deriv)) (fea_symbolic_nvrtc_sparse.cpp:55-101)."""

    class RefCoord(sp.Function):
        nargs = space.dim

        def fdiff(self, argindex=1):
            return space.Jinv[index, argindex - 1]

    RefCoord.__name__ = name
    return RefCoord


class FunctionSpace:
    """P1 Lagrange on simplices: shape functions (r, s, 1-r-s) in 2-D (:264-269), (r, s, t, 1-r-s-t) in 3-D."""

    nd = 1

    def __init__(self, dim, family="Lagrange", order=1):
        if family != "Lagrange" or order != 1 or dim not in (2, 3):
            raise ValueError("only P1 Lagrange simplices (dim 2 or 3) are supported")
        self.dim, self.nn = dim, dim + 1
        names = "xyz"[:dim]
        self.coords = sp.symbols(" ".join(names), real=True)
        self.x, self.y = self.coords[0], self.coords[1]
        self.z = self.coords[2] if dim == 3 else None
        self.ref = sp.symbols("r s t"[: 2 * dim - 1], real=True)
        self.nodes = [sp.symbols(" ".join(f"{c}{k + 1}" for k in range(self.nn)), real=True) for c in names]
        bary = list(self.ref) + [1 - sum(self.ref)]
        # affine map fx = x1 r + x2 s + x3 (1-r-s)  (:259-261)
        self.trans = [sum(self.nodes[c][k] * bary[k] for k in range(self.nn)) for c in range(dim)]
        J = sp.Matrix(dim, dim, lambda c, a: sp.diff(self.trans[c], self.ref[a]))
        self.jac = sp.factor(J.det())                              # getJac() (:281-289)
        self.Jinv = (J.adjugate() / self.jac).applyfunc(sp.simplify)
        self._funcs = [_make_ref_coord("sf" + "RST"[a], a, self)(*self.coords) for a in range(dim)]
        self.shape = list(self._funcs) + [1 - sum(self._funcs)]    # getShapeFunctions()

    def basis(self):
        """Basis functions in dof order (li = node)."""
        return list(self.shape)

    def to_reference(self, expr):
        """subs(sfr==r, sfs==s) then subs(x==fx, y==fy)  (:337)."""
        expr = sp.sympify(expr)
        expr = expr.subs({f: r for f, r in zip(self._funcs, self.ref)})
        return expr.subs({c: t for c, t in zip(self.coords, self.trans)})


class VectorFunctionSpace(FunctionSpace):
    """dim-vector-valued P1 space: dof li = nd*node + component, basis phi_node * e_component."""

    def __init__(self, dim):
        super().__init__(dim)
        self.nd = dim

    def basis(self):
        out = []
        for a in range(self.nn):
            for c in range(self.nd):
                v = sp.zeros(self.dim, 1)
                v[c] = self.shape[a]
                out.append(v)
        return out


_space = None


def grad(f):
    """Gradient with respect to the physical coordinates (scalar → column vector, vector → Jacobian matrix)."""
    cs = _space.coords
    if isinstance(f, sp.MatrixBase):
        return sp.Matrix(f.shape[0], len(cs), lambda i, j: sp.diff(f[i], cs[j]))
    return sp.Matrix([sp.diff(f, c) for c in cs])


def dot(a, b):
    return sum(x * y for x, y in zip(list(a), list(b)))


def div(u):
    return sum(sp.diff(u[i], _space.coords[i]) for i in range(len(_space.coords)))


def sym(m):
    return (m + m.T) / 2


def inner(a, b):
    return sum(x * y for x, y in zip(list(a), list(b)))


class WeakForm:
    """WeakForm::build (:307-356): entries lhs[j][i] = a(phi_j, phi_i)*jac as C strings."""

    def __init__(self, space):
        self.space = space
        self.entries = None
        self.rhs = None
        self.prologue = ""

    def build(self, lhs, rhs=None, cse=True):
        global _space
        fs = self.space
        _space = fs
        try:
            basis = fs.basis()
            n = len(basis)
            exprs = []
            for li in range(n):            # row  = test function  v = phi_li
                for lj in range(n):        # col  = trial function u = phi_lj
                    exprs.append(fs.to_reference(lhs(basis[lj], basis[li])) * fs.jac)
            if rhs is not None:
                for li in range(n):
                    exprs.append(fs.to_reference(rhs(basis[li])) * fs.jac)
        finally:
            _space = None
        exprs = [sp.together(e) for e in exprs]
        if cse:
            repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("w_"), optimizations="basic")
        else:
            repl, red = [], exprs
        # the prologue is evaluated once per element, before the quadrature loop: only
        # sub-expressions free of the quadrature point may live there, the others are inlined back
        ref = set(fs.ref)
        inline, keep = {}, []
        for sym_, e in repl:
            e = e.subs(inline)
            if e.free_symbols & ref:
                inline[sym_] = e
            else:
                keep.append((sym_, e))
        red = [e.subs(inline) for e in red]
        self.prologue = "".join(f"const real {s} = {self._c(e)};\n  " for s, e in keep)
        strs = [self._c(e) for e in red]
        self.entries = [strs[li * n:(li + 1) * n] for li in range(n)]
        self.rhs = strs[n * n:] if rhs is not None else None
        return self

    @staticmethod
    def _c(e):
        # csrc_float's counterpart; literals are cast through `real` so that fp32 forms stay fp32
        from sympy.printing.c import C99CodePrinter

        class P(C99CodePrinter):
            def _print_Float(self, x):
                return f"real({super()._print_Float(x)})"

            def _print_Rational(self, x):
                return f"(real({x.p}.0)/real({x.q}.0))"

            def _print_Integer(self, x):
                return f"real({x.p}.0)"

            def _print_Pow(self, x):
                b, ex = x.as_base_exp()
                if ex.is_Integer and 0 < ex <= 4:
                    return "(" + "*".join([f"({self._print(b)})"] * int(ex)) + ")"
                if ex.is_Integer and -4 <= ex < 0:
                    return "(real(1.0)/(" + "*".join([f"({self._print(b)})"] * int(-ex)) + "))"
                return super()._print_Pow(x)

        return P().doprint(e)

    def compile(self, ctx, dtype=F64, offline=False, rule=None, fmad=True):
        """NVRTC-compile the generated strings (reference semantics: integrand, weighted and summed)."""
        fs = self.space
        return Form(ctx, fs.dim, builtin=CUSTOM, nd=fs.nd, dtype=dtype, entries=self.entries,
                    prologue=self.prologue or None, rhs=self.rhs, rule=rule, fmad=fmad, offline=offline)
