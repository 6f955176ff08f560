"""CPU tier: the C-ABI library loads, exports every symbol of include/femx.h, the
emitter + NVRTC produce sm_100a cubins offline, and errors are reported, not fatal."""
import os
import re
import subprocess

import pytest

import femx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "femx.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(femx_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported():
    syms = header_symbols()
    assert len(syms) >= 25
    L = femx.lib()
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(femx.SYMBOLS) == syms


def test_no_libcuda_link_dependency():
    out = subprocess.check_output(["ldd", femx.LIB_PATH]).decode()
    assert "libcuda.so" not in out
    assert "libnvrtc" in out


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under cuda-fem_b200/ or include/ names it."""
    for base in ("cuda-fem_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert not re.search(r"import oracle|from oracle|femx_oracle|orc_[a-z]", txt), fn


def test_context_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(femx.FemxError) as ei:
        femx.Context(0)
    assert ei.value.status == 2 and "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("dim,builtin,nd", [(2, femx.POISSON, 1), (2, femx.POISSON_MASS, 1), (2, femx.MASS, 1),
                                           (3, femx.POISSON, 1), (3, femx.POISSON_MASS, 1),
                                           (2, femx.ELASTICITY, 2), (3, femx.ELASTICITY, 3)])
@pytest.mark.parametrize("dtype", [femx.F64, femx.F32])
def test_offline_jit_all_builtin_forms(dim, builtin, nd, dtype):
    f = femx.Form(None, dim, builtin, nd=nd, dtype=dtype, params=(0.6, 0.4), offline=True)
    for k in ("coo", "csr", "csr_x", "csr_s") + (("coo_e",) if nd == 1 else ()):
        cb = f.cubin(k)
        assert cb[:4] == b"\x7fELF"
    n = (dim + 1) * nd
    assert all(f.entry(i, j) for i in range(n) for j in range(n))
    assert f.entry(n, 0) is None
    assert "jac" in f.prologue and "femx_c" in f.source
    f.close()


def test_emitter_matches_reference_orientation():
    # entry (li, lj) = a(u=phi_lj, v=phi_li)*jac; Poisson is symmetric in the names
    f = femx.Form(None, 2, femx.POISSON, offline=True)
    # d_a = jac * grad(phi_a); kq = (sum of weights) / jac  →  (d_lj . d_li) * kq
    # (every multiply-add spelled out: the dot product is a chain of fma)
    assert f.entry(0, 1) == "femx_mul(fma(d2y,d1y,femx_mul(d2x,d1x)),kq)"
    assert f.entry(2, 0) == "femx_mul(fma(d1y,d3y,femx_mul(d1x,d3x)),kq)"
    assert "kq = femx_mul(real(0.50000001" in f.prologue  # the reference's 8-digit weights, summed
    f.close()


def test_custom_strings_reference_integrand_compiles(golden_dir):
    import json
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    f = femx.Form(None, 2, entries=j["integrand"], offline=True)  # powf(x,2.0) accepted
    assert f.cubin("csr")[:4] == b"\x7fELF"
    f.close()


def test_bad_integrand_reports_nvrtc_log():
    bad = ["x1+"] + ["0.0"] * 8
    with pytest.raises(femx.FemxError) as ei:
        femx.Form(None, 2, entries=bad, offline=True)
    assert ei.value.status == 3
    assert "error" in str(ei.value).lower() and "femx_coo" in str(ei.value)


def test_invalid_descriptors():
    with pytest.raises(femx.FemxError) as ei:
        femx.Form(None, 2, femx.ELASTICITY, nd=3, offline=True)
    assert ei.value.status == 1
    with pytest.raises(femx.FemxError) as ei:
        femx.Form(None, 2, femx.POISSON, nd=2, offline=True)
    assert ei.value.status == 1
    with pytest.raises(femx.FemxError):
        femx.Form(None, 2, rule=([0.5], None, None), offline=True)


def test_sass_uses_fp64_pipe_and_no_local_memory():
    f = femx.Form(None, 2, femx.POISSON, offline=True)
    path = "/tmp/femx_test_csr.cubin"
    open(path, "wb").write(f.cubin("csr"))
    res = subprocess.check_output(["cuobjdump", "-res-usage", path]).decode()
    assert "LOCAL:0" in res
    assert "sm_100a" in subprocess.check_output(["cuobjdump", "-lelf", path]).decode()
    sass = subprocess.check_output(["cuobjdump", "-sass", path]).decode()
    assert "DFMA" in sass and "STS" in sass
    f.close()


def test_offline_jit_rhs_kernel(golden_dir):
    import json
    j = json.load(open(os.path.join(golden_dir, "ref_integrand_strings.json")))
    f = femx.Form(None, 2, entries=j["integrand"], rhs=j["rhs"], offline=True)
    assert f.cubin("rhs")[:4] == b"\x7fELF" and "femx_rhs" in f.source
    f.close()
    f = femx.Form(None, 3, femx.ELASTICITY, nd=3, params=(1.0, 0.5), rhs_vec=(0.0, 0.0, -9.81), offline=True)
    assert f.cubin("rhs")[:4] == b"\x7fELF"
    f.close()


def test_cxx_clients_build_and_link():
    """examples/: plain C++ clients of the C ABI (the reference's main(), the weak-form front end, the multi-GPU CG) compile
    against include/femx.h and link against libfemx.so only; without arguments the multi-GPU client prints its usage."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "examples")])
    for exe in ("femx_sparse2", "femx_weakform_demo", "femx_dist_cg"):
        assert os.access(os.path.join(root, "examples", exe), os.X_OK), exe
    r = subprocess.run([os.path.join(root, "examples", "femx_dist_cg")], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
