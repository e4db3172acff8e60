"""CPU: the C-ABI shared library builds, loads and exports exactly the symbols include/fgp.h declares, and the ctypes
binding (the stand-in for the Rust `extern "C"` block) names the same set.  No compute call is made here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fgp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from friedrich_b200 import build
    return build.build()


def test_header_declares_the_reference_seam():
    syms = _header_symbols()
    for s in ("fgp_fit", "fgp_refit", "fgp_add_samples", "fgp_predict_mean", "fgp_predict_var", "fgp_predict_mean_var",
              "fgp_predict_cov", "fgp_likelihood", "fgp_lml_gradient", "fgp_mean_pair_distance", "fgp_download_factor"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib], text=True)
    exported = set(re.findall(r"\bT (fgp_[a-z0-9_]+)", out))
    missing = [s for s in _header_symbols() if s not in exported]
    assert not missing, f"declared in include/fgp.h but not exported: {missing}"
    # nothing but the C-ABI leaks out of the library
    leaked = [l.split()[-1] for l in out.splitlines() if " T " in l and not l.split()[-1].startswith("fgp_")]
    assert not leaked, leaked


def test_ctypes_binding_matches_header(built_lib):
    from friedrich_b200 import _native as N
    assert sorted(N.SIGNATURES) == _header_symbols()
    lib = N.lib()  # resolves every symbol or raises
    assert b"sm_100a" in lib.fgp_version()


def test_no_cpu_fallback_without_device(built_lib):
    """Without a CUDA device fgp_create must fail loudly (this container has no GPU)."""
    import ctypes as C
    from friedrich_b200 import _native as N
    probe = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True) if os.path.exists("/usr/bin/nvidia-smi") else None
    if probe is not None and probe.returncode == 0 and "GPU" in probe.stdout:
        pytest.skip("a GPU is present")
    with pytest.raises(N.FgpError):
        N.Handle(0)
    h = C.c_void_p()
    assert N.lib().fgp_create(0, C.byref(h)) == N.FGP_ERR_CUDA
    assert not h.value


def test_library_does_not_link_blas_or_torch(built_lib):
    out = subprocess.check_output(["ldd", built_lib], text=True)
    for banned in ("cublas", "cusolver", "torch", "oracle"):
        assert banned not in out.lower(), out


def test_kernel_desc_layout_matches_header():
    import ctypes as C
    from friedrich_b200 import _native as N
    from oracle import oracle as O
    assert C.sizeof(N.KernelDesc) == 4 + 4 * 15 + 8 * 24 == C.sizeof(O.KernelDesc)
    assert N.KernelDesc.param.offset == 64


def test_branch_free_exp_is_accurate_to_two_ulp():
    """csrc/kernel_eval.cuh exp_nonpos (host twin through fgp_dbg_exp): the exp every device kernel evaluation uses."""
    import math
    import numpy as np
    from friedrich_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(7)
    xs = np.concatenate([-rng.random(20000) * 50.0, -rng.random(5000) * 708.0, -np.logspace(-300, 0, 2000),
                         [0.0, -0.0, -0.5 * math.log(2.0), -math.log(2.0), -708.0, -1e-320]])
    worst = 0.0
    for x in xs:
        got, ref = lib.fgp_dbg_exp(float(x)), math.exp(float(x))
        worst = max(worst, abs(got - ref) / np.spacing(ref))
    assert worst <= 2.0, worst
    assert lib.fgp_dbg_exp(-709.0) == 0.0 and lib.fgp_dbg_exp(-float("inf")) == 0.0
    assert math.isnan(lib.fgp_dbg_exp(float("nan")))


def test_table_assisted_exp_is_accurate_to_two_ulp():
    """csrc/kernel_eval.cuh exp_nonpos_tab (host twin through fgp_dbg_exp_tab): the exp of the Gram / cross-covariance
    interior tiles (2^(j/256) table + degree-4 polynomial, integer-side clamp)."""
    import math
    import numpy as np
    from friedrich_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(11)
    ln2_256 = math.log(2.0) / 256.0
    xs = np.concatenate([-rng.random(30000) * 50.0, -rng.random(8000) * 599.0, -np.logspace(-300, 0, 2000),
                         # reduction boundaries: x near (k + 1/2) ln2/256, where |r| is largest
                         -(rng.integers(0, 150000, 4000) + 0.5) * ln2_256 * (1.0 + 1e-15 * rng.standard_normal(4000)),
                         [0.0, -0.0, -0.5 * math.log(2.0), -math.log(2.0), -599.999, -1e-320]])
    worst = 0.0
    for x in xs:
        got, ref = lib.fgp_dbg_exp_tab(float(x)), math.exp(float(x))
        worst = max(worst, abs(got - ref) / np.spacing(ref))
    assert worst <= 1.5, worst
    assert lib.fgp_dbg_exp_tab(-600.0) == 0.0 and lib.fgp_dbg_exp_tab(-1e9) == 0.0
    assert lib.fgp_dbg_exp_tab(-float("inf")) == 0.0
    assert math.isnan(lib.fgp_dbg_exp_tab(float("nan")))


def test_gemm_cta_shape_rule():
    """Host rule of csrc/gemm_nt.cu: 32-row CTAs exactly where they lower the heaviest SM's load on a 148-SM part —
    i = 2 * tiles 64-row items: i <= 74 (0.5 vs 1 unit per SM) or 148 < i <= 222 (1.5 vs 2)."""
    from friedrich_b200 import _native as N
    lib = N.lib()
    rows = lambda M, Nn, lower=0: lib.fgp_dbg_gemm_cta_rows(M, Nn, lower, 148)
    assert rows(4096, 128) == 32          # panel solve at m = 4096: 32 tiles -> 64 items
    assert rows(4736, 128) == 32          # 37 tiles -> 74 items: last size of the first window
    assert rows(4864, 128) == 64          # 38 tiles -> 76 items: one 64-row CTA per SM is as balanced
    assert rows(9472, 128) == 64          # 74 tiles -> 148 items
    assert rows(9600, 128) == 32          # 75 tiles -> 150 items: second window (1.5 vs 2 units)
    assert rows(14208, 128) == 32         # 111 tiles -> 222 items
    assert rows(14336, 128) == 64         # 112 tiles -> 224 items
    assert rows(16384, 128) == 64
    assert rows(1024, 128) == 32 and rows(128, 128) == 32     # the solve steps of predict, single tiles
    assert rows(16384, 16384, 1) == 64    # trailing updates: the throughput shape
    assert rows(1024, 1024, 1) == 32      # 36 tiles of a small triangle
