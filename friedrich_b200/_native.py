"""ctypes binding of the C-ABI in include/fgp.h (libfgp_sm100.so).

This is the exact surface a Rust `extern "C"` block would bind; the Python host layer in
`friedrich_b200.gp` calls nothing else.  There is no CPU fallback: if the shared library is missing or no CUDA device
is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfgp_sm100.so")

MAX_OPS = 15
MAX_PARAMS = 24

FGP_OPT_LOOKAHEAD = 1
FGP_OPT_HEAD = 2
FGP_OPT_TCGEN05 = 3
FGP_OPT_SHARD_PIPE = 4
FGP_COMM_ID_BYTES = 128
FGP_OK, FGP_ERR_NOT_POSDEF, FGP_ERR_BAD_ARG, FGP_ERR_BAD_KERNEL, FGP_ERR_CUDA, FGP_ERR_NOT_FITTED, FGP_ERR_COMM = range(7)


class KernelDesc(C.Structure):
    """Binary twin of `fgp_kernel_desc` (include/fgp_kernel_desc.h)."""
    _fields_ = [("n_ops", C.c_int32), ("op", C.c_int32 * MAX_OPS), ("param", C.c_double * MAX_PARAMS)]

    @classmethod
    def make(cls, ops, params):
        if len(ops) > MAX_OPS or len(params) > MAX_PARAMS:
            raise ValueError("kernel expression too large for fgp_kernel_desc")
        k = cls()
        k.n_ops = len(ops)
        for i, o in enumerate(ops):
            k.op[i] = int(o)
        for i, p in enumerate(params):
            k.param[i] = float(p)
        return k


class FgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfgp_sm100 error {code}: {msg}")
        self.code = code


class NotPositiveDefinite(FgpError, ArithmeticError):
    pass


_dp = C.POINTER(C.c_double)
_i64 = C.c_int64
_h = C.c_void_p
_kd = C.POINTER(KernelDesc)

# name -> (restype, argtypes); mirrors include/fgp.h one to one (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "fgp_create": (C.c_int, [C.c_int, C.POINTER(_h)]),
    "fgp_destroy": (C.c_int, [_h]),
    "fgp_last_error": (C.c_char_p, [_h]),
    "fgp_version": (C.c_char_p, []),
    "fgp_set_inputs": (C.c_int, [_h, _dp, _i64, _i64, _i64]),
    "fgp_fit": (C.c_int, [_h, _dp, _i64, _i64, _i64, _dp, _kd, C.c_double, C.c_int, C.c_double]),
    "fgp_refit": (C.c_int, [_h, _kd, C.c_double, C.c_int, C.c_double]),
    "fgp_set_outputs": (C.c_int, [_h, _dp, _i64]),
    "fgp_add_samples": (C.c_int, [_h, _dp, _i64, _i64, _dp, _kd, C.c_double, C.c_int, C.c_double]),
    "fgp_failed_column": (_i64, [_h]),
    "fgp_num_samples": (_i64, [_h]),
    "fgp_num_dims": (_i64, [_h]),
    "fgp_predict_mean": (C.c_int, [_h, _kd, _dp, _i64, _i64, _dp]),
    "fgp_predict_var": (C.c_int, [_h, _kd, _dp, _i64, _i64, _dp]),
    "fgp_predict_mean_var": (C.c_int, [_h, _kd, _dp, _i64, _i64, _dp, _dp]),
    "fgp_predict_cov": (C.c_int, [_h, _kd, _dp, _i64, _i64, C.c_int, _dp, _i64, _dp]),
    "fgp_likelihood": (C.c_int, [_h, _kd, C.c_double, _dp]),
    "fgp_lml_gradient": (C.c_int, [_h, _kd, C.c_double, C.c_int, _dp, _dp]),
    "fgp_mean_pair_distance": (C.c_int, [_h, _dp]),
    "fgp_download_factor": (C.c_int, [_h, _dp, _i64]),
    "fgp_download_alpha": (C.c_int, [_h, _dp]),
    "fgp_factor_digest": (C.c_int, [_h, _dp]),
    "fgp_upload_state": (C.c_int, [_h, _dp, _i64, _i64, _i64, _dp, _dp, _i64]),
    "fgp_inverse_columns": (C.c_int, [_h, C.POINTER(_i64), _i64, _dp, _i64]),
    "fgp_last_device_ms": (C.c_double, [_h]),
    "fgp_last_launch_count": (_i64, [_h]),
    "fgp_set_profiling": (C.c_int, [_h, C.c_int]),
    "fgp_set_option": (C.c_int, [_h, C.c_int, _i64]),
    "fgp_profile_summary": (C.c_int, [_h, _dp, _dp, C.POINTER(_i64)]),
    "fgp_stage_queries": (C.c_int, [_h, _dp, _i64, _i64]),
    "fgp_predict_staged": (C.c_int, [_h, _kd, C.c_int, C.c_int]),
    "fgp_fetch_predictions": (C.c_int, [_h, _dp, _dp]),
    "fgp_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "fgp_comm_init_rank": (C.c_int, [_h, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "fgp_comm_destroy": (C.c_int, [_h]),
    "fgp_fit_sharded": (C.c_int, [_h, _dp, _i64, _i64, _i64, _dp, _kd, C.c_double, C.c_int, C.c_double]),
    "fgp_refit_sharded": (C.c_int, [_h, _kd, C.c_double, C.c_int, C.c_double]),
    "fgp_lml_gradient_sharded": (C.c_int, [_h, _kd, C.c_double, C.c_int, _dp, _dp]),
    "fgp_shard_plan": (C.c_int, [_i64, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), _dp]),
    "fgp_comm_last_bytes": (C.c_double, [_h]),
    "fgp_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "fgp_free_pinned": (None, [C.c_void_p]),
    "fgp_cholesky_lower": (C.c_int, [C.c_int, _dp, _i64, _i64, C.POINTER(_i64)]),
    "fgp_dbg_lower_tiles": (_i64, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _i64]),
    "fgp_dbg_lower_tiles_skip": (_i64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        _i64]),
    "fgp_dbg_potrf_head": (C.c_int, [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_double, C.POINTER(C.c_int), C.c_int, _dp]),
    "fgp_dbg_shard_pieces": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "fgp_dbg_exp": (C.c_double, [C.c_double]),
    "fgp_dbg_exp_tab": (C.c_double, [C.c_double]),
    "fgp_dbg_gemm_occupancy": (C.c_int, [C.c_int]),
    "fgp_dbg_gemm_occupancy32": (C.c_int, [C.c_int]),
    "fgp_dbg_gemm_cta_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fgp_dbg_gemm_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "fgp_dbg_gemm_nt": (C.c_int, [C.c_int, _dp, _i64, _dp, _i64, _dp, _i64, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_int, C.c_int]),
    "fgp_dbg_ozaki_syrk": (C.c_int, [C.c_int, _dp, _i64, _dp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int]),
    "fgp_linear_prior_fit": (C.c_int, [_h, _dp, _dp, _dp]),
    "fgp_dbg_ozaki_experiment": (None, [C.c_int]),
    "fgp_dbg_ozaki_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
}

_lib = None


def lib():
    """Load libfgp_sm100.so (built in-tree by `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(_dp)


def fcol(a, copy=False):
    """float64 column-major array — the layout of nalgebra's DMatrix that the ABI takes."""
    a = np.asarray(a, dtype=np.float64)
    if copy or not a.flags.f_contiguous:
        a = np.array(a, dtype=np.float64, order="F", copy=True)
    return a


class Handle:
    """Owns one `fgp_model*` (one GPU, one stream)."""

    def __init__(self, device=0):
        self._h = _h()
        rc = lib().fgp_create(int(device), C.byref(self._h))
        if rc != FGP_OK:
            raise FgpError(rc, f"fgp_create(device={device}) failed: no usable CUDA device (there is no CPU fallback)")
        self.device = int(device)

    def check(self, rc):
        if rc == FGP_OK:
            return
        msg = lib().fgp_last_error(self._h).decode("utf-8", "replace")
        if rc == FGP_ERR_NOT_POSDEF:
            raise NotPositiveDefinite(rc, msg)
        raise FgpError(rc, msg)

    @property
    def ptr(self):
        return self._h

    def close(self):
        if self._h:
            lib().fgp_destroy(self._h)
            self._h = _h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_device_ms(self):
        return float(lib().fgp_last_device_ms(self._h))

    def last_launch_count(self):
        return int(lib().fgp_last_launch_count(self._h))
