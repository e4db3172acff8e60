"""Check and time the tcgen05 exact-integer trailing update (csrc/ozaki.cu) against numpy and the f64 DMMA kernel.
Usage: python tools/ozaki_check.py [check|bench|all]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402

lib = N.lib()
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def run(M, K, lower, row_skip=0, tpc=0, lbo=0, sbo=0, seed=1, exact=False, scale_rows=False):
    rng = np.random.default_rng(seed)
    if exact:
        A = rng.integers(-1000, 1001, size=(M, K)).astype(np.float64) / 1024.0
        Cm = rng.integers(-1000, 1001, size=(M, M)).astype(np.float64)
    else:
        A = rng.standard_normal((M, K))
        Cm = rng.standard_normal((M, M)) * 10
    if scale_rows:
        A *= np.exp2(rng.integers(-40, 40, size=(M, 1)).astype(np.float64))
    Af = np.asfortranarray(A)
    Cf = np.asfortranarray(Cm)
    ref = Cm - (A.astype(np.longdouble) @ A.T.astype(np.longdouble)).astype(np.float64) if M <= 1024 else Cm - A @ A.T
    rc = lib.fgp_dbg_ozaki_syrk(0, ptr(Cf), M, ptr(Af), M, M, K, lower, row_skip, tpc, lbo, sbo)
    got = np.array(Cf)
    if lower:
        T = M // 128
        mask = np.zeros((M, M), bool)
        for ti in range(T):
            for tj in range(T):
                if tj <= ti and ti >= row_skip:
                    blk = np.ones((128, 128), bool)
                    if ti == tj:
                        blk = np.tril(blk)
                    mask[ti * 128:(ti + 1) * 128, tj * 128:(tj + 1) * 128] = blk
        untouched = np.array_equal(got[~mask], Cm[~mask])
    else:
        mask = np.ones((M, M), bool)
        untouched = True
    scale = np.abs(A).max(axis=1)
    # error budget per element: the product to ~2^-53 of rowmax * colmax * K, plus the two roundings of the additions into C
    bound = np.outer(scale, scale) * K + 4.0 * np.abs(ref)
    err = np.abs(got - ref)[mask]
    rel = (np.abs(got - ref) / np.maximum(bound, 1e-300))[mask].max()
    return {"M": M, "K": K, "lower": lower, "row_skip": row_skip, "tpc": tpc, "lbo": lbo, "sbo": sbo, "rc": rc, "exact_inputs": exact,
            "scaled_rows": scale_rows, "max_abs_err": float(err.max()), "max_err_over_budget": float(rel),
            "untouched_ok": bool(untouched), "bit_equal": bool(np.array_equal(got[mask], ref[mask]))}


def check():
    ok = True
    first = run(256, 128, 0, exact=True)
    print(json.dumps(first), flush=True)
    if first["rc"] != 0:
        return False
    if first["max_err_over_budget"] > 1e-14:
        alt = run(256, 128, 0, exact=True, lbo=128, sbo=2048)
        print(json.dumps({"alt_descriptor": alt}), flush=True)
        return False
    for (M, K, lower, skip, tpc, exact, sr) in [(256, 512, 0, 0, 0, True, False), (512, 512, 1, 0, 1, False, False),
                                                (1024, 512, 1, 0, 3, False, True), (1024, 384, 1, 4, 2, False, False),
                                                (2048, 512, 1, 0, 0, False, True), (4096, 256, 1, 4, 4, False, False)]:
        r = run(M, K, lower, skip, tpc, exact=exact, scale_rows=sr)
        print(json.dumps(r), flush=True)
        ok &= r["rc"] == 0 and r["untouched_ok"] and r["max_err_over_budget"] < 4e-16
    return ok


def bench():
    for (M, K) in [(15872, 512), (8192, 512), (4096, 512), (2048, 512), (32256, 512), (15872, 256)]:
        if M == 32256 and len(sys.argv) > 2:
            continue
        for tpc in (1, 2, 3, 4, 6):
            mu, msl = C.c_double(0), C.c_double(0)
            rc = lib.fgp_dbg_ozaki_bench(0, M, K, 5, tpc, C.byref(mu), C.byref(msl))
            tiles = (M // 128) * (M // 128 + 1) // 2
            fl = 2.0 * 128 * 128 * tiles * K
            md, fd = C.c_double(0), C.c_double(0)
            lib.fgp_dbg_gemm_bench(0, M, M, K, 1, 1, 3, C.byref(md), C.byref(fd))
            print(json.dumps({"M": M, "K": K, "tpc": tpc, "rc": rc, "update_ms": mu.value, "slice_ms": msl.value,
                              "f64_equiv_tflops": fl / mu.value * 1e-9 if mu.value else None,
                              "int8_tops": 36 * fl / mu.value * 1e-9 if mu.value else None, "dmma_ms": md.value,
                              "speedup_vs_dmma": md.value / mu.value if mu.value else None}), flush=True)


def experiments():
    """which part bounds the update kernel: the same launch with parts switched off"""
    base = 32 if (len(sys.argv) > 2 and sys.argv[2] == "pairs") else 0
    for flags, what in [(0, "production"), (1, "no epilogue"), (2, "no operand copies"), (3, "MMAs only"), (4, "no MMAs"),
                        (5, "copies only"), (6, "epilogue only")]:
        lib.fgp_dbg_ozaki_experiment(flags | base)
        mu = C.c_double(0)
        rc = lib.fgp_dbg_ozaki_bench(0, 15872, 512, 3, 0 if base else 2, C.byref(mu), None)
        print(json.dumps({"experiment": what, "pairs": bool(base), "flags": flags, "rc": rc, "update_ms": mu.value, "us_per_tile": mu.value * 148 / 7750 * 1e3}), flush=True)
    lib.fgp_dbg_ozaki_experiment(0)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    good = True
    if what == "pairs":   # the CTA-pair kernel (cta_group::2): same checks, then the timings
        lib.fgp_dbg_ozaki_experiment(32)
        what = "all"
    if what in ("check", "all"):
        good = check()
        print("CHECK", "OK" if good else "FAILED", flush=True)
    if what in ("bench", "all") and good:
        bench()
    if what == "exp":
        experiments()
