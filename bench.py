#!/usr/bin/env python
"""bench.py — the hot path of friedrich on B200: GP fit (Gram + Cholesky, f64) TFLOP/s and predict queries/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload metric|c2|c4]

One "step" = one pass of the fit hot path (Gram assembly + blocked Cholesky + alpha solve) over the synthetic training
set of the workload.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is produced.

 * value          fit TFLOP/s, inputs resident in HBM (fgp_refit), device time from CUDA events on the library's stream
 * e2e            same metric through the C-ABI with HOST (pinned) buffers: fgp_fit(X, y) + fgp_predict_mean_var(Xq) with
                  the H2D of X, y, Xq and the D2H of mean/var inside the timed region
 * roofline       dominant kernel = ozaki_update_kernel (trailing updates on tcgen05: exact int8 digit products, int32 TMEM
                  accumulators, 36 integer GEMMs per f64 GEMM): int8 tensor ops of its launches / their summed CUDA-event
                  durations, against 2 x the measured bf16 tensor peak (MEASURED_PEAKS.json; the int8 rate of the tcgen05 pipe
                  is twice the bf16 rate); `dmma` = the f64 DMMA kernel (panel solves, small updates) against the fp64 peak
 * cpu_baseline   the oracle (C restatement of the reference's algorithm, 1 thread like the reference) on a bounded sample
 * --impl reference   the reference arm: the same oracle timed on the host CPU (the Rust crate cannot be built here)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP64_PEAK_TFLOPS = 37.07  # measured on this pool's B200: tools/microbench/fp64_peak.cu, profiles/fp64_peak_r01.jsonl
WORKLOADS = {
    # name: (n, d, q, kernel tag, description)
    "metric": (16384, 16, 1024, "SquaredExp", "fit: Gram+Cholesky n=16384 d=16 SquaredExp f64; predict mean+variance q=1024"),
    "c2": (4096, 8, 1024, "SquaredExp", "fit n=4096 d=8 SquaredExp f64 + predict 1024 queries"),
    "c4": (32768, 32, 1024, "SquaredExp", "fit n=32768 d=32 SquaredExp f64; predict mean+variance q=1024"),
}


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel, averaged over the launches of
# one fit: read from the committed ncu capture of the SHIPPED build (tools/ncu_launch_csv.py writes the CSV; profiles/ is the
# judged copy).
GEMM_TRAFFIC_CSV = {"metric": "profiles/tcgen05_dram_fit16k_r02.csv"}


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written): bf16 burst / sustained TFLOP/s and the HBM copy GB/s; fallback = the profiling
    recipe's figures."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        return float(mp["bf16_tflops"]), float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"])), "MEASURED_PEAKS.json"
    except Exception:
        return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def gemm_traffic_bytes_per_launch(workload, kernel="ozaki_update_kernel"):
    """(bytes per launch, launches, file) from the ncu CSV rows `kernel,dram_bytes_read,dram_bytes_write` of `kernel`, or
    (None, 0, file) when no capture of this build is committed."""
    rel = GEMM_TRAFFIC_CSV.get(workload)
    path = os.path.join(ROOT, rel) if rel else None
    if not path or not os.path.exists(path):
        return None, 0, rel
    import csv
    total, launches = 0.0, 0
    with open(path, newline="") as f:
        for row in csv.DictReader(l for l in f if not l.startswith("#")):
            if kernel not in row.get("kernel", ""):
                continue
            total += float(row["dram_bytes_read"]) + float(row["dram_bytes_write"])
            launches += 1
    return (total / launches if launches else None), launches, rel


def fit_flops(n, d):
    return n ** 3 / 3.0 + float(n) * n * d


def predict_flops(n, d, q):
    return float(n) * n * q + 2.0 * n * q * d + 4.0 * n * q  # forward TRSM + cross-covariance + mean/var reductions


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(N, shape, order="F"):
    count = int(np.prod(shape))
    ptr = N.lib().fgp_alloc_pinned(count * 8)
    if not ptr:
        raise MemoryError("fgp_alloc_pinned failed")
    buf = (C.c_double * count).from_address(ptr)
    return np.ndarray(shape, dtype=np.float64, buffer=buf, order=order), ptr


# ----------------------------------------------------------------------------------------------------------------------
# oracle timing (cpu_baseline leg and the --impl reference arm)

def oracle_step(O, kdesc, X, y, Xq):
    gp = O.OracleGaussianProcess(O.ZeroPrior(), kdesc, 0.1, None, X, y)
    if Xq is not None:
        gp.predict_mean_variance(Xq)
    return gp


def time_oracle(n, d, q, steps, warmup):
    from friedrich_b200.synthetic import make_dataset, make_inputs
    from oracle import oracle as O
    X, y = make_dataset(0x5EED0001, n, d)
    Xq = make_inputs(0x5EED0002, q, d) if q else None
    kdesc = O.KernelDesc.make([O.K_SQUARED_EXP], [math.sqrt(d / 6.0), 1.0])
    for _ in range(warmup):
        oracle_step(O, kdesc, X, y, Xq)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(O, kdesc, X, y, Xq)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    # the same ALGORITHMIC count as the GPU arm's e2e figure (the reference executes about twice the predict flops: it
    # re-solves K^-1 K_nq with both triangular factors on every call, mod.rs:235, :298 — that shows up as time, not as work)
    flops = fit_flops(n, d) + (predict_flops(n, d, q) if q else 0.0)
    return dt, flops


def cpu_port_scaling(d, sizes=(1024, 2048, 4096)):
    """Oracle fit at a few sizes: seconds, flop rate, and the exponent of t ~ n^p fitted over them (SURVEY §8d asks for the
    n^3 extrapolation to be shown, not assumed)."""
    from friedrich_b200.synthetic import make_dataset
    from oracle import oracle as O
    rows = []
    for ns in sizes:
        X, y = make_dataset(0x5EED0001, ns, d)
        kdesc = O.KernelDesc.make([O.K_SQUARED_EXP], [math.sqrt(d / 6.0), 1.0])
        t0 = time.perf_counter()
        O.OracleGaussianProcess(O.ZeroPrior(), kdesc, 0.1, None, X, y)
        dt = time.perf_counter() - t0
        rows.append({"n": ns, "seconds": dt, "gflops": fit_flops(ns, d) / dt * 1e-9})
    p = float(np.polyfit(np.log([r["n"] for r in rows]), np.log([r["seconds"] for r in rows]), 1)[0])
    return rows, p


def run_reference(args, rank, world):
    """The reference arm.  friedrich is a Rust crate and this image has no cargo/rustc, so `oracle/_ref` cannot exist;
    the arm times the oracle port (the reference's algorithm, single-threaded like the crate) on the host CPU.  `config`
    is the workload of our arm (same object); each timed step is a BOUNDED SAMPLE of it, described in `sample`."""
    if rank != 0:
        return
    n, d, q, config = workload_config(args.workload, world, world > 1)
    ns, qs = 2048, 256  # bounded sample of the workload: ~1.5 s per step
    dt, flops = time_oracle(ns, d, qs, args.steps, args.warmup)
    val = flops / dt * 1e-12
    rows, p = cpu_port_scaling(d, (1024, 2048, 4096))
    t_last, n_last = rows[-1]["seconds"], rows[-1]["n"]
    sample = (f"oracle fit n={ns} d={d} + predict_mean_variance q={qs} per step, 1 thread (the crate is single-threaded); "
              f"value = flop rate of the sample on the same algorithmic count")
    line = {
        "impl": "reference", "metric": "gp_fit_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.workload == "metric" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "sample": {"n": ns, "q": qs, "d": d, "what": sample},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": 1, "kind": "port", "sample": sample,
                         "fit_seconds_by_n": rows, "fitted_exponent": p,
                         "extrapolated_fit_seconds_full_config": t_last * (n / n_last) ** 3,
                         "extrapolation": f"t(n={n}) = t(n={n_last}) x (n/{n_last})^3; measured exponent over the sizes above: {p:.2f}"},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------

def workload_config(workload, world, use_sharded, lookahead=True):
    """The `config` object of the JSON line — built in ONE place so that both arms print the same thing."""
    n, d, q, _, desc = WORKLOADS[workload]
    if world > 1 and workload == "metric":
        n = weak_n(world)
        desc = f"fit: Gram+Cholesky n={n} d={d} SquaredExp f64 sharded over {world} GPUs (n = 16384*N^(1/3)); predict q={q}"
    return n, d, q, {"workload": desc, "n": n, "d": d, "q": q, "noise": 0.1, "kernel": "SquaredExp(ls=sqrt(d/6), ampl=1)",
                     "l2": "inputs larger than L2 (factor = %.2f GB)" % (8.0 * n * n / 1e9),
                     "multi_gpu": ("block-cyclic 512-column panels, each broadcast (NCCL) in row pieces so that solve / transfer / digit slicing / look-ahead overlap, replicated factor; queries sharded")
                     if use_sharded else "single GPU", "lookahead": lookahead}


def weak_n(world):
    """Weak scaling: per-GPU Cholesky work n^3/(3N) stays that of n=16384 on one GPU; n is rounded to a whole panel
    (512 columns).  N=1 -> 16384 (the metric's size), N=2 -> 20480, N=4 -> 26112, N=8 -> 32768 (= config C4's n)."""
    return int(round(16384.0 * world ** (1.0 / 3.0) / 512.0)) * 512


def run_ours(args, rank, local_rank, world):
    dist = None
    if world > 1:
        import torch  # before the native library: see csrc/nccl_dyn.cuh
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("gloo")  # plumbing only: NCCL id exchange, barriers, max over ranks
    from friedrich_b200 import _native as N
    from friedrich_b200 import sharded
    from friedrich_b200.kernels import SquaredExp
    from friedrich_b200.synthetic import make_dataset, make_inputs

    use_sharded = world > 1 or args.sharded
    n, d, q, config = workload_config(args.workload, world, use_sharded, not args.no_lookahead)
    lib = N.lib()
    h = N.Handle(local_rank)
    if use_sharded:
        sharded.comm_init(h, rank, world, dist)
    noise = 0.1
    Xs, ys = make_dataset(0x5EED0001, n, d)
    qr = q // world  # queries are independent: each rank predicts its slice against the replicated factor
    Xqs = make_inputs(0x5EED0002, q, d)[rank * qr:(rank + 1) * qr]
    X, pX = pinned_array(N, (n, d))
    y, py = pinned_array(N, (n,))
    Xq, pq = pinned_array(N, (qr, d))
    X[...], y[...], Xq[...] = Xs, ys, Xqs
    mean, pm = pinned_array(N, (qr,))
    var, pv = pinned_array(N, (qr,))
    kd = SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
    noise = 0.1

    def fit_host():
        if use_sharded:  # rank 0's host X, y -> H2D -> ncclBroadcast -> sharded factorisation
            h.check(lib.fgp_fit_sharded(h.ptr, N.dptr(X) if rank == 0 else None, n, n, d, N.dptr(y) if rank == 0 else None,
                                        C.byref(kd), noise, 0, 0.0))
        else:
            h.check(lib.fgp_fit(h.ptr, N.dptr(X), n, n, d, N.dptr(y), C.byref(kd), noise, 0, 0.0))

    def refit():
        if use_sharded:
            h.check(lib.fgp_refit_sharded(h.ptr, C.byref(kd), noise, 0, 0.0))
        else:
            h.check(lib.fgp_refit(h.ptr, C.byref(kd), noise, 0, 0.0))

    def predict_host():
        h.check(lib.fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(Xq), qr, qr, N.dptr(mean), N.dptr(var)))

    def predict_staged():
        h.check(lib.fgp_predict_staged(h.ptr, C.byref(kd), 1, 1))

    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib.fgp_set_profiling(h.ptr, 0)
    trailing = ("tcgen05.mma kind::i8 on exact base-128 digit slices (8 slices, 36 products, int32 TMEM accumulators) while >= 1024 "
                "rows are left below the panel, f64 DMMA below that; results are f64")
    if args.no_tcgen05:
        lib.fgp_set_option(h.ptr, N.FGP_OPT_TCGEN05, 0)
        trailing = "f64 DMMA kernel everywhere (--no-tcgen05)"
    if args.no_lookahead:
        lib.fgp_set_option(h.ptr, N.FGP_OPT_LOOKAHEAD, 0)
    fit_host()  # first touch: allocations, H2D
    for _ in range(args.warmup):
        refit()
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- timed region 1: K fit steps, inputs resident in HBM (the factor is 8 n^2 bytes >> L2, no flush needed) ----
    # per-launch profiling events are OFF here (they cost 2-10 % of a step in CPU launch overhead); the per-kernel
    # breakdown behind `roofline` comes from K more steps with profiling on, right after
    launches = 0
    barrier()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        refit()
        dev_ms += h.last_device_ms()
        launches += h.last_launch_count()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # collective steps start together (the first panel broadcast synchronises the ranks), so the slowest rank's device
    # time per step is the job's time per step; wall_ms (barrier to barrier, host clock) is reported beside it
    fit_ms = max_over_ranks(dev_ms / args.steps)
    wall_ms = max_over_ranks(wall_ms)

    # ---- profiling pass: same K steps with CUDA events around every launch -------------------------------------------
    # On one GPU the pass runs the SINGLE-STREAM schedule (look-ahead off): every launch then has the GPU to itself and its
    # event duration is the kernel's own; with look-ahead on, launches of the two streams share SMs and each one's duration
    # is inflated by the other's work (the flops and the launches are identical either way, results are bit-identical).
    pms, pfl, pcnt = np.zeros(6), np.zeros(6), np.zeros(6, dtype=np.int64)   # classes: include/fgp.h fgp_set_profiling
    tms, tfl, tcnt = np.zeros(6), np.zeros(6), np.zeros(6, dtype=np.int64)
    prof_dev_ms = 0.0
    if not args.no_profile:
        lib.fgp_set_profiling(h.ptr, 1)
        if not use_sharded:
            lib.fgp_set_option(h.ptr, N.FGP_OPT_LOOKAHEAD, 0)
        for _ in range(args.steps):
            refit()
            prof_dev_ms += h.last_device_ms()
            lib.fgp_profile_summary(h.ptr, N.dptr(pms), N.dptr(pfl), pcnt.ctypes.data_as(C.POINTER(C.c_int64)))
            tms += pms
            tfl += pfl
            tcnt += pcnt
        lib.fgp_set_profiling(h.ptr, 0)
        if not use_sharded and not args.no_lookahead:
            lib.fgp_set_option(h.ptr, N.FGP_OPT_LOOKAHEAD, 1)
        barrier()
    # the dominant launch shape alone, back to back: the first trailing update of this fit (m = n - K, K = panel width, lower)
    iso = None
    if rank == 0 and not args.no_profile:
        ms_i, fl_i = C.c_double(0), C.c_double(0)
        k_iso = 1024 if (n >= 24576 and not use_sharded) else 512  # csrc/potrf.cuh panel_tiles(): panel width = update depth
        m_iso = (n // 128) * 128 - k_iso
        if m_iso >= k_iso and lib.fgp_dbg_gemm_bench(local_rank, m_iso, m_iso, k_iso, 1, 1, 3, C.byref(ms_i), C.byref(fl_i)) == 0:
            iso = {"shape": f"C({m_iso}x{m_iso}, lower) -= A A^T, K={k_iso}", "ms": ms_i.value,
                   "tflops": fl_i.value / ms_i.value * 1e-9, "frac": fl_i.value / ms_i.value * 1e-9 / FP64_PEAK_TFLOPS}
    iso_tc = None
    if rank == 0 and not args.no_profile:
        ms_u, ms_s = C.c_double(0), C.c_double(0)
        m_tc = (n // 128) * 128 - 512   # the sharded schedule and the tcgen05 path always use 512-column panels
        if m_tc >= 2048 and lib.fgp_dbg_ozaki_bench(local_rank, m_tc, 512, 3, 0, C.byref(ms_u), C.byref(ms_s)) == 0:
            fl_tc = 2.0 * 128 * 128 * ((m_tc // 128) * (m_tc // 128 + 1) // 2) * 512
            iso_tc = {"shape": f"C({m_tc}x{m_tc}, lower) -= P P^T, K=512, P in 8 int8 digit slices", "ms": ms_u.value,
                      "int8_tops": 36 * fl_tc / ms_u.value * 1e-9, "f64_equivalent_tflops": fl_tc / ms_u.value * 1e-9,
                      "digit_slicing_ms": ms_s.value}

    # ---- predict throughput, queries resident ------------------------------------------------------------------------
    h.check(lib.fgp_stage_queries(h.ptr, N.dptr(Xq), qr, qr))
    for _ in range(args.warmup):
        predict_staged()
    pred_ms = 0.0
    for _ in range(args.steps):
        predict_staged()
        pred_ms += h.last_device_ms()
    pred_ms = max_over_ranks(pred_ms / args.steps)
    # single-query latency (mean + variance of ONE staged query; the wavefront-solve path of csrc/fgp_api.cu predict_small)
    h.check(lib.fgp_stage_queries(h.ptr, N.dptr(Xq), qr, 1))
    q1_ms = []
    for _ in range(args.warmup + args.steps):
        predict_staged()
        q1_ms.append(h.last_device_ms())
    q1_ms = float(np.median(q1_ms[args.warmup:]))
    h.check(lib.fgp_stage_queries(h.ptr, N.dptr(Xq), qr, qr))

    # ---- timed region 2: end to end through the C-ABI with host buffers -------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        fit_host()
        predict_host()
    barrier()
    t0 = time.perf_counter()
    e2e_fit_ms = 0.0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        fit_host()
        e2e_fit_ms += (time.perf_counter() - t1) * 1e3
        predict_host()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()
    bcast_mb = lib.fgp_comm_last_bytes(h.ptr) / 1e6 if use_sharded else 0.0

    # ---- parity of what was just timed (outside every timed region) -----------------------------------------------------
    def digest(handle):
        out = np.zeros(3)
        handle.check(lib.fgp_factor_digest(handle.ptr, N.dptr(out)))
        return out

    def all_ranks_equal(v):
        if dist is None:
            return True
        import torch
        lo, hi = torch.tensor(v), torch.tensor(v)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))

    refit()
    dg = digest(h)
    parity = {"factor_digest": [float(v) for v in dg], "digest_equal_across_ranks": all_ranks_equal(dg)}
    # K alpha = y on probe rows: rows of K recomputed on the host with numpy (independent of the library and of the oracle)
    alpha = np.zeros(n)
    h.check(lib.fgp_download_alpha(h.ptr, N.dptr(alpha)))
    idx = np.random.default_rng(1234).choice(n, 32, replace=False)
    ls2 = d / 6.0
    d2 = ((Xs[idx] ** 2).sum(1)[:, None] + (Xs ** 2).sum(1)[None, :] - 2.0 * Xs[idx] @ Xs.T).clip(min=0.0)
    Krows = np.exp(-d2 / (2.0 * ls2))
    Krows[np.arange(32), idx] = 1.0 + noise ** 2
    parity["solve_residual_max"] = float(np.abs(Krows @ alpha - ys[idx]).max())
    parity["solve_residual_what"] = "max |K[rows] alpha - y[rows]| on 32 probe rows, K rows recomputed on the host (numpy)"
    parity["alpha_equal_across_ranks"] = all_ranks_equal(alpha[idx].copy())
    if use_sharded and world > 1 and rank == 0:
        # the same fit on ONE GPU: a second handle on rank 0's device, factor digests compared bit for bit
        h1 = N.Handle(local_rank)
        h1.check(lib.fgp_fit(h1.ptr, N.dptr(X), n, n, d, N.dptr(y), C.byref(kd), noise, 0, 0.0))
        d1 = digest(h1)
        single_ms = h1.last_device_ms()
        parity["bitwise_equal_to_single_gpu_fit"] = bool(np.array_equal(d1, dg))
        parity["single_gpu_digest_rel_diff"] = float(np.abs(d1 - dg).max() / np.abs(d1).max())
        parity["single_gpu_same_n_ms"] = single_ms
        h1.close()
    barrier()

    # ---- the north star's strong-scaling experiment (C4: n = 32768, d = 32) beside the weak-scaling series ---------------
    strong = None
    if world > 1 and args.workload == "metric" and not args.no_strong:
        n4, d4 = WORKLOADS["c4"][0], WORKLOADS["c4"][1]
        X4s, y4s = make_dataset(0x5EED0004, n4, d4)
        kd4 = SquaredExp(math.sqrt(d4 / 6.0), 1.0).device_desc()
        X4 = np.asfortranarray(X4s)
        h.check(lib.fgp_fit_sharded(h.ptr, N.dptr(X4) if rank == 0 else None, n4, n4, d4, N.dptr(y4s) if rank == 0 else None,
                                    C.byref(kd4), noise, 0, 0.0))
        barrier()
        ms4 = 0.0
        for _ in range(2):
            h.check(lib.fgp_refit_sharded(h.ptr, C.byref(kd4), noise, 0, 0.0))
            ms4 += h.last_device_ms()
        ms4 = max_over_ranks(ms4 / 2)
        dg4 = digest(h)
        eq4 = all_ranks_equal(dg4)
        one_ms = None
        if rank == 0:  # the 1-GPU time of the same problem, measured in this run on rank 0's GPU while the others wait
            h1 = N.Handle(local_rank)
            h1.check(lib.fgp_fit(h1.ptr, N.dptr(X4), n4, n4, d4, N.dptr(y4s), C.byref(kd4), noise, 0, 0.0))
            h1.check(lib.fgp_refit(h1.ptr, C.byref(kd4), noise, 0, 0.0))
            one_ms = h1.last_device_ms()
            h1.close()
        barrier()
        if rank == 0:
            f4 = fit_flops(n4, d4)
            strong = {"config": "C4: RBF n=32768 d=32 fit (Gram + Cholesky), strong scaling", "n": n4, "d": d4, "n_gpus": world,
                      "ms_per_step": ms4, "tflops": f4 / (ms4 * 1e-3) * 1e-12, "one_gpu_ms_same_run": one_ms,
                      "one_gpu_tflops_same_run": f4 / (one_ms * 1e-3) * 1e-12, "speedup": one_ms / ms4,
                      "efficiency": one_ms / ms4 / world, "frac_of_fp64_peak": f4 / (ms4 * 1e-3) * 1e-12 / (world * FP64_PEAK_TFLOPS),
                      "digest_equal_across_ranks": eq4}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = fit_flops(n, d) / (fit_ms * 1e-3) * 1e-12
    traffic, traffic_launches, traffic_file = gemm_traffic_bytes_per_launch(args.workload) if world == 1 else (None, 0, None)
    traffic_src = (f"{traffic_file}: dram__bytes_read.sum + dram__bytes_write.sum over the {traffic_launches} ozaki_update_kernel "
                   f"launches of one fit of this build / launches") if traffic else "no ncu DRAM capture of this build committed for this workload"
    gemm_tflops = tfl[0] / (tms[0] * 1e-3) * 1e-12 if tms[0] > 0 else None
    tc_f64 = tfl[4] / (tms[4] * 1e-3) * 1e-12 if tms[4] > 0 else None   # f64-equivalent TFLOP/s of the tcgen05 launches
    bf16_burst, bf16_sustained, peak_file = measured_peaks()
    int8_peak = 2.0 * bf16_burst
    dmma_block = {"kernel": "gemm_nt_kernel", "bound": "tensor", "achieved": gemm_tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                  "frac": (gemm_tflops / FP64_PEAK_TFLOPS) if gemm_tflops else None, "launches": int(tcnt[0]),
                  "share_of_step": float(tms[0] / max(prof_dev_ms, 1e-9)), "isolated": iso,
                  "what": "f64 DMMA kernel: panel solves A21 W^T, next-diagonal-block updates, trailing updates with < 1024 rows left",
                  "peak_source": "fp64 DMMA m8n8k4 register-resident burst measured on this pool (profiles/fp64_peak_r01.jsonl; "
                                 "MEASURED_PEAKS.json has no fp64 figure; nominal 148 SM x 128 flop/clk x 1.965 GHz = 37.2)"}
    if tc_f64 is not None and tms[4] >= tms[0]:
        roofline = {"kernel": "ozaki_update_kernel", "bound": "tensor", "achieved": 36.0 * tc_f64, "peak": int8_peak,
                    "unit": "TFLOP/s", "frac": 36.0 * tc_f64 / int8_peak,
                    "op": "int8 tensor operations (2 per multiply-add) of tcgen05.mma kind::i8: 36 digit-slice products of "
                          "2*128*128*K per 128x128 tile (SURVEY 8(d)'s 2 m n k per f64 GEMM x 36)",
                    "f64_equivalent_tflops": tc_f64, "f64_equivalent_over_dmma_peak": tc_f64 / FP64_PEAK_TFLOPS,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "launches": int(tcnt[4]), "share_of_step": float(tms[4] / max(prof_dev_ms, 1e-9)),
                    "frac_of_sustained_peak": 36.0 * tc_f64 / (2.0 * bf16_sustained),
                    "note": "rank 0; achieved = int8 tensor ops of ALL tcgen05 trailing-update launches of a fit / sum of their "
                            "CUDA-event durations, from a separate pass of the same steps with per-launch events on and, on one "
                            "GPU, the single-stream schedule so launches do not share SMs; `isolated` = the dominant launch "
                            "shape alone, measured live; the kernel runs into the 1 kW power cap (tools/ozaki_clocks.py: SM "
                            "clock 1965 -> ~1650 MHz when it runs back to back), which is what the sustained figure reflects",
                    "isolated": iso_tc, "profiled_ms_per_step": prof_dev_ms / args.steps,
                    "peak_source": f"2 x bf16_tflops of {peak_file} (burst {bf16_burst:.1f}, sustained {bf16_sustained:.1f} TFLOP/s; "
                                   f"the int8 rate of the tcgen05 pipe is twice its bf16 rate; nominal 4500)",
                    "dmma": dmma_block}
    else:
        roofline = dict(dmma_block)
        roofline.update({"traffic": None, "traffic_source": "no capture for this configuration",
                         "profiled_ms_per_step": prof_dev_ms / args.steps,
                         "note": "rank 0; the tcgen05 path is off or minor in this configuration: f64 DMMA kernel, achieved = "
                                 "algorithmic flops of all its launches / sum of their CUDA-event durations (profiling pass)",
                         "tcgen05": {"f64_equivalent_tflops": tc_f64, "launches": int(tcnt[4]),
                                     "share_of_step": float(tms[4] / max(prof_dev_ms, 1e-9)), "isolated": iso_tc}})
    line = {
        "metric": "gp_fit_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": fit_ms, "higher_is_better": True,
        # the --gpus N series of the metric workload is a WEAK-scaling series (n = 16384 N^(1/3): equal Cholesky work per GPU)
        # whose N = 1 point is the metric's own configuration; the strong-scaling experiment of the north star (C4) is the
        # `strong_c4` block of every N > 1 line
        "scaling": "weak" if args.workload == "metric" else "strong", "vs_baseline": None, "dtype": "f64",
        "trailing_updates": trailing,
        "dtype_detail": "inputs, outputs and storage f64; O(n^3) trailing updates / solve updates computed as exact int8 x int8 -> int32 "
                        "digit-slice products on tcgen05 (36 per f64 product, two f64 roundings per element and launch), panel chain "
                        "in f64 on the DMMA pipe; same tolerances as a pure f64 path (tests/test_gpu_parity.py)",
        "data": "synthetic", "config": config,
        "frac_of_fp64_peak": value / (world * FP64_PEAK_TFLOPS),
        "predict_qps": q / (pred_ms * 1e-3), "predict_ms": pred_ms, "predict_q1_latency_ms": q1_ms,
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": (fit_flops(n, d) + world * predict_flops(n, d, qr)) / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s",
                "ms_per_step": e2e_ms, "fit_ms": e2e_fit_ms / args.steps,
                "predict_qps": q / max((e2e_ms - e2e_fit_ms / args.steps) * 1e-3, 1e-9),
                "h2d_bytes_per_step": 8 * (n * d + n + q * d), "d2h_bytes_per_step": 8 * 2 * q,
                "what": ("fgp_fit_sharded(rank-0 host X, y) + " if use_sharded else "fgp_fit(host X, y) + ") +
                        "fgp_predict_mean_var(host Xq) -> host mean, var"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernel_ms_per_step": {"tcgen05_update": tms[4] / args.steps, "digit_slicing": tms[5] / args.steps,
                               "gemm_nt": tms[0] / args.steps, "potrf_head": tms[1] / args.steps, "gram": tms[2] / args.steps},
        # what bounds a step: summed kernel time per class from the profiling pass (on one GPU: single-stream schedule, so
        # the sum IS the step) next to the overlapped step; the head kernels are the serial chain, the rest is GEMM
        "step_breakdown_ms": {"overlapped_step": fit_ms, "serialised_step": prof_dev_ms / max(args.steps, 1),
                              "tcgen05_updates": tms[4] / args.steps, "digit_slicing": tms[5] / args.steps,
                              "gemm": tms[0] / args.steps, "panel_heads": tms[1] / args.steps, "gram": tms[2] / args.steps,
                              "bcast_as_owner": (tms[3] / args.steps) if use_sharded else 0.0,
                              "panel_head_launches": int(tcnt[1] // max(args.steps, 1))},
        "parity": parity,
        "bcast_as_owner": ({"ms_per_step": tms[3] / args.steps, "launches": int(tcnt[3]),
                            "GBps": (tfl[3] / max(tms[3], 1e-9)) * 1e-6} if use_sharded and tcnt[3] else None),
        "clocks": clocks,
    }
    if use_sharded:
        line["nccl_bcast_mb_per_step"] = bcast_mb
    if strong is not None:
        line["strong_c4"] = strong
    if world == 1 and not args.no_cpu_baseline:
        ns, qs = (4096, 256) if n >= 4096 else (n, min(q, 256))
        dt, flops = time_oracle(ns, d, qs, 1, 0)
        rows, pexp = cpu_port_scaling(d, (1024, 2048))
        rows.append({"n": ns, "seconds": dt, "gflops": flops / dt * 1e-9, "note": f"includes predict q={qs}"})
        line["cpu_baseline"] = {"value": flops / dt * 1e-12, "unit": "TFLOP/s", "cores": 1, "kind": "port",
                                "host_cores": os.cpu_count(), "seconds": dt,
                                "sample": f"oracle fit n={ns} d={d} + predict_mean_variance q={qs}, 1 step, 1 thread "
                                          f"(the reference is single-threaded)",
                                "fit_seconds_by_n": rows, "fitted_exponent_1024_2048": pexp,
                                "extrapolated_fit_seconds_full_config": dt * (n / ns) ** 3,
                                "extrapolation": f"t(n={n}) = t(n={ns}) x (n/{ns})^3 (the rate falls out of cache above n~3000, "
                                                 f"so the cubic law is a lower bound on the CPU time)"}
    print(json.dumps(line), flush=True)
    for p in (pX, py, pq, pm, pv):
        lib.fgp_free_pinned(p)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true", help="A/B: single-stream Cholesky schedule")
    ap.add_argument("--no-profile", action="store_true", help="A/B: no per-launch CUDA events (roofline fields become null)")
    ap.add_argument("--no-tcgen05", action="store_true", help="A/B: f64 DMMA kernel for every trailing update (round-1 arithmetic)")
    ap.add_argument("--sharded", action="store_true", help="use the collective entry points even on one GPU")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling C4 block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
