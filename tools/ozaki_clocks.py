"""SM clock / power while the tcgen05 update kernel runs back to back (is the int8 tensor pipe power-capped?).
Usage: python tools/ozaki_clocks.py [M] [reps]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402

lib = N.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
mu = C.c_double(0)
lib.fgp_dbg_ozaki_bench(0, 2048, 512, 1, 0, C.byref(mu), None)  # context + warm-up
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                      "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.time()
rc = lib.fgp_dbg_ozaki_bench(0, M, 512, reps, 0, C.byref(mu), None)
t1 = time.time()
time.sleep(0.1)
p.terminate()
rows = [l.strip().split(", ") for l in p.stdout.read().strip().splitlines() if l.strip()]
clk = [float(r[0]) for r in rows]
pw = [float(r[1]) for r in rows]
tiles = (M // 128) * (M // 128 + 1) // 2
print(json.dumps({"M": M, "reps": reps, "rc": rc, "update_ms": mu.value, "f64_equiv_tflops": 2.0 * 128 * 128 * tiles * 512 / mu.value * 1e-9,
                  "wall_s": t1 - t0, "samples": len(clk), "sm_mhz_min": min(clk), "sm_mhz_median": sorted(clk)[len(clk) // 2], "sm_mhz_max": max(clk),
                  "power_w_max": max(pw), "power_cap_active_samples": sum(1 for r in rows if "Active" in r[2] and "Not" not in r[2]),
                  "sm_mhz_series": clk[:: max(1, len(clk) // 40)]}))
