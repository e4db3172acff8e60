"""The driver's contract for `bench.py --impl reference` (the CPU arm: the oracle port timed on the host cores), checked without
a GPU: one JSON line, same `config` object as our arm would print, `e2e` with zero copies, a `cpu_baseline` describing the run;
under torchrun (N > 1) rank 0 alone prints and every rank exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check_reference_line(d, world):
    assert d["impl"] == "reference" and d["metric"] == "gp_fit_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == world and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "oracle fit n=2048" in cb["sample"]
    assert [r["n"] for r in cb["fit_seconds_by_n"]] == [1024, 2048, 4096] and 2.0 < cb["fitted_exponent"] < 4.0
    assert d["gpu_launches"] == 0


def test_reference_arm_prints_one_contract_line():
    sys.path.insert(0, ROOT)
    import bench
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)
    # the same config object our arm prints for this workload (built in one place: bench.workload_config)
    assert lines[0]["config"] == bench.workload_config("metric", 1, False)[3]
    assert lines[0]["config"]["n"] == 16384 and lines[0]["config"]["d"] == 16


def test_reference_arm_under_torchrun_prints_on_rank0_only():
    sys.path.insert(0, ROOT)
    import bench
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29877", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 2)
    assert lines[0]["config"] == bench.workload_config("metric", 2, True)[3]
    assert lines[0]["config"]["n"] == bench.weak_n(2) == 20480
