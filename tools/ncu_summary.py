"""Summarise an .ncu-rep (read here, without a GPU) into a small JSON-lines file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_<name>_rNN.jsonl [--stalls]

One line per profiled launch: kernel, grid, block, duration, DRAM bytes read / written, achieved DRAM GB/s, L2 hit rate,
registers, resident-CTA limits, fp64 tensor (DMMA) pipe activity, issue activity — and with --stalls the warp-stall sample
histogram of the launch (source page)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_active_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_inst_pct_of_peak",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__cycles_active.avg": "sm_cycles_active",
    "smsp__inst_executed.sum": "warp_insts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    # tcgen05 kernels (csrc/ozaki.cu): the int8 tensor sub-pipe, the SM clock the launch actually ran at (power cap), shared memory
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg": "imma_subpipe_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_ghz",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic_kb",
    "launch__cluster_size": "cluster_size",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, body = rows[0], rows[1], rows[2:]
    H = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in body:
        d = {"kernel": r[H["Kernel Name"]].split("(")[0], "grid": r[H["Grid Size"]], "block": r[H["Block Size"]]}
        for k, nm in WANT.items():
            if k not in H:   # this ncu version prefixes some metrics with the unit they are collected in (TPC.TriageCompute. ...)
                alt = [h for h in H if h.endswith("." + k)]
                if alt:
                    H[k] = H[alt[0]]
            if k in H and r[H[k]] not in ("", "n/a"):
                v = float(r[H[k]].replace(",", ""))
                u = units[H[k]]
                d[nm] = v * SCALE.get(u, 1.0) if nm in ("duration", "dram_read", "dram_write") else v
        if d.get("duration") and "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0.0)
            d["dram_GBps"] = d["dram_bytes"] / d["duration"] * 1e-9
        lines.append(d)
    if "--stalls" in sys.argv:
        src = page(rep, "source", ("--print-source", "sass"))
        nsec = sum(1 for r in src if r and r[0] == "Kernel Name")
        per = max(1, nsec // max(1, len(lines)))  # the source page repeats each launch once per view
        sec, hdr2 = -1, None
        for r in src:
            if r and r[0] == "Kernel Name":
                sec += 1
                hdr2 = None
                continue
            k = sec // per
            use = (sec % per == 0) and k < len(lines)
            if hdr2 is None:
                hdr2 = {h: i for i, h in enumerate(r)}
                if use:
                    lines[k]["stall_samples"] = {h: 0 for h in r if h.startswith("stall_") and "Not Issued" not in h}
                    lines[k]["samples"] = 0
                continue
            if use:
                lines[k]["samples"] += int(r[hdr2["# Samples"]])
                for s in lines[k]["stall_samples"]:
                    lines[k]["stall_samples"][s] += int(r[hdr2[s]])
        for d in lines:
            if "stall_samples" in d:
                d["stall_samples"] = {s[6:]: v for s, v in sorted(d["stall_samples"].items(), key=lambda x: -x[1]) if v}
    with open(dst, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
    print(f"{len(lines)} launches -> {dst}")


if __name__ == "__main__":
    main()
