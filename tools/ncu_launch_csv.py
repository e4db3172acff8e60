#!/usr/bin/env python
"""Turn the raw `ncu --csv --log-file` output of a metrics pass into the small files kept under profiles/.

    python tools/ncu_launch_csv.py shares gpurun_out/launches.csv  > profiles/launch_shares_<what>_rNN.txt
        per-kernel launch count, summed gpu__time_duration and share of the total (serialised, cold-cache times)
    python tools/ncu_launch_csv.py dram   gpurun_out/gemm_dram.csv > profiles/gemm_dram_<what>_rNN.csv
        one row per launch: kernel,grid,dram_bytes_read,dram_bytes_write   (what bench.py reads for roofline.traffic)
"""
import collections
import csv
import re
import sys


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("fgp::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")


def main():
    mode, path = sys.argv[1], sys.argv[2]
    data = rows(path)
    if mode == "shares":
        ms, cnt = collections.Counter(), collections.Counter()
        for r in data:
            if r["Metric Name"] != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
            ms[short(r["Kernel Name"])] += v * scale
            cnt[short(r["Kernel Name"])] += 1
        total = sum(ms.values())
        print("# kernel, launches, summed gpu__time_duration ms (serialised, cold cache), share")
        for k, v in ms.most_common():
            print(f"{k}, {cnt[k]}, {v:.3f}, {100 * v / total:.1f}%")
        print(f"total, {sum(cnt.values())}, {total:.3f}, 100%")
    elif mode == "dram":
        per = collections.OrderedDict()
        for r in data:
            key = r["ID"]
            d = per.setdefault(key, {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"].replace(",", " ")})
            v = float(r["Metric Value"].replace(",", ""))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
            if r["Metric Name"] == "dram__bytes_read.sum":
                d["dram_bytes_read"] = v * scale
            elif r["Metric Name"] == "dram__bytes_write.sum":
                d["dram_bytes_write"] = v * scale
        print("# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum: one row per kernel launch")
        print("kernel,grid,dram_bytes_read,dram_bytes_write")
        for d in per.values():
            print(f"{d['kernel']},\"{d['grid']}\",{d.get('dram_bytes_read', 0.0):.0f},{d.get('dram_bytes_write', 0.0):.0f}")
    else:
        sys.exit("mode must be shares or dram")


if __name__ == "__main__":
    main()
