#!/usr/bin/env python
"""Per-kernel count of the SASS opcodes that prove which hardware paths libfgp_sm100.so uses (fp64 tensor pipe, TMA, bulk
copies, mbarriers): `cuobjdump -sass` of the in-tree library, grouped by function.  Output: profiles/sass_opcodes_rNN.txt.

    python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "friedrich_b200", "libfgp_sm100.so")
PATTERNS = ["UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "DMMA", "DFMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF",
            "UBLKPF", "SYNCS", "MUFU.RSQ64H", "SHFL", "ATOMG", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "BAR", "NANOSLEEP", "UTC",
            "HMMA", "IMMA", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                counts[cur][p] += 1
                break
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — opcode counts per kernel (instructions in the binary, not executed)")
    print("# UTCIMMA = tcgen05.mma kind::i8 (exact digit-slice products of the trailing updates, csrc/ozaki.cu), LDTM = tcgen05.ld, UTCBAR = "
          "tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc; DMMA.8x8x4 = the f64 tensor pipe (tcgen05 has no .kind::f64, DESIGN.md 4)")
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        name = demangle(fn)
        name = name.replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        items = ", ".join(f"{p}={c[p]}" for p in PATTERNS if c[p])
        print(f"{name}: total={c['_total']}; {items}")
        tot.update(c)
    print("ALL KERNELS: " + ", ".join(f"{p}={tot[p]}" for p in PATTERNS if tot[p]))


if __name__ == "__main__":
    sys.exit(main())
