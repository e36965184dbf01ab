#!/usr/bin/env python3
"""SASS census of libfluxcalc_b200.so: per kernel, registers / stack / shared memory (cuobjdump -res-usage) and an opcode
histogram of the mnemonics that matter for this path (bulk copies and mbarrier traffic of the TMA ring, FP64 pipe,
MUFU seeds, shared / global / local memory accesses, atomics).  Runs without a GPU:
    python profiles/sass_census.py > profiles/r2_sass_census.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "components", "flux_calculator_b200", "libfluxcalc_b200.so")
WATCH = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMG", "RED", "SHFL", "BAR", "NANOSLEEP",
         "ACQBULK", "CCTL", "ERRBAR", "MEMBAR"]

res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
hist, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1
        hist[cur]["__all__"] += 1


def demangle(n):
    m = re.search(r"flux_spec_kernelILi(\d)ELi(\d)ELi(\d)ELb(\d)", n)
    if m:
        return "flux_spec_kernel<%s, S=%s, DIAG=%s, %s>" % ("BULK" if m.group(1) == "0" else "RCO", m.group(2), m.group(3), "dynamic" if m.group(4) == "1" else "static")
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", out)[:70]


print("# SASS census of libfluxcalc_b200.so (sm_100a)\n")
print("No tensor-core / TMEM mnemonics anywhere (the path is elementwise FP64). `UBLKCP` = `cp.async.bulk` (TMA bulk copy),")
print("`SYNCS` = mbarrier arrive / try_wait / expect_tx, `MUFU` = RCP64H / RSQ64H seeds of the lock-step division and square root.")
print("Counts are per kernel entry INCLUDING its out-of-line subroutines (the IEEE recompute paths and the diagnostics fold). The")
print("one-surface-type `flux_spec_kernel`s touch local memory only around the calls of those routines, never in the tile loops;")
print("`fused_step_kernel<0,0,true>`: about 30 of its 210 local-memory instructions sit on the hot path (before the first `EXIT`), the")
print("rest in `fused_cold_warp` and the IEEE division / exp / pow calls.\n")
print("| kernel | regs | stack B | static smem B | instr | " + " | ".join(WATCH) + " |")
print("|---|---|---|---|---|" + "---|" * len(WATCH))
for name in sorted(hist, key=demangle):
    h = hist[name]
    u = usage.get(name, (0, 0, 0))
    print("| %s | %d | %d | %d | %d | %s |" % (demangle(name), u[0], u[1], u[2], h["__all__"], " | ".join(str(h[w]) for w in WATCH)))
