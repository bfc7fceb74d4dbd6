#!/usr/bin/env python
"""Cut the LAST complete step (from one to_bf16/normalize/dist_matrix head kernel to the next) out of an ncu launch list
(--metrics gpu__time_duration.sum --csv) and print it: duration, grid, kernel.  Usage: launch_steps.py list.csv [head-regex]"""
import csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
head = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"to_bf16|qtable_minmax")
out = [(r[ki], float(r[vi].replace(",", "")) / 1000.0, r[gi], r[bi]) for r in rows[1:]]
idx = [i for i, o in enumerate(out) if head.search(o[0])]
# steps start where a head kernel follows a non-head kernel
starts = [i for i in idx if i == 0 or not head.search(out[i - 1][0])]
if len(starts) < 2:
    print("no complete step found; kernels:", sorted({o[0][:60] for o in out})); sys.exit(1)
a, b = starts[-2], starts[-1]
tot = 0.0
for name, us, grid, blk in out[a:b]:
    short = re.sub(r"^.*lgpu\d*_GLOBAL__N__\w+?_cu_\w+?(\d+)", "", name)
    short = re.sub(r"\(.*$", "", re.sub(r"^void |lgpu::|\(anonymous namespace\)::|<unnamed>::", "", name))[:60]
    print(f"{us:9.1f} us  {grid:>16s} {blk:>12s}  {short}")
    tot += us
print(f"{tot:9.1f} us  total of {b - a} launches")
