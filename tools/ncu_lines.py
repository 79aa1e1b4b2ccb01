#!/usr/bin/env python
"""Per-source-line instruction counts, sampling share and active threads per instruction of one kernel.
usage: ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
cur, out, seen = None, [], set()
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if cur in seen:
            cur = None  # a later launch of the same kernel
        else:
            seen.add(cur)
        continue
    if cur is None or len(r) < 12:
        continue
    try:
        ln, ie, te, smp = int(r[0]), int(r[7] or 0), int(r[8] or 0), int(r[6] or 0)
    except ValueError:
        continue
    out.append((cur, ln, ie, te, smp, r[1].strip()[:100]))
tot = sum(o[2] for o in out) or 1
tots = sum(o[4] for o in out) or 1
print(f"total warp instructions {tot:.3e}, samples {tots}")
for o in sorted(out, key=lambda x: -x[2])[:top]:
    print(f"{100*o[2]/tot:5.1f}% inst {100*o[4]/tots:5.1f}% smp  thr/inst {o[3]/max(o[2],1):5.1f}  n={o[2]:.2e}  {o[0]}:{o[1]:<4d} {o[5]}")
