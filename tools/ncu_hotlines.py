#!/usr/bin/env python
"""Top source lines of a kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_hotlines.py report.ncu-rep kernel_regex [topN]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
inst = collections.Counter()
samp = collections.Counter()
text = {}
cur_file = None
seen_files = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        seen_files.setdefault(cur_file, 0)
        seen_files[cur_file] += 1
        continue
    if r[0] in ("Function Name", "Line No", "Kernel Name", "File Name") or len(r) < 9:
        continue
    if seen_files.get(cur_file, 0) > 1:  # second launch of the same kernel
        continue
    try:
        line = int(r[0])
        ie = int(r[7] or 0)
        sm = int(r[6] or 0)
    except ValueError:
        continue
    if r[0] == "":  # SASS row under a source line
        continue
    key = (cur_file, line)
    inst[key] += ie
    samp[key] += sm
    text[key] = r[1].strip()[:110]
tot_i = sum(inst.values()) or 1
tot_s = sum(samp.values()) or 1
print(f"total warp instructions {tot_i:.3e}, samples {tot_s}")
for key, v in inst.most_common(top):
    print(f"{100*v/tot_i:5.1f}% inst {100*samp[key]/tot_s:5.1f}% smp  {key[0]}:{key[1]:<4d} {text[key]}")
