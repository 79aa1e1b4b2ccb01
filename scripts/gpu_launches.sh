#!/bin/bash
# per-kernel time of a bench step from an ncu launch list; EXTRA="--selaln" (default) or EXTRA=""
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --oracle-sample 0 --legs none > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sa_collect|pack_reads|kmer_mask|work_class|k2_class|hits_to_mappings|merge_|selaln|ksw|copy_out" -c 400 --csv --log-file gpurun_out/launches${TAG:-_sel}.csv \
   python bench.py --steps 1 --warmup 1 --chunks 4 --e2e-depth 1 --legs none --no-cpu-baseline --oracle-sample 0 ${EXTRA---selaln} > gpurun_out/ncu_launch_sel.json 2> gpurun_out/ncu_launch_sel.log
python - <<PY
import csv, collections
rows=list(csv.reader(l for l in open("gpurun_out/launches${TAG:-_sel}.csv") if not l.startswith("==")))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=vi: continue
    a=agg.setdefault(r[ki][:60],[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(",",""))
for k,(n,t) in agg.items(): print(f"{k:60s} n={n:3d} total={t/1e6:9.3f} ms  avg={t/1e6/n:8.3f} ms")
PY
