#!/bin/bash
# A/B: bench every library under rapmap_b200/_build/ab/ (and the default build) on the same box; EXTRA="--selaln" etc.
mkdir -p gpurun_out
run() { # name, lib
  RAPMAP_B200_LIB=$2 timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --oracle-sample 5000 --legs none $EXTRA > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1.json")); print("$1", round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_chunk"].items()}, d["parity_checked_vs_oracle"])
except Exception as e: print("$1", "ERR", e, open("gpurun_out/ab_$1.log").read()[-400:])
PY
}
run default ""
for lib in rapmap_b200/_build/ab/lib_*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  run $n $PWD/$lib
done
