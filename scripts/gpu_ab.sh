#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
run() { # name, lib, extra flags
  RAPMAP_B200_LIB=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --oracle-sample 2000 $3 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1.json")); print("$1", round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["parity_checked_vs_oracle"])
except Exception as e: print("$1", "ERR", e, open("gpurun_out/ab_$1.log").read()[-400:])
PY
}
run tab "" ""
run notab $PWD/rapmap_b200/_build/ab/lib_notab.so ""
run tab_sel "" "--selaln"
run notab_sel $PWD/rapmap_b200/_build/ab/lib_notab.so "--selaln"
