#!/bin/bash
mkdir -p gpurun_out
run() { # name, flags
  timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --oracle-sample 2000 $2 > gpurun_out/e_$1.json 2> gpurun_out/e_$1.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e_$1.json")); print("$1", round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["parity_checked_vs_oracle"])
except Exception as e: print("$1", "ERR", e, open("gpurun_out/e_$1.log").read()[-600:])
PY
}
run m2 "--e2e-mappers 2"
run m3 "--e2e-mappers 3"
run m4 "--e2e-mappers 4"
run m3b "--e2e-mappers 3 --batch 524288"
