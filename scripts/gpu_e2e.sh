#!/bin/bash
mkdir -p gpurun_out
run() { # name, env, flags
  env $2 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --oracle-sample 2000 $3 > gpurun_out/e_$1.json 2> gpurun_out/e_$1.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e_$1.json")); print("$1", round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["parity_checked_vs_oracle"])
except Exception as e: print("$1", "ERR", e, open("gpurun_out/e_$1.log").read()[-600:])
PY
}
run m1 "A=1" "--e2e-mappers 1"
run m2 "A=1" "--e2e-mappers 2"
run m3 "A=1" "--e2e-mappers 3"
run l2f32 "RAPMAP_B200_L2_FETCH=32" "--e2e-mappers 2"
run l2f128 "RAPMAP_B200_L2_FETCH=128" "--e2e-mappers 2"
run sel_m2 "A=1" "--e2e-mappers 2 --selaln"
run sel_l2f32 "RAPMAP_B200_L2_FETCH=32" "--e2e-mappers 2 --selaln"
