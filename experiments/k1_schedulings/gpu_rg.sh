#!/bin/bash
# regrouped SA-lookup kernel: parity test + bench against the lane kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "alternative_sa" > gpurun_out/pytest_rg.log 2>&1; tail -3 gpurun_out/pytest_rg.log
run() { # name env flags
  env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --oracle-sample 2000 --e2e-mappers 2 $3 > gpurun_out/rg_$1.json 2> gpurun_out/rg_$1.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/rg_$1.json")); print("$1", round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["parity_checked_vs_oracle"])
except Exception as e: print("$1", "ERR", e, open("gpurun_out/rg_$1.log").read()[-600:])
PY
}
run lane A=1 ""
run rg128 RAPMAP_B200_K1=regroup128 ""
run rg256 RAPMAP_B200_K1=regroup256 ""
