#!/bin/bash
# the edit-measure loop: GPU parity tests, then a short bench (all legs, no CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 2 --no-cpu-baseline ${EXTRA:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.log
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_quick.json"))
    print("headline", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_chunk"].items()}, d["parity_checked_vs_oracle"])
    for k,v in d["legs"].items():
        print(k, round(v["value"]/1e6,2), "e2e", round(v["e2e"]["value"]/1e6,2), v.get("parity"), {a:round(b,2) for a,b in (v.get("stage_ms_per_chunk") or v["roofline"]["stage_ms_per_chunk"]).items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_quick.log").read()[-3000:])
PY
