#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosel.json 2> gpurun_out/bench_nosel.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --selaln > gpurun_out/bench_sel.json 2> gpurun_out/bench_sel.log
python - <<'PY'
import json
for f in ("bench_nosel","bench_sel"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["parity_checked_vs_oracle"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.log").read()[-600:])
PY
