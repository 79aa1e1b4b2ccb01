#!/bin/bash
# end-of-round evidence run on one GPU: parity tests, smoke, both bench arms as the driver runs them, launch list of the default path
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv >> gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_ref.json")); print("reference arm", round(d["value"]/1e6,3), "M pairs/s", d.get("cpu_baseline",{}).get("cores"))
except Exception as e: print("ref ERR", e)
try:
    d=json.load(open("gpurun_out/bench_n1.json"))
    print("headline", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_chunk"].items()}, d["parity_checked_vs_oracle"], d["clocks"], d["cpu_baseline"])
    for k,v in d["legs"].items():
        print(k, round(v["value"]/1e6,2), "e2e", round(v["e2e"]["value"]/1e6,2), v.get("parity"), {a:round(b,2) for a,b in (v.get("stage_ms_per_chunk") or v["roofline"]["stage_ms_per_chunk"]).items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_n1.log").read()[-3000:])
PY
TAG=_nosel EXTRA="" bash scripts/gpu_launches.sh | tail -14
