#!/bin/bash
# full GPU check: parity tests, bench (default + -s), ncu launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_nosel.json 2> gpurun_out/bench_nosel.log
echo "bench exit $?"
timeout 900 python bench.py --steps 6 --warmup 3 --selaln > gpurun_out/bench_sel.json 2> gpurun_out/bench_sel.log
python - <<'PY'
import json
for f in ("bench_nosel","bench_sel"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d.get("cpu_baseline"), d["parity_checked_vs_oracle"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.log").read()[-600:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --oracle-sample 0 > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.log
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect|hits_to_mappings" -s 2 -c 2 -o gpurun_out/prof_r01c -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 > /dev/null 2> gpurun_out/ncu_full.log
echo "ncu full exit $?"
ls -la gpurun_out | tail -12
