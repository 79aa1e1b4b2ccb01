#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "cli or sample" > gpurun_out/pytest_cli.log 2>&1; tail -3 gpurun_out/pytest_cli.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosel.json 2> gpurun_out/bench_nosel.log
timeout 900 python bench.py --steps 6 --warmup 3 --selaln > gpurun_out/bench_sel.json 2> gpurun_out/bench_sel.log
tail -2 gpurun_out/bench_sel.log
python - <<'PY'
import json
for f in ("bench_nosel","bench_sel"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]/1e6,2),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,2), d["roofline"]["stage_ms_per_step"], d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect|hits_to_mappings" -s 2 -c 2 -o gpurun_out/prof_r01c -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 > /dev/null 2> gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ksw_extz|selaln_prepare|hits_to_mappings" -s 3 -c 3 -o gpurun_out/prof_r01c_sel -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 --selaln > /dev/null 2> gpurun_out/ncu_full_sel.log
ls -la gpurun_out | tail -8
