#!/bin/bash
# full GPU evidence run: parity tests, bench (default + -s, both arms), ncu launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv >> gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
timeout 1200 python bench.py > gpurun_out/bench_nosel.json 2> gpurun_out/bench_nosel.log
echo "bench exit $?"
timeout 900 python bench.py --steps 12 --warmup 3 --selaln > gpurun_out/bench_sel.json 2> gpurun_out/bench_sel.log
python - <<'PY'
import json
for f in ("bench_ref","bench_nosel","bench_sel"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]/1e6,3),"M pairs/s e2e", round(d["e2e"]["value"]/1e6,3), {k:round(v,2) for k,v in d.get("roofline",{}).get("stage_ms_per_step",{}).items()}, d.get("cpu_baseline"), d.get("parity_checked_vs_oracle"), d.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.log").read()[-600:])
PY
K='sa_collect|pack_reads|hits_to_mappings|merge_|selaln|ksw'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-mappers 1 > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 200 --csv --log-file gpurun_out/launches_sel.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-mappers 1 --selaln > gpurun_out/ncu_launch_bench_sel.json 2> gpurun_out/ncu_launch_sel.log
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect_lane|hits_to_mappings" -s 3 -c 3 -o gpurun_out/prof_full -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-mappers 1 > /dev/null 2> gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect_lane|ksw_extz_lane|selaln_prepare|hits_to_mappings_kernel" -s 4 -c 4 -o gpurun_out/prof_full_sel -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-mappers 1 --selaln > /dev/null 2> gpurun_out/ncu_full_sel.log
echo "ncu full exit $?"
ls -la gpurun_out | tail -14
