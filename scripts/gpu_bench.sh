#!/bin/bash
# bench at full scale + ncu launch list + ncu full capture of the two top kernels
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.log
echo "bench exit $?" >> gpurun_out/bench_ours.log
tail -4 gpurun_out/bench_ours.log; cat gpurun_out/bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sa_collect|hits_to_mappings|merge_|build_table" -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --oracle-sample 0 > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.log
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect|hits_to_mappings" -s 2 -c 4 -o gpurun_out/prof_r01 -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --oracle-sample 0 > gpurun_out/ncu_full_bench.json 2> gpurun_out/ncu_full.log
echo "ncu full exit $?"
ls -la gpurun_out
