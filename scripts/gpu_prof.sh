#!/bin/bash
# ncu --set full of the hot kernels; NAME=tag KERNELS=regex EXTRA="--selaln"
mkdir -p gpurun_out
NAME=${NAME:-prof}; KERNELS=${KERNELS:-"sa_collect_lane|hits_to_mappings"}
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --oracle-sample 0 --legs none > /dev/null 2>&1   # builds + caches the index outside the profiler
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s ${SKIP:-2} -c ${COUNT:-2} -o gpurun_out/$NAME -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-depth 1 --legs none --chunks 2 $EXTRA > /dev/null 2> gpurun_out/ncu_$NAME.log
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_$NAME.log
