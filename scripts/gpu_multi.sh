#!/bin/bash
# bench under torchrun on N GPUs of one box: N=${N:-2}
N=${N:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 $EXTRA > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log
echo "torchrun exit $?"
tail -4 gpurun_out/bench_n$N.log | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n$N.json")); print(d["n_gpus"], round(d["value"]/1e6,1), "M pairs/s; e2e", round(d["e2e"]["value"]/1e6,1), d["clocks"], {k: (round(v["value"]/1e6,1), round(v["e2e"]["value"]/1e6,1)) for k, v in d["legs"].items()})
except Exception as e: print("no JSON:", e)
PY
dmesg 2>/dev/null | tail -5
