#!/bin/bash
# N-GPU diagnosis of the resident leg's scaling: waiting mode of the host thread x clock-sampler interval
N=${N:-8}
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --legs none --oracle-sample 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('$label', d['n_gpus'], 'resident', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ms/step', round(d['ms_per_step'],2), d['clocks'])"
}
run "poll + sampler 250ms " RAPMAP_B200_WAIT=poll RAPMAP_BENCH_CLOCK_MS=250
run "block + sampler 250ms" RAPMAP_BENCH_CLOCK_MS=250
run "poll + sampler 20ms  " RAPMAP_B200_WAIT=poll
