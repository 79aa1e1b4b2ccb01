#!/bin/bash
# end-to-end leg A/B: copy-engine + tail-kernel copy-out (default) against the all-kernel copy-out, then the in/out placement matrix
mkdir -p gpurun_out
one() { # label, env, args
  env $2 python bench.py --steps ${STEPS:-4} --warmup 2 --no-cpu-baseline --oracle-sample 2000 --legs none $3 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('$1', 'resident', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k:round(v,2) for k,v in d['e2e']['per_chunk_ms_on_its_streams'].items()})"
}
one "copy engines + tail kernel   " "X=1" ""
one "all-kernel copy-out          " "RAPMAP_B200_COPYOUT=kernel" ""
one "all-kernel, out device       " "RAPMAP_B200_COPYOUT=kernel" "--e2e-out device"
one "all-kernel, in device        " "RAPMAP_B200_COPYOUT=kernel" "--e2e-in device"
one "copy engines, -s             " "X=1" "--selaln"
