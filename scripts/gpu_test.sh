#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "sample or edge or (synth_flagsets and (selaln-True or selaln_w5 or default))" -p no:cacheprovider > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
tail -5 gpurun_out/memcheck.log
timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.log
tail -3 gpurun_out/bench_quick.log; cat gpurun_out/bench_quick.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'])"
