#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck over the kernels of the default and -s paths (incl. the ksw2 pair kernel and the copy-out)
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "test_synth_flagsets_match_oracle_and_golden_sam and (default or selaln-) or test_edge_case or nosensitive" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -4
done
