#!/bin/bash
# measurement-only evidence: chunk-size sweep, compute-sanitizer racecheck / synccheck / memcheck on the parity tests' kernels
mkdir -p gpurun_out
timeout 900 python tools/batch_sweep.py > gpurun_out/batch_sweep.json 2> gpurun_out/batch_sweep.log; tail -9 gpurun_out/batch_sweep.log
SAN="-k synth_flagsets_match_oracle_and_golden_sam and (default or selaln or nosensitive or fuzzy_recover)"
for tool in racecheck synccheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "test_synth_flagsets_match_oracle_and_golden_sam and (default or selaln-) or test_edge_case or nosensitive" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -4
done
