#!/bin/bash
mkdir -p gpurun_out
EXTRA=--selaln STEPS=4 bash scripts/gpu_ab.sh
bash scripts/gpu_launches_sel.sh 2>&1 | grep -E "hits_to|ksw|prepare|sa_collect"
