#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "longer_reads or alternative_sa" > gpurun_out/pytest_new.log 2>&1; tail -25 gpurun_out/pytest_new.log
