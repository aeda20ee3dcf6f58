#!/bin/bash
# round 2, GPU call N (1 GPU): the sharded step as a single-process group of handles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_local.py -m gpu -q --timeout 400 -x > gpurun_out/r2n_pytest_local.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest_local.log
grep -E "^E  |passed|failed|rc=" gpurun_out/r2n_pytest_local.log | head -30
