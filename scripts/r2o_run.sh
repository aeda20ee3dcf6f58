#!/bin/bash
# round 2, GPU call O (1 GPU): the sharded step as a lockstep group of handles on one stream
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded_local.py -m gpu -q --timeout 200 > gpurun_out/r2o_pytest_local.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest_local.log
grep -E "^E  |passed|failed|rc=" gpurun_out/r2o_pytest_local.log | head -40
