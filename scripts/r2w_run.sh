#!/bin/bash
# round 2, GPU call W (1 GPU, the last seconds of the budget): the heavy-tailed local group of 8 with an empty rank
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_gpu_sharded_local.py -m gpu -q --timeout 25 -k "heavy" > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?"
grep -E "^E  |passed|failed" gpurun_out/r2w_pytest.log | head -12
