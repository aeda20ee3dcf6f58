#!/bin/bash
# round 2, GPU call U (1 GPU): local groups with heavy vertices / hub rows and other dimensions
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_sharded_local.py -m gpu -q --timeout 60 -k "heavy or 30000-16" > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?"
grep -E "^E  |passed|failed" gpurun_out/r2u_pytest.log | head -30
