#!/bin/bash
# round 2, GPU call Q (1 GPU): pairs grouped by destination with match.any; local groups + the fast GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded_local.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q --timeout 300 > gpurun_out/r2q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest.log
grep -E "^E  |passed|failed|rc=" gpurun_out/r2q_pytest.log | head -40
