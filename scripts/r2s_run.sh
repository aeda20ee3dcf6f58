#!/bin/bash
# round 2, GPU call S (1 GPU): final tree - smoke() and the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s_smoke.log
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 --durations=8 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|s call" gpurun_out/r2s_pytest.log | head -30
