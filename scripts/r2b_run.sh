#!/bin/bash
# round 2, GPU call B: first run of the pair-list pipeline (tests under timeout: a wrong byte count of a bulk copy would spin on its mbarrier)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_multi.py -k "not at_size and not baseline_sizes and not headline_size" > gpurun_out/r2b_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> gpurun_out/r2b_pytest_fast.log
tail -5 gpurun_out/r2b_pytest_fast.log
if grep -q "rc=0" gpurun_out/r2b_pytest_fast.log; then
  timeout 900 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_full_size.py -m gpu -x -q --timeout 600 -s > gpurun_out/r2b_pytest_size.log 2>&1; echo "pytest size rc=$?" >> gpurun_out/r2b_pytest_size.log
  tail -5 gpurun_out/r2b_pytest_size.log
fi
timeout 300 python scripts/gpu_trajectory.py c3 3000 50 > gpurun_out/r2b_traj_c3.log 2>&1; tail -3 gpurun_out/r2b_traj_c3.log
WB_SKIN_MAX=0 timeout 300 python scripts/gpu_trajectory.py c3 200 50 > gpurun_out/r2b_traj_c3_noskin.log 2>&1; tail -2 gpurun_out/r2b_traj_c3_noskin.log
timeout 200 python scripts/gpu_trajectory.py c2 3000 100 > gpurun_out/r2b_traj_c2.log 2>&1; tail -2 gpurun_out/r2b_traj_c2.log
timeout 300 python scripts/gpu_locality_probe.py > gpurun_out/r2b_locality.log 2>&1; cat gpurun_out/r2b_locality.log
WB_MORTON_BITS=3 timeout 200 python scripts/gpu_trajectory.py c3 150 50 > gpurun_out/r2b_traj_c3_bits3.log 2>&1; tail -4 gpurun_out/r2b_traj_c3_bits3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 600 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
