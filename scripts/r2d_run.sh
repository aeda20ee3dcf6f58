#!/bin/bash
# round 2, GPU call D: per-term fp64 sums, two-tier row sort, edge-weight / occupancy A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_multi.py -k "not at_size and not baseline_sizes and not headline_size" > gpurun_out/r2d_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> gpurun_out/r2d_pytest_fast.log
tail -5 gpurun_out/r2d_pytest_fast.log
timeout 400 python scripts/gpu_ab.py c3 20 60 wembed_b200/lib/variants/libwb_mb4_ws1.so wembed_b200/lib/variants/libwb_mb3_ws1.so wembed_b200/lib/variants/libwb_mb4_ws0.so wembed_b200/lib/variants/libwb_mb3_ws0.so > gpurun_out/r2d_ab.log 2>&1; cat gpurun_out/r2d_ab.log
timeout 300 python scripts/gpu_trajectory.py c3 3000 50 > gpurun_out/r2d_traj_c3.log 2>&1; head -4 gpurun_out/r2d_traj_c3.log; tail -8 gpurun_out/r2d_traj_c3.log
timeout 200 python scripts/gpu_trajectory.py c2 3000 100 > gpurun_out/r2d_traj_c2.log 2>&1; tail -2 gpurun_out/r2d_traj_c2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
if grep -q "rc=0" gpurun_out/r2d_pytest_fast.log; then
  timeout 900 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_full_size.py -m gpu -x -q --timeout 600 -s > gpurun_out/r2d_pytest_size.log 2>&1; echo "pytest size rc=$?" >> gpurun_out/r2d_pytest_size.log
  tail -5 gpurun_out/r2d_pytest_size.log
fi
