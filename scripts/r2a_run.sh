#!/bin/bash
# round 2, GPU call A: full GPU test suite (incl. the at-size parity tests), A/B of the prepared variants, whole-run trajectories, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
nproc >> gpurun_out/r2a_smi.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for v in staged pointhalf hitbatch both; do
  echo "== $v" >> gpurun_out/r2a_ab.log
  WB_LIB=$PWD/wembed_b200/lib/variants/libwb_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r2a_ab.log
done
timeout 300 python scripts/gpu_ab.py c3 20 60 wembed_b200/lib/variants/libwb_default.so wembed_b200/lib/variants/libwb_staged.so wembed_b200/lib/variants/libwb_pointhalf.so wembed_b200/lib/variants/libwb_hitbatch.so wembed_b200/lib/variants/libwb_both.so >> gpurun_out/r2a_ab.log 2>&1
timeout 300 python scripts/gpu_trajectory.py c3 3000 50 > gpurun_out/r2a_traj_c3.log 2>&1
timeout 200 python scripts/gpu_trajectory.py c2 3000 50 > gpurun_out/r2a_traj_c2.log 2>&1
timeout 200 python scripts/gpu_e2e_probe.py > gpurun_out/r2a_e2e.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_ab.log | tail -12; tail -2 gpurun_out/r2a_traj_c3.log; cat gpurun_out/r2a_e2e.log
