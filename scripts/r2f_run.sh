#!/bin/bash
# round 2, GPU call F: full GPU suite, bench with all legs, ncu captures of the two hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 300 python scripts/gpu_trajectory.py c3 3000 50 > gpurun_out/r2f_traj_c3.log 2>&1; head -5 gpurun_out/r2f_traj_c3.log; tail -4 gpurun_out/r2f_traj_c3.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 2500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 600 gpurun_out/r2f_bench_ref.json
# ncu: full captures of the walk and the fused kernel at step 15 (the bench window) and step 100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step_fused|k_repulse_pairs" -s 42 -c 3 -o gpurun_out/r2f_prof_step15 python scripts/profile_step.py 1000000 8 15 > gpurun_out/r2f_ncu15.log 2>&1; tail -2 gpurun_out/r2f_ncu15.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step_fused|k_repulse_pairs" -s 297 -c 3 -o gpurun_out/r2f_prof_step100 python scripts/profile_step.py 1000000 8 100 > gpurun_out/r2f_ncu100.log 2>&1; tail -2 gpurun_out/r2f_ncu100.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 10 --warmup 20 --no-cpu --no-secondary > gpurun_out/r2f_ncu_bench.log 2>&1; tail -2 gpurun_out/r2f_ncu_bench.log
