#!/bin/bash
# round 2, GPU call G: step graph, 64-bit pair-list offsets, c5 on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_multi.py -k "not at_size and not baseline_sizes and not headline_size" > gpurun_out/r2g_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> gpurun_out/r2g_pytest_fast.log
tail -6 gpurun_out/r2g_pytest_fast.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; python -c "
import json; t=open('gpurun_out/r2g_bench.json').read(); d=json.loads(t[t.index('{\"metric\"'):].splitlines()[0]); print(d['ms_per_step'], d['steps_per_s'], d['e2e']['steps_per_s'], d['phases_ms'], d['gpu_launches'])"; tail -3 gpurun_out/r2g_bench.err
WB_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2g_bench_nograph.json 2> gpurun_out/r2g_bench_nograph.err; python -c "
import json; t=open('gpurun_out/r2g_bench_nograph.json').read(); d=json.loads(t[t.index('{\"metric\"'):].splitlines()[0]); print('nograph', d['ms_per_step'], d['steps_per_s'], d['e2e']['steps_per_s'])"
timeout 200 python scripts/gpu_trajectory.py c2 3000 100 > gpurun_out/r2g_traj_c2.log 2>&1; head -3 gpurun_out/r2g_traj_c2.log; tail -2 gpurun_out/r2g_traj_c2.log
timeout 900 python scripts/gpu_trajectory.py c5 30 1 > gpurun_out/r2g_traj_c5.log 2>&1; cat gpurun_out/r2g_traj_c5.log | cut -c1-230
