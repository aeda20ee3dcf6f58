#!/bin/bash
# round 2, GPU call J (1 GPU): fast suite on the current tree, A/B of the bulk-copy prefetch of the Adam moments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_multi.py -k "not at_size and not baseline_sizes and not headline_size" > gpurun_out/r2j_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> gpurun_out/r2j_pytest_fast.log
tail -4 gpurun_out/r2j_pytest_fast.log
timeout 300 python scripts/gpu_ab.py c3 20 60 wembed_b200/lib/variants/libwb_bulkmv0.so wembed_b200/lib/variants/libwb_bulkmv1.so wembed_b200/lib/variants/libwb_bulkmv0.so wembed_b200/lib/variants/libwb_bulkmv1.so > gpurun_out/r2j_ab.log 2>&1; cat gpurun_out/r2j_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; python -c "
import json; t=open('gpurun_out/r2j_bench.json').read(); d=json.loads(t[t.index('{\"metric\"'):].splitlines()[0]); print(d['ms_per_step'], d['steps_per_s'], d['e2e']['steps_per_s'], d['phases_ms'], d['gpu_launches'])"; tail -3 gpurun_out/r2j_bench.err
