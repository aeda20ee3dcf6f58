#!/bin/bash
# round 2, GPU call E (2 GPUs): split-loop fused kernel, sharded step over IPC peer mappings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_multi.py -k "not at_size and not baseline_sizes and not headline_size" > gpurun_out/r2e_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> gpurun_out/r2e_pytest_fast.log
tail -4 gpurun_out/r2e_pytest_fast.log
timeout 300 python scripts/gpu_ab.py c3 20 60 wembed_b200/lib/variants/libwb_mb4.so wembed_b200/lib/variants/libwb_mb3.so > gpurun_out/r2e_ab.log 2>&1; cat gpurun_out/r2e_ab.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 500 > gpurun_out/r2e_pytest_multi.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/r2e_pytest_multi.log
tail -25 gpurun_out/r2e_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/multi_gpu_check.py 1000000 8 60 > gpurun_out/r2e_multi_c3.log 2>&1; tail -5 gpurun_out/r2e_multi_c3.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; tail -c 400 gpurun_out/r2e_bench2.json; tail -5 gpurun_out/r2e_bench2.err
