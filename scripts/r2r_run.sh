#!/bin/bash
# round 2, GPU call R (8 GPUs): the 8-rank step after the match.any fix: equality with one GPU + the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/multi_gpu_check.py 1000000 8 40 > gpurun_out/r2r_check.log 2>&1
grep "world=" gpurun_out/r2r_check.log | cut -c1-300; grep -E "Error|error" gpurun_out/r2r_check.log | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r2r_bench$N.json 2> gpurun_out/r2r_bench$N.err
tail -c 3000 gpurun_out/r2r_bench$N.json
