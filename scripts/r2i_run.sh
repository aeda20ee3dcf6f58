#!/bin/bash
# round 2, GPU call I (N GPUs): debug run of the sharded step at c3 with host-side milestones on stderr
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
WB_DEBUG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/multi_gpu_check.py 1000000 8 30 > gpurun_out/r2i_multi_c3_$N.log 2>&1
grep -E "world=|\[wb rank|\[check rank 0\]|gave up" gpurun_out/r2i_multi_c3_$N.log | head -60
