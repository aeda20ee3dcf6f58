#!/bin/bash
# round 2, GPU call M (8 GPUs): why does the 8-rank step lose pairs?  three short runs of the check script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
for delay in 0 300; do
  WB_XCHG_DELAY_US=$delay WB_DEBUG=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$((delay/100)) scripts/multi_gpu_check.py 1000000 8 4 > gpurun_out/r2m_delay$delay.log 2>&1
  echo "== delay $delay us"; grep -E "\[check rank 0\]" gpurun_out/r2m_delay$delay.log | tr '[' '\n' | grep "check rank 0" | cut -c1-90; grep "world=" gpurun_out/r2m_delay$delay.log | cut -c1-250
done
