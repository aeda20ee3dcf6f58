#!/bin/bash
# round 2, GPU call P (1 GPU): which vertices does a local group of 8 get wrong? + two variants of emit_pair
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python scripts/local_group_probe.py 60000 8 8 > gpurun_out/r2p_probe.log 2>&1
WB_LIB=$PWD/wembed_b200/lib/variants/libwb_ev1.so timeout 200 python scripts/local_group_probe.py 60000 8 8 2>&1 | head -3 > gpurun_out/r2p_probe_ev1.log
WB_LIB=$PWD/wembed_b200/lib/variants/libwb_ev2.so timeout 200 python scripts/local_group_probe.py 60000 8 8 2>&1 | head -3 > gpurun_out/r2p_probe_ev2.log
timeout 200 python scripts/local_group_probe.py 60000 8 7 2>&1 | head -3 > gpurun_out/r2p_probe_w7.log
timeout 200 python scripts/local_group_probe.py 60000 8 6 2>&1 | head -3 > gpurun_out/r2p_probe_w6.log
timeout 200 python scripts/local_group_probe.py 60000 8 4 2>&1 | head -3 > gpurun_out/r2p_probe_w4.log
head -50 gpurun_out/r2p_probe.log | cut -c1-400; cat gpurun_out/r2p_probe_ev1.log gpurun_out/r2p_probe_ev2.log gpurun_out/r2p_probe_w7.log gpurun_out/r2p_probe_w6.log gpurun_out/r2p_probe_w4.log | cut -c1-300
