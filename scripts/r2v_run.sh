#!/bin/bash
# round 2, GPU call V (1 GPU): heavy-tailed graph in a local group of 8 (one rank owns no vertices): where do the layouts differ?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python scripts/local_group_probe.py 20000 8 8 heavy > gpurun_out/r2v_probe.log 2>&1; echo "rc=$?"
cut -c1-330 gpurun_out/r2v_probe.log | grep -v "^v=" | head -60
