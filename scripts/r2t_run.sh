#!/bin/bash
# round 2, GPU call T (1 GPU): the bench line of the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
head -c 700 gpurun_out/r2t_bench.json; echo; tail -3 gpurun_out/r2t_bench.err
