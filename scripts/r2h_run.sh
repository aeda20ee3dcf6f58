#!/bin/bash
# round 2, GPU call H (N GPUs): sharded step - bit equality, c3 bench, c5 secondary
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_edge_cases.py::test_graph_replay_equals_direct_launches -m gpu -x -q --timeout 500 > gpurun_out/r2h_pytest_multi.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/r2h_pytest_multi.log
tail -6 gpurun_out/r2h_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/multi_gpu_check.py 1000000 8 60 > gpurun_out/r2h_multi_c3_$N.log 2>&1; grep "world=" gpurun_out/r2h_multi_c3_$N.log; tail -2 gpurun_out/r2h_multi_c3_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_bench$N.json 2> gpurun_out/r2h_bench$N.err; python -c "
import json; t=open('gpurun_out/r2h_bench$N.json').read(); d=json.loads(t[t.index('{\"metric\"'):].splitlines()[0]); print(d['n_gpus'], d['ms_per_step'], d['steps_per_s'], d['e2e']['steps_per_s'], d['phases_ms'], d.get('ranks_identical'), d.get('secondary'))"; tail -4 gpurun_out/r2h_bench$N.err
