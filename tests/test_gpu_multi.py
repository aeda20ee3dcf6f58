"""Vertex-sharded multi-GPU step (wb_comm_init): results must equal the single-GPU step.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("args,port", [(["60000", "4", "24"], "29533"), (["20000", "8", "12", "heavy"], "29534")])
def test_sharded_step_equals_single_gpu_step(device_lib, args, port):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "scripts", "multi_gpu_check.py"), *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("world=")][-1]
    assert "replicas identical across ranks: True" in line and "max rel coord err 0.000e+00" in line, line
