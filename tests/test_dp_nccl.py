"""N > 1 on real hardware: 2 NCCL ranks through the CUDA hot path (skipped below 2 GPUs; the CPU/gloo twin is
tests/test_dp_gloo.py).  SURVEY.md:279 - "2/4/8-GPU all-reduce equals single-GPU gradient on the concatenated batch"."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_nccl_allreduced_arena_equals_single_gpu_gradient(cuda_lib, tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dp_nccl_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    print(res)
    assert res["world"] == world and res["ranks_identical"]
    # same bar as everything else: 1e-5 max-norm relative, per parameter tensor and on the whole arena
    assert res["arena_vs_single_gpu"] <= 1e-5, res
    assert max(res["per_parameter"].values()) <= 1e-5, res
    # two epochs of Adam on the same global batches: DP == single process
    assert res["train_loss_rel"] <= 1e-4 and res["train_weights_rel"] <= 1e-3, res
