"""Multi-GPU checks on real hardware (skipped on boxes with one GPU): the data-parallel dynamics training step over NCCL and
over the fused NVLink peer-memory kernel, one rank per GPU under torchrun (tools/dist_train_check.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_training_on_nccl_and_peer_memory():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29431", os.path.join(ROOT, "tools", "dist_train_check.py"), "50"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == n
    assert out["allreduced_vs_full_batch_rel"] < 1e-5           # slices + NCCL sum == the full-batch gradient
    assert out["nccl_ranks_identical"]
    assert "p2p_error" not in out, out.get("p2p_error")
    assert out["p2p_ranks_identical"] and out["p2p_status"] == 0
    assert out["p2p_loss_vs_full"] < 1e-5
    assert out["p2p_vs_nccl_weights_after_1_step"]["n_gt_1e-6"] <= 16  # same update up to first-step Adagrad sign flips
    for a, b in zip(out["p2p_losses"], out["nccl_losses"]):
        assert abs(a - b) <= 1e-4 * abs(b)
