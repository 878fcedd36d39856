# coding=utf-8
"""Multi-GPU test (needs >= 2 CUDA devices; skipped otherwise): NCCL data-parallel training step
equals the single-GPU full-batch step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_step_equals_full_batch():
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "ddp_check.py")]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
  assert "DDP_CHECK" in r.stdout
