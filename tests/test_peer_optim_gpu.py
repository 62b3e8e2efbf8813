"""The data-parallel optimizer step over peer memory (csrc/peer_optim.cu) needs at least two GPUs:
when the box has them, tests/peer_optim_check.py is run under torchrun (peer path against the NCCL
path: parameters, replicas, skip flag, checkpoint gathering); skipped on a single-GPU box, where
bench.py's dp_check covers the same ground whenever the driver runs it with N > 1."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_optimizer_matches_nccl_on_two_gpus():
    env = dict(os.environ, VITK_PEER_CHECK_QUICK="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29531", str(ROOT / "tests" / "peer_optim_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "replicas identical: True" in r.stdout
