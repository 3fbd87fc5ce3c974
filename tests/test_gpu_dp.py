"""Multi-GPU consistency (needs >= 2 visible GPUs; skipped otherwise): the class-sharded text tower under data
parallelism gives the gradients of the replicated one and of a single rank over the global batch."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_class_sharded_text_tower_matches_replicated_and_single_rank():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", str(REPO / "tools" / "gpu_dp_check.py")],
                       cwd=str(REPO), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
