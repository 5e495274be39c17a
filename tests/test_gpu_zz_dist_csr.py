"""Multi-GPU parity of the row-sharded float64 CSR solver (needs >= 2 GPUs): launches
tests/dist_gpu_worker_csr.py with torchrun, one rank per GPU over NCCL.  Named to run last."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_row_sharded_csr_classes_match_oracle():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "dist_gpu_worker_csr.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_CSR_OK" in res.stdout
