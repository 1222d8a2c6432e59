"""Data-parallel numerics over NCCL (needs >= 2 visible GPUs; the driver's 1-GPU test box skips it -- the log of a
`gpurun --gpus 2` run is committed under profiles/): tools/nccl_parity.py compares two ranks x B/2 with the CPU oracle
evaluated the way the reference's DataParallel does (per-chunk BatchNorm statistics, loss on the gathered batch:
models/unetbaseline_model.py:52-55, train.py:642-669)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_matches_dataparallel_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(REPO, "tools", "nccl_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    print(r.stdout[-4000:])
    print(r.stderr[-4000:])
    assert r.returncode == 0 and "nccl parity ok" in r.stdout
