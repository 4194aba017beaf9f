"""Two-GPU run of the sharded filter (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu`)."""
import os
import subprocess
import sys

import pytest

from host_common import ROOT, build_host


@pytest.mark.gpu
def test_sharded_filter_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    build_host()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "dist_check room181_n30: ok" in res.stdout
