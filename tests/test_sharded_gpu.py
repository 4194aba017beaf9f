"""The sharded filter on real GPUs: two ranks, peer-memory exchange of the scan-match rows, pull migration of resampled
parents -- on two GPUs when the box has them, otherwise both ranks on GPU 0 (CUDA IPC works within one device, so the
whole exchange path runs on a one-GPU box too)."""
import os
import subprocess
import sys

import pytest

from host_common import ROOT, build_host


def run_dist_check(port, extra_env, timeout=600):
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["GM_DIST_SAME_GPU"] = "1"
        env["GM_B200_DEVICE"] = "0"
    env.update(extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.gpu
def test_sharded_filter_two_ranks():
    build_host()
    res = run_dist_check(29611, {})
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "dist_check room181_n30: ok" in res.stdout and "dist_check room181_k2_lskip: ok" in res.stdout


@pytest.mark.gpu
def test_sharded_filter_whose_maps_outgrow_their_directories():
    """room181_resize starts from an 8 x 8 m map: Map::resize on the first scans, the directory stride grows on both ranks (the
    exported directory buffers are retired, not freed: a peer may still have them mapped), the handles are published again and
    the resampling that follows pulls parents through the re-opened mappings"""
    build_host()
    res = run_dist_check(29613, {"GM_DIST_CASES": "room181_resize"})
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "dist_check room181_resize: ok" in res.stdout


@pytest.mark.gpu
def test_pool_exhaustion_during_migration_is_loud():
    """a pool too small for the imported parents: the run must die with the pool message on every rank, not continue with holes"""
    build_host()
    res = run_dist_check(29612, {"GM_B200_POOL_PATCHES": "700", "GM_B200_DIST_TIMEOUT_S": "30", "GM_DIST_CASES": "room181_n30"}, timeout=300)
    assert res.returncode != 0
    assert "exhausted" in res.stderr, res.stderr[-3000:]
