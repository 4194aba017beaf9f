"""The C++ host layer on a machine WITHOUT a GPU: GridSlamProcessor::processScan's orchestration, the trajectory tree, the
resample bookkeeping, the host map mirror, ScanMatcher's standalone entry points and the .gfs writer run against
tests/mock/gm_mock.c -- the C ABI of include/gm_b200.h implemented with the oracle and injected with LD_PRELOAD (test
infrastructure: the product never loads it, and without the preload it still refuses to run without a CUDA device).

The mock's arithmetic is the oracle's, which is bit-for-bit the reference's, and the host layer's own arithmetic (motion
model, drand48 stream, weights) restates the reference's: the driver written against the reference's API, linked with our
host library, must therefore reproduce the golden traces of the unmodified reference EXACTLY (tolerance 0)."""
import subprocess
import sys

import pytest

from common import CASES
from host_common import ROOT, check_trace, free_port, free_run, mock_env

import os


@pytest.mark.parametrize("case", CASES)
def test_free_running_trace_is_the_reference_trace_bit_for_bit(case):
    g, tr = free_run(case, env=mock_env())
    check_trace(case, g, tr, exact=True)


@pytest.mark.parametrize("scenario", ["clone", "order", "cli", "lazy_tree", "long:120", "n32", "adapt"])
def test_host_scenarios_on_the_mock(scenario):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_scenarios.py"), scenario], capture_output=True, text=True,
                         timeout=900, env=mock_env())
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]


def test_sharded_filter_two_ranks_on_the_mock(tmp_path):
    """the N > 1 host path on CPU: two torchrun ranks (gloo), each with its slice of the particles behind the mock, follow the
    reference trace bit for bit; resampled parents cross the rank boundary (tests/dist_check.py asserts that they do)"""
    env = mock_env()
    env["GM_MOCK_DIST_DIR"] = str(tmp_path)
    port = free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "dist_check room181_n30: ok" in res.stdout and "dist_check room181_k2_lskip: ok" in res.stdout


def test_building_blocks_of_the_host_layer_are_the_references(tmp_path):
    """oracle/ref_units.cpp -- written against the reference's headers: GridLineTraversal::gridLine, uniform_resampler,
    sampleGaussian, MotionModel::drawFromMotion, ScanMatcherMap::world2map / map2world / isInside -- compiled UNCHANGED against
    our include tree and libraries writes the same bytes as when it is linked with the reference (sha256 in tests/golden/units.npz).
    No device is involved: none of these calls reaches the C ABI."""
    import hashlib
    import numpy as np
    from host_common import HOST, PKG, build_host
    build_host()
    exe, out = str(tmp_path / "units_b200"), str(tmp_path / "units.bin")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-w", "-I", os.path.join(HOST, "include"), "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "oracle", "ref_units.cpp"), "-L", HOST, "-lgridfastslam_b200", "-L", PKG, "-lgm_b200",
                           "-Wl,-rpath," + HOST, "-Wl,-rpath," + PKG])
    subprocess.check_call([exe, out])
    z = np.load(os.path.join(ROOT, "tests", "golden", "units.npz"))
    data = open(out, "rb").read()
    assert len(data) == int(z["units_bytes"]) and hashlib.sha256(data).hexdigest() == str(z["units_sha256"])


def test_mock_implements_the_whole_c_abi():
    """the CPU mock must keep up with include/gm_b200.h: every declared entry point is defined by it (gm_dist_plan, host-only
    planning code without device work, is the product's own)"""
    import ctypes
    import re
    from host_common import build_mock
    hdr = open(os.path.join(ROOT, "include", "gm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gm_[a-z_]+)\s*\(", hdr)) - {"gm_ctx", "gm_dist_plan"})
    out = subprocess.run(["nm", "-D", "--defined-only", build_mock()], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (gm_[a-z_]+)", out))
    missing = [f for f in declared if f not in defined]
    assert not missing, missing
