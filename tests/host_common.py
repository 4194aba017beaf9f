"""Build helpers for the C++ host layer tests: the reference-facing driver oracle/ref_driver.cpp (written against the
reference's headers and API) is compiled UNCHANGED against our headers and linked with our libraries."""
from __future__ import annotations

import hashlib
import importlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "slam-gmapping-learn_b200")
HOST = os.path.join(PKG, "host")
BUILD = os.path.join(ROOT, "tests", "_build")


def build_host():
    importlib.import_module("slam-gmapping-learn_b200.build").build_all()
    return os.path.join(HOST, "libgridfastslam_b200.so")


def build_driver():
    """oracle/ref_driver.cpp + our include tree + our libraries -> tests/_build/ref_driver_b200"""
    build_host()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "ref_driver_b200")
    src = os.path.join(ROOT, "oracle", "ref_driver.cpp")
    deps = [src, os.path.join(HOST, "libgridfastslam_b200.so"), os.path.join(PKG, "libgm_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-w", "-I", os.path.join(HOST, "include"), "-I", os.path.join(ROOT, "include"),
                               "-o", exe, src, "-L", HOST, "-lgridfastslam_b200", "-L", PKG, "-lgm_b200",
                               "-Wl,-rpath," + HOST, "-Wl,-rpath," + PKG])
    return exe


def driver_args(p):
    """command-line flags of ref_driver for a golden fixture's parameter dict"""
    cmd = []
    for k in ("particles", "seed", "kernelSize", "iterations", "lskip"):
        cmd += ["-" + k, str(int(p[k]))]
    for k in ("xmin", "ymin", "xmax", "ymax", "delta", "sigma", "maxrange", "maxUrange", "lstep", "astep", "lsigma", "ogain",
              "srr", "srt", "str", "stt", "linearUpdate", "angularUpdate", "resampleThreshold", "minimumScore", "period"):
        cmd += ["-" + k, repr(float(p[k]))]
    if not p["generate_map"]:
        cmd.append("-no-generateMap")
    return cmd


def build_mock():
    """tests/mock/gm_mock.c + oracle/gm_oracle.c -> tests/_build/libgm_mock.so: the C ABI of include/gm_b200.h implemented on
    the CPU with the oracle.  TEST INFRASTRUCTURE: preloaded (LD_PRELOAD) into the driver by the CPU tests of the host
    layer so that GridSlamProcessor's host-side logic runs without a GPU; never part of the product."""
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libgm_mock.so")
    srcs = [os.path.join(ROOT, "tests", "mock", "gm_mock.c"), os.path.join(ROOT, "oracle", "gm_oracle.c")]
    deps = srcs + [os.path.join(ROOT, "oracle", "gm_oracle.h"), os.path.join(ROOT, "include", "gm_b200.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wno-unused-parameter",
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-o", lib] + srcs + ["-lm"])
    return lib


def mock_env():
    """environment that makes a process using the host layer resolve the gm_* C ABI to the CPU mock"""
    env = dict(os.environ)
    env["LD_PRELOAD"] = build_mock()
    return env


def golden_log(g, path):
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    laser = synth.LaserSpec(g.B, 1.0 if g.B == 181 else 0.25, float(g.ranges.max()))
    log = synth.SynthLog(laser, g.ranges, g.odom, g.odom, g.stamps, {})
    synth.write_carmen(log, path)


def free_run(case, env=None, extra=()):
    from common import Golden
    from oracle.trace import read_trace
    """the reference-facing driver, linked with our libraries, over a golden log -> (fixture, trace)"""
    exe = build_driver()
    g = Golden(case)
    with tempfile.TemporaryDirectory() as td:
        lp, tp = os.path.join(td, "in.log"), os.path.join(td, "trace.bin")
        golden_log(g, lp)
        gp = os.path.join(td, "trace.gfs")
        cmd = [exe, "-mode", "trace", "-log", lp, "-out", tp, "-gfs", gp, "-probe", "2", "-maps-every", str(g.p["maps_every"])] + driver_args(g.p) + list(extra)
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert res.returncode == 0, res.stderr[-2000:]
        tr = read_trace(tp)
        tr["gfs"] = open(gp, "rb").read()  # what GridSlamProcessor::outputStream() received
    return g, tr


def check_trace(case, g, tr, exact=False):
    from common import digest_rows
    """trace of the driver linked with our libraries against the golden trace of the unmodified reference.  exact=True (CPU
    mock of the C ABI, whose arithmetic is the oracle's): floating-point results must be the reference's bit for bit."""
    ptol, rtol3, rtol6, rtol4 = (0.0, 0.0, 0.0, 0.0) if exact else (1e-4, 1e-3, 1e-6, 1e-4)
    R = tr["readings"]
    assert len(R) == g.S
    assert np.allclose(tr["angles"], g.angles, rtol=0, atol=0)
    k_dig = 0
    dig_scans = list(g.dig_scan)
    for i, r in enumerate(R):
        assert np.array_equal(r["ranges"], g.ranges[i])
        assert r["processed"] == bool(g.processed[i]), "%s[%d] gate" % (case, i)
        assert np.abs(r["pre_pose"] - g.pre_pose[i]).max() <= ptol, "%s[%d] motion-model poses" % (case, i)
        if not r["processed"]:
            continue
        assert r["resampled"] == bool(g.resampled[i]), "%s[%d] resample decision" % (case, i)
        if "smat" in r:
            ref = g.smat[i]
            assert np.abs(r["smat"][:, 10:13] - ref[:, 10:13]).max() <= ptol, "%s[%d] scan-matched poses" % (case, i)
            np.testing.assert_allclose(r["smat"][:, 8], ref[:, 8], rtol=rtol3, err_msg="%s[%d] log-likelihood" % (case, i))
            np.testing.assert_allclose(r["smat"][:, 3], ref[:, 3], rtol=rtol6, err_msg="%s[%d] optimize score via m_matcher" % (case, i))
            np.testing.assert_allclose(r["probes"][:, 4:7], g.probes[i][:len(r["probes"]), 4:7], rtol=rtol6, atol=0 if exact else 1e-9)
            # ScanMatcher::likelihood(lmax, mean, cov, ...) and optimize(mean, cov, ...) on particle 0's map
            np.testing.assert_allclose(r["like"][:5], g.like[i][:5], rtol=rtol6, atol=0 if exact else 1e-9, err_msg="%s[%d] likelihood()" % (case, i))
            np.testing.assert_allclose(r["like"][5:11], g.like[i][5:11], rtol=rtol4, atol=0 if exact else 1e-12, err_msg="%s[%d] likelihood() covariance" % (case, i))
            np.testing.assert_allclose(r["like"][11:15], g.like[i][11:15], rtol=rtol6, atol=0 if exact else 1e-9, err_msg="%s[%d] optimize(mean, cov)" % (case, i))
            np.testing.assert_allclose(r["like"][15:21], g.like[i][15:21], rtol=rtol4, atol=0 if exact else 1e-12, err_msg="%s[%d] optimize(mean, cov) covariance" % (case, i))
        if r["resampled"]:
            assert np.array_equal(r["indexes"], g.ridx[i]), "%s[%d] resample indexes" % (case, i)
            np.testing.assert_allclose(r["neff_at_resample"], g.rneff[i], rtol=rtol6)
        st = r["state"]
        assert np.abs(st[:, :3] - g.state[i][:, :3]).max() <= ptol, "%s[%d] poses" % (case, i)
        np.testing.assert_allclose(st[:, 3:5], g.state[i][:, 3:5], rtol=rtol3, atol=0 if exact else 1e-6, err_msg="%s[%d] weight / weightSum" % (case, i))
        assert np.array_equal(st[:, 5], g.state[i][:, 5]), "%s[%d] previousIndex" % (case, i)
        np.testing.assert_allclose(r["neff"], g.neff[i], rtol=rtol6)
        # the trajectory tree as the second updateTreeWeights left it (gridslamprocessor_tree.cpp:235-346): chain lengths, visit
        # counters and child counts exactly, accumulated weights like the weights they are sums of
        assert np.array_equal(r["tree"][:, [0, 4, 5]], g.tree[i][:, [0, 4, 5]]), "%s[%d] trajectory tree shape / visit counters" % (case, i)
        np.testing.assert_allclose(r["tree"][:, 1:4], g.tree[i][:, 1:4], rtol=rtol4, atol=0 if exact else 1e-9, err_msg="%s[%d] TNode::accWeight" % (case, i))
        if "digests" in r and i in dig_scans:
            k = dig_scans.index(i)
            got, want = digest_rows(r["digests"]), digest_rows(g.digests[k])
            assert np.array_equal(got[:, :5], want[:, :5]), "%s[%d] integer map digests through the host mirror" % (case, i)
            assert np.array_equal(r["geometry"], g.geometry[k]), "%s[%d] map geometry" % (case, i)
            k_dig += 1
    fin = tr["final"]
    assert np.array_equal(digest_rows(fin["digests"])[:, :5], digest_rows(g.digests[-1])[:, :5])
    assert fin["best"] == int(g.best)
    # getTrajectories() (gridslamprocessor_tree.cpp:57-147): the deep copy has the reference's shape (chain lengths, child counts,
    # readings, sharing between neighbouring particles, no aliasing of the live tree) and carries its poses and weights
    got, want = fin["trajectories"], g.trajectories
    assert np.array_equal(got[:, [0, 5, 6, 7, 8]], want[:, [0, 5, 6, 7, 8]]) and not got[:, 8].any(), "%s: getTrajectories() tree shape" % case
    np.testing.assert_allclose(got[:, 1:5], want[:, 1:5], rtol=rtol4, atol=0 if exact else 1e-3, err_msg="%s: getTrajectories() poses / accWeight" % case)
    # SensorLog::load + boundingBox (log/sensorlog.cpp): readings in the log, extent (with the reference's ymin quirk), first pose
    assert np.array_equal(fin["sensorlog_bbox"], g.sensorlog_bbox), "%s: SensorLog::boundingBox %r != %r" % (case, fin["sensorlog_bbox"], g.sensorlog_bbox)
    # ScanMatcherProcessor (scanmatcher/scanmatcherprocessor.cpp): incremental matching against one map
    assert fin["smp_poses"].shape == g.smp_poses.shape
    assert np.abs(fin["smp_poses"] - g.smp_poses).max() <= ptol, "%s: ScanMatcherProcessor poses" % case
    assert np.array_equal(digest_rows(fin["smp_digest"])[:, :5], digest_rows(g.smp_digest)[:, :5]), "%s: ScanMatcherProcessor map" % case
    assert np.array_equal(fin["smp_geometry"], g.smp_geometry), "%s: ScanMatcherProcessor map geometry" % case
    # ScanMatcher::icpOptimize / icpStep (scanmatcher.cpp:356-375, scanmatcher.h:88-165) on that map from two poses: the pose
    # handed back (heading wrapped) and the score, which is score() at the initial pose
    got, want = fin["icp"], g.icp
    assert got.shape == want.shape == (2, 8), "%s: icp record" % case
    np.testing.assert_allclose(got[:, [0, 1, 2, 4, 5, 6]], want[:, [0, 1, 2, 4, 5, 6]], rtol=0, atol=0 if exact else 1e-12, err_msg="%s: icpOptimize / icpStep poses" % case)
    np.testing.assert_allclose(got[:, [3, 7]], want[:, [3, 7]], rtol=0 if exact else 1e-9, atol=0, err_msg="%s: icpOptimize / icpStep scores" % case)
    # ScanMatcher::registerScan's return value (scanmatcher.cpp:317-341): the reference sums the entropy change update by update,
    # the device forms entropy(after) - entropy(before) over the map -- equal up to the rounding of ~1e5 additions (and 0 without
    # free-space tracing, where the reference adds nothing)
    np.testing.assert_allclose(fin["register_gain"], g.register_gain, rtol=1e-9, atol=1e-9, err_msg="%s: registerScan return value" % case)
    # integrateScanSequence (gridslamprocessor_tree.cpp:149-224) on a second, small filter: poses, weights, trajectory depth, the
    # leaves' poses and the (resized) map geometry of every particle
    got, want = fin["integrate"], g.integrate
    assert got.shape == want.shape and want[0, 15] >= 2, "%s: integrateScanSequence record" % case
    assert np.array_equal(got[:, 5], want[:, 5]) and np.array_equal(got[:, 9:], want[:, 9:]), "%s: integrateScanSequence depth / map geometry" % case
    np.testing.assert_allclose(got[:, :5], want[:, :5], rtol=0, atol=0 if exact else 1e-12, err_msg="%s: integrateScanSequence poses / weights" % case)
    np.testing.assert_allclose(got[:, 6:9], want[:, 6:9], rtol=0, atol=0 if exact else 1e-12, err_msg="%s: integrateScanSequence leaf poses" % case)
    # the .gfs trace (gridslamprocessor.cpp:454-475, 570-606; gridslamprocessor.hxx:159-167): the same records in the same order;
    # with the reference's arithmetic underneath (mock) the same bytes
    tags = " ".join(l.split(b" ", 1)[0].decode() for l in tr["gfs"].splitlines() if l)
    assert tags == str(g.gfs_tags), "%s: .gfs record sequence" % case
    if exact:
        assert len(tr["gfs"]) == int(g.gfs_bytes) and hashlib.sha256(tr["gfs"]).hexdigest() == str(g.gfs_sha256), "%s: .gfs bytes" % case
    # the ROS node's updateMap (slam_gmapping.cpp:866-993) as the driver restates it: a standalone ScanMatcher replays the best
    # trajectory into a fresh host ScanMatcherMap (computeActiveArea + registerScan), every cell is read back and thresholded
    got, want = digest_rows(fin["ros_map_digest"]), digest_rows(g.ros_map_digest)
    assert np.array_equal(got[:, :5], want[:, :5]), "%s: map rebuilt from the best trajectory, integer digest" % case
    if exact:
        assert np.array_equal(got, want), "%s: map rebuilt from the best trajectory, float accumulators" % case
    assert np.array_equal(fin["ros_map_info"], g.ros_map_info), "%s: occupancy grid of the rebuilt map (size, origin, cell counts, hash)" % case
    assert k_dig >= 1


def free_port():
    """a TCP port nobody listens on right now (torch.distributed rendezvous of the multi-process CPU tests)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]
