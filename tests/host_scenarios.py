"""Scenarios of the C++ host layer driven through slam-gmapping-learn_b200/hostapi.py (ctypes view of GridSlamProcessor) and the
gfs_b200 CLI.  The `-m gpu` tests of tests/test_host_layer.py call them in process on the B200; tests/test_host_layer_cpu.py
runs this file in a subprocess with the CPU mock of the C ABI preloaded (tests/mock/gm_mock.c), where everything the host
layer computes itself must still match the unmodified reference's trace.

    python tests/host_scenarios.py clone | order | cli | lazy_tree | long[:SCANS]
"""
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from common import Golden, digest_rows  # noqa: E402
from host_common import HOST, build_host, golden_log  # noqa: E402

synth = importlib.import_module("slam-gmapping-learn_b200.synth")


def scenario_cli():
    build_host()
    g = Golden("room181_resize")
    with tempfile.TemporaryDirectory() as td:
        lp, gp = os.path.join(td, "in.log"), os.path.join(td, "out.gfs")
        golden_log(g, lp)
        res = subprocess.run([os.path.join(HOST, "gfs_b200"), "-filename", lp, "-outfilename", gp, "-particles", "12", "-randseed", "12345",
                              "-xmin", "-4", "-ymin", "-4", "-xmax", "4", "-ymax", "4"], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        info = json.loads(res.stdout.strip().splitlines()[-1])
        assert info["processed"] == int(g.processed.sum()) and info["read"] == g.S
        txt = open(gp).read()
        for tag in ("ODOM ", "ODO_UPDATE 12 ", "FRAME ", "LASER_READING 181 ", "SM_UPDATE 12 ", "NEFF "):
            assert tag in txt, tag
        neff = [float(l.split()[1]) for l in txt.splitlines() if l.startswith("NEFF")]
        want = g.neff[g.processed & ~np.isnan(g.smat[:, 0, 0])]
        # the .gfs NEFF record is written before resampling: compare with the reference where no resample happened
        assert len(neff) == len(want)
        # gfs_nogui's option names and -autosize (gui/gsp_thread.cpp:97-144, 315-324, 634-642): the map is sized from the poses
        # in the log (SensorLog::boundingBox, golden) grown by 3 x maxrange, and the filter starts at the first range pose
        res = subprocess.run([os.path.join(HOST, "gfs_b200"), "-filename", lp, "-particles", "4", "-randseed", "12345", "-autosize", "-maxrange", "20",
                              "-maxUrange", "20", "-lobsGain", "3", "-generateMap", "-scans", "3", "-linearOdometryReliability", "0", "-llsamplerange", "0.01"],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        info = json.loads(res.stdout.strip().splitlines()[-1])
        bb = g.sensorlog_bbox
        assert np.allclose(info["map"], [bb[1] - 60, bb[2] - 60, bb[3] + 60, bb[4] + 60], atol=1e-3) and info["processed"] == 3


def scenario_clone():
    """GridSlamProcessor::clone() (gridslamprocessor.cpp:38-97): the copy, fed the same readings and the same drand48
    stream, stays bit-identical to the original -- poses, weights, resample indexes and every particle map."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    libc = ctypes.CDLL("libc.so.6")
    libc.seed48.restype = ctypes.POINTER(ctypes.c_ushort * 3)
    g = Golden("room181_n30")
    p = g.p
    a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"], seed=p["seed"],
                          **{k: p[k] for k in hostapi.PARAM_ORDER})
    for i in range(12):
        a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    before = a.check_maps()
    b = a.clone()
    assert np.array_equal(a.state(), b.state()) and np.array_equal(a.map_digests(), b.map_digests())
    # getTrajectories() (what the copy was made with: each node notes its copy under a stamp, no map): the original's tree and the
    # clone's have the same number of distinct nodes and the same rows seen from every particle, call after call
    na, nb = a.copy_trajectories(), b.copy_trajectories()
    assert na == nb == a.copy_trajectories() and na >= 12 + 1 and np.array_equal(a.tree_rows(), b.tree_rows())
    # the clone is a second set of directories over the same patches: no patch was copied, every count is consistent
    assert a.check_maps() == before and b.check_maps() == before
    saved = (ctypes.c_ushort * 3)(*libc.seed48((ctypes.c_ushort * 3)(0, 0, 0)).contents)  # read the drand48 state ...
    libc.seed48(saved)                                                                      # ... and put it back
    for i in range(12, 24):
        a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    resampled = a.resamples()
    libc.seed48(saved)
    for i in range(12, 24):
        b.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    assert np.array_equal(a.state(), b.state()), "clone diverged"
    assert np.array_equal(a.map_digests(), b.map_digests())
    shared = a.check_maps()  # both filters wrote their own copies of the patches they touched (copy-on-write), counts still consistent
    assert b.check_maps() == shared
    assert b.resamples() == resampled and a.neff() == b.neff()
    assert np.abs(a.state()[:, :3] - g.state[23][:, :3]).max() < 1e-4  # and both still follow the reference
    # occupancy grid of the best particle through the C++ API
    grid, origin = a.occupancy_grid(a.best(), 0.25)
    assert grid.shape == (4000, 4000) and abs(origin[0] + 100.0) < 1e-9 and abs(origin[1] + 100.0) < 1e-9
    assert (grid == 100).sum() > 200 and (grid == 0).sum() > 20000 and (grid == -1).sum() > 15000000
    x, y = g.state[23][a.best(), 0], g.state[23][a.best(), 1]
    assert grid[int(round((y + 100) / 0.05)), int(round((x + 100) / 0.05))] == 0  # the robot stands in free space
    assert a.occupancy(a.best(), x, y) < 0.1
    b.close()  # the clone gives its references back; what only it referenced returns to the free list
    alone = a.check_maps()
    assert alone <= shared
    a.close()


def scenario_order():
    """setearlyRegistration(false) keeps the reference's order (copy the resampled particles, then register the scan in
    every child); the default registers first and copies afterwards.  Same drand48 stream in, bit-identical filter out:
    poses, weights, resample indexes, every map -- and both follow the reference trace."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    g = Golden("room181_k2_lskip")
    p = g.p
    runs = []
    # early registration (the default), the reference's order, and early registration queued on the device right behind the scan
    # match (setregisterWithScanMatch: the default of a sharded filter)
    for early, fused in ((True, False), (False, False), (True, True)):
        a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                              seed=p["seed"], early_registration=early, register_with_scanmatch=fused, **{k: p[k] for k in hostapi.PARAM_ORDER})
        idx = []
        for i in range(len(g.ranges)):
            a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
            if a.last_resampled():
                idx.append(a.indexes().copy())
        runs.append((a.state(), a.map_digests(), a.resamples(), a.neff(), idx))
        a.close()
    (s0, d0, r0, n0, i0) = runs[0]
    for (s1, d1, r1, n1, i1) in runs[1:]:
        assert r0 == r1 and r0 > 0 and n0 == n1
        assert np.array_equal(s0, s1) and np.array_equal(d0, d1)
        assert len(i0) == len(i1) and all(np.array_equal(x, y) for x, y in zip(i0, i1))
    assert np.abs(s0[:, :3] - g.state[len(g.ranges) - 1][:, :3]).max() < 1e-4


def scenario_adapt():
    """processScan(reading, adaptParticles) (gridslamprocessor.hxx:153-156 -> particlefilter.h:196-203): the resampling at a reading
    continues with another number of particles, fewer (30 -> 20) and then more (20 -> 34, room reserved with setmaxParticles).
    The index vectors are the reference's arithmetic (pinned by tests/golden/units.npz: resampleIndexes(w, nparticles)); here: the
    filter carries on, child j continues parent idx[j] (pose, weightSum, previousIndex, map), the reference counts of the patch
    pool stay consistent, and the three registration orders (early, the reference's, queued behind the scan match) agree bit
    for bit.  (The unmodified reference cannot serve as the oracle here: with fewer particles its bookkeeping leaks the dead
    leaves beyond the new count and its own propagateWeights() assertion fires; with more it reads past the particle vector.)"""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    g = Golden("room181_n30")
    p = g.p
    schedule = {6: 20, 7: 20, 12: 34}  # (34 > the 30 the filter starts with: the particle vector must have been reserved for it)
    runs = []
    for early, fused in ((True, False), (False, False), (True, True)):
        par = dict({k: p[k] for k in hostapi.PARAM_ORDER}, resampleThreshold=2.0)  # Neff < 2 N: every processed reading resamples
        a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                              seed=p["seed"], early_registration=early, register_with_scanmatch=fused, max_particles=36, **par)
        counts, trail = [], []
        for i in range(16):
            before = a.state()
            done = a.process_scan(g.ranges[i], g.odom[i], g.stamps[i], adapt=schedule.get(i, 0))
            st = a.state()
            counts.append(len(st))
            if done and a.last_resampled():
                idx = a.indexes()
                want = schedule.get(i, 0) or len(before)
                assert len(idx) == want == len(st), (i, len(idx), want, len(st))
                assert np.array_equal(st[:, 5], idx) and np.all(idx < len(before)) and np.all(np.diff(idx.astype(np.int64)) >= 0)
                assert np.all(st[:, 3] == 0)  # setWeight(0) after a resampling
                trail.append(idx.copy())
            assert a.check_maps() > 0
        assert counts[5] == 30 and counts[6] == 20 and counts[11] == 20 and counts[12] == 34 and counts[-1] == 34, counts
        assert len(a.map_digests()) == 34 and len(a.tree_rows()) == 34
        runs.append((a.state(), a.map_digests(), trail, a.neff()))
        a.close()
    s0, d0, t0, n0 = runs[0]
    for s1, d1, t1, n1 in runs[1:]:
        assert np.array_equal(s0, s1) and np.array_equal(d0, d1) and n0 == n1
        assert len(t0) == len(t1) and all(np.array_equal(x, y) for x, y in zip(t0, t1))
    return "30 -> 20 -> 34 particles, %d resamplings" % len(t0)


def scenario_lazy_tree():
    """setlazyTreeWeights(true): processScan only normalises the weights; TNode::accWeight is brought up to date on demand.
    The filter itself must not notice (same poses, weights, resampling, maps), and after refreshTreeWeights() the tree is the
    eager one -- which, per the golden trace, is the reference's."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    g = Golden("room181_k2_lskip")
    p = g.p
    runs = []
    for lazy in (False, True):
        a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                              seed=p["seed"], lazy_tree=lazy, **{k: p[k] for k in hostapi.PARAM_ORDER})
        for i in range(len(g.ranges)):
            a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
        stale = a.tree_rows()
        a.refresh_tree()
        runs.append((a.state(), a.map_digests(), a.resamples(), a.neff(), stale, a.tree_rows()))
        a.close()
    (s0, d0, r0, n0, stale0, t0), (s1, d1, r1, n1, stale1, t1) = runs
    assert r0 == r1 and r0 > 0 and n0 == n1 and np.array_equal(s0, s1) and np.array_equal(d0, d1)
    assert np.array_equal(stale0, t0)            # eager: nothing left to refresh
    assert not np.array_equal(stale1[:, 1:4], t1[:, 1:4])  # lazy: the weights in the tree were stale until asked for
    assert np.array_equal(t0, t1)
    last = len(g.ranges) - 1
    assert np.array_equal(t0[:, [0, 4, 5]], g.tree[last][:, [0, 4, 5]])
    np.testing.assert_allclose(t0[:, 1:4], g.tree[last][:, 1:4], rtol=1e-4, atol=1e-9)


def scenario_long(limit=None, fixture="room181_long"):
    """BASELINE.json configs[0]/[1]: 30 particles, 181-beam scans, 1000 scans, free running through the C++ API against the
    unmodified reference (fixture: Neff and flags for every scan, all poses / log-weights every 10th scan): the filter never
    leaves the reference's trajectory -- same resampling scans, same indexes, poses within 1e-4, log-weights within 1e-3.
    Other fixtures of the same kind: room1081_c3 (BASELINE.json configs[2], 512 particles x 1081 beams, 56 scans of the
    benchmark's log with the benchmark's parameters) and room181_n32 (32 particles, the sharded-parity prelude of bench.py)."""
    build_host()
    selfcheck = importlib.import_module("slam-gmapping-learn_b200.selfcheck")
    r = selfcheck.follow(fixture, limit)
    return r["worst_pose"], r["worst_weight"], r["resamples"]


def scenario_c3(limit=None):
    return scenario_long(limit, "room1081_c3")


def scenario_n32(limit=None):
    return scenario_long(limit, "room181_n32")


if __name__ == "__main__":
    name, _, arg = sys.argv[1].partition(":")
    fn = {"cli": scenario_cli, "clone": scenario_clone, "order": scenario_order, "long": scenario_long, "lazy_tree": scenario_lazy_tree, "c3": scenario_c3,
          "n32": scenario_n32, "adapt": scenario_adapt}[name]
    out = fn(int(arg)) if arg else fn()
    print("scenario %s ok %s" % (sys.argv[1], "" if out is None else out))
