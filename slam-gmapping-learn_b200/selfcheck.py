"""Follow a compact golden trace of the UNMODIFIED reference (tests/golden/*.npz made by tools/make_golden.py `run_long`:
Neff and the resampling decision for every scan, every resample index vector, poses / log-weights of all particles at the
kept scans, the final integer map digests) with the product's GridSlamProcessor -- single GPU or sharded over ranks.

Used by the tests (tests/host_scenarios.py) and, before its timed region on more than one GPU, by bench.py: a bench line
is only printed for a filter whose sharded path (peer-memory row exchange, pull migration of resampled parents) has just
reproduced the reference on this very box.  No oracle, no reference code: the fixture is data."""
from __future__ import annotations

import importlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
DIGEST_KEYS = ("patches", "cells", "sum_n", "sum_visits", "hash_counts")


def load_fixture(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    p = json.loads(str(z["params"]))
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    laser = getattr(synth, str(z["laser"])) if "laser" in z.files else synth.LMS200
    step = float(z["step"]) if "step" in z.files else 0.25
    log = synth.make_log("room", laser, scans=int(z["scans"]), step=step, seed=int(z["log_seed"]))
    assert np.array_equal(log.laser.angles(), z["angles"]), "the synthetic log generator no longer reproduces the fixture's log"
    return z, p, log


def follow(name, limit=None, rank=0, world=1, session_id=None, pose_tol=1e-4, weight_tol=1e-3, **kw):
    """run the filter over the fixture's log and assert, scan by scan, that it stays on the reference's trajectory.
    Returns dict(worst_pose, worst_weight, scans, resamples, migrated) for this rank."""
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    z, p, log = load_fixture(name)
    f = hostapi.Processor(z["angles"], p["particles"], z["head_init"], bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                          seed=p["seed"], rank=rank, world=world, nccl_id=session_id, **dict({k: p[k] for k in hostapi.PARAM_ORDER}, **kw))
    kept = {int(s): k for k, s in enumerate(z["kept"])}
    rsc = {int(s): k for k, s in enumerate(z["ridx_scans"])}
    worst_pose = worst_w = 0.0
    migrated = resamples = 0
    total = int(z["scans"])
    scans = total if limit is None else min(int(limit), total)
    for i in range(scans):
        done = f.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
        assert done == bool(z["processed"][i]), "%s scan %d: update gate" % (name, i)
        assert f.last_resampled() == bool(z["resampled"][i]), "%s scan %d: resampling decision" % (name, i)
        assert abs(f.neff() - z["neff"][i]) <= 1e-6 * z["neff"][i], "%s scan %d: Neff %r vs %r" % (name, i, f.neff(), float(z["neff"][i]))
        if i in rsc:
            resamples += 1
            assert np.array_equal(f.indexes(), z["ridx"][rsc[i]]), "%s scan %d: resample indexes" % (name, i)
            if world > 1:
                migrated += int(f.stats()["migrated_particles"])
        if i in kept:
            st = f.state(); k = kept[i]
            worst_pose = max(worst_pose, float(np.abs(st[:, :3] - z["poses"][k]).max()))
            worst_w = max(worst_w, float(np.abs(st[:, 3:5] - z["weights"][k]).max() / max(1.0, np.abs(z["weights"][k]).max())))
            assert np.array_equal(st[:, 5], z["prev"][k]), "%s scan %d: previousIndex" % (name, i)
    assert worst_pose < pose_tol and worst_w < weight_tol, (name, worst_pose, worst_w)
    if scans == total:
        lo, hi = f.local_range()
        d = f.map_digests()
        want = np.stack([z["digests"][k].astype(np.uint64) for k in DIGEST_KEYS], axis=-1)[lo:hi]
        assert np.array_equal(d[:, :5], want), "%s: final maps, (n, visits) of every cell of every local particle" % name
        assert f.best() == int(z["best"])
    f.close()
    return dict(worst_pose=worst_pose, worst_weight=worst_w, scans=scans, resamples=resamples, migrated=migrated)
