#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref/ref_driver).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tools/make_golden.py

For every case: write a seeded synthetic CARMEN log (slam-gmapping-learn_b200/synth.py), run the
reference driver in trace mode on it, and store the per-stage trace (poses after the motion model,
optimize/likelihoodAndScore results, weights, Neff, resample indexes, per-particle map digests and
geometry) as one compressed npz.  The fixtures are what pins the oracle (tests/test_oracle_vs_reference.py)
and what the GPU parity tests compare against; the GPU box never sees /root/reference.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("slam-gmapping-learn_b200.synth")
from oracle.trace import read_trace  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# name -> (log kwargs, filter parameters).  Parameters not listed keep gfs-carmen's defaults
# (reference: openslam_gmapping/gfs-carmen/gfs-carmen.cpp:52-84), see oracle/ref_driver.cpp Opts.
CASES = {
    # ROS-default motion noise => weights spread => resampling fires; every scan passes the gate
    "room181_n30": (dict(world="room", laser="LMS200", scans=40, step=0.5),
                    dict(particles=30, linearUpdate=0.4, angularUpdate=0.2, srr=0.1, srt=0.2, str=0.1, stt=0.2, maps_every=5)),
    # small initial map => Map::resize on the first scans; linearUpdate 1.0 => ungated readings in between
    "room181_resize": (dict(world="room", laser="LMS200", scans=36, step=0.5),
                       dict(particles=12, xmin=-4, ymin=-4, xmax=4, ymax=4, linearUpdate=1.0, angularUpdate=0.5, maps_every=3)),
    # the benchmark's scan shape: 1081 beams, 0.25 deg, beams beyond maxUrange are traced free without a hit
    "room1081_n8": (dict(world="room", laser="UTM30", scans=12, step=0.5),
                    dict(particles=8, maxrange=30.0, maxUrange=29.0, linearUpdate=0.4, angularUpdate=0.2,
                         srr=0.1, srt=0.2, str=0.1, stt=0.2, maps_every=4)),
    # kernelSize 2, likelihoodSkip 2, endpoint-only maps (generateMap false), minimumScore rejects
    "room181_k2_lskip": (dict(world="room", laser="LMS200", scans=16, step=0.5),
                         dict(particles=6, kernelSize=2, lskip=2, generate_map=False, minimumScore=60.0,
                              linearUpdate=0.4, angularUpdate=0.2, srr=0.05, srt=0.1, str=0.05, stt=0.1, maps_every=4)),
    # map frame with its lower corner half a cell off the origin and the room around (20, 20): the reference's
    # ipfree = world2map(pfree - phit) (scanmatcher.h:208-210, 320-322) then lands ON the neighbourhood of the hit, so the
    # free-cell test reads allocated, visited cells, and score()'s free-cell offset depends on the beam direction
    "room181_halfcell": (dict(world="room", laser="LMS200", scans=14, step=0.5, shift=(20.0, 20.0, 0.0)),
                         dict(particles=6, xmin=0.025, ymin=0.025, xmax=40.025, ymax=40.025, linearUpdate=0.4, angularUpdate=0.2,
                              srr=0.05, srt=0.1, str=0.05, stt=0.1, maps_every=4)),
    # BASELINE.json's fifth configuration uses a 0.025 m grid
    "room181_d025": (dict(world="room", laser="LMS200", scans=12, step=0.5),
                     dict(particles=6, delta=0.025, xmin=-50., ymin=-50., xmax=50., ymax=50., linearUpdate=0.4, angularUpdate=0.2,
                          srr=0.05, srt=0.1, str=0.05, stt=0.1, maps_every=4)),
}

DEFAULTS = dict(particles=30, xmin=-100., ymin=-100., xmax=100., ymax=100., delta=0.05, sigma=0.05, maxrange=80., maxUrange=80.,
                lstep=0.05, astep=0.05, lsigma=0.075, ogain=3., kernelSize=1, iterations=5, lskip=0,
                srr=0.01, srt=0.01, str=0.01, stt=0.01, linearUpdate=1.0, angularUpdate=0.5, resampleThreshold=0.5,
                minimumScore=0.0, period=-1.0, seed=12345, generate_map=True, maps_every=0)


def run_case(name, log_kw, par):
    p = dict(DEFAULTS); p.update(par)
    lk = dict(log_kw); laser = getattr(synth, lk.pop("laser"))
    shift = lk.pop("shift", None)
    log = synth.make_log(laser=laser, **lk)
    if shift is not None:  # the same room somewhere else in the world
        log.odom = np.round(log.odom + np.asarray(shift)[None, :], 6)
    with tempfile.TemporaryDirectory() as td:
        lp, tp = os.path.join(td, "in.log"), os.path.join(td, "trace.bin")
        synth.write_carmen(log, lp)
        gp = os.path.join(td, "trace.gfs")
        cmd = [REF, "-mode", "trace", "-log", lp, "-out", tp, "-gfs", gp, "-probe", "2", "-maps-every", str(p["maps_every"])]
        for k in ("particles", "seed", "kernelSize", "iterations", "lskip"):
            cmd += ["-" + k, str(int(p[k]))]
        for k in ("xmin", "ymin", "xmax", "ymax", "delta", "sigma", "maxrange", "maxUrange", "lstep", "astep", "lsigma", "ogain",
                  "srr", "srt", "str", "stt", "linearUpdate", "angularUpdate", "resampleThreshold", "minimumScore", "period"):
            cmd += ["-" + k, repr(float(p[k]))]
        if not p["generate_map"]:
            cmd.append("-no-generateMap")
        out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
        tr = read_trace(tp)
        gfs = open(gp, "rb").read()  # the .gfs file the reference itself writes (GridSlamProcessor::outputStream())
    R = tr["readings"]
    S, N, B = len(R), tr["head"]["particles"], tr["head"]["beams"]
    a = dict(
        params=np.array(json.dumps(p)), head_init=tr["head"]["init"], angles=tr["angles"],
        odom=np.stack([r["odom"] for r in R]), stamps=np.array([r["stamp"] for r in R]),
        ranges=np.stack([r["ranges"] for r in R]), pre_pose=np.stack([r["pre_pose"] for r in R]),
        processed=np.array([r["processed"] for r in R]), resampled=np.array([r["resampled"] for r in R]),
    )
    assert a["ranges"].shape == (S, B) and a["pre_pose"].shape == (S, N, 3)
    smat = np.full((S, N, 15), np.nan); probes = np.full((S, 10, 8), np.nan); like = np.full((S, 22), np.nan)
    state = np.full((S, N, 6), np.nan); neff = np.full(S, np.nan); weights = np.full((S, N), np.nan); tree = np.full((S, N, 6), np.nan)
    ridx = np.zeros((S, N), dtype=np.uint32); rw = np.full((S, N), np.nan); rneff = np.full(S, np.nan)
    dig_scan, digs, geos = [], [], []
    for i, r in enumerate(R):
        if "smat" in r: smat[i] = r["smat"]
        if "probes" in r and len(r["probes"]): probes[i, :len(r["probes"])] = r["probes"]
        if "like" in r: like[i] = r["like"]
        if "state" in r: state[i] = r["state"]; neff[i] = r["neff"]; weights[i] = r["weights"]; tree[i] = r["tree"]
        if "indexes" in r: ridx[i] = r["indexes"]; rw[i] = r["w_at_resample"]; rneff[i] = r["neff_at_resample"]
        if "digests" in r: dig_scan.append(i); digs.append(r["digests"]); geos.append(r["geometry"])
    # the driver also dumps all maps after the last reading
    dig_scan.append(S - 1); digs.append(tr["final"]["digests"]); geos.append(tr["final"]["geometry"])
    a.update(smat=smat, probes=probes, like=like, state=state, tree=tree, neff=neff, weights=weights, ridx=ridx, rw=rw, rneff=rneff,
             dig_scan=np.array(dig_scan), digests=np.stack(digs), geometry=np.stack(geos),
             best=np.array(tr["final"]["best"]), ros_map_digest=tr["final"]["ros_map_digest"], ros_map_info=tr["final"]["ros_map_info"],
             trajectories=tr["final"]["trajectories"], smp_poses=tr["final"]["smp_poses"], smp_digest=tr["final"]["smp_digest"],
             smp_geometry=tr["final"]["smp_geometry"], sensorlog_bbox=tr["final"]["sensorlog_bbox"],
             integrate=tr["final"]["integrate"], icp=tr["final"]["icp"], register_gain=tr["final"]["register_gain"])
    a.update(gfs_sha256=np.array(hashlib.sha256(gfs).hexdigest()), gfs_bytes=np.array(len(gfs)),
             gfs_tags=np.array(" ".join(l.split(b" ", 1)[0].decode() for l in gfs.splitlines() if l)))
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **a)
    print("%-18s readings %3d processed %3d resamples %2d  %6.1f KB   %s" % (
        name, S, int(a["processed"].sum()), int(a["resampled"].sum()), os.path.getsize(path) / 1024, out.strip()[:80]))


def run_long(name="room181_long", scans=1000, particles=30, laser="LMS200", log_seed=7, every=10, step=0.25, **par):
    """BASELINE.json configs[0]/[1]: 30 particles, 181 beams, 1000 scans.  Too long for a full per-stage trace in the
    repository: the fixture keeps, per reading, the gate / resample flags, Neff, the resample indexes and the poses +
    log-weights of all particles (float64), plus the final map digests."""
    p = dict(DEFAULTS); p.update(dict(particles=particles, linearUpdate=0.2, angularUpdate=0.1, srr=0.05, srt=0.1, str=0.05, stt=0.1, maps_every=0))
    p.update(par)
    log = synth.make_log(world="room", laser=getattr(synth, laser), scans=scans, step=step, seed=log_seed)
    with tempfile.TemporaryDirectory() as td:
        lp, tp = os.path.join(td, "in.log"), os.path.join(td, "trace.bin")
        synth.write_carmen(log, lp)
        cmd = [REF, "-mode", "trace", "-log", lp, "-out", tp, "-probe", "0", "-maps-every", "0"]
        for k in ("particles", "seed", "kernelSize", "iterations", "lskip"):
            cmd += ["-" + k, str(int(p[k]))]
        for k in ("xmin", "ymin", "xmax", "ymax", "delta", "sigma", "maxrange", "maxUrange", "lstep", "astep", "lsigma", "ogain",
                  "srr", "srt", "str", "stt", "linearUpdate", "angularUpdate", "resampleThreshold", "minimumScore", "period"):
            cmd += ["-" + k, repr(float(p[k]))]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        tr = read_trace(tp)
    R = tr["readings"]
    S, N = len(R), particles
    state = np.stack([r["state"] for r in R]); neff = np.array([r["neff"] for r in R])
    resampled = np.array([r["resampled"] for r in R]); processed = np.array([r["processed"] for r in R])
    ridx = np.zeros((S, N), dtype=np.uint16)
    for i, r in enumerate(R):
        if "indexes" in r: ridx[i] = r["indexes"]
    path = os.path.join(GOLD, name + ".npz")
    # Neff and the flags for every scan; poses / weights every 10th scan, at the last one and where the filter resampled
    rs = np.nonzero(resampled)[0]
    keep = np.union1d(np.union1d(np.arange(0, S, every), [S - 1]), rs)
    np.savez_compressed(path, params=np.array(json.dumps(p)), head_init=tr["head"]["init"], angles=tr["angles"],
                        log_seed=np.array(log_seed), laser=np.array(laser), step=np.array(step), scans=np.array(scans), processed=processed, resampled=resampled, neff=neff,
                        ridx_scans=rs, ridx=ridx[rs], kept=keep,
                        poses=state[keep, :, :3], weights=state[keep, :, 3:5].astype(np.float64), prev=state[keep, :, 5].astype(np.int16),
                        digests=tr["final"]["digests"], geometry=tr["final"]["geometry"], best=np.array(tr["final"]["best"]))
    print("%-18s readings %4d processed %4d resamples %3d  %7.1f KB" % (name, S, int(processed.sum()), int(resampled.sum()), os.path.getsize(path) / 1024))


def run_units(name="units"):
    """Known-answer vectors of the building blocks (Bresenham lines, systematic resampling, Gaussian sampler, motion model,
    world2map) straight from the reference: oracle/_ref/ref_units (oracle/ref_units.cpp) -> tests/golden/units.npz."""
    import struct
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_units")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    with tempfile.TemporaryDirectory() as td:
        bp = os.path.join(td, "units.bin")
        subprocess.run([exe, bp], check=True)
        buf = open(bp, "rb").read()
    off = 0
    out = {"units_sha256": np.array(hashlib.sha256(buf).hexdigest()), "units_bytes": np.array(len(buf))}
    rs_n, rs_seed, rs_w, rs_idx = [], [], [], []
    ra_hdr, ra_w, ra_idx = [], [], []
    while off < len(buf):
        tag = buf[off:off + 4].decode(); n = struct.unpack_from("<Q", buf, off + 4)[0]
        p = buf[off + 12:off + 12 + n]; off += 12 + n
        if tag == "LINE": out["lines"] = np.frombuffer(p, dtype="<i8").reshape(-1, 6).copy()
        elif tag == "RSHD": h = np.frombuffer(p, dtype="<f8"); rs_n.append(int(h[0])); rs_seed.append(int(h[1]))
        elif tag == "RSWT": rs_w.append(np.frombuffer(p, dtype="<f8").copy())
        elif tag == "RSIX": rs_idx.append(np.frombuffer(p, dtype="<u4").copy())
        elif tag == "RAHD": ra_hdr.append(np.frombuffer(p, dtype="<f8").astype(np.int64))
        elif tag == "RAWT": ra_w.append(np.frombuffer(p, dtype="<f8").copy())
        elif tag == "RAIX": ra_idx.append(np.frombuffer(p, dtype="<u4").copy())
        elif tag == "GAUS": out["gauss"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "MOTN": out["motion"] = np.frombuffer(p, dtype="<f8").copy()
        elif tag == "W2MP": out["w2m"] = np.frombuffer(p, dtype="<f8").reshape(-1, 8).copy()
        elif tag == "HMAP": out["host_map"] = np.frombuffer(p, dtype="<f8").reshape(-1, 16).copy()
        elif tag == "PTOP": out["point_ops"] = np.frombuffer(p, dtype="<f8").reshape(-1, 23).copy()
        else: raise ValueError(tag)
    out.update(rs_n=np.array(rs_n), rs_seed=np.array(rs_seed), rs_w=np.concatenate(rs_w), rs_idx=np.concatenate(rs_idx))
    out.update(ra_hdr=np.stack(ra_hdr), ra_w=np.concatenate(ra_w), ra_idx=np.concatenate(ra_idx))  # rows of ra_hdr: n, n_out, seed
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-18s lines %d resample cases %d gaussians %d motion rows %d world2map rows %d  %6.1f KB" % (
        name, len(out["lines"]), len(rs_n), len(out["gauss"]) - 1, (len(out["motion"]) - 1) // 12, len(out["w2m"]), os.path.getsize(path) / 1024))


def main():
    if "long" in sys.argv[1:]:
        run_long(); return
    if "c3" in sys.argv[1:]:
        # BASELINE.json configs[2] (SURVEY.md 8d "a >= 50-scan prefix of C3"): 512 particles x 1081 beams, the benchmark's log and
        # parameters (ROS-default motion noise: the filter resamples); ~12 minutes of the reference on one core
        # lobsGain: the weights are softmax(l / (ogain * N)) (gridslamprocessor.hxx:93), so gfs-carmen's ogain 3, tuned for 30
        # particles, never lets Neff drop at 512; the benchmark keeps the 30-particle configuration's effective gain, ogain * N = 90
        run_long("room1081_c3", scans=56, particles=512, laser="UTM30", log_seed=1, every=8, maxrange=30.0, maxUrange=29.0,
                 srr=0.1, srt=0.2, str=0.1, stt=0.2, ogain=90.0 / 512); return
    if "n32" in sys.argv[1:]:
        # 32 particles (divisible by 2, 4 and 8 ranks) for the sharded-parity prelude of bench.py
        run_long("room181_n32", scans=40, particles=32, laser="LMS200", log_seed=1, every=5, step=0.5, linearUpdate=0.4, angularUpdate=0.2,
                 srr=0.1, srt=0.2, str=0.1, stt=0.2); return
    if "units" in sys.argv[1:]:
        run_units(); return
    if not os.path.exists(REF):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    os.makedirs(GOLD, exist_ok=True)
    for name, (lk, par) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run_case(name, lk, par)


if __name__ == "__main__":
    main()
