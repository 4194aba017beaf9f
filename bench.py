#!/usr/bin/env python
"""bench.py -- scans/sec of the per-scan particle update at 4096 particles x 1081 beams (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--particles P] [--impl reference]

A "step" is one processScan: scan matching of every particle (hill-climb + likelihood), weight
normalisation / Neff, resampling when Neff drops, and the grid update (registerScan) -- on the next scan of
a seeded synthetic 1081-beam log (18 x 12 m room, 0.05 m grid, 200 x 200 m map).  Prints ONE JSON line.

  value      scans/s from device time only (CUDA events around the kernels of each stage, maps resident in HBM)
  e2e        scans/s through the C ABI with HOST buffers (ranges + poses in, poses + likelihoods out, every
             step), wall clock bracketed by synchronisation
  roofline   the scan-match kernel: beam-evaluations x 144 B (3x3 window of 16-byte cells)
             over its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline   the unmodified reference (oracle/_ref/ref_driver) on one host core, bounded sample,
             scaled linearly in the particle count (the reference is one loop over particles)

`--impl reference` times the reference's CPU implementation on all host cores (independent particle
slices, one single-threaded reference process per core -- the reference itself has no threading).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scans/sec at 4096 particles x 1081 beams"
BYTES_PER_BEAM_EVAL = 144  # (2k+1)^2 x 16 B at kernelSize 1
PARAMS = dict(maxrange=30.0, maxUrange=29.0, sigma=0.05, kernelSize=1, lstep=0.05, astep=0.05, iterations=5, lsigma=0.075,
              ogain=3.0, lskip=0, srr=0.01, srt=0.01, str=0.01, stt=0.01, resampleThreshold=0.5, delta=0.05,
              xmin=-100.0, ymin=-100.0, xmax=100.0, ymax=100.0)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.stop = False; self.t = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True); self.t.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) >= 6 and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_log(scans, seed=1):
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    return synth, synth.make_log("room", synth.UTM30, scans=scans, step=0.25, seed=seed)


class HostFilter:
    """The host side of processScan over the C ABI (reference: gridslamprocessor.cpp:425-662): motion-model
    sampling, weights bookkeeping and the resample decision stay on the host; the four heavy stages are gm_* calls."""

    def __init__(self, gm, particles, angles, init_pose, device=0, seed=12345):
        self.n = particles
        # pool: ~160 patches per particle covers the 18 x 12 m room with private copies of the active area
        self.ctx = gm.GmContext(particles, device=device, pool_patches=max(8192, particles * 256))
        self.ctx.set_laser(angles)
        self.ctx.set_matching(urange=PARAMS["maxUrange"], range_=PARAMS["maxrange"], sigma=PARAMS["sigma"], kernel=PARAMS["kernelSize"],
                              lopt=PARAMS["lstep"], aopt=PARAMS["astep"], iterations=PARAMS["iterations"], lsigma=PARAMS["lsigma"],
                              lskip=PARAMS["lskip"], generate_map=True)
        self.ctx.init_particles(particles, bounds=(PARAMS["xmin"], PARAMS["ymin"], PARAMS["xmax"], PARAMS["ymax"]), delta=PARAMS["delta"])
        self.pose = np.repeat(np.asarray(init_pose, dtype=np.float64)[None, :], particles, axis=0)
        self.weight = np.zeros(particles); self.weight_sum = np.zeros(particles)
        self.rng = np.random.default_rng(seed)
        self.odo = None
        self.count = 0
        self.resamples = 0

    def draw_from_motion(self, pnew, pold):
        """MotionModel::drawFromMotion (motionmodel.cpp:40-60), vectorised over particles."""
        srr, srt, str_, stt = PARAMS["srr"], PARAMS["srt"], PARAMS["str"], PARAMS["stt"]
        d = pnew - pold
        dth = np.arctan2(np.sin(d[2]), np.cos(d[2]))
        s, c = np.sin(pold[2]), np.cos(pold[2])
        dx, dy = c * d[0] + s * d[1], -s * d[0] + c * d[1]
        sxy = 0.3 * srr
        n = self.n
        nx = dx + self.rng.normal(0, 1, n) * (srr * abs(dx) + str_ * abs(dth) + sxy * abs(dy))
        ny = dy + self.rng.normal(0, 1, n) * (srr * abs(dy) + str_ * abs(dth) + sxy * abs(dx))
        nth = dth + self.rng.normal(0, 1, n) * (stt * abs(dth) + srt * np.hypot(dx, dy))
        nth = np.fmod(nth, 2 * np.pi); nth = np.where(nth > np.pi, nth - 2 * np.pi, nth)
        s1, c1 = np.sin(self.pose[:, 2]), np.cos(self.pose[:, 2])
        self.pose = np.stack([c1 * nx - s1 * ny + self.pose[:, 0], s1 * nx + c1 * ny + self.pose[:, 1], nth + self.pose[:, 2]], axis=1)

    def process_scan(self, ranges, odom):
        """-> dict of device stage times (ms) for this scan; every scan of the bench log passes the update gate."""
        ctx = self.ctx
        if self.odo is None:
            self.odo = odom.copy()
        self.draw_from_motion(odom, self.odo)
        self.odo = odom.copy()
        t = dict(scanmatch=0.0, weights=0.0, remap=0.0, register=0.0, beam_evals=0, score_evals=0)
        if self.count == 0:
            ctx.register(ranges, self.pose)
        else:
            self.pose, score, s, l, c, ev = ctx.scanmatch(ranges, self.pose, 0.0)
            self.weight += l; self.weight_sum += l
            w, neff = ctx.normalize(self.weight, PARAMS["ogain"])
            st = ctx.stats()
            t["scanmatch"] = st["ms_scanmatch"]; t["weights"] = st["ms_weights"]
            t["beam_evals"] = st["beam_evals"]; t["score_evals"] = st["score_evals"]
            if neff < PARAMS["resampleThreshold"] * self.n:
                idx, _ = ctx.resample_indexes(w, float(self.rng.random()))
                self.pose = self.pose[idx]; self.weight_sum = self.weight_sum[idx]; self.weight = np.zeros(self.n)
                ctx.register(ranges, self.pose, idx)
                self.resamples += 1
            else:
                ctx.register(ranges, self.pose)
        st = ctx.stats()
        t["register"] = st["ms_register"]; t["remap"] = st["ms_remap"]
        self.count += 1
        return t


def run_cpu_reference(particles, scans, procs, log_path):
    """oracle/_ref/ref_driver -mode time, `procs` independent single-threaded processes.  Returns
    (particle-scans per second summed over processes, per-process ms per matched scan)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(exe):
        return None
    cmd = [exe, "-mode", "time", "-log", log_path, "-particles", str(particles), "-scans", str(scans), "-maxrange", repr(PARAMS["maxrange"]),
           "-maxUrange", repr(PARAMS["maxUrange"]), "-linearUpdate", "0.2", "-angularUpdate", "0.1"]
    t0 = time.time()
    ps = [subprocess.Popen(cmd + ["-seed", str(12345 + k)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for k in range(procs)]
    outs = [p.communicate()[0] for p in ps]
    wall = time.time() - t0
    ms = []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("{"):
                ms.append(json.loads(line)["ms_per_matched_scan"])
    if not ms:
        return None
    rate = sum(particles / (m / 1000.0) for m in ms)  # particle-scans/s, matched scans only
    return rate, float(np.mean(ms)), wall


def reference_arm(args):
    """--impl reference: the reference's own CPU code on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    pp = 16  # particles per process
    scans = min(1 + args.warmup + args.steps, 40)  # the same log prefix the GPU arm runs (first scan registers only)
    synth, log = make_log(scans)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "bench.log"); synth.write_carmen(log, lp)
        r = run_cpu_reference(pp, scans, procs, lp)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver is not built"})); return
    rate, ms, wall = r
    value = rate / args.particles
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d particles x 1081 beams, 0.05 m grid, 200x200 m map, synthetic room log" % args.particles,
                       "particles": args.particles, "beams": 1081},
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": procs, "kind": "reference",
                             "sample": "%d processes x %d particles x %d matched scans of the same log (%.0f ms per scan per process, %.1f s wall); "
                                       "scaled linearly to %d particles" % (procs, pp, scans - 1, ms, wall, args.particles)},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args); return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        raise SystemExit("multi-GPU sharding is not wired into bench.py yet")
    torch.cuda.set_device(local)
    gm = importlib.import_module("slam-gmapping-learn_b200")
    K, W = args.steps, args.warmup
    synth, log = make_log(1 + W + K)
    f = HostFilter(gm, args.particles, log.laser.angles(), log.odom[0], device=local)
    f.process_scan(log.ranges[0], log.odom[0])  # first scan: registration only
    for i in range(1, 1 + W):
        f.process_scan(log.ranges[i], log.odom[i])
    launches0 = f.ctx.stats()["launches"]
    torch.cuda.synchronize()
    stage = []
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for i in range(1 + W, 1 + W + K):
            stage.append(f.process_scan(log.ranges[i], log.odom[i]))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    launches = f.ctx.stats()["launches"] - launches0
    wall_ms = (t1 - t0) * 1000.0 / K
    dev_ms = float(np.mean([s["scanmatch"] + s["weights"] + s["remap"] + s["register"] for s in stage]))
    sm_ms = float(np.mean([s["scanmatch"] for s in stage]))
    beam_evals = float(np.mean([s["beam_evals"] for s in stage]))
    peak, peak_src = measured_peak()
    achieved = beam_evals * BYTES_PER_BEAM_EVAL / (sm_ms * 1e-3) / 1e9
    n, B = args.particles, log.laser.beams
    h2d = 2 * (B * 8 + n * 24) + n * 8  # ranges+poses for scanmatch and register, weights for normalize
    d2h = n * (24 + 8 * 3 + 4 * 2) + n * 8 + 8
    line = {"metric": METRIC, "value": 1000.0 / dev_ms, "unit": "scans/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%d particles x %d beams (270 deg, 0.25 deg), 0.05 m grid, 200x200 m map, synthetic 18x12 m room log" % (n, B),
                       "particles": n, "beams": B, "l2": "inputs larger than L2 (per-scan working set ~%.1f GB of map patches)" % (n * 70 * 16384 / 1e9),
                       "resamples_in_timed_region": f.resamples, "params": "gfs-carmen defaults, maxUrange 29 / maxrange 30"},
            "beam_evals_per_s": beam_evals / (dev_ms * 1e-3), "beam_evals_per_scan": beam_evals,
            "score_evals_per_particle": float(np.mean([s["score_evals"] for s in stage])) / n,
            "stage_ms": {k: float(np.mean([s[k] for s in stage])) for k in ("scanmatch", "weights", "remap", "register")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "k_scanmatch", "peak_source": peak_src, "bytes_per_beam_eval": BYTES_PER_BEAM_EVAL},
            "e2e": {"value": 1000.0 / wall_ms, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": wall_ms},
            "gpu_launches": launches, "clocks": clk.summary()}
    if not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as td:
            ns = min(1 + W + K, 40)
            lp = os.path.join(td, "bench.log"); synth.write_carmen(synth.make_log("room", synth.UTM30, scans=ns, step=0.25, seed=1), lp)
            r = run_cpu_reference(16, ns, 1, lp)
        if r is not None:
            rate, ms, wall = r
            line["cpu_baseline"] = {"value": rate / n, "unit": "scans/s", "cores": 1, "kind": "reference",
                                    "sample": "unmodified reference, 16 particles x %d matched scans of the same log on 1 core "
                                              "(%.0f ms per scan, %.1f s wall), scaled linearly to %d particles" % (ns - 1, ms, wall, n)}
    print(json.dumps(line))
    f.ctx.close()


if __name__ == "__main__":
    main()
