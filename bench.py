#!/usr/bin/env python
"""bench.py -- scans/sec of the per-scan particle update at 4096 particles x 1081 beams (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--particles P] [--impl reference] [--noise ros|carmen]

A "step" is one processScan: scan matching of every particle (hill-climb + likelihood), weight normalisation / Neff,
resampling when Neff drops (particle copy by patch reference, on several GPUs with migration of the parents' maps), and the
grid update (registerScan) -- on the next scan of a seeded synthetic 1081-beam log (18 x 12 m room, 0.05 m grid,
200 x 200 m map).  The motion-model noise is slam_gmapping's default (srr .1 / srt .2 / str .1 / stt .2), so the filter
resamples inside the timed region; `--noise carmen` gives round 1's workload (gfs-carmen's .01, no resampling).
Prints ONE JSON line.

  value      scans/s from device time only (CUDA events around the kernels of each stage, maps resident in HBM)
  e2e        scans/s through the reference-facing API, GMapping::GridSlamProcessor::processScan of the C++ host layer,
             with HOST buffers (ranges + odometry in; poses, weights, Neff out every step), wall clock bracketed by
             barrier + synchronisation, max over ranks
  roofline   the scan-match kernel.  What binds it is instruction issue (ncu: profiles/kernel_facts.json), so `frac` is
             warp-instructions issued per second over the SMs' issue capacity at the clock sampled during the run; the
             HBM figure SURVEY.md 8d defines (beam-evaluations x 144 B over the CUDA-event time against the measured
             copy bandwidth) is reported next to it as `hbm_algorithmic`, and `traffic` is the kernel's DRAM bytes per
             launch from the ncu capture, scaled to the particles per GPU
  cpu_baseline   the unmodified reference (oracle/_ref/ref_driver) on one host core, bounded sample,
             scaled linearly in the particle count (the reference is one loop over particles)
  sharded_parity (N > 1)  before the timed region the sharded filter follows a golden trace of the reference
             (tests/golden/room181_n32.npz) on this box: decisions and maps exact, parents migrated between the GPUs

`--impl reference` times the reference's CPU implementation on the host cores (independent particle slices, one
single-threaded reference process per core -- the reference itself has no threading), same log and parameters.
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
PATCH_BYTES = 4096                        # visit counters of one 32 x 32 patch: all a free-space patch owns
HIT_PLANES_BYTES = 16384 + 16384 + 128    # cells + means + occupancy bitmap, owned by the patches that hold end points
CPU_BASELINE_PARTICLES = 48  # bounded sample of the cpu_baseline leg: ~15 s of single-core reference work over the bench's scans
REFERENCE_PROCS = 16         # processes of the reference arm: the same on every box, so the per-N ratios share one denominator
LINEARITY_PARTICLES = 512    # one measured reference point at a large particle count next to the extrapolated one
NOISE = {"ros": dict(srr=0.1, srt=0.2, str=0.1, stt=0.2),          # gmapping/src/slam_gmapping.cpp defaults
         "carmen": dict(srr=0.01, srt=0.01, str=0.01, stt=0.01)}  # gfs-carmen.cpp:52-84 (round 1's workload)
# lobsGain: the weights are softmax(l / (ogain * N)) (gridslamprocessor.hxx:93), so gfs-carmen's ogain 3, tuned for its 30
# particles, never lets Neff drop at 4096 (round 1: no resampling in the timed region).  The benchmark keeps that
# configuration's EFFECTIVE gain, ogain * N = 90, at every particle count -- in both arms
GAIN_TIMES_PARTICLES = 90.0
PARAMS = dict(maxrange=30.0, maxUrange=29.0, sigma=0.05, kernelSize=1, lstep=0.05, astep=0.05, iterations=5, lsigma=0.075,
              lskip=0, resampleThreshold=0.5, delta=0.05, xmin=-100.0, ymin=-100.0, xmax=100.0, ymax=100.0,
              linearUpdate=0.2, angularUpdate=0.1)


def workload_config(particles, noise):
    """the `config` of both arms: the workload and nothing else"""
    return {"workload": "%d particles x 1081 beams (270 deg, 0.25 deg), 0.05 m grid, 200x200 m map, synthetic 18x12 m room log, "
                        "every scan processed, motion noise %s" % (particles, noise),
            "particles": particles, "beams": 1081, "noise": noise,
            "params": "gfs-carmen matcher defaults, maxUrange 29 / maxrange 30, lobsGain x particles = %g, " % GAIN_TIMES_PARTICLES +
                      ", ".join("%s %g" % kv for kv in NOISE[noise].items())}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def kernel_facts():
    """per-kernel constants taken from committed ncu captures (profiles/kernel_facts.json, written by tools/ncu_facts.py)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_facts.json")))
    except Exception:
        return {}


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


def reference_cmd(particles, scans, log_path, noise):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    cmd = [exe, "-mode", "time", "-log", log_path, "-particles", str(particles), "-scans", str(scans), "-maxrange", repr(PARAMS["maxrange"]),
           "-maxUrange", repr(PARAMS["maxUrange"]), "-linearUpdate", repr(PARAMS["linearUpdate"]), "-angularUpdate", repr(PARAMS["angularUpdate"])]
    for k, v in NOISE[noise].items():
        cmd += ["-" + k, repr(v)]
    cmd += ["-ogain", repr(GAIN_TIMES_PARTICLES / particles)]  # a slice of the filter weighs like the whole filter
    return exe, cmd


def run_cpu_reference(particles, scans, procs, log_path, noise, extra=None):
    """oracle/_ref/ref_driver -mode time, `procs` independent single-threaded processes (+ optionally one more process with
    `extra` = (particles, scans)).  Returns (particle-scans per second summed over the processes, mean ms per matched scan
    per process, wall seconds, ms per matched scan of the extra process or None)."""
    exe, cmd = reference_cmd(particles, scans, log_path, noise)
    if not os.path.exists(exe):
        return None
    t0 = time.time()
    ps = [subprocess.Popen(cmd + ["-seed", str(12345 + k)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for k in range(procs)]
    px = None
    if extra is not None:
        px = subprocess.Popen(reference_cmd(extra[0], extra[1], log_path, noise)[1] + ["-seed", "12345"], stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, text=True)
    outs = [p.communicate()[0] for p in ps]
    wall = time.time() - t0
    ms = []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("{"):
                ms.append(json.loads(line)["ms_per_matched_scan"])
    ms_extra = None
    if px is not None:
        for line in px.communicate()[0].splitlines():
            if line.startswith("{"):
                ms_extra = json.loads(line)["ms_per_matched_scan"]
    if not ms:
        return None
    rate = sum(particles / (m / 1000.0) for m in ms)  # particle-scans/s, matched scans only
    return rate, float(np.mean(ms)), wall, ms_extra


def reference_arm(args):
    """--impl reference: the reference's own CPU code on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, REFERENCE_PROCS))
    pp = CPU_BASELINE_PARTICLES  # particles per process: every core runs the same bounded sample as the cpu_baseline leg
    scans = min(1 + args.warmup + args.steps, 40)  # the same log prefix the GPU arm runs (first scan registers only)
    synth, log = make_log(scans)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "bench.log"); synth.write_carmen(log, lp)
        # alongside, when a core is free: ONE process at 512 particles x 3 matched scans -- a measured point far up the
        # particle axis that checks the linear extrapolation of the 48-particle samples
        extra = (LINEARITY_PARTICLES, 4) if cores > procs else None
        r = run_cpu_reference(pp, scans, procs, lp, args.noise, extra)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver is not built"})); return
    rate, ms, wall, ms_extra = r
    value = rate / args.particles
    sample = ("%d processes x %d particles x %d matched scans of the same log (%.0f ms per scan per process, %.1f s wall), "
              "scaled linearly to %d particles; %d cores available, the process count is pinned so that every N has the same "
              "denominator; the real %d-particle working set (~5 GB of patches) would not stay in cache like the sample's: "
              "the extrapolation flatters the reference" % (procs, pp, scans - 1, ms, wall, args.particles, cores, args.particles))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.particles, args.noise),
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": procs, "kind": "reference", "sample": sample,
                             "cores_available": cores, "one_core_scans_per_s": value / procs},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if ms_extra:
        line["cpu_baseline"]["linearity"] = {
            "particles": LINEARITY_PARTICLES, "matched_scans": 3, "measured_ms_per_scan_one_core": ms_extra,
            "extrapolated_ms_per_scan_one_core": ms * LINEARITY_PARTICLES / pp,
            "note": "one process at %d particles ran beside the %d sample processes" % (LINEARITY_PARTICLES, procs)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--noise", default="ros", choices=sorted(NOISE))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prelude", action="store_true", help="skip the sharded-parity prelude (N > 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args); return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    K, W = args.steps, args.warmup
    synth, log = make_log(1 + W + K)
    n, B = args.particles, log.laser.beams
    if n % world:
        raise SystemExit("--particles must be a multiple of the number of GPUs")

    def session_id():  # rank 0 draws the id of the filter's shared segment, torch.distributed carries it to the others
        if world == 1:
            return None
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(hostapi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    # ---- sharded-parity prelude (untimed): this box, these GPUs, the reference's golden trace
    prelude = None
    if world > 1 and not args.no_prelude:
        selfcheck = importlib.import_module("slam-gmapping-learn_b200.selfcheck")
        r = selfcheck.follow("room181_n32", rank=rank, world=world, session_id=session_id())
        tot = torch.tensor([float(r["migrated"])], dtype=torch.float64, device="cuda"); dist.all_reduce(tot)
        if tot.item() <= 0:
            raise SystemExit("sharded-parity prelude: no parent crossed a rank boundary, the migration path did not run")
        prelude = {"sharded_parity": "ok", "parents_migrated": int(tot.item()), "fixture": "tests/golden/room181_n32.npz",
                   "resamples": r["resamples"], "worst_pose_error": r["worst_pose"]}

    # every scan of the log passes the update gate (0.25 m per step >= linearUpdate)
    noise = NOISE[args.noise]
    proc = hostapi.Processor(log.laser.angles(), n, log.odom[0], bounds=(PARAMS["xmin"], PARAMS["ymin"], PARAMS["xmax"], PARAMS["ymax"]),
                             delta=PARAMS["delta"], seed=12345, rank=rank, world=world, nccl_id=session_id(),
                             maxUrange=PARAMS["maxUrange"], maxrange=PARAMS["maxrange"], sigma=PARAMS["sigma"], kernelSize=PARAMS["kernelSize"],
                             lstep=PARAMS["lstep"], astep=PARAMS["astep"], iterations=PARAMS["iterations"], lsigma=PARAMS["lsigma"],
                             ogain=GAIN_TIMES_PARTICLES / n, lskip=PARAMS["lskip"], srr=noise["srr"], srt=noise["srt"], str=noise["str"], stt=noise["stt"],
                             linearUpdate=PARAMS["linearUpdate"], angularUpdate=PARAMS["angularUpdate"], resampleThreshold=PARAMS["resampleThreshold"])
    for i in range(0, 1 + W):  # first scan registers only; then W warm-up steps
        assert proc.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
    launches0 = proc.stats()["launches"]; resamples0 = proc.resamples()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    barrier()
    s0 = proc.stats()  # running totals; reading them completes any registration still in flight
    migrated = [0, 0]
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for i in range(1 + W, 1 + W + K):
            # host buffers in, poses / weights out; the registration kernels of scan i overlap the host-side tail of scan i
            # and the head of scan i + 1 (gm_set_async_register), everything else is synchronous
            proc.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
            if world > 1 and proc.last_resampled():
                migrated[0] += 1  # tallied after the timed region from the running totals below
        barrier()  # device synchronised: the last registration is complete inside the timed region
        t1 = time.perf_counter()
    s1 = proc.stats()
    keys = ("ms_scanmatch", "ms_weights", "ms_remap", "ms_register", "ms_migrate", "ms_allgather")
    hkeys = ("host_ms_motion", "host_ms_scanmatch", "host_ms_treeweights", "host_ms_resample")
    per = {k: (s1["total_" + k] - s0["total_" + k]) / K for k in keys + hkeys}
    # device time of a step: the kernels of every stage; ms_allgather is host time spent waiting for the other ranks' rows
    # (their kernels, measured on their own clocks) and is reported, not added
    # the filter queues the registration behind the scan match and normalises the weights on a second stream meanwhile
    # (setregisterWithScanMatch, the default of a plain GridSlamProcessor and of a sharded one): k_normalize then overlaps
    # k_register and is not on the device's critical path -- it is listed in stage_ms, not added
    dev_keys = ("ms_scanmatch", "ms_remap", "ms_register", "ms_migrate")
    dev_ms_local = sum(per[k] for k in dev_keys)
    wall_ms_local = (t1 - t0) * 1000.0 / K
    sm_ms = per["ms_scanmatch"]
    beam_local = (s1["total_beam_evals"] - s0["total_beam_evals"]) / K
    score_local = (s1["total_score_evals"] - s0["total_score_evals"]) / K
    launches_local = s1["launches"] - launches0
    mig_local = [s1["total_migrated_particles"] - s0["total_migrated_particles"], s1["total_migrated_bytes"] - s0["total_migrated_bytes"]]
    if dist is not None:
        mx = torch.tensor([dev_ms_local, wall_ms_local, sm_ms], dtype=torch.float64, device="cuda"); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = torch.tensor([beam_local, launches_local, score_local] + mig_local, dtype=torch.float64, device="cuda"); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, sm_ms_max = mx.tolist(); beam_evals, launches, score_evals, mig_p, mig_b = sm.tolist()
    else:
        dev_ms, wall_ms, sm_ms_max = dev_ms_local, wall_ms_local, sm_ms
        beam_evals, launches, score_evals = beam_local, launches_local, score_local
        mig_p, mig_b = mig_local
    if rank == 0:
        peak, peak_src = measured_peak()
        clocks = clk.summary()
        facts = kernel_facts().get("k_scanmatch", {})
        nl = n // world
        algorithmic = beam_local * BYTES_PER_BEAM_EVAL / (sm_ms * 1e-3) / 1e9  # rank 0's kernel on rank 0's GPU
        ipb = facts.get("warp_instructions_per_beam_eval")
        sm_hz = (clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0) * 1e6
        issue_peak = 4 * 148 * sm_hz / 1e9  # warp-instructions per ns the 148 SMs' four schedulers can issue
        roofline = {"kernel": "k_scanmatch", "bound": "issue", "unit": "Gwarp-inst/s",
                    "achieved": (ipb * beam_local / 32.0 / (sm_ms * 1e-3) / 1e9) if ipb else None, "peak": issue_peak,
                    "traffic": (facts["dram_bytes_per_beam_eval"] * beam_local) if facts.get("dram_bytes_per_beam_eval") else None,
                    "facts_from": facts.get("source"),
                    "hbm_algorithmic": {"achieved": algorithmic, "peak": peak, "unit": "GB/s", "frac": algorithmic / peak, "peak_source": peak_src,
                                        "bytes_per_beam_eval": BYTES_PER_BEAM_EVAL,
                                        "note": "SURVEY.md 8d: 144 B per window search actually run / CUDA-event time; not a bandwidth the kernel "
                                                "draws -- an occupancy bitmap and per-cell means replace the nine 16-byte cell reads and what it "
                                                "reads is served from L1/L2 (`traffic` = DRAM bytes per launch)"}}
        roofline["frac"] = roofline["achieved"] / issue_peak if roofline["achieved"] else None
        rfacts = kernel_facts().get("k_register", {})
        free_upd, hit_upd, cow = s1["free_updates"], s1["hit_updates"], s1["cow_copies"]
        # last registration of the timed region (SURVEY.md 8d grid update); a copy-on-write copy moves at least the 4 KB of visit
        # counters each way (16 KB of cells more for a patch that holds end points: not counted)
        reg_bytes = free_upd * 8 + hit_upd * 32 + cow * 2 * 4096
        reg_ms = s1["ms_register"]
        roofline_register = {"kernel": "k_register", "bound": "hbm", "unit": "GB/s", "achieved": reg_bytes / (reg_ms * 1e-3) / 1e9 if reg_ms else None,
                             "peak": peak, "frac": reg_bytes / (reg_ms * 1e-3) / 1e9 / peak if reg_ms else None,
                             "algorithmic_bytes": reg_bytes, "traffic": rfacts.get("dram_bytes_per_particle", 0) * nl or None,
                             # what actually binds it (ncu: issue slots busy 57 %, DRAM 11 % of peak): warp instructions per second
                             "issue": ({"achieved": rfacts["warp_instructions_per_particle"] * nl / (reg_ms * 1e-3) / 1e9, "peak": issue_peak, "unit": "Gwarp-inst/s",
                                        "frac": rfacts["warp_instructions_per_particle"] * nl / (reg_ms * 1e-3) / 1e9 / issue_peak}
                                       if reg_ms and rfacts.get("warp_instructions_per_particle") else None),
                             "note": "last registration of the timed region: 8 B per free-cell update, 32 B per hit, 2 x 4 KB per copy-on-write patch copy"}
        h2d = world * (2 * (B * 8 + nl * 24) + n * 8)
        d2h = world * (n * 48 + 8 + 2 * 160)
        config = workload_config(n, args.noise)
        line = {"metric": METRIC, "value": 1000.0 / dev_ms, "unit": "scans/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "details": {"particles_per_gpu": nl,
                            "l2": "inputs larger than L2 (%.1f GB of map patches resident per GPU, every particle's active area is read and written every scan)" % ((s1["pool_in_use"] * PATCH_BYTES + s1["pool_hit_in_use"] * HIT_PLANES_BYTES) / 1e9),
                            "api": "GMapping::GridSlamProcessor::processScan (C++ host layer over the C ABI)"},
                "resamples_in_timed_region": proc.resamples() - resamples0,
                "migrated_particles": int(mig_p), "migrated_bytes": int(mig_b),
                "beam_evals_per_s": beam_evals / (dev_ms * 1e-3), "beam_evals_per_scan": beam_evals, "score_evals_per_particle": score_evals / n,
                "stage_ms": {k[3:]: per[k] for k in keys},
                "stages_in_value": [k[3:] for k in dev_keys],
                "host_ms": {k[8:]: per[k] for k in hkeys},
                "roofline": roofline, "roofline_register": roofline_register,
                "e2e": {"value": 1000.0 / wall_ms, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": wall_ms},
                "gpu_launches": int(launches), "clocks": clocks}
        if prelude:
            line.update(prelude)
        if not args.no_cpu_baseline and world == 1:
            with tempfile.TemporaryDirectory() as td:
                ns = min(1 + W + K, 40)
                lp = os.path.join(td, "bench.log"); synth.write_carmen(synth.make_log("room", synth.UTM30, scans=ns, step=0.25, seed=1), lp)
                r = run_cpu_reference(CPU_BASELINE_PARTICLES, ns, 1, lp, args.noise)
            if r is not None:
                rate, ms, wall, _ = r
                line["cpu_baseline"] = {"value": rate / n, "unit": "scans/s", "cores": 1, "kind": "reference",
                                        "sample": "unmodified reference, %d particles x %d matched scans of the same log on 1 core "
                                                  "(%.0f ms per scan, %.1f s wall), scaled linearly to %d particles" % (CPU_BASELINE_PARTICLES, ns - 1, ms, wall, n)}
        print(json.dumps(line))
    proc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
