#!/usr/bin/env python
"""Run one of BASELINE.json's configurations through GMapping::GridSlamProcessor (C++ host layer) and print one JSON line.

    python tools/run_config.py --config C3                       # 512 x 1081, 0.05 m, 200 m map, 1 GPU
    torchrun --nproc-per-node 8 tools/run_config.py --config C5  # 16384 x 1081, 0.025 m, 400 m corridor loop, 8 GPUs

These are sizing / behaviour runs (pool use, directory size, migration volume, stage times), not the bench line."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CONFIGS = {
    "C2": dict(particles=30, world="room", laser="LMS200", delta=0.05, half=100.0, urange=80.0, rng=80.0),
    "C3": dict(particles=512, world="room", laser="UTM30", delta=0.05, half=100.0, urange=29.0, rng=30.0),
    "C4": dict(particles=4096, world="room", laser="UTM30", delta=0.05, half=100.0, urange=29.0, rng=30.0),
    "C5": dict(particles=16384, world="corridor", laser="UTM30", delta=0.025, half=200.0, urange=29.0, rng=30.0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scans", type=int, default=30)
    ap.add_argument("--particles", type=int, default=0, help="override the particle count (e.g. one GPU's slice of C5)")
    ap.add_argument("--noise", type=float, default=0.01, help="motion noise srr=srt=str=stt")
    ap.add_argument("--ogain", type=float, default=3.0, help="likelihood gain (obsSigmaGain): small values spread the weights and force real resampling")
    ap.add_argument("--resample-threshold", type=float, default=0.5, help="resample when Neff < threshold * N (> 1: every scan)")
    ap.add_argument("--steady", action="store_true", help="do not read the statistics after every scan (that completes the registration "
                    "in flight): wall_ms_per_scan_after_setup is then the steady-state wall clock, the per-stage means come from running totals")
    args = ap.parse_args()
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    c = dict(CONFIGS[args.config])
    if args.particles:
        c["particles"] = args.particles
    log = synth.make_log(c["world"], getattr(synth, c["laser"]), scans=args.scans, step=0.25, seed=1)
    nccl_id = None
    if world > 1:
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(hostapi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().tolist())
    h = c["half"]
    proc = hostapi.Processor(log.laser.angles(), c["particles"], log.odom[0], bounds=(-h, -h, h, h), delta=c["delta"], seed=12345,
                             rank=rank, world=world, nccl_id=nccl_id, maxUrange=c["urange"], maxrange=c["rng"], linearUpdate=0.2, angularUpdate=0.1,
                             srr=args.noise, srt=args.noise, str=args.noise, stt=args.noise, resampleThreshold=args.resample_threshold, ogain=args.ogain)
    stats = []
    t0 = time.perf_counter()
    t_steady = None
    for i in range(args.scans):
        if i == 3:  # scans 0-2 carry one-off costs: registration only, NCCL connection set-up, first pool growth
            torch.cuda.synchronize(); t_steady = time.perf_counter()
        proc.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
        if not args.steady or i == 2 or i == args.scans - 1:
            stats.append(proc.stats())  # (completes the registration in flight)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    wall_steady = (time.perf_counter() - t_steady) / (args.scans - 3) if t_steady is not None and args.scans > 3 else None
    st = proc.state()
    best = proc.best()
    truth_err = float(np.hypot(*(st[best, :2] - log.truth[-1, :2])))
    free_b, total_b = torch.cuda.mem_get_info()
    keys = ("ms_scanmatch", "ms_weights", "ms_remap", "ms_register", "ms_migrate", "ms_allgather")
    if args.steady:  # two snapshots (after scan 2, after the last scan): per-stage means from the running totals
        a, b = stats[0], stats[-1]
        mean = {k: (b["total_" + k] - a["total_" + k]) / max(1, args.scans - 3) for k in keys}
        stats = [None, None, None, dict(b, **mean)]
    out = {"config": args.config, "rank": rank, "world": world, "particles": c["particles"], "beams": log.laser.beams, "delta": c["delta"],
           "scans": args.scans, "resamples": proc.resamples(), "wall_ms_per_scan": 1000 * wall / args.scans,
           "wall_ms_per_scan_after_setup": None if wall_steady is None else 1000 * wall_steady,
           "device_ms_per_scan": float(np.mean([sum(s[k] for k in keys) for s in stats[3:]])),
           "device_ms_per_scan_median": float(np.median([sum(s[k] for k in keys) for s in stats[3:]])),
           "migrate_ms_median_when_resampling": float(np.median([s["ms_migrate"] for s in stats[3:] if s["ms_migrate"] > 0] or [0.0])),  # scans 0-2: registration only / NCCL connection setup
           "stage_ms": {k[3:]: float(np.mean([s[k] for s in stats[3:]])) for k in keys},
           "pool_patches_in_use": stats[-1]["pool_in_use"], "pool_hit_patches_in_use": stats[-1]["pool_hit_in_use"],
           "patch_gb_in_use": (stats[-1]["pool_in_use"] * 4096 + stats[-1]["pool_hit_in_use"] * 32896) / 2**30, "patches_per_particle": stats[-1]["pool_in_use"] / max(1, (proc.local_range()[1] - proc.local_range()[0])),
           "hbm_used_gb": (total_b - free_b) / 2**30, "migrated_particles": float(stats[-1]["total_migrated_particles"]),
           "migrated_mb": float(stats[-1]["total_migrated_bytes"]) / 2**20, "neff_last": proc.neff(),
           "distinct_parents_last": int(len(set(proc.indexes().tolist()))) if proc.resamples() else None,
           "best_particle_error_m": truth_err}
    if dist is not None:
        outs = [None] * world
        dist.all_gather_object(outs, out)
        if rank == 0:
            out["per_rank_pool"] = [o["pool_patches_in_use"] for o in outs]
            out["per_rank_hbm_gb"] = [round(o["hbm_used_gb"], 2) for o in outs]
            out["migrated_particles"] = float(sum(o["migrated_particles"] for o in outs))
            out["migrated_mb"] = float(sum(o["migrated_mb"] for o in outs))
    if rank == 0:
        print(json.dumps(out))
    proc.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
