#!/usr/bin/env python
"""Host-side cost of GridSlamProcessor::processScan (motion model, trajectory tree, resample bookkeeping) per scan, measured
WITHOUT a GPU: the C ABI is the CPU mock of the tests (tests/mock/gm_mock.c, LD_PRELOAD), the scans have few beams so that
the mock's scan matcher is cheap, and only the host timers of the C++ layer are reported.  A development aid for the part
of the wall clock that is replicated on every rank of a multi-GPU run.

    python tools/host_bench.py [--particles 4096] [--beams 16] [--scans 30]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--beams", type=int, default=16)
    ap.add_argument("--scans", type=int, default=30)
    ap.add_argument("--noise", type=float, default=0.01, help="motion-model noise (srr = stt; srt = str = 2x); larger values make the filter resample")
    args = ap.parse_args()
    if "libgm_mock" not in os.environ.get("LD_PRELOAD", ""):
        from host_common import build_host, mock_env
        build_host()
        os.execve(sys.executable, [sys.executable] + sys.argv, mock_env())
    import numpy as np
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    laser = synth.LaserSpec(args.beams, 270.0 / max(1, args.beams - 1), 30.0)
    log = synth.make_log("room", laser, scans=args.scans + 1, step=0.25, seed=1)
    f = hostapi.Processor(laser.angles(), args.particles, log.odom[0], maxUrange=29.0, maxrange=30.0, linearUpdate=0.2, angularUpdate=0.1,
                          srr=args.noise, stt=args.noise, srt=2 * args.noise, str=2 * args.noise)
    f.process_scan(log.ranges[0], log.odom[0], log.stamps[0])
    s0 = f.stats()
    for i in range(1, args.scans + 1):
        f.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
    s1 = f.stats()
    per = {k[len("total_host_ms_"):]: (s1[k] - s0[k]) / args.scans for k in s1 if k.startswith("total_host_ms_")}
    print("particles %d beams %d scans %d resamples %d | host ms per scan: %s" % (
        args.particles, args.beams, args.scans, f.resamples(), "  ".join("%s %.3f" % kv for kv in per.items())))
    f.close()


if __name__ == "__main__":
    main()
