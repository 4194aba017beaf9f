"""Shared helpers for the tests: golden fixtures (tests/golden/*.npz, made by tools/make_golden.py from
the unmodified reference build) and construction of the oracle filter from a fixture's parameters."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["room181_n30", "room181_resize", "room1081_n8", "room181_k2_lskip", "room181_halfcell", "room181_d025"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLD, name + ".npz"))
        self.name = name
        self.p = json.loads(str(z["params"]))
        for k in z.files:
            if k != "params":
                setattr(self, k, z[k])
        self.S, self.N = self.pre_pose.shape[:2]
        self.B = self.ranges.shape[1]

    def matcher(self):
        p = self.p
        return orc.make_matcher(self.angles, (0, 0, 0), urange=p["maxUrange"], range_=p["maxrange"], sigma=p["sigma"],
                                kernel=p["kernelSize"], lopt=p["lstep"], aopt=p["astep"], iterations=p["iterations"],
                                lsigma=p["lsigma"], lskip=p["lskip"], generate_map=p["generate_map"])

    def oracle_filter(self):
        p = self.p
        return orc.Filter(p["particles"], self.matcher(), bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                          init_pose=tuple(self.head_init), motion=(p["srr"], p["srt"], p["str"], p["stt"]),
                          linear_update=p["linearUpdate"], angular_update=p["angularUpdate"],
                          resample_thr=p["resampleThreshold"], period=p["period"], minimum_score=p["minimumScore"],
                          ogain=p["ogain"], seed=p["seed"])


def digest_rows(digests):
    """structured digest array -> int rows (patches, cells, sum_n, sum_visits, hash_counts, hash_acc)"""
    return np.stack([digests[k].astype(np.uint64) for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc")], axis=-1)
