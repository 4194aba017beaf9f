"""Helpers for the `-m gpu` tests: a CUDA context configured from a golden fixture's parameters."""
from __future__ import annotations

import importlib

import numpy as np

from common import digest_rows  # noqa: F401

gm = importlib.import_module("slam-gmapping-learn_b200")


def context_for(g, pool_patches=None, max_particles=None):
    p = g.p
    n = p["particles"]
    ctx = gm.GmContext(max_particles or n, pool_patches=pool_patches or max(4096, n * 400))
    ctx.set_laser(g.angles)
    ctx.set_matching(urange=p["maxUrange"], range_=p["maxrange"], sigma=p["sigma"], kernel=p["kernelSize"], lopt=p["lstep"],
                     aopt=p["astep"], iterations=p["iterations"], lsigma=p["lsigma"], lskip=p["lskip"],
                     generate_map=p["generate_map"])
    ctx.init_particles(n, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"])
    return ctx


def gpu_digest_rows(ctx):
    d = ctx.digests()
    return np.stack([d[k].astype(np.uint64) for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc")], axis=-1)


def map_cells_world(info, coords, cells):
    """dict-free dense comparison aid: -> (wx, wy, n, visits, ax, ay) rows of touched cells, sorted."""
    rows = []
    for (px, py), c in zip(coords, cells):
        c = c.reshape(32, 32)
        lx, ly = np.nonzero((c["visits"] != 0) | (c["n"] != 0))
        wx = px * 32 + lx - info[0]; wy = py * 32 + ly - info[1]
        rows.append(np.stack([wx, wy, c["n"][lx, ly], c["visits"][lx, ly]], axis=1).astype(np.int64))
    r = np.concatenate(rows) if rows else np.zeros((0, 4), dtype=np.int64)
    order = np.lexsort((r[:, 1], r[:, 0]))
    acc = np.concatenate([np.stack([c.reshape(32, 32)["ax"][np.nonzero((c.reshape(32, 32)["visits"] != 0) | (c.reshape(32, 32)["n"] != 0))],
                                    c.reshape(32, 32)["ay"][np.nonzero((c.reshape(32, 32)["visits"] != 0) | (c.reshape(32, 32)["n"] != 0))]], axis=1)
                          for c in cells]) if len(cells) else np.zeros((0, 2), dtype=np.float32)
    return r[order], acc[order]
