"""Multi-GPU host-side logic on CPU: two gloo ranks plan the migration of resampled parents (gm_dist_plan, the planner
gm_dist_register uses) for the same global index vector and must agree on who sends what to whom."""
import ctypes
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from host_common import PKG, build_host, free_port
from oracle import orc


def _plan(lib, n_local, rank, world, idx):
    imp = np.zeros(n_local, dtype=np.uint32); loc = np.zeros(n_local, dtype=np.uint32)
    k = lib.gm_dist_plan(n_local, rank, world, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                         imp.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), loc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert k >= 0
    return imp[:k].copy(), loc


def _worker(rank, world, port, n_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ctypes.CDLL(os.path.join(PKG, "libgm_b200.so"))
    lib.gm_dist_plan.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_uint32)] * 3
    N = n_local * world
    ok = True
    for trial in range(6):
        # every rank computes the same weights -> the same systematic-resampling indexes (the filter replicates this state)
        rng = np.random.default_rng(100 + trial)
        logw = rng.normal(0, 3.0 + 4 * trial, size=N)
        w, _ = orc.normalize(logw, 1.0)
        idx, k = orc.resample_indexes(w, 0.3 + 0.1 * trial)
        imp, loc = _plan(lib, n_local, rank, world, idx)
        first = rank * n_local
        # every local child resolves to its parent: a local slot or an import slot
        for j in range(n_local):
            p = idx[first + j]
            if first <= p < first + n_local:
                ok &= loc[j] == p - first
            else:
                ok &= loc[j] >= n_local and imp[loc[j] - n_local] == p
        ok &= bool(np.all(np.diff(imp.astype(np.int64)) > 0))  # ascending, unique: grouped by owner
        ok &= not np.any((imp >= first) & (imp < first + n_local))
        # what I import from rank g is exactly what g derives it must send me (g runs the planner for my rank)
        mine = [None] * world
        dist.all_gather_object(mine, imp.tolist())
        for g in range(world):
            if g == rank:
                continue
            theirs_from_me = [p for p in mine[g] if p // n_local == rank]
            derived, _ = _plan(lib, n_local, g, world, idx)
            ok &= theirs_from_me == [int(p) for p in derived if p // n_local == rank]
        # systematic resampling keeps order: imports are a small fraction of the slice
        ok &= len(imp) <= n_local
    t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


def test_migration_plan_two_ranks_gloo():
    build_host()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1
