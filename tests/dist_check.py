"""Run under torchrun with 2 ranks: the sharded filter (GridSlamProcessor::setDistributed: particle slices, all-gather of the
scan-match results, migration of resampled parents between ranks) against the golden trace of the reference.

  * on a 2-GPU box (tests/test_sharded_gpu.py): NCCL all-gather + patch migration between the two GPUs;
  * on a CPU box (tests/test_host_layer_cpu.py): the same host logic over the CPU mock of the C ABI (tests/mock/gm_mock.c,
    LD_PRELOAD; its gm_dist_* exchange through files), torch.distributed on gloo -- results must then be the reference's bit
    for bit."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import Golden, digest_rows  # noqa: E402

hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")


def make(g, rank, world, nccl_id):
    p = g.p
    return hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                             seed=p["seed"], rank=rank, world=world, nccl_id=nccl_id,
                             **{k: p[k] for k in hostapi.PARAM_ORDER})


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    mock = "libgm_mock" in os.environ.get("LD_PRELOAD", "")
    # GM_DIST_SAME_GPU: both ranks drive GPU 0 (a one-GPU box): the filter's own exchange is peer memory through CUDA IPC and
    # does not care; torch.distributed, which only carries the session id and the final tallies, falls back to gloo
    same_gpu = bool(os.environ.get("GM_DIST_SAME_GPU"))
    dev = "cpu" if (mock or same_gpu) else "cuda"
    ptol, rtol = (0.0, 0.0) if mock else (1e-4, 1e-3)
    if mock or same_gpu:
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_id():  # an NCCL unique id makes exactly one communicator
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.tensor(list(hostapi.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    for case in os.environ.get("GM_DIST_CASES", "room181_n30,room181_k2_lskip").split(","):
        g = Golden(case)
        sharded = make(g, rank, world, fresh_id())
        lo, hi = sharded.local_range()
        assert hi - lo == g.N // world
        migrated = 0
        dig = list(g.dig_scan)
        for i in range(g.S):
            done = sharded.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
            assert done == bool(g.processed[i])
            if not done:
                continue
            st = sharded.state()
            assert np.abs(st[:, :3] - g.state[i][:, :3]).max() <= ptol, (case, i)
            np.testing.assert_allclose(st[:, 3:5], g.state[i][:, 3:5], rtol=rtol, atol=0 if mock else 1e-6)
            assert np.array_equal(st[:, 5], g.state[i][:, 5]), (case, i, "previousIndex")
            assert sharded.last_resampled() == bool(g.resampled[i])
            if g.resampled[i]:
                assert np.array_equal(sharded.indexes(), g.ridx[i])
                st_ = sharded.stats()
                migrated += st_["migrated_particles"]
                print("dist_check %s rank %d scan %d: %d parents imported, pool in use %d" % (case, rank, i, st_["migrated_particles"], st_["pool_in_use"]), flush=True)
            if i in dig:
                want = digest_rows(g.digests[dig.index(i)])[lo:hi]
                assert np.array_equal(sharded.map_digests()[:, :5], want[:, :5]), (case, i, "map digests of the local slice")
        tot = torch.tensor([migrated], dtype=torch.float64, device=dev); dist.all_reduce(tot)
        assert tot.item() > 0, "no parent ever crossed the rank boundary: the migration path was not exercised"
        # host mirror of a local particle through Particle::map (lazy download)
        occ = sharded.occupancy(lo, float(g.state[-1][lo, 0]), float(g.state[-1][lo, 1]))
        assert -1.0 <= occ <= 1.0
        if rank == 0:
            print("dist_check %s: ok, %d parents migrated" % (case, int(tot.item())))
        sharded.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
