"""ctypes view of the C++ host layer (host/src/capi_shim.cpp): GMapping::GridSlamProcessor driven from Python for the
benchmark and the tests.  Plumbing only -- every call lands in the C++ API, which lands in the CUDA library."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(HERE, "host", "libgridfastslam_b200.so")
STAT_KEYS = ["ms_scanmatch", "ms_weights", "ms_remap", "ms_register", "ms_migrate", "ms_allgather", "beam_evals", "score_evals",
             "cow_copies", "patch_allocs", "free_updates", "hit_updates", "pool_in_use", "launches", "migrated_particles", "migrated_bytes",
             "host_ms_motion", "host_ms_scanmatch", "host_ms_treeweights", "host_ms_resample",
             "total_ms_scanmatch", "total_ms_weights", "total_ms_remap", "total_ms_register", "total_ms_migrate", "total_ms_allgather",
             "total_host_ms_motion", "total_host_ms_scanmatch", "total_host_ms_treeweights", "total_host_ms_resample",
             "total_beam_evals", "total_score_evals", "total_scanmatches", "total_migrated_particles", "total_migrated_bytes", "pool_hit_in_use"]
PARAM_ORDER = ["maxUrange", "maxrange", "sigma", "kernelSize", "lstep", "astep", "iterations", "lsigma", "ogain", "lskip", "srr", "srt", "str",
               "stt", "linearUpdate", "angularUpdate", "resampleThreshold", "period", "generate_map", "minimumScore"]
DEFAULTS = dict(maxUrange=80., maxrange=80., sigma=0.05, kernelSize=1, lstep=0.05, astep=0.05, iterations=5, lsigma=0.075, ogain=3., lskip=0,
                srr=0.01, srt=0.01, str=0.01, stt=0.01, linearUpdate=1.0, angularUpdate=0.5, resampleThreshold=0.5, period=-1.0,
                generate_map=True, minimumScore=0.0)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_LIB):
        raise RuntimeError("%s is missing: run `python slam-gmapping-learn_b200/build.py` (there is no CPU fallback)" % HOST_LIB)
    L = C.CDLL(HOST_LIB)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.gmh_create.restype = vp; L.gmh_create.argtypes = [C.c_int]
    L.gmh_destroy.argtypes = [vp]
    L.gmh_dist_unique_id.argtypes = [vp]
    L.gmh_set_distributed.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gmh_set_early_registration.argtypes = [vp, C.c_int]
    L.gmh_set_register_with_scanmatch.argtypes = [vp, C.c_int]
    L.gmh_set_lazy_tree.argtypes = [vp, C.c_int]
    L.gmh_refresh_tree.argtypes = [vp]
    L.gmh_tree_rows.argtypes = [vp, dp]
    L.gmh_setup.argtypes = [vp, C.c_uint, dp, dp, dp]
    L.gmh_init.argtypes = [vp, C.c_uint] + [C.c_double] * 5 + [dp, C.c_ulong]
    L.gmh_process_scan.argtypes = [vp, dp, C.c_uint, dp, C.c_double]
    L.gmh_process_scan_adapt.argtypes = [vp, dp, C.c_uint, dp, C.c_double, C.c_int]
    L.gmh_set_max_particles.argtypes = [vp, C.c_uint]
    for f in ("gmh_particles", "gmh_local_particles", "gmh_first_local"):
        getattr(L, f).restype = C.c_uint; getattr(L, f).argtypes = [vp]
    L.gmh_neff.restype = C.c_double; L.gmh_neff.argtypes = [vp]
    for f in ("gmh_best", "gmh_resamples", "gmh_last_resampled"):
        getattr(L, f).argtypes = [vp]
    L.gmh_state.argtypes = [vp, dp]
    L.gmh_indexes.argtypes = [vp, vp]
    L.gmh_device_stats.argtypes = [vp, dp]
    L.gmh_map_digests.argtypes = [vp, vp]
    L.gmh_clone.restype = vp; L.gmh_clone.argtypes = [vp]
    L.gmh_check_maps.restype = C.c_ulonglong; L.gmh_check_maps.argtypes = [vp]
    L.gmh_copy_trajectories.restype = C.c_ulonglong; L.gmh_copy_trajectories.argtypes = [vp, C.c_int]
    L.gmh_occupancy_grid.argtypes = [vp, C.c_uint, C.c_double, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]
    L.gmh_occupancy.restype = C.c_double; L.gmh_occupancy.argtypes = [vp, C.c_uint, C.c_double, C.c_double]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().gmh_dist_unique_id(buf)
    if rc:
        raise RuntimeError("gm_dist_unique_id failed (%d)" % rc)
    return buf.raw


class Processor:
    """GMapping::GridSlamProcessor (C++) behind ctypes."""

    def __init__(self, angles, particles, init_pose, bounds=(-100., -100., 100., 100.), delta=0.05, seed=12345, laser_pose=(0., 0., 0.),
                 rank=0, world=1, nccl_id=None, verbose=False, early_registration=True, lazy_tree=False, register_with_scanmatch=None,
                 max_particles=0, **params):
        L = load()
        self.L = L
        self.h = L.gmh_create(1 if verbose else 0)
        if world > 1:
            self._id = C.create_string_buffer(nccl_id, 128)
            L.gmh_set_distributed(self.h, rank, world, self._id)
        if lazy_tree:
            L.gmh_set_lazy_tree(self.h, 1)  # TNode::accWeight is brought up to date on demand (refresh_tree / getTrajectories)
        if register_with_scanmatch is not None:  # default: on (like a plain GridSlamProcessor object; subclasses with hooks: off unless sharded)
            L.gmh_set_register_with_scanmatch(self.h, 1 if register_with_scanmatch else 0)
        if max_particles:
            L.gmh_set_max_particles(self.h, int(max_particles))  # room for process_scan(..., adapt=) to grow the particle set
        if not early_registration:
            L.gmh_set_early_registration(self.h, 0)  # the reference's order: copy the resampled particles, then register
        p = dict(DEFAULTS); p.update(params)
        pv = np.array([float(p[k]) for k in PARAM_ORDER], dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float64); lp = np.ascontiguousarray(laser_pose, dtype=np.float64)
        L.gmh_setup(self.h, len(a), _dp(a), _dp(lp), _dp(pv))
        ip = np.ascontiguousarray(init_pose, dtype=np.float64)
        L.gmh_init(self.h, particles, bounds[0], bounds[1], bounds[2], bounds[3], delta, _dp(ip), seed)
        self.n = particles
        self.beams = len(a)

    def close(self):
        if getattr(self, "h", None):
            self.L.gmh_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_scan(self, ranges, odom, stamp, adapt=0):
        """adapt: GridSlamProcessor::processScan's adaptParticles -- a resampling at this reading continues with that many particles"""
        r = np.ascontiguousarray(ranges, dtype=np.float64); o = np.ascontiguousarray(odom, dtype=np.float64)
        assert r.shape == (self.beams,)
        if not adapt:
            return bool(self.L.gmh_process_scan(self.h, _dp(r), len(r), _dp(o), float(stamp)))
        done = bool(self.L.gmh_process_scan_adapt(self.h, _dp(r), len(r), _dp(o), float(stamp), int(adapt)))
        self.n = int(self.L.gmh_particles(self.h))
        return done

    def state(self):
        a = np.zeros((self.n, 6)); self.L.gmh_state(self.h, _dp(a)); return a

    def tree_rows(self):
        """per particle, walking its trajectory leaf -> root: depth, leaf accWeight, sum accWeight, weighted sum, sum visitCounter, sum childs"""
        a = np.zeros((self.n, 6)); self.L.gmh_tree_rows(self.h, _dp(a)); return a

    def refresh_tree(self):
        self.L.gmh_refresh_tree(self.h)

    def indexes(self):
        a = np.zeros(self.n, dtype=np.uint32); self.L.gmh_indexes(self.h, a.ctypes.data); return a

    def neff(self):
        return self.L.gmh_neff(self.h)

    def best(self):
        return self.L.gmh_best(self.h)

    def resamples(self):
        return self.L.gmh_resamples(self.h)

    def last_resampled(self):
        return bool(self.L.gmh_last_resampled(self.h))

    def local_range(self):
        f = self.L.gmh_first_local(self.h); return f, f + self.L.gmh_local_particles(self.h)

    def stats(self):
        a = np.zeros(36); self.L.gmh_device_stats(self.h, _dp(a)); return dict(zip(STAT_KEYS, a.tolist()))

    def map_digests(self):
        nl = self.L.gmh_local_particles(self.h)
        a = np.zeros((nl, 6), dtype=np.uint64); self.L.gmh_map_digests(self.h, a.ctypes.data); return a

    def copy_trajectories(self, count_nodes=True):
        """GridSlamProcessor::getTrajectories() (the deep copy of the trajectory tree a clone() makes), freed again; -> nodes copied
        (count_nodes=False: only the copy and its release, for timing; returns the number of leaves)"""
        return int(self.L.gmh_copy_trajectories(self.h, 1 if count_nodes else 0))

    def check_maps(self):
        """MAP_CONSISTENCY_CHECK on the device (aborts the process on a mismatch); -> distinct patches referenced in the pool"""
        return int(self.L.gmh_check_maps(self.h))

    def clone(self):
        """GridSlamProcessor::clone(): deep copy incl. trajectory tree and (device-side) particle maps"""
        other = object.__new__(Processor)
        other.L = self.L; other.n = self.n; other.beams = self.beams
        other.h = self.L.gmh_clone(self.h)
        return other

    def occupancy_grid(self, particle, thresh=0.25):
        w, h = C.c_int(), C.c_int(); origin = np.zeros(2)
        self.L.gmh_occupancy_grid(self.h, particle, thresh, None, C.byref(w), C.byref(h), _dp(origin))
        data = np.zeros((h.value, w.value), dtype=np.int8)
        self.L.gmh_occupancy_grid(self.h, particle, thresh, data.ctypes.data, C.byref(w), C.byref(h), _dp(origin))
        return data, origin

    def occupancy(self, particle, wx, wy):
        return self.L.gmh_occupancy(self.h, particle, wx, wy)
