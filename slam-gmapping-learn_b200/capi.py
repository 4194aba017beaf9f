"""ctypes binding of libgm_b200.so (include/gm_b200.h).  Plumbing only: every call goes to the CUDA
library; if the library or a B200 is missing this raises -- there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgm_b200.so")

EXPORTS = ["gm_create", "gm_destroy", "gm_last_error", "gm_set_laser", "gm_set_matching", "gm_default_matching",
           "gm_init_particles", "gm_score", "gm_likelihood_and_score", "gm_scanmatch", "gm_normalize_neff",
           "gm_resample_indexes", "gm_resample_indexes_n", "gm_copy_particles", "gm_copy_particles_n", "gm_register", "gm_map_info_get", "gm_download_map", "gm_map_digest", "gm_map_entropy",
           "gm_get_stats", "gm_particles", "gm_optimize", "gm_compute_active_area", "gm_upload_map", "gm_export_occupancy", "gm_clone", "gm_set_async_register", "gm_synchronize",
           "gm_dist_unique_id", "gm_dist_init", "gm_dist_register", "gm_dist_copy_particles", "gm_dist_plan", "gm_scanmatch_weights", "gm_check_refcounts", "gm_set_fused_register", "gm_scanmatch_weights_begin", "gm_scanmatch_weights_end"]


class GmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gm_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_particles", C.c_uint32), ("pool_patches", C.c_uint64),
                ("block_threads", C.c_uint32), ("flags", C.c_uint32), ("global_particles", C.c_uint32), ("reserved", C.c_uint32)]


class MatchParams(C.Structure):
    _fields_ = [("laser_max_range", C.c_double), ("usable_range", C.c_double), ("gaussian_sigma", C.c_double),
                ("likelihood_sigma", C.c_double), ("kernel_size", C.c_int32), ("generate_map", C.c_int32),
                ("opt_linear_delta", C.c_double), ("opt_angular_delta", C.c_double),
                ("opt_recursive_iterations", C.c_uint32), ("likelihood_skip", C.c_uint32),
                ("enlarge_step", C.c_double), ("fullness_threshold", C.c_double),
                ("angular_odometry_reliability", C.c_double), ("linear_odometry_reliability", C.c_double),
                ("free_cell_ratio", C.c_double), ("initial_beams_skip", C.c_uint32), ("reserved", C.c_uint32)]


class MapInfo(C.Structure):
    _fields_ = [("center_x", C.c_double), ("center_y", C.c_double), ("delta", C.c_double),
                ("size_x2", C.c_int32), ("size_y2", C.c_int32), ("patches_x", C.c_int32), ("patches_y", C.c_int32),
                ("allocated_patches", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("pool_capacity", C.c_uint64), ("pool_in_use", C.c_uint64), ("score_evals", C.c_uint64),
                ("beam_evals", C.c_uint64), ("cow_copies", C.c_uint64), ("patch_allocs", C.c_uint64),
                ("free_updates", C.c_uint64), ("hit_updates", C.c_uint64), ("resizes", C.c_uint64),
                ("ms_scanmatch", C.c_float), ("ms_register", C.c_float), ("ms_remap", C.c_float), ("ms_weights", C.c_float),
                ("ms_migrate", C.c_float), ("ms_allgather", C.c_float), ("migrated_particles", C.c_uint64), ("migrated_bytes", C.c_uint64),
                ("launches", C.c_uint32), ("reserved", C.c_uint32), ("register_phase_cycles", C.c_uint64 * 8),
                ("max_active_patches", C.c_uint32), ("register_tile_cap", C.c_uint32),
                ("total_ms", C.c_double * 6), ("total_beam_evals", C.c_uint64), ("total_score_evals", C.c_uint64),
                ("total_scanmatches", C.c_uint64), ("total_registers", C.c_uint64),
                ("total_migrated_particles", C.c_uint64), ("total_migrated_bytes", C.c_uint64),
                ("pool_hit_capacity", C.c_uint64), ("pool_hit_in_use", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("reserved", "register_phase_cycles", "total_ms")}
        d["register_phase_cycles"] = list(self.register_phase_cycles)
        d["total_ms"] = list(self.total_ms)
        return d


CELL_DTYPE = np.dtype([("ax", "<f4"), ("ay", "<f4"), ("n", "<i4"), ("visits", "<i4")])
DIGEST_DTYPE = np.dtype([("patches", "<i8"), ("cells", "<i8"), ("sum_n", "<i8"), ("sum_visits", "<i8"),
                         ("hash_counts", "<u8"), ("hash_acc", "<u8")])

_lib = None


def load_library(path=LIB_PATH):
    """dlopen the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise GmError(-4, "%s is missing: build it with `python slam-gmapping-learn_b200/build.py` (no CPU fallback exists)" % path)
    L = C.CDLL(path)
    dp, up, vp = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_void_p
    L.gm_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.gm_destroy.argtypes = [vp]
    L.gm_last_error.restype = C.c_char_p; L.gm_last_error.argtypes = [vp]
    L.gm_set_laser.argtypes = [vp, C.c_uint32, dp, dp]
    L.gm_set_matching.argtypes = [vp, C.POINTER(MatchParams)]
    L.gm_default_matching.restype = None; L.gm_default_matching.argtypes = [C.POINTER(MatchParams)]
    L.gm_init_particles.argtypes = [vp, C.c_uint32, dp, C.c_double, C.c_double, C.c_double]
    L.gm_score.argtypes = [vp, C.c_uint32, up, dp, dp, dp]
    L.gm_likelihood_and_score.argtypes = [vp, C.c_uint32, up, dp, dp, dp, dp, up]
    L.gm_scanmatch.argtypes = [vp, dp, dp, C.c_double, dp, dp, dp, up, up]
    L.gm_normalize_neff.argtypes = [vp, dp, C.c_uint32, C.c_double, dp, dp]
    L.gm_resample_indexes.argtypes = [vp, dp, C.c_uint32, C.c_double, up, up]
    L.gm_resample_indexes_n.argtypes = [vp, dp, C.c_uint32, C.c_uint32, C.c_double, up, up]
    L.gm_copy_particles_n.argtypes = [vp, up, C.c_uint32]
    L.gm_register.argtypes = [vp, dp, dp, up]
    L.gm_copy_particles.argtypes = [vp, up]
    L.gm_map_info_get.argtypes = [vp, C.c_uint32, C.POINTER(MapInfo)]
    L.gm_download_map.argtypes = [vp, C.c_uint32, C.POINTER(MapInfo), vp, vp, C.c_uint32]
    L.gm_upload_map.argtypes = [vp, C.c_uint32, C.POINTER(MapInfo), vp, vp, C.c_uint32]
    L.gm_map_digest.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.gm_map_entropy.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.gm_export_occupancy.argtypes = [vp, C.c_uint32, C.c_double, vp, C.c_uint32, C.c_uint32]
    L.gm_optimize.argtypes = [vp, C.c_uint32, up, dp, dp, dp, dp]
    L.gm_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.gm_particles.restype = C.c_uint32; L.gm_particles.argtypes = [vp]
    L.gm_scanmatch_weights.argtypes = [vp, dp, dp, C.c_double, dp, C.c_double, dp, dp, dp]
    L.gm_check_refcounts.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GmContext:
    """One gm_ctx: a particle set with its maps on one GPU."""

    def __init__(self, max_particles, device=0, pool_patches=0, block_threads=0, global_particles=0):
        self.L = load_library()
        self.h = C.c_void_p()
        cfg = Config(device, max_particles, pool_patches, block_threads, 0, global_particles, 0)
        rc = self.L.gm_create(C.byref(self.h), C.byref(cfg))
        if rc:
            raise GmError(rc, self.L.gm_last_error(None).decode())
        self.nbeams = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.gm_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise GmError(rc, self.L.gm_last_error(self.h).decode())

    # ---- configuration
    def set_laser(self, angles, laser_pose=(0., 0., 0.)):
        a = _f64(angles); lp = _f64(laser_pose)
        self._ck(self.L.gm_set_laser(self.h, len(a), _dp(a), _dp(lp)))
        self.nbeams = len(a)

    def default_matching(self):
        p = MatchParams(); self.L.gm_default_matching(C.byref(p)); return p

    def set_matching(self, urange=80., range_=80., sigma=0.05, kernel=1, lopt=0.05, aopt=0.05, iterations=5, lsigma=0.075,
                     lskip=0, generate_map=True, **extra):
        """ScanMatcher::setMatchingParameters (reference scanmatcher.cpp:872-883) + setgenerateMap."""
        p = self.default_matching()
        p.usable_range = urange; p.laser_max_range = range_; p.gaussian_sigma = sigma; p.kernel_size = kernel
        p.opt_linear_delta = lopt; p.opt_angular_delta = aopt; p.opt_recursive_iterations = iterations
        p.likelihood_sigma = lsigma; p.likelihood_skip = lskip; p.generate_map = 1 if generate_map else 0
        for k, v in extra.items():
            setattr(p, k, v)
        self._ck(self.L.gm_set_matching(self.h, C.byref(p)))
        self.params = p

    def init_particles(self, n, bounds=(-100., -100., 100., 100.), delta=0.05):
        xmin, ymin, xmax, ymax = bounds
        c = _f64([(xmin + xmax) * .5, (ymin + ymax) * .5])
        self._ck(self.L.gm_init_particles(self.h, n, _dp(c), xmax - xmin, ymax - ymin, delta))
        self.n = n

    # ---- stages
    def _ranges(self, ranges):
        r = _f64(ranges)
        if not self.nbeams:
            raise GmError(-2, "gm_set_laser has not been called")
        if r.shape != (self.nbeams,):
            raise GmError(-2, "expected %d ranges, got %r" % (self.nbeams, r.shape))
        return r

    def score(self, particle_idx, poses, ranges):
        idx = np.ascontiguousarray(particle_idx, dtype=np.uint32); p = _f64(poses).reshape(-1, 3); r = self._ranges(ranges)
        out = np.zeros(len(idx))
        self._ck(self.L.gm_score(self.h, len(idx), _up(idx), _dp(p), _dp(r), _dp(out)))
        return out

    def likelihood_and_score(self, particle_idx, poses, ranges):
        idx = np.ascontiguousarray(particle_idx, dtype=np.uint32); p = _f64(poses).reshape(-1, 3); r = self._ranges(ranges)
        s = np.zeros(len(idx)); l = np.zeros(len(idx)); c = np.zeros(len(idx), dtype=np.uint32)
        self._ck(self.L.gm_likelihood_and_score(self.h, len(idx), _up(idx), _dp(p), _dp(r), _dp(s), _dp(l), _up(c)))
        return s, l, c

    def scanmatch(self, ranges, poses, minimum_score=0.0):
        """-> (poses, score, s, l, c, evals)"""
        r = self._ranges(ranges); p = _f64(poses).reshape(self.n, 3).copy()
        sc = np.zeros(self.n); s = np.zeros(self.n); l = np.zeros(self.n)
        c = np.zeros(self.n, dtype=np.uint32); ev = np.zeros(self.n, dtype=np.uint32)
        self._ck(self.L.gm_scanmatch(self.h, _dp(r), _dp(p), float(minimum_score), _dp(sc), _dp(s), _dp(l), _up(c), _up(ev)))
        return p, sc, s, l, c, ev

    def scanmatch_weights(self, ranges, poses, log_weights, minimum_score=0.0, obs_sigma_gain=1.0):
        """the fused call: -> (rows [N, 5] = x y theta score l, weights [N], neff)"""
        r = self._ranges(ranges); p = _f64(poses).reshape(self.n, 3); lw = _f64(log_weights)
        rows = np.zeros((len(lw), 5)); w = np.zeros(len(lw)); ne = C.c_double()
        self._ck(self.L.gm_scanmatch_weights(self.h, _dp(r), _dp(p), float(minimum_score), _dp(lw), float(obs_sigma_gain), _dp(rows), _dp(w), C.byref(ne)))
        return rows, w, ne.value

    def check_refcounts(self):
        """MAP_CONSISTENCY_CHECK: raises when a patch's reference count differs from the directories; -> referenced patches"""
        bad = C.c_uint64(); used = C.c_uint64()
        self._ck(self.L.gm_check_refcounts(self.h, C.byref(bad), C.byref(used)))
        return used.value

    def normalize(self, log_weights, obs_sigma_gain=1.0):
        lw = _f64(log_weights); w = np.zeros_like(lw); ne = C.c_double()
        self._ck(self.L.gm_normalize_neff(self.h, _dp(lw), len(lw), float(obs_sigma_gain), _dp(w), C.byref(ne)))
        return w, ne.value

    def resample_indexes(self, weights, u01, n_out=None):
        """n_out: uniform_resampler::resampleIndexes' `nparticles` (processScan's adaptParticles); None = as many as weights"""
        w = _f64(weights); n_out = len(w) if n_out is None else int(n_out)
        idx = np.zeros(n_out, dtype=np.uint32); em = C.c_uint32()
        self._ck(self.L.gm_resample_indexes_n(self.h, _dp(w), len(w), n_out, float(u01), _up(idx), C.byref(em)))
        return idx, em.value

    def register(self, ranges, poses, idx=None):
        r = self._ranges(ranges); p = _f64(poses).reshape(self.n, 3)
        ip = None
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.uint32); assert idx.shape == (self.n,)
            ip = _up(idx)
        self._ck(self.L.gm_register(self.h, _dp(r), _dp(p), ip))

    def copy_particles(self, idx, n_new=None):
        """n_new: the new generation's particle count when the resampling changes it (len(idx) == n_new)"""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        if n_new is None:
            assert idx.shape == (self.n,)
            self._ck(self.L.gm_copy_particles(self.h, _up(idx)))
        else:
            assert idx.shape == (n_new,)
            self._ck(self.L.gm_copy_particles_n(self.h, _up(idx), int(n_new)))
            self.n = int(n_new)

    # ---- read-back
    def map_info(self, particle):
        mi = MapInfo(); self._ck(self.L.gm_map_info_get(self.h, particle, C.byref(mi))); return mi

    def download_map(self, particle):
        mi = self.map_info(particle)
        n = mi.allocated_patches
        coords = np.zeros((n, 2), dtype=np.int32); cells = np.zeros((n, 1024), dtype=CELL_DTYPE)
        if n:
            self._ck(self.L.gm_download_map(self.h, particle, C.byref(mi), coords.ctypes.data, cells.ctypes.data, n))
        return mi, coords, cells

    def upload_map(self, particle, info, coords, cells):
        """the inverse of download_map: replace one particle's map by host content"""
        coords = np.ascontiguousarray(coords, dtype=np.int32); cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        assert coords.shape[0] == cells.shape[0]
        self._ck(self.L.gm_upload_map(self.h, particle, C.byref(info), coords.ctypes.data, cells.ctypes.data, coords.shape[0]))

    def digests(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.zeros(count, dtype=DIGEST_DTYPE)
        self._ck(self.L.gm_map_digest(self.h, first, count, out.ctypes.data))
        return out

    def occupancy_grid(self, particle, occ_thresh=0.25):
        """-> int8 [height, width] (row y, column x): -1 unknown, 0 free, 100 occupied"""
        mi = self.map_info(particle)
        w, h = mi.patches_x * 32, mi.patches_y * 32
        out = np.zeros((h, w), dtype=np.int8)
        self._ck(self.L.gm_export_occupancy(self.h, particle, float(occ_thresh), out.ctypes.data, w, h))
        return out

    def optimize(self, particle_idx, poses, ranges):
        idx = np.ascontiguousarray(particle_idx, dtype=np.uint32); p = _f64(poses).reshape(-1, 3); r = self._ranges(ranges)
        out = np.zeros_like(p); sc = np.zeros(len(idx))
        self._ck(self.L.gm_optimize(self.h, len(idx), _up(idx), _dp(p), _dp(r), _dp(out), _dp(sc)))
        return out, sc

    def stats(self):
        s = Stats(); self._ck(self.L.gm_get_stats(self.h, C.byref(s))); return s.as_dict()
