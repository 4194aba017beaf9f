"""ctypes binding of oracle/libgm_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgm_oracle.so")
MAXBEAMS = 2048


class Matcher(C.Structure):
    _fields_ = [
        ("beams", C.c_uint32),
        ("angles", C.c_double * MAXBEAMS),
        ("laser_pose", C.c_double * 3),
        ("laser_max_range", C.c_double), ("usable_range", C.c_double),
        ("gaussian_sigma", C.c_double), ("likelihood_sigma", C.c_double),
        ("kernel_size", C.c_int32),
        ("opt_angular_delta", C.c_double), ("opt_linear_delta", C.c_double),
        ("opt_recursive_iterations", C.c_uint32), ("likelihood_skip", C.c_uint32),
        ("generate_map", C.c_int32),
        ("enlarge_step", C.c_double), ("fullness_threshold", C.c_double),
        ("angular_odometry_reliability", C.c_double), ("linear_odometry_reliability", C.c_double),
        ("free_cell_ratio", C.c_double),
        ("initial_beams_skip", C.c_uint32),
        ("score_calls", C.c_uint64), ("beam_evals", C.c_uint64),
    ]


class Digest(C.Structure):
    _fields_ = [("patches", C.c_int64), ("cells", C.c_int64), ("sum_n", C.c_int64), ("sum_visits", C.c_int64),
                ("hash_counts", C.c_uint64), ("hash_acc", C.c_uint64), ("sum_accx", C.c_double), ("sum_accy", C.c_double)]

    def key(self):
        return (self.patches, self.cells, self.sum_n, self.sum_visits, self.hash_counts, self.hash_acc)


CELL_DTYPE = np.dtype([("ax", "<f4"), ("ay", "<f4"), ("n", "<i4"), ("visits", "<i4")])
DIGEST_DTYPE = np.dtype([("patches", "<i8"), ("cells", "<i8"), ("sum_n", "<i8"), ("sum_visits", "<i8"),
                         ("hash_counts", "<u8"), ("hash_acc", "<u8"), ("sum_accx", "<f8"), ("sum_accy", "<f8")])

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH) for f in ("gm_oracle.c", "gm_oracle.h")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.orc_matcher_defaults.argtypes = [C.POINTER(Matcher)]
    L.orc_map_new.restype = vp; L.orc_map_new.argtypes = [C.c_double] * 5
    L.orc_map_clone.restype = vp; L.orc_map_clone.argtypes = [vp]
    L.orc_map_free.argtypes = [vp]
    L.orc_map_resize.argtypes = [vp] + [C.c_double] * 4
    L.orc_map_patch_count.restype = C.c_int64; L.orc_map_patch_count.argtypes = [vp]
    L.orc_map_export.restype = C.c_int64; L.orc_map_export.argtypes = [vp, vp, vp, C.c_int64]
    L.orc_map_digest.argtypes = [vp, C.POINTER(Digest)]
    L.orc_map_world2map.argtypes = [vp, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_score.restype = C.c_double; L.orc_score.argtypes = [C.POINTER(Matcher), vp, dp, dp]
    L.orc_likelihood_and_score.restype = C.c_uint32
    L.orc_likelihood_and_score.argtypes = [C.POINTER(Matcher), vp, dp, dp, dp, dp]
    L.orc_optimize.restype = C.c_double; L.orc_optimize.argtypes = [C.POINTER(Matcher), vp, dp, dp, dp]
    L.orc_compute_active_area.argtypes = [C.POINTER(Matcher), vp, dp, dp]
    L.orc_register_scan.argtypes = [C.POINTER(Matcher), vp, dp, dp]
    L.orc_grid_line.restype = C.c_int; L.orc_grid_line.argtypes = [C.c_int] * 4 + [vp, C.c_int]
    L.orc_normalize.restype = C.c_double; L.orc_normalize.argtypes = [dp, C.c_uint32, C.c_double, dp]
    L.orc_resample_indexes.restype = C.c_uint32; L.orc_resample_indexes.argtypes = [dp, C.c_uint32, C.c_double, vp]
    L.orc_resample_indexes_n.restype = C.c_uint32; L.orc_resample_indexes_n.argtypes = [dp, C.c_uint32, C.c_uint32, C.c_double, vp]
    L.orc_seed.argtypes = [C.c_long]
    L.orc_sample_gaussian.restype = C.c_double; L.orc_sample_gaussian.argtypes = [C.c_double]
    L.orc_draw_from_motion.argtypes = [dp, dp, dp] + [C.c_double] * 4 + [dp]
    L.orc_filter_new.restype = vp
    L.orc_filter_new.argtypes = [C.c_uint32] + [C.c_double] * 5 + [dp, C.POINTER(Matcher)]
    L.orc_filter_free.argtypes = [vp]
    L.orc_filter_set_motion.argtypes = [vp] + [C.c_double] * 4
    L.orc_filter_set_update.argtypes = [vp] + [C.c_double] * 6
    L.orc_filter_process.restype = C.c_int; L.orc_filter_process.argtypes = [vp, dp, dp, C.c_double]
    L.orc_filter_size.restype = C.c_uint32; L.orc_filter_size.argtypes = [vp]
    L.orc_filter_state.argtypes = [vp, dp]
    L.orc_filter_pre_poses.argtypes = [vp, dp]
    L.orc_filter_match_rows.argtypes = [vp, dp]
    L.orc_filter_neff.restype = C.c_double; L.orc_filter_neff.argtypes = [vp]
    L.orc_filter_weights.argtypes = [vp, dp]
    L.orc_filter_last_resampled.restype = C.c_int; L.orc_filter_last_resampled.argtypes = [vp, vp, dp, dp]
    L.orc_filter_map.restype = vp; L.orc_filter_map.argtypes = [vp, C.c_uint32]
    L.orc_filter_best.restype = C.c_int; L.orc_filter_best.argtypes = [vp]
    L.orc_filter_matcher.restype = C.POINTER(Matcher); L.orc_filter_matcher.argtypes = [vp]
    L.orc_filter_last_u01.restype = C.c_double; L.orc_filter_last_u01.argtypes = [vp]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_matcher(angles, laser_pose=(0, 0, 0), urange=80., range_=80., sigma=0.05, kernel=1, lopt=0.05, aopt=0.05,
                 iterations=5, lsigma=0.075, lskip=0, generate_map=True):
    """setLaserParameters (scanmatcher.cpp:699-709) + setMatchingParameters (scanmatcher.cpp:872-883)."""
    m = Matcher()
    lib().orc_matcher_defaults(C.byref(m))
    angles = np.asarray(angles, dtype=np.float64)
    assert len(angles) < MAXBEAMS  # scanmatcher.cpp:704
    m.beams = len(angles)
    for i, a in enumerate(angles):
        m.angles[i] = a
    for i in range(3):
        m.laser_pose[i] = laser_pose[i]
    m.usable_range = urange; m.laser_max_range = range_; m.gaussian_sigma = sigma; m.kernel_size = kernel
    m.opt_linear_delta = lopt; m.opt_angular_delta = aopt; m.opt_recursive_iterations = iterations
    m.likelihood_sigma = lsigma; m.likelihood_skip = lskip; m.generate_map = 1 if generate_map else 0
    return m


class Map:
    def __init__(self, cx=0., cy=0., w=200., h=200., delta=0.05, handle=None, own=True):
        self.h = handle if handle is not None else lib().orc_map_new(cx, cy, w, h, delta)
        self.own = own

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            lib().orc_map_free(self.h); self.h = None

    def clone(self):
        return Map(handle=lib().orc_map_clone(self.h))

    def digest(self):
        d = Digest(); lib().orc_map_digest(self.h, C.byref(d)); return d

    def export(self):
        n = lib().orc_map_patch_count(self.h)
        coords = np.zeros((n, 2), dtype=np.int32)
        cells = np.zeros((n, 1024), dtype=CELL_DTYPE)
        lib().orc_map_export(self.h, coords.ctypes.data, cells.ctypes.data, n)
        return coords, cells

    def world2map(self, x, y):
        ix, iy = C.c_int(), C.c_int()
        lib().orc_map_world2map(self.h, x, y, C.byref(ix), C.byref(iy)); return ix.value, iy.value

    def geometry(self):
        """(size_x2, size_y2, npx, npy) read from the C struct."""
        class _M(C.Structure):
            _fields_ = [("cx", C.c_double), ("cy", C.c_double), ("ww", C.c_double), ("wh", C.c_double), ("delta", C.c_double),
                        ("map_w", C.c_int), ("map_h", C.c_int), ("sx2", C.c_int), ("sy2", C.c_int), ("npx", C.c_int), ("npy", C.c_int)]
        m = C.cast(self.h, C.POINTER(_M)).contents
        return m.sx2, m.sy2, m.npx, m.npy


def score(m, mp, pose, ranges):
    pose = np.ascontiguousarray(pose, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    return lib().orc_score(C.byref(m), mp.h, _dp(pose), _dp(ranges))


def likelihood_and_score(m, mp, pose, ranges):
    pose = np.ascontiguousarray(pose, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    s, l = C.c_double(), C.c_double()
    c = lib().orc_likelihood_and_score(C.byref(m), mp.h, _dp(pose), _dp(ranges), C.byref(s), C.byref(l))
    return s.value, l.value, c


def optimize(m, mp, pose, ranges):
    pose = np.ascontiguousarray(pose, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    out = np.zeros(3)
    sc = lib().orc_optimize(C.byref(m), mp.h, _dp(pose), _dp(ranges), _dp(out))
    return sc, out


def register(m, mp, pose, ranges):
    """invalidateActiveArea + computeActiveArea + registerScan for one map."""
    pose = np.ascontiguousarray(pose, dtype=np.float64); ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    lib().orc_compute_active_area(C.byref(m), mp.h, _dp(pose), _dp(ranges))
    lib().orc_register_scan(C.byref(m), mp.h, _dp(pose), _dp(ranges))


def grid_line(x0, y0, x1, y1, cap=20000):
    pts = np.zeros((cap, 2), dtype=np.int32)
    n = lib().orc_grid_line(x0, y0, x1, y1, pts.ctypes.data, cap)
    return pts[:n].copy()


def normalize(logw, gain=1.0):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    w = np.zeros_like(logw)
    neff = lib().orc_normalize(_dp(logw), len(logw), gain, _dp(w))
    return w, neff


def resample_indexes(w, u01, n_out=None):
    """n_out: the `nparticles` argument of uniform_resampler::resampleIndexes (None: as many indexes as weights)"""
    w = np.ascontiguousarray(w, dtype=np.float64)
    n_out = len(w) if n_out is None else int(n_out)
    idx = np.zeros(n_out, dtype=np.uint32)
    k = lib().orc_resample_indexes_n(_dp(w), len(w), n_out, u01, idx.ctypes.data)
    return idx, k


class Filter:
    """Whole-filter oracle (processScan restated)."""

    def __init__(self, particles, matcher, bounds=(-100., -100., 100., 100.), delta=0.05, init_pose=(0., 0., 0.),
                 motion=(0.01, 0.01, 0.01, 0.01), linear_update=1.0, angular_update=0.5, resample_thr=0.5,
                 period=-1.0, minimum_score=0.0, ogain=3.0, seed=12345):
        L = lib()
        ip = np.asarray(init_pose, dtype=np.float64)
        self.h = L.orc_filter_new(particles, bounds[0], bounds[1], bounds[2], bounds[3], delta, _dp(ip), C.byref(matcher))
        L.orc_filter_set_motion(self.h, *motion)
        L.orc_filter_set_update(self.h, linear_update, angular_update, resample_thr, period, minimum_score, ogain)
        L.orc_seed(seed)
        self.n = particles

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_filter_free(self.h); self.h = None

    def process(self, ranges, odom, stamp):
        ranges = np.ascontiguousarray(ranges, dtype=np.float64); odom = np.ascontiguousarray(odom, dtype=np.float64)
        return bool(lib().orc_filter_process(self.h, _dp(ranges), _dp(odom), float(stamp)))

    def state(self):
        a = np.zeros((self.n, 6)); lib().orc_filter_state(self.h, _dp(a)); return a

    def pre_poses(self):
        a = np.zeros((self.n, 3)); lib().orc_filter_pre_poses(self.h, _dp(a)); return a

    def match_rows(self):
        a = np.zeros((self.n, 8)); lib().orc_filter_match_rows(self.h, _dp(a)); return a

    def neff(self):
        return lib().orc_filter_neff(self.h)

    def weights(self):
        a = np.zeros(self.n); lib().orc_filter_weights(self.h, _dp(a)); return a

    def last_resampled(self):
        idx = np.zeros(self.n, dtype=np.uint32); w = np.zeros(self.n); ne = C.c_double()
        r = lib().orc_filter_last_resampled(self.h, idx.ctypes.data, C.byref(ne), _dp(w))
        return (idx, ne.value, w) if r else None

    def map(self, i):
        return Map(handle=lib().orc_filter_map(self.h, i), own=False)

    def best(self):
        return lib().orc_filter_best(self.h)

    def matcher(self):
        return lib().orc_filter_matcher(self.h).contents

    def last_u01(self):
        return lib().orc_filter_last_u01(self.h)
