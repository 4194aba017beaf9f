"""GPU parity (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
  (1) the golden traces of the unmodified reference (tests/golden), teacher-forced: every stage gets the
      reference's inputs (motion-model poses, weights, the drand48 draw) and must reproduce its outputs --
      integer cell counts (n, visits) and resample indexes bit-exact, poses within 1e-4 m / 1e-4 rad,
      log-likelihoods within 1e-3 relative (BASELINE.json north_star);
  (2) the oracle on seeded random queries, including the edge cases (zero / NaN / out-of-range beams,
      empty scans, degenerate weights).
"""
import numpy as np
import pytest

from common import CASES, Golden, digest_rows
from oracle import orc

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4   # metres / radians (north_star)
LOGW_RTOL = 1e-3  # relative (north_star)
# what the kernels actually achieve given identical inputs: FP64 with libm-level (1-2 ulp) differences
TIGHT = 1e-9


def check_maps(ctx, f, want_rows, what):
    """(n, visits) bit-exact via the order-independent digest; float acc exact or, if a (float) rounding of an
    endpoint differed in the last place, within 1e-6 of the oracle's cells."""
    from gpu_common import gpu_digest_rows, map_cells_world
    got = gpu_digest_rows(ctx)
    assert np.array_equal(got[:, :5], want_rows[:, :5]), "%s: integer map digest (patches, cells, n, visits)" % what
    bad = np.nonzero(got[:, 5] != want_rows[:, 5])[0]
    for j in bad:
        info, coords, cells = ctx.download_map(int(j))
        ocoords, ocells = f.map(int(j)).export()
        ogeo = f.map(int(j)).geometry()
        r1, a1 = map_cells_world((info.size_x2, info.size_y2), coords, cells)
        r2, a2 = map_cells_world(ogeo[:2], ocoords, ocells)
        assert np.array_equal(r1, r2)
        np.testing.assert_allclose(a1, a2, rtol=1e-6, atol=1e-5, err_msg="%s: acc of particle %d" % (what, j))


def test_teacher_forced_trace_with_few_visit_tiles(monkeypatch):
    """the same trace with room for 24 visit tiles only: the active area no longer fits one pass, so the multi-pass ray
    counting (and the chunked copy-on-write work list) must give the same maps"""
    monkeypatch.setenv("GM_B200_REGISTER_TILES", "24")
    test_teacher_forced_trace("room1081_n8")


@pytest.mark.parametrize("fraction", ["1.0", "0.25"])
def test_teacher_forced_trace_with_other_pool_splits(monkeypatch, fraction):
    """the patch pool's split into hit-capable ids and free-space ids (visit counters only) must not show: with every id
    hit-capable (1.0: the layout without the split, no promotion ever) and with few of them (0.25: patches start on the cheap
    side, move when their first end point arrives, and the free-space stack is what new patches come from) the trace is the same"""
    monkeypatch.setenv("GM_B200_HIT_FRACTION", fraction)
    test_teacher_forced_trace("room181_n30")
    test_teacher_forced_trace("room181_k2_lskip")  # end points only (generateMap off): every patch is promoted


def test_free_space_stack_runs_dry(monkeypatch):
    """a pool whose free-space side is tiny: new patches and copies fall back to hit-capable ids, nothing changes in the maps"""
    monkeypatch.setenv("GM_B200_HIT_FRACTION", "0.97")
    test_teacher_forced_trace("room181_resize")


@pytest.mark.parametrize("case", CASES)
def test_teacher_forced_trace(case):
    from gpu_common import context_for
    g = Golden(case)
    ctx = context_for(g)
    f = g.oracle_filter()  # runs alongside: supplies the drand48 draw of each resample and full maps for diagnostics
    first = True
    n_match = n_res = n_dig = 0
    for i in range(g.S):
        processed = f.process(g.ranges[i], g.odom[i], g.stamps[i])
        assert processed == bool(g.processed[i])
        if not processed:
            continue
        ranges = g.ranges[i]
        if first:
            ctx.register(ranges, g.state[i][:, :3])
            first = False
        else:
            poses, sc, s, l, c, ev = ctx.scanmatch(ranges, g.pre_pose[i], g.p["minimumScore"])
            ref = g.smat[i]
            np.testing.assert_allclose(sc, ref[:, 3], rtol=TIGHT, atol=TIGHT, err_msg="%s[%d] optimize score" % (case, i))
            assert np.abs(poses - ref[:, 10:13]).max() <= POSE_TOL, "%s[%d] scan-matched pose" % (case, i)
            np.testing.assert_allclose(poses, ref[:, 10:13], rtol=0, atol=TIGHT, err_msg="%s[%d] pose (tight)" % (case, i))
            np.testing.assert_allclose(l, ref[:, 8], rtol=LOGW_RTOL, err_msg="%s[%d] log-likelihood" % (case, i))
            np.testing.assert_allclose(l, ref[:, 8], rtol=TIGHT, atol=TIGHT, err_msg="%s[%d] log-likelihood (tight)" % (case, i))
            np.testing.assert_allclose(s, ref[:, 7], rtol=TIGHT, atol=TIGHT)
            assert np.array_equal(c, ref[:, 9].astype(np.uint32)), "%s[%d] match count" % (case, i)
            assert np.array_equal(ev, f.match_rows()[:, 7].astype(np.uint32)), "%s[%d] score() evaluations" % (case, i)
            n_match += 1
            # weights: normalize the reference's accumulated log-weights
            w, neff = ctx.normalize(ref[:, 13], g.p["ogain"])
            if g.resampled[i]:
                np.testing.assert_allclose(w, g.rw[i], rtol=1e-12, atol=0)
                np.testing.assert_allclose(neff, g.rneff[i], rtol=1e-12)
                idx, emitted = ctx.resample_indexes(g.rw[i], f.last_u01())
                assert np.array_equal(idx, g.ridx[i]), "%s[%d] resample indexes" % (case, i)
                assert emitted == g.N
                ctx.register(ranges, g.state[i][:, :3], idx)
                # the reference's MAP_CONSISTENCY_CHECK (gridslamprocessor.cpp:82-143) after every particle copy: reference counts
                # recounted from the directories, "referenced" against "off the free stack"
                assert ctx.check_refcounts() == ctx.stats()["pool_in_use"]
                n_res += 1
            else:
                np.testing.assert_allclose(w, g.weights[i], rtol=1e-12, atol=0)
                np.testing.assert_allclose(neff, g.neff[i], rtol=1e-12)
                ctx.register(ranges, g.state[i][:, :3])
        for k in np.nonzero(g.dig_scan == i)[0]:
            check_maps(ctx, f, digest_rows(g.digests[k]), "%s[%d]" % (case, i))
            geo = np.array([[m.size_x2, m.size_y2, m.patches_x, m.patches_y] for m in (ctx.map_info(j) for j in range(g.N))])
            assert np.array_equal(geo[:, :2], g.geometry[k][:, :2]) and np.array_equal(geo[:, 2:], g.geometry[k][:, 4:6])
            n_dig += 1
    assert n_match >= 10 and n_dig >= 2
    assert n_res == int(g.resampled.sum())
    st = ctx.stats()
    assert st["launches"] > 0 and st["pool_in_use"] > 0
    assert ctx.check_refcounts() == st["pool_in_use"]
    ctx.close()


def _random_world(seed, beams=181, particles=5, scans=6):
    """A few scans registered at jittered poses in both the oracle maps and the GPU maps."""
    import importlib
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    gm = importlib.import_module("slam-gmapping-learn_b200")
    laser = synth.LMS200 if beams == 181 else synth.UTM30
    log = synth.make_log("room", laser, scans=scans, step=0.4, seed=seed)
    rng = np.random.default_rng(seed)
    m = orc.make_matcher(laser.angles(), urange=laser.max_range - 1.0, range_=laser.max_range)
    ctx = gm.GmContext(particles, pool_patches=4096)
    ctx.set_laser(laser.angles())
    ctx.set_matching(urange=laser.max_range - 1.0, range_=laser.max_range)
    ctx.init_particles(particles)
    maps = [orc.Map() for _ in range(particles)]
    poses = None
    for k in range(scans):
        poses = log.odom[k][None, :] + rng.normal(0, [0.02, 0.02, 0.01], size=(particles, 3))
        r = log.ranges[k].copy()
        if k == 2:  # edge cases inside a scan: zero, NaN, beyond max range, exactly usable range
            r[3] = 0.0; r[7] = np.nan; r[11] = laser.max_range + 5; r[15] = laser.max_range - 1.0
        for j in range(particles):
            orc.register(m, maps[j], poses[j], r)
        ctx.register(r, poses)
    return log, m, ctx, maps, poses, rng


@pytest.mark.parametrize("beams", [181, 1081])
def test_score_and_likelihood_vs_oracle(beams):
    log, m, ctx, maps, poses, rng = _random_world(7, beams=beams)
    n = len(maps)
    got = np.stack([np.array(ctx.digests()[k]) for k in ("patches", "cells", "sum_n", "sum_visits")], axis=1)
    want = np.array([[d.patches, d.cells, d.sum_n, d.sum_visits] for d in (mp.digest() for mp in maps)])
    assert np.array_equal(got.astype(np.int64), want)
    r = log.ranges[-1].copy()
    r[5] = 0.0  # score() skips r == 0, likelihoodAndScore() does not (scanmatcher.h:195 vs :303-308)
    r[9] = np.nan  # neither skips NaN: the x86 reference turns it into cell INT_MIN (unknown)
    q = 64
    idx = rng.integers(0, n, size=q).astype(np.uint32)
    qp = poses[idx] + rng.normal(0, [0.05, 0.05, 0.03], size=(q, 3))
    s_gpu = ctx.score(idx, qp, r)
    s2, l2, c2 = ctx.likelihood_and_score(idx, qp, r)
    for k in range(q):
        s = orc.score(m, maps[idx[k]], qp[k], r)
        so, lo, co = orc.likelihood_and_score(m, maps[idx[k]], qp[k], r)
        assert abs(s_gpu[k] - s) <= TIGHT * max(1.0, abs(s))
        assert abs(s2[k] - so) <= TIGHT * max(1.0, abs(so)) and abs(l2[k] - lo) <= TIGHT * max(1.0, abs(lo)) and c2[k] == co
    assert s_gpu.max() > 10  # the queries really hit the map
    ctx.close()


def test_scanmatch_vs_oracle_random():
    log, m, ctx, maps, poses, rng = _random_world(11, beams=181, particles=16, scans=5)
    r = log.ranges[-1]
    init = poses + rng.normal(0, [0.08, 0.08, 0.04], size=poses.shape)
    out, sc, s, l, c, ev = ctx.scanmatch(r, init, 0.0)
    for j in range(len(maps)):
        m.score_calls = 0
        so, po = orc.optimize(m, maps[j], init[j], r)
        assert ev[j] == m.score_calls
        assert abs(sc[j] - so) <= TIGHT * max(1.0, abs(so))
        assert np.abs(out[j] - po).max() <= TIGHT
        s1, l1, c1 = orc.likelihood_and_score(m, maps[j], po, r)
        assert abs(l[j] - l1) <= TIGHT * max(1.0, abs(l1)) and c[j] == c1
    st = ctx.stats()
    assert st["score_evals"] == int(ev.sum()) and st["beam_evals"] > 0
    ctx.close()


def test_empty_and_degenerate_scans():
    """all-zero scan: nothing matches, nothing is registered; minimum score keeps the odometry pose."""
    log, m, ctx, maps, poses, rng = _random_world(3, beams=181, particles=4, scans=3)
    before = ctx.digests()
    zero = np.zeros(181)
    out, sc, s, l, c, ev = ctx.scanmatch(zero, poses, 0.0)
    assert np.array_equal(out, poses) and np.all(sc == 0) and np.all(ev == 1 + 6 * (m.opt_recursive_iterations + 1))
    # likelihoodAndScore still evaluates r == 0 beams (they match the robot's own cell or nothing)
    for j in range(4):
        s1, l1, c1 = orc.likelihood_and_score(m, maps[j], poses[j], zero)
        assert abs(l[j] - l1) <= TIGHT * max(1.0, abs(l1)) and c[j] == c1
    ctx.register(zero, poses)
    after = ctx.digests()
    assert np.array_equal(before, after)
    # a huge minimum score rejects every match
    out, sc, *_ = ctx.scanmatch(log.ranges[2], poses + 0.03, 1e9)
    assert np.array_equal(out, poses + 0.03) and np.all(sc > 0)
    ctx.close()


def test_normalize_and_resample_vs_oracle():
    import importlib
    gm = importlib.import_module("slam-gmapping-learn_b200")
    ctx = gm.GmContext(4096, pool_patches=64)
    ctx.set_laser(np.linspace(-1, 1, 181)); ctx.set_matching(); ctx.init_particles(4096)
    rng = np.random.default_rng(5)
    for n in (1, 2, 30, 513, 4096):
        for spread in (0.0, 1.0, 50.0, 5000.0):
            lw = rng.normal(0, 1, size=n) * spread - 300.0
            w, neff = ctx.normalize(lw, 3.0)
            wo, neffo = orc.normalize(lw, 3.0)
            np.testing.assert_allclose(w, wo, rtol=1e-13, atol=0)
            assert abs(neff - neffo) <= 1e-12 * neffo
            for u in (0.0, 0.25, 0.999999, float(rng.random())):
                idx, em = ctx.resample_indexes(wo, u)
                io, ko = orc.resample_indexes(wo, u)
                assert np.array_equal(idx, io), (n, spread, u)
                assert em == min(ko, n)
    # all the weight on one particle
    w = np.zeros(64); w[17] = 1.0
    idx, em = ctx.resample_indexes(w, 0.5)
    assert np.all(idx == 17) and em == 64
    ctx.close()


def test_resample_and_copy_with_a_new_particle_count():
    """processScan's adaptParticles on the device: k_resample with `nparticles` against the reference's own index vectors
    (tests/golden/units.npz, made by oracle/ref_units.cpp from uniform_resampler::resampleIndexes(w, nparticles)), then
    gm_copy_particles_n to fewer and to more particles: child j's map is parent idx[j]'s, the reference counts stay consistent,
    and the next registration works on the new generation"""
    import ctypes
    import importlib
    import os
    gm = importlib.import_module("slam-gmapping-learn_b200")
    units = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "units.npz"))
    libc = ctypes.CDLL("libc.so.6"); libc.drand48.restype = ctypes.c_double; libc.srand48.argtypes = [ctypes.c_long]
    ctx = gm.GmContext(4096, pool_patches=64)
    ctx.set_laser(np.linspace(-1, 1, 181)); ctx.set_matching(); ctx.init_particles(4096)
    w_off = i_off = 0
    for n, n_out, seed in units["ra_hdr"]:
        w = units["ra_w"][w_off:w_off + n]; want = units["ra_idx"][i_off:i_off + n_out]
        w_off += n; i_off += n_out
        libc.srand48(int(seed))
        idx, em = ctx.resample_indexes(w, libc.drand48(), n_out=int(n_out))
        assert np.array_equal(idx, want) and em == n_out, (n, n_out, seed)
    ctx.close()
    # 8 particles -> 5 -> 11 (the context has room for 12)
    log, m, _, maps, poses, rng = _random_world(13, beams=181, particles=8, scans=3)
    _.close()
    laser = importlib.import_module("slam-gmapping-learn_b200.synth").LMS200
    ctx = gm.GmContext(12, pool_patches=4096)
    ctx.set_laser(laser.angles()); ctx.set_matching(urange=laser.max_range - 1.0, range_=laser.max_range); ctx.init_particles(8)
    rng = np.random.default_rng(13)
    for k in range(3):  # the same registrations as _random_world's (same seed)
        ps = log.odom[k][None, :] + rng.normal(0, [0.02, 0.02, 0.01], size=(8, 3))
        r = log.ranges[k].copy()
        if k == 2:
            r[3] = 0.0; r[7] = np.nan; r[11] = laser.max_range + 5; r[15] = laser.max_range - 1.0
        ctx.register(r, ps)
    base = ctx.digests()
    assert [int(x) for x in base["hash_counts"]] == [mp.digest().hash_counts for mp in maps]
    gen = list(range(8))
    for idx in (np.array([0, 2, 2, 5, 7], dtype=np.uint32), np.array([0, 0, 1, 1, 1, 2, 3, 3, 4, 4, 4], dtype=np.uint32)):
        ctx.copy_particles(idx, n_new=len(idx))
        gen = [gen[i] for i in idx]
        d = ctx.digests()
        assert len(d["hash_counts"]) == len(idx)
        assert [int(x) for x in d["hash_counts"]] == [int(base["hash_counts"][g]) for g in gen]
        ctx.check_refcounts()
    # the new generation registers a scan like any other: against the oracle maps of the ancestors
    r = log.ranges[1]
    ps = log.odom[1][None, :] + rng.normal(0, [0.02, 0.02, 0.01], size=(len(gen), 3))
    ctx.register(r, ps)
    d = ctx.digests()
    for j, g in enumerate(gen):
        mp = maps[g].clone(); orc.register(m, mp, ps[j], r)
        assert int(d["hash_counts"][j]) == mp.digest().hash_counts and int(d["hash_acc"][j]) == mp.digest().hash_acc
    ctx.check_refcounts()
    ctx.close()


def test_copy_on_write_refcounts():
    """resample to a single parent: children share patches (no deep copy), the next registration makes the
    touched patches private again, and the pool returns to its steady size."""
    log, m, ctx, maps, poses, rng = _random_world(9, beams=181, particles=6, scans=3)
    in_use0 = ctx.stats()["pool_in_use"]
    idx = np.zeros(6, dtype=np.uint32) + 2
    r = log.ranges[2]
    p2 = np.repeat(poses[2:3], 6, axis=0)
    ctx.register(r, p2, idx)
    st = ctx.stats()
    d = ctx.digests()
    assert len(set(d["hash_counts"].tolist())) == 1  # identical maps
    assert st["cow_copies"] > 0
    # every child copied only the active patches; the parent's other patches are shared
    patches = int(d["patches"][0])
    assert st["pool_in_use"] < 6 * patches + 1 and st["pool_in_use"] >= patches
    # oracle: clone map 2 six times and register
    want = []
    for j in range(6):
        mp = maps[2].clone(); orc.register(m, mp, p2[j], r); want.append(mp.digest())
    assert all(int(d["hash_counts"][j]) == want[j].hash_counts for j in range(6))
    assert in_use0 > 0
    assert ctx.check_refcounts() == st["pool_in_use"]  # shared patches are counted once, with the right number of owners
    ctx.close()


def test_fused_scanmatch_weights_equals_the_separate_calls():
    """gm_scanmatch_weights (one stream-ordered sequence, one synchronisation) against gm_scanmatch followed by the host's
    `weight += l` and gm_normalize_neff: the same bits -- poses, scores, log-likelihoods, weights, Neff"""
    log, m, ctx, maps, poses, rng = _random_world(21, beams=181, particles=7, scans=4)
    r = log.ranges[3]
    init = poses + rng.normal(0, [0.03, 0.03, 0.01], size=poses.shape)
    logw = rng.normal(-40.0, 25.0, size=7)
    p1, sc, s, l, c, ev = ctx.scanmatch(r, init, 0.0)
    w1, neff1 = ctx.normalize(logw + l, 3.0)
    rows, w2, neff2 = ctx.scanmatch_weights(r, init, logw, 0.0, 3.0)
    assert np.array_equal(rows[:, :3], p1) and np.array_equal(rows[:, 3], sc) and np.array_equal(rows[:, 4], l)
    assert np.array_equal(w1, w2) and neff1 == neff2
    # and against the oracle, like everything else
    for j in range(7):
        so, po = orc.optimize(m, maps[j], init[j], r)
        assert np.abs(rows[j, :3] - po).max() < 1e-9 and abs(rows[j, 3] - so) <= 1e-9 * so
    wo, neffo = orc.normalize(logw + l, 3.0)
    np.testing.assert_allclose(w2, wo, rtol=1e-13, atol=0)
    ctx.close()


@pytest.mark.parametrize("beams", [181, 1081])
def test_wide_scanmatch_shape_gives_the_same_bits(monkeypatch, beams):
    """k_scanmatch's wide shape (2 x 288 threads per particle, chosen for few particles per GPU: GM_B200_SM_WIDE_BELOW) sums the
    same beams in the same order as the narrow one: poses, scores, likelihoods, match counts and evaluation counts are
    bit-identical, in score() mode (hill climb) and likelihoodAndScore() mode (final evaluation)"""
    out = []
    for below in ("0", "1000000"):
        monkeypatch.setenv("GM_B200_SM_WIDE_BELOW", below)
        log, m, ctx, maps, poses, rng = _random_world(33, beams=beams, particles=9, scans=4)
        r = log.ranges[3]
        init = poses + rng.normal(0, [0.04, 0.04, 0.015], size=poses.shape)
        logw = rng.normal(-40.0, 25.0, size=9)
        res = ctx.scanmatch(r, init, 0.0)
        rows, w, neff = ctx.scanmatch_weights(r, init, logw, 0.0, 3.0)
        out.append(tuple(res) + (rows, w, np.float64(neff)))
        ctx.close()
    for a, b in zip(*out):
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_register_then_copy_equals_copy_then_register():
    """GridSlamProcessor's early registration: registering the scan in every particle at its own pose and then copying
    particle j <- idx[j] gives, bit for bit, the maps of the reference's order (copy, then register each child at its
    parent's pose) -- including the derived planes gm_map_digest verifies."""
    def world():
        return _random_world(11, beams=181, particles=8, scans=3)
    idx = np.array([0, 0, 2, 2, 2, 5, 7, 7], dtype=np.uint32)
    log, m, a, maps, poses, rng = world()
    r = log.ranges[2]
    own = poses[:8] + rng.normal(0, [0.02, 0.02, 0.01], size=(8, 3))  # each particle's scan-matched pose
    a.register(r, own[idx], idx)  # the reference's order
    da = a.digests()
    log, m, b, maps_b, poses_b, rng_b = world()
    b.register(r, own)  # early registration ...
    b.copy_particles(idx)  # ... then the resampling copies
    db = b.digests()
    for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc"):
        assert np.array_equal(da[k], db[k]), k
    # and a further scan registers identically on both (copy-on-write of the shared patches)
    nxt = own[idx] + 0.01
    a.register(log.ranges[1], nxt); b.register(log.ranges[1], nxt)
    da, db = a.digests(), b.digests()
    for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc"):
        assert np.array_equal(da[k], db[k]), k
    a.close(); b.close()


def test_download_upload_round_trip():
    """gm_download_map -> gm_upload_map into another context: same digest (which also proves the derived planes --
    visit plane, means, occupancy bitmap -- rebuilt on install), same scan-matching results, and the uploaded maps keep
    registering like the originals."""
    import importlib
    gm = importlib.import_module("slam-gmapping-learn_b200")
    log, m, a, maps, poses, rng = _random_world(13, beams=181, particles=4, scans=4)
    b = gm.GmContext(4, pool_patches=4096)
    b.set_laser(log.laser.angles())
    b.set_matching(urange=80.0, range_=81.0)
    b.init_particles(4)
    for j in range(4):
        info, coords, cells = a.download_map(j)
        assert coords.shape[0] > 0
        b.upload_map(j, info, coords, cells)
    da, db = a.digests(), b.digests()
    for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc"):
        assert np.array_equal(da[k], db[k]), k
    init = poses + rng.normal(0, [0.03, 0.03, 0.01], size=(4, 3))
    ra, rb = a.scanmatch(log.ranges[3], init, 0.0), b.scanmatch(log.ranges[3], init, 0.0)
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    a.register(log.ranges[3], ra[0]); b.register(log.ranges[3], rb[0])
    da, db = a.digests(), b.digests()
    for k in ("patches", "cells", "sum_n", "sum_visits", "hash_counts", "hash_acc"):
        assert np.array_equal(da[k], db[k]), k
    a.close(); b.close()


def test_errors_are_loud():
    import importlib
    gm = importlib.import_module("slam-gmapping-learn_b200")
    ctx = gm.GmContext(4, pool_patches=8)  # far too small a pool
    with pytest.raises(gm.GmError):
        ctx.scanmatch(np.zeros(3), np.zeros((4, 3)))  # laser not set
    ctx.set_laser(np.linspace(-1.5, 1.5, 181)); ctx.set_matching(); ctx.init_particles(4)
    with pytest.raises(gm.GmError) as e:
        ctx.register(np.full(181, 20.0), np.zeros((4, 3)))
    assert e.value.code == -3  # GM_ERR_POOL
    with pytest.raises(gm.GmError):
        ctx.set_laser(np.zeros(2048))  # beams must be < 2048 (reference asserts, scanmatcher.cpp:704)
    ctx.close()


def test_occupancy_export_and_standalone_optimize():
    """gm_export_occupancy (the ROS wrapper's grid fill, slam_gmapping.cpp:956-993) and gm_optimize against the oracle."""
    log, m, ctx, maps, poses, rng = _random_world(5, beams=181, particles=3, scans=4)
    grid = ctx.occupancy_grid(1, 0.25)
    coords, cells = maps[1].export()
    sx2, sy2, npx, npy = maps[1].geometry()
    want = np.full((npy * 32, npx * 32), -1, dtype=np.int8)
    for (px, py), c in zip(coords, cells):
        c = c.reshape(32, 32)  # [lx, ly]
        v = np.where(c["visits"] == 0, -1, np.where(c["n"] / np.maximum(c["visits"], 1) > 0.25, 100, 0)).astype(np.int8)
        want[py * 32:(py + 1) * 32, px * 32:(px + 1) * 32] = v.T
    assert grid.shape == want.shape and np.array_equal(grid, want)
    assert (grid == 100).sum() > 50 and (grid == 0).sum() > 1000
    init = poses + rng.normal(0, [0.05, 0.05, 0.02], size=poses.shape)
    out, sc = ctx.optimize(np.arange(3, dtype=np.uint32), init, log.ranges[-1])
    for j in range(3):
        so, po = orc.optimize(m, maps[j], init[j], log.ranges[-1])
        assert abs(sc[j] - so) <= TIGHT * max(1.0, abs(so)) and np.abs(out[j] - po).max() <= TIGHT
    ctx.close()


@pytest.mark.parametrize("variant", ["laser_offset", "odo_gain", "beams_skip", "kernel0", "kernel3", "threshold", "beams2047", "lskip3",
                                     "delta025", "delta03", "corner_origin", "corner_origin_k2", "half_cell_centre", "centre_offset"])
def test_parameter_variants_vs_oracle(variant):
    """Paths that the default configuration never takes: laser mounted off-centre (scanmatcher.h:177-180), the odometry
    prior of optimize() (scanmatcher.cpp:497-509), initialBeamsSkip, other kernel sizes, a non-default fullnessThreshold
    (generic occupancy test), the largest beam count the reference allows (2047), likelihoodSkip; other grid resolutions
    (0.025 m as in BASELINE.json's fifth configuration; 0.03 m, whose reciprocal is not a short binary fraction); and map
    frames whose centre is not the origin.  The centre matters because of the reference's `ipfree = world2map(pfree - phit)`
    (scanmatcher.h:208-210, 320-322): world2map of a DIFFERENCE adds sizeX2 - centre/delta, so with the map's lower corner at
    the origin ("corner_origin": world 0..40 m, the room around (20, 20)) the free cell really is the cell in front of the hit
    and the scan matcher's free-cell test reads allocated, visited cells for every beam -- with a centred map it lands half a
    map away on unknown cells.  (In that frame score()'s own free cell, 3.5 mm in front of the hit, IS the hit cell, so score() is
    zero everywhere and only likelihoodAndScore matches; "half_cell_centre" moves the frame by half a cell, where the rounding
    puts score()'s free cell one cell in front of or on the hit depending on the beam direction.)"""
    import importlib
    gm = importlib.import_module("slam-gmapping-learn_b200")
    synth = importlib.import_module("slam-gmapping-learn_b200.synth")
    laser = synth.LaserSpec(2047, 270.0 / 2046, 30.0) if variant == "beams2047" else synth.LMS200
    log = synth.make_log("room", laser, scans=5, step=0.4, seed=21)
    kw = dict(urange=laser.max_range - 1.0, range_=laser.max_range)
    lp = (0.0, 0.0, 0.0)
    extra = {}
    if variant == "laser_offset": lp = (0.2, 0.05, 0.1)
    if variant == "odo_gain": extra = dict(angular_odometry_reliability=2.0, linear_odometry_reliability=5.0)
    if variant == "beams_skip": extra = dict(initial_beams_skip=7)
    if variant == "kernel0": kw["kernel"] = 0
    if variant == "kernel3": kw["kernel"] = 3
    if variant == "threshold": extra = dict(fullness_threshold=0.25)
    if variant == "lskip3": kw["lskip"] = 3
    bounds, delta, shift = (-100., -100., 100., 100.), 0.05, np.zeros(3)
    if variant == "delta025": delta = 0.025
    if variant == "delta03": delta = 0.03
    if variant.startswith("corner_origin"): bounds, shift = (0., 0., 40., 40.), np.array([20., 20., 0.])
    if variant == "corner_origin_k2": kw["kernel"] = 2
    if variant == "half_cell_centre": bounds, shift = (0.025, 0.025, 40.025, 40.025), np.array([20., 20., 0.])
    if variant == "centre_offset": bounds, shift = (-30., -10., 50., 90.), np.array([7.3, 38.1, 0.])
    log.odom = log.odom + shift[None, :]  # the same room somewhere else in the world
    n = 4
    m = orc.make_matcher(laser.angles(), lp, **kw)
    for k, v in extra.items():
        setattr(m, k, v)
    ctx = gm.GmContext(n, pool_patches=16384 if delta < 0.05 else 4096)
    ctx.set_laser(laser.angles(), lp)
    ctx.set_matching(**kw, **extra)
    ctx.init_particles(n, bounds=bounds, delta=delta)
    maps = [orc.Map((bounds[0] + bounds[2]) * .5, (bounds[1] + bounds[3]) * .5, bounds[2] - bounds[0], bounds[3] - bounds[1], delta) for _ in range(n)]
    rng = np.random.default_rng(3)
    for k in range(4):
        poses = log.odom[k][None, :] + rng.normal(0, [0.02, 0.02, 0.01], size=(n, 3))
        for j in range(n):
            orc.register(m, maps[j], poses[j], log.ranges[k])
        ctx.register(log.ranges[k], poses)
    d = ctx.digests()
    for j in range(n):
        od = maps[j].digest()
        assert (int(d["patches"][j]), int(d["cells"][j]), int(d["sum_n"][j]), int(d["sum_visits"][j]), int(d["hash_counts"][j])) == \
               (od.patches, od.cells, od.sum_n, od.sum_visits, od.hash_counts), variant
    r = log.ranges[4]
    init = log.odom[4][None, :] + rng.normal(0, [0.05, 0.05, 0.03], size=(n, 3))
    out, sc, s, l, c, ev = ctx.scanmatch(r, init, 0.0)
    for j in range(n):
        m.score_calls = 0
        so, po = orc.optimize(m, maps[j], init[j], r)
        assert ev[j] == m.score_calls, variant
        assert abs(sc[j] - so) <= TIGHT * max(1.0, abs(so)) and np.abs(out[j] - po).max() <= TIGHT, variant
        s1, l1, c1 = orc.likelihood_and_score(m, maps[j], po, r)
        assert abs(l[j] - l1) <= TIGHT * max(1.0, abs(l1)) and c[j] == c1 and abs(s[j] - s1) <= TIGHT * max(1.0, abs(s1)), variant
    if variant.startswith("corner_origin"):
        assert sc.max() == 0 and c.max() > 20  # score() cannot match in this frame (see above), likelihoodAndScore does
    else:
        assert sc.max() > 5
    if variant.startswith("corner_origin") or variant == "half_cell_centre":
        # the point of this variant: the free-cell test must have rejected matches that a centred map accepts
        mc = orc.make_matcher(laser.angles(), lp, **kw)
        centred = [orc.Map() for _ in range(n)]
        rng2 = np.random.default_rng(3)
        for k in range(4):
            poses = (log.odom[k] - shift)[None, :] + rng2.normal(0, [0.02, 0.02, 0.01], size=(n, 3))
            for j in range(n):
                orc.register(mc, centred[j], poses[j], log.ranges[k])
        s_c, _, c_c = orc.likelihood_and_score(mc, centred[0], init[0] - shift, r)
        s_o, _, c_o = orc.likelihood_and_score(m, maps[0], init[0], r)
        assert c_o != c_c or abs(s_o - s_c) > 1e-6, "the free-cell test made no difference: the variant does not exercise it"
    ctx.close()
