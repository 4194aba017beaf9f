"""The C++ host layer (slam-gmapping-learn_b200/host): GMapping::GridSlamProcessor / ScanMatcher / ScanMatcherMap with the
reference's API over the C ABI.

CPU part: the libraries build, export the C ABI declared in include/gm_b200.h, and a consumer written against the
REFERENCE's headers (oracle/ref_driver.cpp: subclasses GridSlamProcessor, overrides the three hooks, reads
m_particles / m_weights / m_indexes, walks map.storage().m_cells[..][..] and autoptr::m_reference) compiles unchanged
against ours.

GPU part (-m gpu): that driver, linked with our libraries, free-runs the golden logs on the B200 and its trace is
compared with the trace of the unmodified reference build (tests/golden): resample decisions and indexes exact,
poses within 1e-4 m / rad, weights within 1e-3 relative, map digests (n, visits) exact."""
import ctypes
import importlib
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from common import CASES, Golden, digest_rows
from host_common import BUILD, HOST, PKG, ROOT, build_driver, build_host, driver_args
from oracle.trace import read_trace

synth = importlib.import_module("slam-gmapping-learn_b200.synth")


def test_c_abi_exports_match_header():
    build_host()
    hdr = open(os.path.join(ROOT, "include", "gm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gm_[a-z_]+)\s*\(", hdr)) - {"gm_ctx"})
    lib = ctypes.CDLL(os.path.join(PKG, "libgm_b200.so"))
    for name in declared:
        assert hasattr(lib, name), "libgm_b200.so does not export %s" % name
    capi = importlib.import_module("slam-gmapping-learn_b200.capi")
    assert set(capi.EXPORTS) <= set(declared)
    assert len(declared) >= 21


def test_no_device_is_a_loud_error():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    gm = importlib.import_module("slam-gmapping-learn_b200")
    with pytest.raises(gm.GmError) as e:
        gm.GmContext(4)
    assert e.value.code == -4 and "no CPU path" in str(e.value)


def test_reference_consumer_compiles_against_our_headers():
    exe = build_driver()
    assert os.path.exists(exe)
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(HOST, "libgridfastslam_b200.so")], capture_output=True, text=True).stdout
    for sym in ("GridSlamProcessor11processScan", "GridSlamProcessor4init", "ScanMatcher12registerScan", "ScanMatcher8optimize",
                "MotionModel14drawFromMotion", "CarmenConfiguration16computeSensorMap", "sampleGaussian"):
        assert sym in out, sym


def test_drand48_stream_matches_libc():
    """the motion model steps the drand48 sequence inline (Drand48Stream): its samples and the state it hands back to
    libc must be exactly those of drand48() / sampleGaussian()"""
    import subprocess
    os.makedirs(os.path.join(ROOT, "tests", "_build"), exist_ok=True)
    exe = os.path.join(ROOT, "tests", "_build", "drand48_stream_check")
    host = os.path.join(ROOT, "slam-gmapping-learn_b200", "host")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(host, "include"), os.path.join(ROOT, "tests", "drand48_stream_check.cpp"),
                    os.path.join(host, "src", "utils.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0, next draw equal 1" in out.stdout and "uniform mismatches 0" in out.stdout


def test_product_does_not_link_the_oracle():
    out = subprocess.run(["ldd", os.path.join(HOST, "libgridfastslam_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libgm_b200.so" in out
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "gm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)


def _golden_log(g, path):
    laser = synth.LaserSpec(g.B, 1.0 if g.B == 181 else 0.25, float(g.ranges.max()))
    log = synth.SynthLog(laser, g.ranges, g.odom, g.odom, g.stamps, {})
    synth.write_carmen(log, path)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_free_running_trace_matches_reference(case):
    exe = build_driver()
    g = Golden(case)
    with tempfile.TemporaryDirectory() as td:
        lp, tp = os.path.join(td, "in.log"), os.path.join(td, "trace.bin")
        _golden_log(g, lp)
        cmd = [exe, "-mode", "trace", "-log", lp, "-out", tp, "-probe", "2", "-maps-every", str(g.p["maps_every"])] + driver_args(g.p)
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        tr = read_trace(tp)
    R = tr["readings"]
    assert len(R) == g.S
    assert np.allclose(tr["angles"], g.angles, rtol=0, atol=0)
    k_dig = 0
    dig_scans = list(g.dig_scan)
    for i, r in enumerate(R):
        assert np.array_equal(r["ranges"], g.ranges[i])
        assert r["processed"] == bool(g.processed[i]), "%s[%d] gate" % (case, i)
        assert np.abs(r["pre_pose"] - g.pre_pose[i]).max() < 1e-4, "%s[%d] motion-model poses" % (case, i)
        if not r["processed"]:
            continue
        assert r["resampled"] == bool(g.resampled[i]), "%s[%d] resample decision" % (case, i)
        if "smat" in r:
            ref = g.smat[i]
            assert np.abs(r["smat"][:, 10:13] - ref[:, 10:13]).max() < 1e-4, "%s[%d] scan-matched poses" % (case, i)
            np.testing.assert_allclose(r["smat"][:, 8], ref[:, 8], rtol=1e-3, err_msg="%s[%d] log-likelihood" % (case, i))
            np.testing.assert_allclose(r["smat"][:, 3], ref[:, 3], rtol=1e-6, err_msg="%s[%d] optimize score via m_matcher" % (case, i))
            np.testing.assert_allclose(r["probes"][:, 4:7], g.probes[i][:len(r["probes"]), 4:7], rtol=1e-6, atol=1e-9)
            # ScanMatcher::likelihood(lmax, mean, cov, ...) and optimize(mean, cov, ...) on particle 0's map
            np.testing.assert_allclose(r["like"][:5], g.like[i][:5], rtol=1e-6, atol=1e-9, err_msg="%s[%d] likelihood()" % (case, i))
            np.testing.assert_allclose(r["like"][5:11], g.like[i][5:11], rtol=1e-4, atol=1e-12, err_msg="%s[%d] likelihood() covariance" % (case, i))
            np.testing.assert_allclose(r["like"][11:15], g.like[i][11:15], rtol=1e-6, atol=1e-9, err_msg="%s[%d] optimize(mean, cov)" % (case, i))
            np.testing.assert_allclose(r["like"][15:21], g.like[i][15:21], rtol=1e-4, atol=1e-12, err_msg="%s[%d] optimize(mean, cov) covariance" % (case, i))
        if r["resampled"]:
            assert np.array_equal(r["indexes"], g.ridx[i]), "%s[%d] resample indexes" % (case, i)
            np.testing.assert_allclose(r["neff_at_resample"], g.rneff[i], rtol=1e-6)
        st = r["state"]
        assert np.abs(st[:, :3] - g.state[i][:, :3]).max() < 1e-4, "%s[%d] poses" % (case, i)
        np.testing.assert_allclose(st[:, 3:5], g.state[i][:, 3:5], rtol=1e-3, atol=1e-6, err_msg="%s[%d] weight / weightSum" % (case, i))
        assert np.array_equal(st[:, 5], g.state[i][:, 5]), "%s[%d] previousIndex" % (case, i)
        np.testing.assert_allclose(r["neff"], g.neff[i], rtol=1e-6)
        if "digests" in r and i in dig_scans:
            k = dig_scans.index(i)
            got, want = digest_rows(r["digests"]), digest_rows(g.digests[k])
            assert np.array_equal(got[:, :5], want[:, :5]), "%s[%d] integer map digests through the host mirror" % (case, i)
            assert np.array_equal(r["geometry"], g.geometry[k]), "%s[%d] map geometry" % (case, i)
            k_dig += 1
    fin = tr["final"]
    assert np.array_equal(digest_rows(fin["digests"])[:, :5], digest_rows(g.digests[-1])[:, :5])
    assert fin["best"] == int(g.best)
    assert k_dig >= 1


@pytest.mark.gpu
def test_cli_runs_and_writes_gfs():
    build_host()
    g = Golden("room181_resize")
    with tempfile.TemporaryDirectory() as td:
        lp, gp = os.path.join(td, "in.log"), os.path.join(td, "out.gfs")
        _golden_log(g, lp)
        res = subprocess.run([os.path.join(HOST, "gfs_b200"), "-filename", lp, "-outfilename", gp, "-particles", "12", "-randseed", "12345",
                              "-xmin", "-4", "-ymin", "-4", "-xmax", "4", "-ymax", "4"], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        info = json.loads(res.stdout.strip().splitlines()[-1])
        assert info["processed"] == int(g.processed.sum()) and info["read"] == g.S
        txt = open(gp).read()
        for tag in ("ODOM ", "ODO_UPDATE 12 ", "FRAME ", "LASER_READING 181 ", "SM_UPDATE 12 ", "NEFF "):
            assert tag in txt, tag
        neff = [float(l.split()[1]) for l in txt.splitlines() if l.startswith("NEFF")]
        want = g.neff[g.processed & ~np.isnan(g.smat[:, 0, 0])]
        # the .gfs NEFF record is written before resampling: compare with the reference where no resample happened
        assert len(neff) == len(want)


def test_host_shim_loads_under_python():
    """The C++ host library must be loadable into a Python process (ctypes) and stream numbers through its own
    iostreams: a toolchain that links libstdc++ statically into the .so crashes exactly here."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    L = hostapi.load()
    h = L.gmh_create(0)
    a = np.linspace(-1.0, 1.0, 181); lp = np.zeros(3)
    pv = np.array([float(hostapi.DEFAULTS[k]) for k in hostapi.PARAM_ORDER])
    L.gmh_setup(h, 181, hostapi._dp(a), hostapi._dp(lp), hostapi._dp(pv))
    L.gmh_destroy(h)


@pytest.mark.gpu
def test_clone_continues_identically_and_occupancy_grid():
    """GridSlamProcessor::clone() (gridslamprocessor.cpp:38-97): the copy, fed the same readings and the same drand48
    stream, stays bit-identical to the original -- poses, weights, resample indexes and every particle map."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    libc = ctypes.CDLL("libc.so.6")
    libc.seed48.restype = ctypes.POINTER(ctypes.c_ushort * 3)
    g = Golden("room181_n30")
    p = g.p
    a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"], seed=p["seed"],
                          **{k: p[k] for k in hostapi.PARAM_ORDER})
    for i in range(12):
        a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    b = a.clone()
    assert np.array_equal(a.state(), b.state()) and np.array_equal(a.map_digests(), b.map_digests())
    saved = (ctypes.c_ushort * 3)(*libc.seed48((ctypes.c_ushort * 3)(0, 0, 0)).contents)  # read the drand48 state ...
    libc.seed48(saved)                                                                      # ... and put it back
    for i in range(12, 24):
        a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    resampled = a.resamples()
    libc.seed48(saved)
    for i in range(12, 24):
        b.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
    assert np.array_equal(a.state(), b.state()), "clone diverged"
    assert np.array_equal(a.map_digests(), b.map_digests())
    assert b.resamples() == resampled and a.neff() == b.neff()
    assert np.abs(a.state()[:, :3] - g.state[23][:, :3]).max() < 1e-4  # and both still follow the reference
    # occupancy grid of the best particle through the C++ API
    grid, origin = a.occupancy_grid(a.best(), 0.25)
    assert grid.shape == (4000, 4000) and abs(origin[0] + 100.0) < 1e-9 and abs(origin[1] + 100.0) < 1e-9
    assert (grid == 100).sum() > 200 and (grid == 0).sum() > 20000 and (grid == -1).sum() > 15000000
    x, y = g.state[23][a.best(), 0], g.state[23][a.best(), 1]
    assert grid[int(round((y + 100) / 0.05)), int(round((x + 100) / 0.05))] == 0  # the robot stands in free space
    assert a.occupancy(a.best(), x, y) < 0.1
    a.close(); b.close()


@pytest.mark.gpu
def test_reference_order_and_early_registration_agree():
    """setearlyRegistration(false) keeps the reference's order (copy the resampled particles, then register the scan in
    every child); the default registers first and copies afterwards.  Same drand48 stream in, bit-identical filter out:
    poses, weights, resample indexes, every map -- and both follow the reference trace."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    g = Golden("room181_k2_lskip")
    p = g.p
    runs = []
    for early in (True, False):
        a = hostapi.Processor(g.angles, p["particles"], g.head_init, bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                              seed=p["seed"], early_registration=early, **{k: p[k] for k in hostapi.PARAM_ORDER})
        idx = []
        for i in range(len(g.ranges)):
            a.process_scan(g.ranges[i], g.odom[i], g.stamps[i])
            if a.last_resampled():
                idx.append(a.indexes().copy())
        runs.append((a.state(), a.map_digests(), a.resamples(), a.neff(), idx))
        a.close()
    (s0, d0, r0, n0, i0), (s1, d1, r1, n1, i1) = runs
    assert r0 == r1 and r0 > 0 and n0 == n1
    assert np.array_equal(s0, s1) and np.array_equal(d0, d1)
    assert len(i0) == len(i1) and all(np.array_equal(x, y) for x, y in zip(i0, i1))
    assert np.abs(s0[:, :3] - g.state[len(g.ranges) - 1][:, :3]).max() < 1e-4


@pytest.mark.gpu
def test_thousand_scans_free_running():
    """BASELINE.json configs[0]/[1]: 30 particles, 181-beam scans, 1000 scans, free running through the C++ API against the
    unmodified reference (fixture: Neff and flags for every scan, all poses / log-weights every 10th scan): the filter never
    leaves the reference's trajectory -- same resampling scans, same indexes, poses within 1e-4, log-weights within 1e-3."""
    build_host()
    hostapi = importlib.import_module("slam-gmapping-learn_b200.hostapi")
    z = np.load(os.path.join(ROOT, "tests", "golden", "room181_long.npz"))
    p = json.loads(str(z["params"]))
    log = synth.make_log("room", synth.LMS200, scans=int(z["scans"]), step=0.25, seed=int(z["log_seed"]))
    assert np.array_equal(log.laser.angles(), z["angles"])
    f = hostapi.Processor(z["angles"], p["particles"], z["head_init"], bounds=(p["xmin"], p["ymin"], p["xmax"], p["ymax"]), delta=p["delta"],
                          seed=p["seed"], **{k: p[k] for k in hostapi.PARAM_ORDER})
    kept = {int(s): k for k, s in enumerate(z["kept"])}
    rsc = {int(s): k for k, s in enumerate(z["ridx_scans"])}
    worst_pose = worst_w = 0.0
    for i in range(int(z["scans"])):
        done = f.process_scan(log.ranges[i], log.odom[i], log.stamps[i])
        assert done == bool(z["processed"][i])
        assert f.last_resampled() == bool(z["resampled"][i]), "scan %d: resampling decision" % i
        assert abs(f.neff() - z["neff"][i]) <= 1e-6 * z["neff"][i], "scan %d: Neff" % i
        if i in rsc:
            assert np.array_equal(f.indexes(), z["ridx"][rsc[i]]), "scan %d: resample indexes" % i
        if i in kept:
            st = f.state(); k = kept[i]
            worst_pose = max(worst_pose, float(np.abs(st[:, :3] - z["poses"][k]).max()))
            worst_w = max(worst_w, float(np.abs(st[:, 3:5] - z["weights"][k]).max() / max(1.0, np.abs(z["weights"][k]).max())))
            assert np.array_equal(st[:, 5], z["prev"][k])
    assert worst_pose < 1e-4 and worst_w < 1e-3, (worst_pose, worst_w)
    d = f.map_digests()
    assert np.array_equal(d[:, :5], digest_rows(z["digests"])[:, :5]), "final maps: (n, visits) of every cell of every particle"
    assert f.best() == int(z["best"])
    f.close()
