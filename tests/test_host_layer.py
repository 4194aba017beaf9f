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
from host_common import BUILD, HOST, PKG, ROOT, build_driver, build_host, check_trace, driver_args, free_run, golden_log
import host_scenarios

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
                "MotionModel14drawFromMotion", "CarmenConfiguration16computeSensorMap", "sampleGaussian",
                "ScanMatcherProcessor11processScan"):
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


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_free_running_trace_matches_reference(case):
    g, tr = free_run(case)
    check_trace(case, g, tr)


@pytest.mark.gpu
def test_cli_runs_and_writes_gfs():
    host_scenarios.scenario_cli()


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
    host_scenarios.scenario_clone()


@pytest.mark.gpu
def test_reference_order_and_early_registration_agree():
    host_scenarios.scenario_order()


@pytest.mark.gpu
def test_adapt_particles_changes_the_particle_count():
    host_scenarios.scenario_adapt()


@pytest.mark.gpu
def test_lazy_tree_weights():
    host_scenarios.scenario_lazy_tree()


@pytest.mark.gpu
def test_thousand_scans_free_running():
    host_scenarios.scenario_long()


@pytest.mark.gpu
def test_c3_prefix_free_running():
    """BASELINE.json configs[2] / SURVEY.md 8d's gate: 512 particles x 1081 beams, 56 scans of the benchmark's log with the
    benchmark's parameters (the filter resamples), free running against the unmodified reference: decisions, indexes and the
    final (n, visits) of every cell of every particle exact, Neff within 1e-6, poses within 1e-4, log-weights within 1e-3"""
    worst_pose, worst_w, resamples = host_scenarios.scenario_c3()
    assert resamples >= 1
    print("C3 prefix: worst pose error %.3g, worst log-weight error %.3g, %d resamplings" % (worst_pose, worst_w, resamples))


@pytest.mark.gpu
def test_n32_fixture_single_gpu():
    """the fixture bench.py's sharded-parity prelude follows on N > 1 GPUs, here on one"""
    assert host_scenarios.scenario_n32()[2] >= 1
