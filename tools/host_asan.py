#!/usr/bin/env python
"""AddressSanitizer + UBSan over the C++ host layer, without a GPU: the host sources, the reference-facing driver
(oracle/ref_driver.cpp) or the API walk (tests/sanitize/api_walk.cpp: processScan, clone(), getTrajectories(),
integrateScanSequence(), a Particle copy read through the host mirror, destruction of both filters) and the CPU mock of the
C ABI (tests/mock/gm_mock.c + the oracle) are compiled into one instrumented binary and run over the golden logs.

    python tools/host_asan.py            # prints one line per run; non-zero exit on any sanitizer report

The sensors allocated by CarmenConfiguration::computeSensorMap() belong to the caller (the reference's drivers never free them
either); those leak records are filtered out.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import CASES, Golden  # noqa: E402
from host_common import driver_args, golden_log  # noqa: E402

HOST = os.path.join(ROOT, "slam-gmapping-learn_b200", "host")
SRCS = [os.path.join(HOST, "src", f) for f in ("utils.cpp", "motionmodel.cpp", "log.cpp", "devicemap.cpp", "scanmatcher.cpp", "scanmatcherprocessor.cpp", "gridslamprocessor.cpp")]
CSRC = [os.path.join(ROOT, "tests", "mock", "gm_mock.c"), os.path.join(ROOT, "oracle", "gm_oracle.c")]
FLAGS = ["-g", "-O1", "-std=c++14", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-w", "-I", os.path.join(HOST, "include"),
         "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle")]


def build(main_cpp, exe):
    subprocess.check_call(["g++"] + FLAGS + ["-o", exe, main_cpp] + SRCS + ["-x", "c"] + CSRC + ["-lm"])


def reports(stderr):
    """sanitizer records that are not the caller-owned sensors of computeSensorMap()"""
    bad = []
    for block in stderr.split("\n\n"):
        if ("ERROR: AddressSanitizer" in block or "runtime error" in block or
                (("Direct leak" in block or "Indirect leak" in block) and "computeSensorMap" not in block and "RangeSensor" not in block)):
            bad.append(block)
    return bad


def main():
    failed = 0
    with tempfile.TemporaryDirectory() as td:
        drv, walk = os.path.join(td, "drv_asan"), os.path.join(td, "walk_asan")
        build(os.path.join(ROOT, "oracle", "ref_driver.cpp"), drv)
        build(os.path.join(ROOT, "tests", "sanitize", "api_walk.cpp"), walk)
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
        for case in CASES:
            g = Golden(case)
            lp = os.path.join(td, case + ".log")
            golden_log(g, lp)
            cmd = [drv, "-mode", "trace", "-log", lp, "-out", os.path.join(td, "t.bin"), "-gfs", os.path.join(td, "t.gfs"), "-probe", "2",
                   "-maps-every", str(g.p["maps_every"])] + driver_args(g.p)
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            bad = reports(r.stderr)
            print("%-18s driver   rc %d  sanitizer reports %d" % (case, r.returncode, len(bad)))
            failed += len(bad) + (r.returncode not in (0, 23))  # 23: LeakSanitizer's exit code for the filtered sensor leaks
            for b in bad[:3]:
                print(b[:1500])
        lp = os.path.join(td, "room181_n30.log")
        r = subprocess.run([walk, lp], capture_output=True, text=True, env=env)
        bad = reports(r.stderr)
        print("%-18s api walk rc %d  sanitizer reports %d" % ("room181_n30", r.returncode, len(bad)))
        failed += len(bad) + (r.returncode not in (0, 23))
        for b in bad[:3]:
            print(b[:1500])
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
