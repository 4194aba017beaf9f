"""Build helpers for the C++ host layer tests: the reference-facing driver oracle/ref_driver.cpp (written against the
reference's headers and API) is compiled UNCHANGED against our headers and linked with our libraries."""
from __future__ import annotations

import importlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "slam-gmapping-learn_b200")
HOST = os.path.join(PKG, "host")
BUILD = os.path.join(ROOT, "tests", "_build")


def build_host():
    importlib.import_module("slam-gmapping-learn_b200.build").build_all()
    return os.path.join(HOST, "libgridfastslam_b200.so")


def build_driver():
    """oracle/ref_driver.cpp + our include tree + our libraries -> tests/_build/ref_driver_b200"""
    build_host()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "ref_driver_b200")
    src = os.path.join(ROOT, "oracle", "ref_driver.cpp")
    deps = [src, os.path.join(HOST, "libgridfastslam_b200.so"), os.path.join(PKG, "libgm_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-w", "-I", os.path.join(HOST, "include"), "-I", os.path.join(ROOT, "include"),
                               "-o", exe, src, "-L", HOST, "-lgridfastslam_b200", "-L", PKG, "-lgm_b200",
                               "-Wl,-rpath," + HOST, "-Wl,-rpath," + PKG])
    return exe


def driver_args(p):
    """command-line flags of ref_driver for a golden fixture's parameter dict"""
    cmd = []
    for k in ("particles", "seed", "kernelSize", "iterations", "lskip"):
        cmd += ["-" + k, str(int(p[k]))]
    for k in ("xmin", "ymin", "xmax", "ymax", "delta", "sigma", "maxrange", "maxUrange", "lstep", "astep", "lsigma", "ogain",
              "srr", "srt", "str", "stt", "linearUpdate", "angularUpdate", "resampleThreshold", "minimumScore", "period"):
        cmd += ["-" + k, repr(float(p[k]))]
    if not p["generate_map"]:
        cmd.append("-no-generateMap")
    return cmd
