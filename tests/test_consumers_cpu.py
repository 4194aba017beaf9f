"""The reference's own consumers of the hot path, compiled UNCHANGED against slam-gmapping-learn_b200/host/include (SURVEY.md 8b:
"links unchanged" = recompiles unchanged against our headers): the ROS node gmapping/src/slam_gmapping.cpp (with a stub ROS1 /
tf / rosbag / boost include tree, tests/stubs/ros1) and gfs-carmen/gfs-carmen.cpp (stub carmen.h; the consumer-side headers that are
out of scope here -- utils/commandline.h, utils/orientedboundingbox.h, configfile/configfile.h, carmenwrapper/carmenwrapper.h --
come from the consumer's tree, our include directory is searched first).  Syntax + semantics only (-fsyntax-only): the
consumers' other dependencies do not exist in this image.  Runs where /root/reference is mounted (the build container)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OURS = os.path.join(ROOT, "slam-gmapping-learn_b200", "host", "include")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "openslam_gmapping")), reason="the reference tree is not mounted here")


def compile_only(src, includes):
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-w"] + [a for i in includes for a in ("-I", i)] + [src]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    deps = subprocess.run(cmd[:2] + ["-M", "-w"] + cmd[4:], capture_output=True, text=True, timeout=300).stdout.split()
    return [d for d in deps if "/gmapping/" in d and d.endswith((".h", ".hxx"))]


def test_ros_node_compiles_unchanged_against_our_headers():
    deps = compile_only(os.path.join(REF, "gmapping", "src", "slam_gmapping.cpp"),
                        [os.path.join(ROOT, "tests", "stubs", "ros1"), OURS, os.path.join(ROOT, "include"), os.path.join(REF, "gmapping", "src")])
    lib = [d for d in deps if "/include/gmapping/" in d]
    assert lib and all(d.startswith(OURS) for d in lib), [d for d in lib if not d.startswith(OURS)]
    assert any(d.endswith("gridfastslam/gridslamprocessor.h") for d in lib) and any(d.endswith("scanmatcher/smmap.h") for d in lib)


def test_gfs_carmen_compiles_unchanged_against_our_headers():
    deps = compile_only(os.path.join(REF, "openslam_gmapping", "gfs-carmen", "gfs-carmen.cpp"),
                        [os.path.join(ROOT, "tests", "stubs", "carmen"), OURS, os.path.join(ROOT, "include"), os.path.join(REF, "openslam_gmapping", "include")])
    theirs = sorted(os.path.relpath(d, os.path.join(REF, "openslam_gmapping", "include", "gmapping")) for d in deps if d.startswith(REF))
    # only the consumer-side headers that are outside the per-scan particle update come from the consumer's tree
    assert theirs == ["carmenwrapper/carmenwrapper.h", "configfile/configfile.h", "utils/commandline.h", "utils/orientedboundingbox.h",
                      "utils/orientedboundingbox.hxx"], theirs
    assert any(d.startswith(OURS) and d.endswith("gridfastslam/gridslamprocessor.h") for d in deps)
