import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; without a device they are skipped rather than failed when
    # somebody runs the whole suite on a CPU box.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    # test development on a CPU box: with the mock of the C ABI preloaded (tests/mock/gm_mock.c) the `gpu` tests can be run
    # to check the TEST code itself (the mock is the oracle, so parity holds trivially); never set by the driver
    if not has_gpu and os.environ.get("GM_TEST_RUN_GPU_TESTS_ON_MOCK") == "1" and "libgm_mock" in os.environ.get("LD_PRELOAD", ""):
        import importlib
        importlib.import_module("slam-gmapping-learn_b200").load_library(path=os.environ["LD_PRELOAD"].split(":")[0])  # ctypes binds per handle
        return
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
