# -*- coding: utf-8 -*-
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libtricolour_b200_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _ensure_built():
    import subprocess
    csrc = os.path.join(ROOT, "tricolour_b200", "csrc")
    if not os.path.exists(EMU_LIB):
        subprocess.check_call(["make", "-C", csrc, "-s", "emu"])
    import oracle
    oracle.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    _ensure_built()


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """Routes tricolour_b200 through the CPU-emulated kernels ("emu", the same
    sources compiled for the host; small cases, runs anywhere) or through the
    real sm_100a library ("cuda", marked gpu)."""
    from tricolour_b200 import _cabi
    if request.param == "emu":
        _cabi._set_library_for_testing(_cabi.load(EMU_LIB))
    else:
        _cabi._set_library_for_testing(None)
        lib = _cabi.load()
        assert lib.tc_is_emulated() == 0
        assert lib.tc_device_count() >= 1, "no CUDA device visible"
    yield request.param
    _cabi._set_library_for_testing(None)


@pytest.fixture
def cuda_lib():
    """the real library only (gpu tests)"""
    from tricolour_b200 import _cabi
    _cabi._set_library_for_testing(None)
    lib = _cabi.load()
    assert lib.tc_is_emulated() == 0
    return lib


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
