# -*- coding: utf-8 -*-
"""The C-ABI library loads, exports every symbol include/*.h declares, and the
product path fails loudly (no CPU fallback) when no GPU is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tricolour_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_binding():
    from tricolour_b200 import _cabi
    syms = header_symbols()
    assert len(syms) >= 30
    assert set(_cabi.EXPORTED_SYMBOLS) == set(syms)


def test_cuda_library_exports_every_symbol():
    from tricolour_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH), "build the CUDA library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    lib.tc_is_emulated.restype = ctypes.c_int
    assert lib.tc_is_emulated() == 0


def test_library_is_sm100a_sass():
    import shutil
    import subprocess
    from tricolour_b200 import _cabi
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly():
    from tricolour_b200 import _cabi
    import tricolour_b200 as tb
    _cabi._set_library_for_testing(None)
    lib = _cabi.load()
    if lib.tc_device_count() > 0:
        pytest.skip("a GPU is visible")
    vis = np.zeros((1, 1, 4, 8), np.complex64)
    flags = np.zeros((1, 1, 4, 8), bool)
    with pytest.raises(RuntimeError):
        tb.flag_nans_and_zeros(vis, flags)
    with pytest.raises(RuntimeError):
        tb.sum_threshold_flagger(vis, flags)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tricolour_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "tricolour_oracle" not in text, f
