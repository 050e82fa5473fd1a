# -*- coding: utf-8 -*-
"""
Loads the UNMODIFIED reference modules from /root/reference under a stub
``tricolour`` package (the real package __init__ needs donfig + dist metadata;
packing/window_statistics need dask/zarr, which are mocked -- only their numba /
numpy kernels are called).  TEST INFRASTRUCTURE: used by
tests/golden/make_golden.py and by the in-container oracle-vs-reference tests.
/root/reference does not exist on the GPU box; ``available()`` is False there.
"""
import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("TRICOLOUR_REFERENCE", "/root/reference")


def available():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "tricolour")):
        return False
    try:
        import numba  # noqa: F401
    except ImportError:
        return False
    return True


def load():
    """Returns (flagging, stokes, packing, window_statistics) reference modules."""
    if not available():
        raise RuntimeError("reference not available at %s" % REFERENCE_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/tricolour_ref_numba_cache")
    if "tricolour" not in sys.modules or not getattr(
            sys.modules["tricolour"], "_graft_stub", False):
        pkg = types.ModuleType("tricolour")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "tricolour")]
        pkg._graft_stub = True
        sys.modules["tricolour"] = pkg
        for m in ("dask", "dask.array", "dask.highlevelgraph", "dask.blockwise",
                  "dask.base", "zarr"):
            if m not in sys.modules:
                try:
                    importlib.import_module(m)
                except ImportError:
                    sys.modules[m] = mock.MagicMock()
    flagging = importlib.import_module("tricolour.flagging")
    stokes = importlib.import_module("tricolour.stokes")
    packing = importlib.import_module("tricolour.packing")
    wstats = importlib.import_module("tricolour.window_statistics")
    return flagging, stokes, packing, wstats
