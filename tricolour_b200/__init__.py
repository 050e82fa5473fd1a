# -*- coding: utf-8 -*-
"""
tricolour_b200 -- B200-native (sm_100a) implementation of tricolour's flagging
hot path behind tricolour's own Python signatures.

    from tricolour_b200 import flagging, stokes, packing, window_statistics

mirror ``tricolour.flagging`` etc.; ``install()`` rebinds the names that
``tricolour.dask_wrappers`` bound at import so that the tricolour application
calls the GPU path.  All numerics run in ``libtricolour_b200.so`` (hand-written
CUDA behind the C ABI of ``include/tricolour_b200.h``); there is no CPU
fallback.
"""
from . import _cabi  # noqa: F401
from . import flagging, stokes, packing, window_statistics, strategy  # noqa: F401
from .flagging import (flag_nans_and_zeros, flag_autos, apply_static_mask,  # noqa: F401
                       sum_threshold_flagger, uvcontsub_flagger, SumThresholdFlagger)
from .stokes import (stokes_corr_map, polarised_intensity,  # noqa: F401
                     unpolarised_intensity, STOKES_TYPES)
from .packing import pack_data, unpack_data, unique_baselines  # noqa: F401
from .window_statistics import (window_stats, combine_window_stats,  # noqa: F401
                                summarise_stats, WindowStatistics, StatsLayout,
                                allreduce_window_stats)
from .strategy import StrategyExecutor  # noqa: F401

__version__ = "0.1.0"


def install():
    """Make the tricolour application call the GPU path: rebinds

    * the numpy-level functions ``tricolour.dask_wrappers`` imported under ``np_*``
      names (tricolour/dask_wrappers.py:9-18; its wrappers look the names up when
      they build a graph),
    * the per-block functions of ``tricolour.packing`` (``_fast_pack_data``
      packing.py:281-292, ``_unpack_data`` 391-415; ``pack_data`` / ``unpack_data``
      look them up per call) and
    * ``tricolour.window_statistics._window_stats`` (window_statistics.py:12-66).

    Call it after ``import tricolour`` and before the application builds its dask
    graphs.  Returns the modules it patched."""
    import importlib
    dw = importlib.import_module("tricolour.dask_wrappers")
    dw.np_flag_nans_and_zeros = flag_nans_and_zeros
    dw.np_sum_threshold_flagger = sum_threshold_flagger
    dw.np_uvcontsub_flagger = uvcontsub_flagger
    dw.np_apply_static_mask = apply_static_mask
    dw.np_flag_autos = flag_autos
    dw.np_polarised_intensity = polarised_intensity
    dw.np_unpolarised_intensity = unpolarised_intensity
    pk = importlib.import_module("tricolour.packing")
    pk._fast_pack_data = packing._fast_pack_data
    pk._unpack_data = packing._unpack_data
    ws = importlib.import_module("tricolour.window_statistics")
    ws._window_stats = window_statistics._window_stats
    return dw, pk, ws
