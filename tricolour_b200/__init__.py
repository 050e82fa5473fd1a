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
                                summarise_stats, WindowStatistics)
from .strategy import StrategyExecutor  # noqa: F401

__version__ = "0.1.0"


def install():
    """Rebind the numpy-level functions that ``tricolour.dask_wrappers`` imported
    under ``np_*`` names (tricolour/dask_wrappers.py:9-18) and the private block
    functions of ``tricolour.packing`` / ``tricolour.window_statistics``."""
    import tricolour.dask_wrappers as dw
    dw.np_flag_nans_and_zeros = flag_nans_and_zeros
    dw.np_sum_threshold_flagger = sum_threshold_flagger
    dw.np_uvcontsub_flagger = uvcontsub_flagger
    dw.np_apply_static_mask = apply_static_mask
    dw.np_flag_autos = flag_autos
    dw.np_polarised_intensity = polarised_intensity
    dw.np_unpolarised_intensity = unpolarised_intensity
    import tricolour.window_statistics as ws
    ws._window_stats = window_statistics._window_stats
    return dw
