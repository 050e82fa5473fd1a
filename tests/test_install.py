# -*- coding: utf-8 -*-
"""``tricolour_b200.install()`` against a stand-in ``tricolour`` module tree (dask is
absent here, so the reference's ``dask_wrappers`` cannot be imported): the stand-ins
bind the reference functions' names the way tricolour/dask_wrappers.py:9-18,
packing.py:281-292/391-415 and window_statistics.py:12-66 do, and call them
through module globals like the reference's wrappers; after ``install()`` every
call must land in the GPU library (here: its CPU-emulated build) and agree with
the oracle."""
import sys
import types

import numpy as np
import pytest

import oracle
import tricolour_b200 as tb
import common

DW_SRC = '''
def _cpu(*a, **k):
    raise AssertionError("the CPU reference function was called")
np_flag_nans_and_zeros = np_sum_threshold_flagger = np_uvcontsub_flagger = _cpu
np_apply_static_mask = np_flag_autos = np_polarised_intensity = np_unpolarised_intensity = _cpu

# the per-block callables the reference hands to dask.blockwise (dask_wrappers.py:23-145)
def sum_threshold_flagger(vis, flag, **kwargs):
    return np_sum_threshold_flagger(vis, flag, **kwargs)
def uvcontsub_flagger(vis, flag, **kwargs):
    return np_uvcontsub_flagger(vis, flag, **kwargs)
def flag_nans_and_zeros(vis_windows, flag_windows):
    return np_flag_nans_and_zeros(vis_windows, flag_windows)
def _apply_static_mask_wrapper(flag, ubl, antspos, masks, spw_chanlabels, spw_chanwidths, **kwargs):
    return np_apply_static_mask(flag, ubl[0], antspos, masks, spw_chanlabels, spw_chanwidths, **kwargs)
def flag_autos(flag, ubl):
    return np_flag_autos(flag, ubl)
def polarised_intensity(vis, stokes_pol):
    return np_polarised_intensity(vis, stokes_pol)
def unpolarised_intensity(vis, stokes_unpol, stokes_pol):
    return np_unpolarised_intensity(vis, stokes_unpol, stokes_pol)
'''

PK_SRC = '''
import numpy as np
def _cpu(*a, **k):
    raise AssertionError("the CPU reference function was called")
_fast_pack_data = _unpack_data = _cpu

def pack_block(time_inv, ubl, ant1, ant2, data, flags, ntime):
    """what dask executes for one row chunk of pack_data (packing.py:306-366): windows are
    created with their defaults (96-98, 116-117) and filled in place by the block function"""
    nbl = sum(b.shape[0] for bl in ubl for b in bl)
    vis_win = np.full((nbl, data.shape[2], ntime, data.shape[1]), np.nan + np.nan * 1j, data.dtype)
    flag_win = np.ones((nbl, data.shape[2], ntime, data.shape[1]), flags.dtype)
    _fast_pack_data(time_inv, ubl, ant1, ant2, data, flags, [vis_win], [flag_win])
    return vis_win, flag_win

def unpack_block(antenna1, antenna2, time_inv, ubl, windows):
    return _unpack_data(antenna1, antenna2, time_inv, ubl, windows)
'''

WS_SRC = '''
def _cpu(*a, **k):
    raise AssertionError("the CPU reference function was called")
_window_stats = _cpu
def window_stats_block(flag_window, ubls, chan_freqs, antenna_names, scan_no, field_name, ddid, nchanbins):
    return _window_stats(flag_window, ubls, chan_freqs, antenna_names, scan_no, field_name, ddid, nchanbins)
'''


@pytest.fixture
def fake_tricolour(backend):
    saved = {k: sys.modules.get(k) for k in ("tricolour", "tricolour.dask_wrappers", "tricolour.packing",
                                             "tricolour.window_statistics")}
    pkg = types.ModuleType("tricolour")
    pkg.__path__ = []
    mods = {}
    for name, src in (("dask_wrappers", DW_SRC), ("packing", PK_SRC), ("window_statistics", WS_SRC)):
        m = types.ModuleType("tricolour." + name)
        exec(src, m.__dict__)
        setattr(pkg, name, m)
        sys.modules["tricolour." + name] = m
        mods[name] = m
    sys.modules["tricolour"] = pkg
    yield mods
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_install_rebinds_every_block_function(fake_tricolour):
    dw, pk, ws = (fake_tricolour[k] for k in ("dask_wrappers", "packing", "window_statistics"))
    with pytest.raises(AssertionError):
        dw.flag_nans_and_zeros(np.zeros((1, 1, 2, 2), np.complex64), np.zeros((1, 1, 2, 2), bool))
    patched = tb.install()
    assert patched == (dw, pk, ws)

    nant, T, F, ncorr = 3, 12, 32, 4
    ubl = common.baselines(nant)
    nbl = ubl.shape[0]
    ants = common.antenna_layout(nant)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    vis, flags = common.make_windows(nbl, ncorr, T, F, seed=41, ubl=ubl)

    # flagging functions through the rebound np_* names
    nz = dw.flag_nans_and_zeros(vis, flags)
    assert np.array_equal(nz, oracle.flag_nans_and_zeros(vis, flags))
    kw = dict(outlier_nsigma=10, background_iterations=2, num_major_iterations=2)
    assert np.array_equal(dw.sum_threshold_flagger(vis, nz, **kw), oracle.sum_threshold_flagger(vis, nz, **kw))
    uv = dw.uvcontsub_flagger(vis, nz, major_cycles=2, taylor_degrees=5, sigma=15.0)
    assert (uv != oracle.uvcontsub_flagger(vis, nz, major_cycles=2, taylor_degrees=5, sigma=15.0)).mean() <= 1e-4
    sm = dw._apply_static_mask_wrapper(flags, [ubl], ants, masks, cf, cw, accumulation_mode="or", uvrange="0~550")
    assert np.array_equal(sm, oracle.apply_static_mask(flags, ubl, ants, masks, cf, cw, "or", "0~550"))
    assert np.array_equal(dw.flag_autos(flags, [ubl]), oracle.flag_autos(flags, [ubl]))
    rows = np.ascontiguousarray(np.nan_to_num(vis).transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr))
    smap = tb.stokes_corr_map([9, 10, 11, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    unpol = tuple(v for k, v in smap.items() if k == 'I')
    assert np.array_equal(dw.polarised_intensity(rows, pol), oracle.polarised_intensity(rows, pol))
    assert np.array_equal(dw.unpolarised_intensity(rows, unpol, pol), oracle.unpolarised_intensity(rows, unpol, pol))

    # packing block functions: two baseline chunks, three row chunks, 7 rows deleted
    # (missing baseline-times keep the window defaults, tests/test_packing.py:28-109)
    a1 = np.tile(ubl[:, 1], T).astype(np.int32)
    a2 = np.tile(ubl[:, 2], T).astype(np.int32)
    tinv = np.repeat(np.arange(T), nbl)
    rflags = np.ascontiguousarray(flags.transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr))
    keep = np.ones(T * nbl, bool)
    keep[np.random.RandomState(3).choice(T * nbl, 7, replace=False)] = False
    a1, a2, tinv, rows, rflags = a1[keep], a2[keep], tinv[keep], rows[keep], rflags[keep]
    ubl_blocks = [[ubl[:4]], [ubl[4:]]]
    want_v, want_f = oracle.pack_data(tinv, ubl, a1, a2, rows, rflags, T)
    vis_win = np.full((nbl, ncorr, T, F), np.nan + np.nan * 1j, np.complex64)
    flag_win = np.ones((nbl, ncorr, T, F), bool)
    cuts = [0, 20, 45, rows.shape[0]]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        out = pk._fast_pack_data(tinv[lo:hi], ubl_blocks, a1[lo:hi], a2[lo:hi], rows[lo:hi], rflags[lo:hi],
                                 [vis_win], [flag_win])
        assert out.shape == (1, 1, 1) and out.all()
    assert np.array_equal(vis_win, want_v, equal_nan=True) and np.array_equal(flag_win, want_f)
    v2, f2 = pk.pack_block(tinv, ubl_blocks, a1, a2, rows, rflags, T)
    assert np.array_equal(v2, want_v, equal_nan=True) and np.array_equal(f2, want_f)
    # unpack from two baseline chunks of the window
    chunks_u = [[ubl[:4]], [ubl[4:]]]
    chunks_w = [[flag_win[:4]], [flag_win[4:]]]
    back = pk.unpack_block(a1, a2, tinv, chunks_u, chunks_w)
    assert back.dtype == flag_win.dtype and np.array_equal(back, rflags)
    backv = pk.unpack_block(a1, a2, tinv, chunks_u, [[vis_win[:4]], [vis_win[4:]]])
    assert np.array_equal(backv, rows)

    # window statistics block function
    names = ["m%03d" % i for i in range(nant)]
    st = ws.window_stats_block(flag_win, [ubl], [cf], names, 2, "fld", 0, 10)
    want = oracle.window_counts(flag_win, ubl, cf, nant)
    assert int(st._counts_per_field["fld"]) == int(flag_win.sum())
    assert [int(st._counts_per_ant[n]) for n in names] == [int(x) for x in want[0]]
    assert np.array_equal(st._counts_per_ddid[0], want[6].astype(np.uint64))
    with pytest.raises(ValueError):
        pk._fast_pack_data(tinv, ubl_blocks, a1, a2, rows[:, :5], rflags[:, :5], [vis_win], [flag_win])
