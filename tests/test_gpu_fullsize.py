# -*- coding: utf-8 -*-
"""Parity against the CPU oracle AT THE PLANE SIZES bench.py times (BASELINE.json
configs[0..3]), through the Python boundary and the C ABI: the code paths that
only long lines / large ranges reach (r = 54/43 filters over 512 dumps, r = 277
over 4096 channels, 32768-sample collecting blocks of the bracket select, 2 M
sample plane medians of uvcontsub, 1024-dump line medians, 256 x 32768 planes).
The oracle runs plane by plane in a ThreadPool (seconds per plane)."""
import numpy as np
import pytest

import oracle
import tricolour_b200 as tb
import common

pytestmark = pytest.mark.gpu


def _setup(F, nant=64, autos=True):
    ubl = common.baselines(nant, autos=autos)
    ants = common.antenna_layout(nant)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    return ubl, ants, cf, cw, masks


def _report(got, want):
    nd = int((got != want).sum())
    where = np.argwhere(got != want)[:5].tolist() if nd else []
    return "%d of %d flags differ (first: %s); flagged %.3f vs %.3f" % (
        nd, got.size, where, float(got.mean()), float(want.mean()))


def test_config1_full_strategy_at_plane_size(cuda_lib):
    """configs[1]: 512 dumps x 4096 channels x 4 corr, all 12 default.yaml tasks; an
    auto-correlation, the longest and the shortest cross baseline (the uvrange mask
    of task 6 selects by baseline length)"""
    T, F, ncorr = 512, 4096, 4
    ubl, ants, cf, cw, masks = _setup(F)
    sel = common.pick_baselines(ubl, ants, 3)
    sub = ubl[sel].copy()
    vis, flags = common.make_windows(len(sel), ncorr, T, F, seed=301, ubl=sub)
    strategies = common.default_strategies()
    ex_ubl = sub.copy()
    ex_ubl[:, 0] = np.arange(len(sel))
    got = tb.StrategyExecutor(ants, ex_ubl, cf, cw, masks, strategies).apply_strategies(flags, vis)
    want = common.run_strategies_planes(oracle, strategies, vis, flags, ex_ubl, ants, masks, cf, cw)
    assert np.array_equal(got, want), _report(got, want)
    assert got[ex_ubl[:, 1] == ex_ubl[:, 2]].all() and got[flags].all()


def test_config0_single_window_step3(cuda_lib):
    """configs[0] exactly: (1, 4, 64, 4096) complex64, default.yaml step 3
    (5 major iterations x 5 background iterations, r = 54/43 on 64-dump lines)"""
    vis, flags = common.make_windows(1, 4, 64, 4096, seed=302)
    kw = dict(common.DEFAULT_STRATEGY_KW["background_flags"])
    got = tb.sum_threshold_flagger(vis, flags, **kw)
    want = oracle.sum_threshold_flagger(vis, flags, nthreads=4, **kw)
    assert np.array_equal(got, want), _report(got, want)
    # uint8 flags in, uint8 flags out (the reference's tests use uint8)
    got8 = tb.sum_threshold_flagger(vis, flags.astype(np.uint8), **kw)
    assert got8.dtype == np.uint8 and np.array_equal(got8.astype(bool), want)


def test_config1_each_sum_threshold_task_at_plane_size(cuda_lib):
    """the four sum_threshold tasks of default.yaml one by one on 512 x 4096 planes
    (step 3: r = 54/43 ... 10/8 with 5 major iterations; step 7: r = 28/277 and the
    [32, 48, 64, 128] windows; step 9: r = 1 in time)"""
    vis, flags = common.make_windows(1, 4, 512, 4096, seed=303)
    flags = oracle.flag_nans_and_zeros(vis, flags)
    for name in ("background_flags", "final_st_very_broad", "final_st_broad", "final_st_narrow"):
        kw = dict(common.DEFAULT_STRATEGY_KW[name])
        got = tb.sum_threshold_flagger(vis, flags, **kw)
        want = oracle.sum_threshold_flagger(vis, flags, nthreads=4, **kw)
        assert np.array_equal(got, want), name + ": " + _report(got, want)


def test_config1_uvcontsub_at_plane_size(cuda_lib):
    """uvcontsub on 2 M-sample planes (plane-wide medians through the sliced
    bracket select), both default.yaml parameter sets"""
    vis, flags = common.make_windows(1, 4, 512, 4096, seed=304)
    flags = oracle.flag_nans_and_zeros(vis, flags)
    for kw in (dict(major_cycles=7, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0),
               dict(major_cycles=10, or_original_from_cycle=0, taylor_degrees=25, sigma=13.0)):
        got = tb.uvcontsub_flagger(vis, flags, **kw)
        want = oracle.uvcontsub_flagger(vis, flags, **kw)
        # numpy's float32 pocketfft against the direct float64 Fourier terms: equal up to
        # threshold ties
        assert (got != want).mean() <= 1e-6, _report(got, want)


def test_config3_uvcontsub_then_very_broad_1024_dumps(cuda_lib):
    """configs[3]: (2, 4, 1024, 4096), default.yaml tasks 4 -> 7 (uvcontsub 7 cycles,
    nan/zero reflag, static mask with uvrange 0~550, final_st_very_broad)"""
    T, F, ncorr = 1024, 4096, 4
    ubl, ants, cf, cw, masks = _setup(F, autos=False)
    sel = common.pick_baselines(ubl, ants, 2)
    sub = ubl[sel].copy()
    sub[:, 0] = np.arange(len(sel))
    vis, flags = common.make_windows(len(sel), ncorr, T, F, seed=305, ubl=sub)
    strategies = common.default_strategies()[3:7]
    got = tb.StrategyExecutor(ants, sub, cf, cw, masks, strategies).apply_strategies(flags, vis)
    want = common.run_strategies_planes(oracle, strategies, vis, flags, sub, ants, masks, cf, cw)
    assert (got != want).mean() <= 1e-6, _report(got, want)


def test_config2_polarised_wideband_full_strategy(cuda_lib):
    """configs[2]: one baseline of 256 dumps x 32768 channels: rows -> polarised
    intensity (Q, U, V) + any(corr) flags -> window (1, 1, 256, 32768) -> the full
    default strategy"""
    T, F = 256, 32768
    ubl, ants, cf, cw, masks = _setup(F)
    b = common.pick_baselines(ubl, ants, 2)[1]            # the longest cross baseline
    sub = ubl[b:b + 1].copy()
    sub[:, 0] = 0
    vis4, flags4 = common.make_windows(1, 4, T, F, seed=306, ubl=sub)
    # MS row order: (row = time, chan, corr)
    rows = np.ascontiguousarray(vis4[0].transpose(1, 2, 0))
    rflags = np.ascontiguousarray(flags4[0].transpose(1, 2, 0))
    smap = tb.stokes_corr_map([9, 10, 11, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    a1 = np.full(T, sub[0, 1], np.int32)
    a2 = np.full(T, sub[0, 2], np.int32)
    tinv = np.arange(T)
    pi = tb.polarised_intensity(rows, pol)
    want_pi = oracle.polarised_intensity(rows, pol)
    assert np.array_equal(pi.view(np.uint32), want_pi.view(np.uint32)), "polarised intensity is not bit-exact"
    pf = rflags.any(axis=2, keepdims=True)
    vw, fw = tb.pack_data(tinv, sub, a1, a2, pi, pf, T)
    assert vw.shape == (1, 1, T, F)
    strategies = common.default_strategies()
    got = tb.StrategyExecutor(ants, sub, cf, cw, masks, strategies).apply_strategies(fw, vw)
    want = common.run_strategies_planes(oracle, strategies, vw, fw, sub, ants, masks, cf, cw)
    assert (got != want).mean() <= 1e-6, _report(got, want)
