# -*- coding: utf-8 -*-
"""Pins the CPU oracle: (1) against outputs of the UNMODIFIED reference stored in
tests/golden (made by tests/golden/make_golden.py in the build container) and
(2) against the known-answer vectors of the reference's own unit tests
(tricolour/tests/test_flagging.py).  Runs without a GPU."""
import numpy as np

import oracle
import common
from conftest import golden


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(a, b, equal_nan=a.dtype.kind in "fc")


def test_golden_sum_threshold_flagger():
    g = golden("sum_threshold_flagger.npz")
    vis, flags = g["vis"], g["flags"]
    for name in ("background_flags", "final_st_very_broad", "final_st_narrow"):
        kw = dict(common.DEFAULT_STRATEGY_KW[name])
        if name == "background_flags":
            kw["num_major_iterations"] = 2
        same(oracle.sum_threshold_flagger(vis, flags, nthreads=4, **kw), g[name])
    same(oracle.sum_threshold_flagger(vis, flags, nthreads=4), g["defaults"])
    same(oracle.sum_threshold_flagger(vis, flags, average_freq=2, windows_freq=[2, 4, 8, 16],
                                      num_major_iterations=1), g["avg2"])


def test_golden_stages():
    g = golden("stages.npz")
    data, fl, ce = g["data"], g["flags"], g["chunk_ends"]
    same(oracle._get_background2d(data, fl, 5, np.array((12.5, 10.0)), 2.0, ce), g["background"])
    mf = np.zeros_like(data)
    oracle.masked_gaussian_filter(data, fl, np.array((12.5, 10.0)), mf)
    same(mf, g["masked_filter"])
    tm, tmf = oracle._time_median(data, fl)
    same(tm, g["time_median"])
    same(tmf, g["time_median_flags"])
    res = data - g["background"]
    same(oracle._sum_threshold(res, fl, 0, np.array([1, 2, 4, 8]), 10, 1.3), g["st_time"])
    same(oracle._sum_threshold(res, fl, 1, np.array([1, 2, 4, 8]), 10, 1.3, ce), g["st_freq"])


def test_golden_uvcontsub_and_companions():
    g = golden("uvcontsub.npz")
    same(oracle.uvcontsub_flagger(g["vis"].copy(), g["flags"], major_cycles=7, or_original_from_cycle=1,
                                  taylor_degrees=20, sigma=15.0), g["cycles7"])
    same(oracle.uvcontsub_flagger(g["vis"].copy(), g["flags"], major_cycles=3, or_original_from_cycle=0,
                                  taylor_degrees=25, sigma=13.0), g["cycles3_or0"])
    c = golden("companions.npz")
    same(oracle.flag_nans_and_zeros(g["vis"], g["flags"]), c["nanzero"])
    for tag, ct in (("lin", [9, 10, 11, 12]), ("circ", [5, 6, 7, 8]), ("mixed", [11, 9, 10, 12])):
        m = oracle.stokes_corr_map(ct)
        pol = tuple(x for k, x in m.items() if k != 'I')
        unpol = tuple(x for k, x in m.items() if k == 'I')
        same(oracle.polarised_intensity(c["rowvis"], pol), c["pol_" + tag])
        same(oracle.unpolarised_intensity(c["rowvis"], unpol, pol), c["unpol_" + tag])
    p = golden("packing.npz")
    vw, fw = oracle.pack_data(p["time_inv"], p["ubl"], p["ant1"], p["ant2"], p["vis"], p["flags"], int(p["ntime"]))
    same(vw, p["vis_win"])
    same(fw, p["flag_win"])
    same(oracle.unpack_data(p["ant1"], p["ant2"], p["time_inv"], p["ubl"], fw), p["unpacked"])
    antc, ants, blc, plane, tot, totsz, bins, edges = oracle.window_counts(fw, p["ubl"], p["chan_freqs"], 6, 10)
    assert antc.tolist() == p["counts_per_ant"].tolist() and ants.tolist() == p["size_per_ant"].tolist()
    assert blc.tolist() == p["counts_per_bl"].tolist()
    assert tot == int(p["counts_field"]) and totsz == int(p["size_scan"])
    assert np.array_equal(bins.astype(np.uint64), p["bins"]) and np.array_equal(edges, p["bin_edges"])


# ---- known-answer vectors of tricolour/tests/test_flagging.py -----------------
def test_average_freq_vectors():
    data = np.arange(30, dtype=np.float32).reshape(1, 5, 6).repeat(2, axis=0)
    flags = np.zeros(data.shape, np.bool_)
    flags[0, 3, :] = 1
    flags[0, :, 4] = 1
    flags[:, 2, 0] = 1
    flags[:, 2, 5] = 1
    d2, f2 = oracle._average_freq(data, flags, oracle._as_min_dtype(2))
    same(d2[0], np.array([[0.5, 2.5, 5.0], [6.5, 8.5, 11.0], [13.0, 14.5, 0.0], [0.0, 0.0, 0.0],
                          [24.5, 26.5, 29.0]], np.float32))
    same(f2[0], np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1], [1, 1, 1], [0, 0, 0]], bool))
    d4, _ = oracle._average_freq(data, flags, oracle._as_min_dtype(4))
    same(d4[1], np.array([[1.5, 4.5], [7.5, 10.5], [14.0, 16.0], [19.5, 22.5], [25.5, 28.5]], np.float32))


def test_median_vectors():
    data = np.array([[2.0, 1.0, 2.0, 5.0], [3.0, 1.0, 8.0, 6.0], [4.0, 1.0, 4.0, 7.0],
                     [5.0, 1.0, 5.0, 6.5], [1.5, 1.0, 1.5, 5.5]], np.float32)
    flags = np.array([[0, 1, 0, 1], [0, 1, 1, 0], [0, 1, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1]], np.bool_)
    od, of = oracle._time_median(data, flags)
    same(od, np.array([[3.0, 0.0, 3.0, 6.0]], np.float32))
    same(of, np.array([[0, 1, 0, 0]], np.bool_))
    d = np.array([[-2.0, -6.0, 4.5], [1.5, 3.3, 0.5]], np.float32)
    f = np.array([[0, 0, 0], [0, 1, 0]], np.uint8)
    assert oracle._median_abs(d, f) == 2.0
    assert np.isnan(oracle._median_abs(d, np.ones_like(f)))
    same(oracle._median_abs_axis0(d, f), np.array([[1.75, 6.0, 2.5]], np.float32))
    f[:, 1] = True
    out = oracle._median_abs_axis0(d, f)
    assert out[0, 0] == 1.75 and np.isnan(out[0, 1]) and out[0, 2] == 2.5


def test_box_filter_vectors():
    a = np.array([50.0, 10.0, 60.0, -70.0, 30.0, 20.0, -15.0], np.float32)
    b = np.empty_like(a)
    oracle._box_gaussian_filter1d(a, 2, b, 1)
    same(b, np.array([24.0, 10.0, 16.0, 10.0, 5.0, -7.0, 7.0], np.float32))
    # impulse response: sum 1, symmetric, std ~ sigma (tests/test_flagging.py:236-251)
    x = np.zeros((1, 200), np.float32)
    x[:, 100] = 1.0
    y = np.empty_like(x)
    oracle._box_gaussian_filter(x, np.array([0.0, 10.0]), y)
    k = np.arange(200) - 100
    np.testing.assert_allclose(1.0, y.sum(), rtol=1e-5)
    np.testing.assert_allclose(0.0, (k * y).sum(), atol=1e-5)
    np.testing.assert_allclose(np.sqrt((k * k * y).sum()), 10.0, atol=1)
    # axes handled consistently
    rs = np.random.RandomState(seed=1)
    data = rs.uniform(size=(77, 53)).astype(np.float32)
    o0, o1 = np.zeros_like(data), np.zeros_like(data.T).copy()
    oracle._box_gaussian_filter(data, np.array([8.0, 0.0]), o0)
    oracle._box_gaussian_filter(np.ascontiguousarray(data.T), np.array([0.0, 8.0]), o1)
    same(o0, np.ascontiguousarray(o1.T))


def test_interpolate_and_background_vectors():
    y = np.array([[np.nan, np.nan, 4.0, np.nan, np.nan, 10.0, np.nan, -2.0, np.nan, np.nan]], np.float32)
    oracle._linearly_interpolate_nans(y)
    np.testing.assert_allclose(y[0], [4.0, 4.0, 4.0, 6.0, 8.0, 10.0, 4.0, -2.0, -2.0, -2.0], rtol=1e-6)
    shape = (95, 86)
    data = np.ones(shape, np.float32) * 7.5
    bg = oracle._get_background2d(data, np.ones(shape, np.uint8), 1, (10.0, 10.0), 2.0, np.array([0, 86]))
    same(bg, np.zeros(shape, np.float32))
    bg = oracle._get_background2d(data, np.zeros(shape, np.uint8), 1, (10.0, 10.0), 2.0, np.array([0, 86]))
    np.testing.assert_allclose(data, bg, rtol=1e-5)
    # interpolation across a fully flagged block (tests/test_flagging.py:382-408)
    data[:, 70:] = 3.0
    flags = np.zeros(shape, np.uint8)
    flags[:, 30:70] = True
    rs = np.random.RandomState(seed=1)
    data[:50, :] += rs.uniform(-0.001, 0.001, data[0:50].shape)
    bg = oracle._get_background2d(data, flags, 1, (2.5, 2.5), 5.0, np.array([0, 86]))
    expected = np.zeros_like(data)
    expected[:, :37] = 7.5
    expected[:, 63:] = 3.0
    expected[:, 37:63] = np.linspace(7.5, 3.0, 26)
    np.testing.assert_allclose(expected[56:], bg[56:], rtol=1e-4)
    np.testing.assert_allclose(expected[:56], bg[:56], rtol=1e-2)


def test_sum_threshold_vectors():
    rs = np.random.RandomState(seed=1)
    data = rs.standard_normal((100, 90)).astype(np.float32) * 3.0
    in_flags = np.zeros(data.shape, np.bool_)
    data[:48] += 1000.0
    in_flags[:48] = True
    data[70, 0], data[70, 1], data[70, 2], data[70, 3] = 12.5, -12.5, 20.0, -20.0
    out = oracle._sum_threshold(data, in_flags, 0, np.array([1, 2, 4, 8]), 5, 1.3)
    np.testing.assert_array_equal([False, False, True, True], out[70, :4])
    small = np.arange(30, dtype=np.float32).reshape(5, 6)
    out = oracle._sum_threshold(small, np.ones(small.shape, bool), 0, np.array([1, 2, 4]), 4.5, 1.3)
    assert not out.any()
    # synthetic RFI recovery (tests/test_flagging.py:444-475)
    for axis in (0, 1):
        rs = np.random.RandomState(seed=1)
        d = rs.standard_normal((100, 90)).astype(np.float32) * 3.0
        rfi = np.zeros_like(d)
        rfi[10, 20] = 100.0
        rfi[80, 80] = -100.0
        rfi[:, 40] = rs.uniform(80.0, 120.0, size=(100,))
        rfi[:, 2] = -rfi[:, 40]
        rfi[:, 60:67] = rs.uniform(15.0, 20.0, size=(100, 7))
        rfi[:, 10:17] = -rfi[:, 60:67]
        expected = rfi != 0
        d += rfi
        fl = np.zeros(d.shape, bool)
        if axis == 0:
            d, fl = d.T.copy(), fl.T.copy()
        o = oracle._sum_threshold(d, fl, axis, np.array([1, 2, 4, 8]), 4.5, 1.3)
        if axis == 0:
            o = o.T
        assert (expected != o).sum() / d.size < 0.01
        for region in (np.s_[8:13, 18:23], np.s_[78:83, 78:83]):
            np.testing.assert_equal(expected[region], o[region])


def test_sum_threshold_flagger_class_behaviour():
    """tests/test_flagging.py:523-649 (shape reduced to keep the CPU suite short)"""
    import scipy.interpolate
    rs = np.random.RandomState(seed=1)
    shape = (1, 234, 345)
    x = np.linspace(0.0, shape[2], 10)
    y = np.ones((1, shape[1], 10)) * 2.34
    y[:, :, 0] = y[:, :, -1] = 0.1
    y += rs.uniform(0.0, 0.1, y.shape)
    background = scipy.interpolate.interp1d(x, y, axis=2, kind='cubic', assume_sorted=True)(np.arange(shape[2]))
    background = background.astype(np.float32)
    data = background + (rs.standard_normal(shape) * 0.1).astype(np.float32)
    rfi = np.zeros(shape, np.float32)
    rfi[:, 12, :] = 1
    rfi[:, 20:25, :] = 1
    rfi[:, :, 17] = 1
    rfi[:, :, 200:220] = 1
    rfi[:, 30, :300] = 1
    rfi[:, 50:, 80] = 1
    rfi[:, 60:65, 100:170] = 1
    rfi[:, 150:200, 150:153] = 1
    expected = rfi.astype(np.bool_)
    expected[:, 30, :] = True
    expected[:, :, 80] = True
    data += rfi * rs.standard_normal(shape) * 3.0
    data[:, :, 260] += 0.2
    expected[:, :, 260] = True
    data[:, 225, 225] = np.nan
    expected[:, 225, 225] = True
    in_flags = np.zeros(shape, np.bool_)
    in_flags[:, :, 185:190] = True
    data[:, :, 185:190] = np.nan
    data = np.abs(data)
    d0, f0 = data.copy(), in_flags.copy()
    out = oracle.SumThresholdFlagger().get_flags(data, in_flags)
    np.testing.assert_equal(d0, data)
    np.testing.assert_equal(f0, in_flags)
    allowed = expected | in_flags
    allowed[:, :-1, :] |= allowed[:, 1:, :]
    allowed[:, 1:, :] |= allowed[:, :-1, :]
    allowed[:, :, :-1] |= allowed[:, :, 1:]
    allowed[:, :, 1:] |= allowed[:, :, :-1]
    allowed[:, :, :40] = True
    allowed[:, :, -40:] = True
    assert 0 == (expected & ~out).sum()
    assert (out & ~allowed).sum() / data.size < 0.03
    z = oracle.SumThresholdFlagger().get_flags(np.zeros((4, 100, 80), np.float32), np.ones((4, 100, 80), bool))
    assert not z.any()
