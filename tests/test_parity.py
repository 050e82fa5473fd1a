# -*- coding: utf-8 -*-
"""
Parity of the CUDA path (through the Python boundary and the C ABI) with the
CPU oracle.  Every test runs twice: ``backend=emu`` executes the same kernel
sources under the CPU SIMT emulator at small sizes (no GPU needed) and
``backend=cuda`` (marked ``gpu``) is the parity test proper on the B200.
Integer / byte results must be bit-exact; float32 intermediates (backgrounds)
are bit-exact as well because the kernels keep the reference's operation order.
The test layout follows tricolour/tests/test_flagging.py.
"""
import numpy as np
import pytest

import oracle
import tricolour_b200 as tb
from tricolour_b200 import flagging as G
import common
from conftest import golden


def assert_same(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    if a.dtype.kind in "fc":
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    else:
        bad = a != b
    assert not bad.any(), "%s: %d of %d differ" % (what, int(bad.sum()), a.size)


def big(backend):
    return backend == "cuda"


# ------------------------------------------------------------------ F1-F3 ----
def test_flag_nans_and_zeros(backend):
    rs = np.random.RandomState(1)
    shape = (6, 4, 10, 64) if big(backend) else (3, 2, 5, 16)
    vis = (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    vis[2, 1, 4, 5] = 0
    vis[0, 1, 2, 7] = np.nan + np.nan * 1j
    vis[1, 0, 1, 1] = np.nan
    vis[1, 0, 1, 2] = 1j * np.nan
    vis[1, 1, 1, 3] = -0.0
    vis[1, 1, 1, 4] = np.inf
    for dt in (np.bool_, np.uint8):
        flags = (rs.uniform(size=shape) < 0.3).astype(dt)
        out = tb.flag_nans_and_zeros(vis, flags)
        assert_same(out, oracle.flag_nans_and_zeros(vis, flags), "nan/zero")
        assert_same(out.astype(bool), flags.astype(bool) | (vis == 0) | np.isnan(vis))
    # ragged size (scalar tail path) and complex128 input
    v7 = vis.reshape(-1)[:77].reshape(1, 1, 7, 11)
    f7 = np.zeros(v7.shape, np.uint8)
    assert_same(tb.flag_nans_and_zeros(v7, f7), oracle.flag_nans_and_zeros(v7, f7))
    v128 = v7.astype(np.complex128)
    v128[0, 0, 0, 0] = 1e-60
    assert tb.flag_nans_and_zeros(v128, f7)[0, 0, 0, 0] == 0
    with pytest.raises(ValueError):
        tb.flag_nans_and_zeros(vis, flags[:1])
    empty = tb.flag_nans_and_zeros(np.zeros((0, 4, 3, 8), np.complex64), np.zeros((0, 4, 3, 8), bool))
    assert empty.shape == (0, 4, 3, 8)


def test_flag_autos_and_static_mask(backend):
    rs = np.random.RandomState(2)
    nant = 7
    ubl = common.baselines(nant)
    ants = common.antenna_layout(nant)
    nbl = ubl.shape[0]
    for nchan in (16, 21):
        flags = rs.uniform(size=(nbl, 2, 5, nchan)) < 0.1
        assert_same(tb.flag_autos(flags, [ubl]), oracle.flag_autos(flags, [ubl]), "autos")
        f8 = flags.astype(np.uint8)
        assert_same(tb.flag_autos(f8, [ubl]), oracle.flag_autos(f8, [ubl]), "autos u8")
        cf = np.linspace(.856e9, 2 * .856e9, nchan)
        cw = np.full(nchan, cf[1] - cf[0])
        m1 = np.array([cf[2] + 128., cf[10]])[:, None]
        m2 = np.array([cf[4] - 64, cf[11] + 64, cf[5] - 128])[:, None]
        for mode in ("or", "override"):
            for uv in ("", "0~2500", "100~4000m", "*"):
                for masks in ([m1], [m1, m2], []):
                    got = tb.apply_static_mask(flags, ubl, ants, masks, cf, cw, accumulation_mode=mode, uvrange=uv)
                    want = oracle.apply_static_mask(flags, ubl, ants, masks, cf, cw, accumulation_mode=mode, uvrange=uv)
                    assert_same(got, want, "mask %s %s" % (mode, uv))
    with pytest.raises(ValueError):
        tb.apply_static_mask(flags, ubl, ants, [m1], cf, cw, accumulation_mode="xor")
    with pytest.raises(ValueError):
        tb.apply_static_mask(flags, ubl, ants, [m1], cf, cw, uvrange="abc")
    with pytest.raises(ValueError):
        tb.flag_autos(flags, [ubl[:3]])


def test_apply_static_mask_reference_case(backend):
    """tricolour/tests/test_flagging_additional.py:103-192 (WSRT-like layout)"""
    nant = 14
    rs = np.random.RandomState(0)
    ants = np.stack([3828763.1 - 16.5 * np.arange(nant), 442449.1 + 143.0 * np.arange(nant),
                     np.full(nant, 5064923.0)], axis=1)
    a1, a2 = np.triu_indices(nant, 0)
    ubl = np.stack([np.arange(a1.size), a1, a2], axis=1)
    ntime, nchan, ncorr = 10, 16, 4
    cf = np.linspace(.856e9, 2 * .856e9, nchan)
    cw = np.zeros_like(cf)
    cw[:-1] = np.diff(cf)
    cw[-1] = cw[0]
    m1 = np.asarray([cf[2] + 128., cf[10]])[:, None]
    m2 = np.asarray([cf[4] - 64, cf[11] + 64, cf[5] - 128])[:, None]
    flags = np.zeros((ubl.shape[0], ncorr, ntime, nchan), np.uint8)
    out = tb.apply_static_mask(flags, ubl, ants, [m1], cf, cw, accumulation_mode="or")
    sel = np.zeros(nchan, bool)
    sel[[2, 10]] = True
    assert np.all(out[:, :, :, sel] == 1) and np.all(out[:, :, :, ~sel] == 0)
    out = tb.apply_static_mask(flags, ubl, ants, [m1, m2], cf, cw, accumulation_mode="override")
    sel[:] = False
    sel[[4, 11, 5]] = True
    assert np.all(out[:, :, :, sel] == 1) and np.all(out[:, :, :, ~sel] == 0)
    out = tb.apply_static_mask(flags, ubl, ants, [m1, m2], cf, cw, uvrange="%f~%f" % (1e3, 2e4))
    d2 = 0.5 * ((ants[a1] - ants[a2]) ** 2).sum(axis=1)
    bl_sel = (d2 > 1e6) & (d2 < 4e8)
    sel[[2, 10]] = True
    assert np.all(out[np.ix_(bl_sel, range(ncorr), range(ntime), sel)] == 1)
    assert np.all(out[np.ix_(~bl_sel, range(ncorr), range(ntime), ~sel)] == 0)


# ------------------------------------------------------------- S1 _average_freq
def test_average_freq_known_answers(backend):
    """exact arrays of tricolour/tests/test_flagging.py:36-130"""
    data = np.arange(30, dtype=np.float32).reshape(1, 5, 6).repeat(2, axis=0)
    flags = np.zeros(data.shape, np.bool_)
    flags[0, 3, :] = 1
    flags[0, :, 4] = 1
    flags[:, 2, 0] = 1
    flags[:, 2, 5] = 1
    d1, f1 = G._average_freq(data, flags, 1)
    exp = data.copy()
    exp[flags] = 0
    assert d1.dtype == np.float32 and f1.dtype == np.bool_
    assert_same(d1, exp)
    assert_same(f1, flags)
    d2, f2 = G._average_freq(data, flags, 2)
    assert_same(d2[0], np.array([[0.5, 2.5, 5.0], [6.5, 8.5, 11.0], [13.0, 14.5, 0.0],
                                 [0.0, 0.0, 0.0], [24.5, 26.5, 29.0]], np.float32))
    assert_same(f2[0], np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1], [1, 1, 1], [0, 0, 0]], bool))
    d4, f4 = G._average_freq(data, flags, 4)
    assert_same(d4[0], np.array([[1.5, 5.0], [7.5, 11.0], [14.0, 0.0], [0.0, 0.0], [25.5, 29.0]], np.float32))
    assert_same(d4[1], np.array([[1.5, 4.5], [7.5, 10.5], [14.0, 16.0], [19.5, 22.5], [25.5, 28.5]], np.float32))
    with pytest.raises(ValueError):
        G._average_freq(data, flags[:1], 1)


def test_average_freq_complex(backend):
    rs = np.random.RandomState(3)
    shape = (4, 64, 1000) if big(backend) else (2, 7, 53)
    d = (rs.standard_normal(shape) * 10 ** rs.uniform(-3, 3, shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    d[0, 2, 3] = np.nan
    d[1, 1, 1] = 0
    fl = rs.uniform(size=shape) < 0.2
    for fac in (1, 2, 3, 4):
        a, b = G._average_freq(d, fl, fac)
        a2, b2 = oracle._average_freq(d, fl, oracle._as_min_dtype(fac))
        assert_same(a, a2, "avg data %d" % fac)
        assert_same(b, b2, "avg flags %d" % fac)


# ------------------------------------------------------------------ medians ---
def test_time_median(backend):
    data = np.array([[2.0, 1.0, 2.0, 5.0], [3.0, 1.0, 8.0, 6.0], [4.0, 1.0, 4.0, 7.0],
                     [5.0, 1.0, 5.0, 6.5], [1.5, 1.0, 1.5, 5.5]], np.float32)
    flags = np.array([[0, 1, 0, 1], [0, 1, 1, 0], [0, 1, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1]], np.bool_)
    od, of = G._time_median(data, flags)
    assert_same(od, np.array([[3.0, 0.0, 3.0, 6.0]], np.float32))
    assert_same(of, np.array([[0, 1, 0, 0]], np.bool_))
    rs = np.random.RandomState(4)
    for T in ((33, 34, 130, 300, 600, 1100) if big(backend) else (33, 34, 130, 300)):
        F = 64 if big(backend) else 20
        d = rs.standard_normal((T, F)).astype(np.float32)
        d[:, 3] = 1.0  # ties
        fl = rs.uniform(size=d.shape) < 0.3
        fl[:, 5] = True
        a, b = G._time_median(d, fl)
        a2, b2 = oracle._time_median(d, fl)
        assert_same(a, a2, "time median T=%d" % T)
        assert_same(b, b2)


def test_line_median_adversarial(backend):
    """the interpolation-search line median on inputs that stress its bracket logic:
    massive ties at the median (more equal keys than the 32-key finishing stage holds),
    two-valued and constant lines, 60 decades of dynamic range, signed data, even / odd
    counts, nearly fully flagged lines, every register tiling (T = 1 ... 1024)"""
    rs = np.random.RandomState(44)
    Ts = ((1, 2, 3, 31, 32, 33, 64, 100, 129, 257, 512, 600, 1000, 1024, 1025, 3277) if big(backend)
          else (1, 2, 3, 33, 64, 100, 257, 600, 1100))
    for T in Ts:
        F = 16
        d = rs.standard_normal((T, F)).astype(np.float32)
        d[:, 1] = np.round(d[:, 1] * 2) / 2                     # heavy ties around the median
        d[:, 2] = 1.0                                           # constant
        d[:, 3] = rs.randint(0, 2, T)                           # two values
        d[:, 4] = (d[:, 4] * 10.0 ** rs.uniform(-30, 30, T)).astype(np.float32)
        d[:, 5] = np.abs(d[:, 5]) + 2.3                         # one binade
        d[:, 6] = np.where(rs.uniform(size=T) < 0.5, 0.75, d[:, 6])   # half the line equal
        d[:, 7] = np.sort(d[:, 7])
        d[:, 8] = -np.abs(d[:, 8])                              # negative only
        d[:, 9] = np.float32(1e-42) * rs.randint(0, 5, T)       # float32 denormals and zeros
        for pf in (0.0, 0.3, 0.9):
            fl = rs.uniform(size=d.shape) < pf
            fl[:, 10] = True
            fl[: T // 2, 11] = True
            if T > 1:
                fl[0, 12] = not fl[1:, 12].sum() % 2               # force an odd / even count
            a, b = G._time_median(d, fl)
            a2, b2 = oracle._time_median(d, fl)
            assert_same(a, a2, "time median T=%d pf=%.1f" % (T, pf))
            assert_same(b, b2)
    # chunked |x| medians behind the SumThreshold thresholds (LM_ST_THRESHOLD), both axes
    T, F = (96, 410) if big(backend) else (24, 70)
    d = (rs.standard_normal((2, T, F)) * 10 ** rs.uniform(-3, 1, (2, T, F))).astype(np.float32)
    d[0, :, 5] = 0.5
    fl = rs.uniform(size=d.shape) < 0.4
    ce = np.linspace(0, F, 4).astype(int)
    for axis, ch in ((0, None), (1, ce)):
        got = G._sum_threshold(d, fl, axis, np.array([1, 2, 4, 8]), 10, 1.3, ch)
        for p in range(2):
            want = oracle._sum_threshold(d[p], fl[p], axis, np.array([1, 2, 4, 8]), 10, 1.3, ch)
            assert_same(got[p], want, "thresholds axis %d" % axis)
    # frequency chunks longer than 1024 channels (the 32768-channel mode): re-reading form
    T, F = (16, 6600) if big(backend) else (3, 2300)
    d = (rs.standard_normal((1, T, F)) * 10 ** rs.uniform(-3, 1, (1, T, F))).astype(np.float32)
    fl = rs.uniform(size=d.shape) < 0.4
    ce = np.array([0, 1100, F])
    got = G._sum_threshold(d, fl, 1, np.array([1, 2, 4, 8]), 10, 1.3, ce)
    want = oracle._sum_threshold(d[0], fl[0], 1, np.array([1, 2, 4, 8]), 10, 1.3, ce)
    assert_same(got[0], want, "long chunk thresholds")


def test_median_abs(backend):
    data = np.array([[-2.0, -6.0, 4.5], [1.5, 3.3, 0.5]], np.float32)
    flags = np.array([[0, 0, 0], [0, 1, 0]], np.uint8)
    assert G._median_abs(data, flags) == 2.0
    assert np.isnan(G._median_abs(data, np.ones_like(flags)))
    rs = np.random.RandomState(5)
    T, F = (128, 400) if big(backend) else (20, 90)
    d = rs.standard_normal((3, T, F)).astype(np.float32)
    fl = rs.uniform(size=d.shape) < 0.3
    fl[1, :, :F // 3] = True
    ce = np.linspace(0, F, 4).astype(int)
    got = G._median_abs(d, fl, ce)
    for p in range(3):
        for k in range(3):
            want = oracle._median_abs(d[p][:, ce[k]:ce[k + 1]], fl[p][:, ce[k]:ce[k + 1]])
            assert (got[p, k] == want) or (np.isnan(got[p, k]) and np.isnan(want))


def test_median_abs_long_ranges(backend):
    """ranges longer than one slice take the multi-block radix select"""
    rs = np.random.RandomState(55)
    shape = (2, 512, 1100) if big(backend) else (1, 80, 500)
    d = (rs.standard_normal(shape) * 10 ** rs.uniform(-2, 1, shape)).astype(np.float32)
    d[0, 3, :7] = 0.25  # ties
    fl = rs.uniform(size=shape) < 0.3
    for odd in (0, 1):
        fl[0, 0, 0] = bool(odd)
        ce = [0, shape[2]] if not big(backend) else [0, 600, shape[2]]
        got = np.atleast_2d(G._median_abs(d, fl, ce))
        for p in range(shape[0]):
            for k in range(len(ce) - 1):
                want = oracle._median_abs(d[p][:, ce[k]:ce[k + 1]], fl[p][:, ce[k]:ce[k + 1]])
                assert got[p, k] == want, (p, k, got[p, k], want)
    # massive ties overflow the bracket buffer -> the sliced radix fallback must take over
    const = np.full((1, 40, 520), 0.75, np.float32)
    const[0, 0, :3] = [0.1, 0.2, 5.0]
    nofl = np.zeros(const.shape, bool)
    assert G._median_abs(const, nofl) == oracle._median_abs(const[0], nofl[0]) == 0.75
    nofl[0, 0, 0] = True   # odd count
    assert G._median_abs(const, nofl) == 0.75
    vis, flags = common.make_windows(1, 1, 64, 600, seed=56)
    got = tb.uvcontsub_flagger(vis, flags, major_cycles=2, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0)
    want = oracle.uvcontsub_flagger(vis.copy(), flags, major_cycles=2, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0)
    assert (got != want).mean() <= 1e-4


# --------------------------------------------------------- filters/background --
def test_linearly_interpolate_nans(backend):
    y = np.array([np.nan, np.nan, 4.0, np.nan, np.nan, 10.0, np.nan, -2.0, np.nan, np.nan], np.float32)
    expected = np.array([4.0, 4.0, 4.0, 6.0, 8.0, 10.0, 4.0, -2.0, -2.0, -2.0], np.float32)
    z = np.stack([y, expected, np.full(10, np.nan, np.float32)])
    z2 = z.copy()
    G._linearly_interpolate_nans(z)
    oracle._linearly_interpolate_nans(z2)
    assert_same(z, z2)
    np.testing.assert_allclose(z[0], expected, rtol=1e-6)
    assert np.all(z[2] == 0)


def test_masked_gaussian_filter(backend):
    rs = np.random.RandomState(6)
    cases = [((77, 53), (8, 2.3)), ((77, 53), (0, 3.)), ((77, 53), (5., 0)), ((20, 100), (30., 50.)),
             ((20, 40), (0., 0.)), ((3, 8), (12.5, 10.))]
    if big(backend):
        cases += [((64, 1200), (62.5, 50.)), ((40, 1500), (32.5, 320.)), ((256, 300), (6.5, 64.))]
    for shape, sig in cases:
        d = (rs.uniform(size=shape) * 10 ** rs.uniform(-1, 1, shape)).astype(np.float32)
        fl = rs.uniform(size=shape) < 0.4
        fl[shape[0] // 4:shape[0] // 2, 5:shape[1] // 2] = True
        o = np.zeros_like(d)
        G.masked_gaussian_filter(d, fl, np.array(sig), o)
        o2 = np.zeros_like(d)
        oracle.masked_gaussian_filter(d, fl, np.array(sig), o2)
        assert_same(o, o2, "masked filter %s %s" % (shape, sig))
    with pytest.raises(ValueError):
        G.masked_gaussian_filter(d, fl[:1], np.array(sig), o)


def _sigma_for_radius(r):
    """a sigma that the reference turns into box radius r (flagging.py:451)"""
    return float(np.sqrt(((2 * r + 1) ** 2 - 1) / 3.0)) + 1e-6 if r > 0 else 0.0


def test_masked_gaussian_filter_radii(backend):
    """every delay-line phase of the lean filter kernels (radius mod 4, ring
    wrap, lines not a multiple of 8, samples not a multiple of 8)"""
    rs = np.random.RandomState(66)
    radii = [(4, 5), (6, 7), (5, 4), (7, 6), (8, 9), (10, 11), (13, 3), (3, 12), (9, 0), (0, 14), (17, 2)]
    if big(backend):
        radii += [(r, r + 1) for r in range(15, 40, 3)] + [(54, 43), (28, 277), (5, 110), (127, 16), (128, 8), (130, 131)]
    for i, (r0, r1) in enumerate(radii):
        if big(backend):
            shape = [(68, 92), (128, 200), (30, 301), (252, 64), (1, 400)][i % 5]
        else:
            shape = [(16, 24), (12, 40), (24, 20), (13, 21), (1, 36)][i % 5]
        sig = np.array((_sigma_for_radius(r0), _sigma_for_radius(r1)))
        d = (rs.uniform(size=shape) * 10 ** rs.uniform(-2, 2, shape)).astype(np.float32)
        d[rs.uniform(size=shape) < 0.01] = 0
        d[rs.uniform(size=shape) < 0.01] *= -1
        fl = rs.uniform(size=shape) < 0.3
        fl[shape[0] // 4:shape[0] // 2, 2:shape[1] // 2] = True
        o = np.zeros_like(d)
        G.masked_gaussian_filter(d, fl, sig, o)
        o2 = np.zeros_like(d)
        oracle.masked_gaussian_filter(d, fl, sig, o2)
        assert_same(o, o2, "masked filter radii (%d, %d) shape %s" % (r0, r1, shape))
    # thread-per-line kernels: n % 16 == 0, every small radius (ring phases 2r mod 4)
    t4 = [(r0, 4 + r0 % 3) for r0 in (1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 14)] + [(3, 2), (6, 3), (5, 5), (1, 1), (1, 9), (18, 5), (23, 6)]
    if big(backend):
        t4 += [(r0, r0 + 2) for r0 in (16, 21, 22, 28, 31, 32, 36)]
    for i, (r0, r1) in enumerate(t4):
        shape = [(160, 96), (64, 272)][i % 2] if big(backend) else [(16, 40), (32, 24)][i % 2]
        sig = np.array((_sigma_for_radius(r0), _sigma_for_radius(r1)))
        d = (rs.uniform(size=shape) * 10 ** rs.uniform(-2, 2, shape)).astype(np.float32)
        d[rs.uniform(size=shape) < 0.01] = 0
        fl = rs.uniform(size=shape) < 0.3
        fl[shape[0] // 4:shape[0] // 2, 2:shape[1] // 2] = True
        o = np.zeros_like(d)
        G.masked_gaussian_filter(d, fl, sig, o)
        o2 = np.zeros_like(d)
        oracle.masked_gaussian_filter(d, fl, sig, o2)
        assert_same(o, o2, "masked filter (thread per line) radii (%d, %d) shape %s" % (r0, r1, shape))


def test_masked_gaussian_filter_thread_per_line(backend):
    """the skewed thread-per-line kernels (k_filter3.cuh): every ring phase (radius mod 2,
    one-vector rings), lines that do not fill the last warp, several planes, the largest
    radii each form takes (54 on the first axis, 34 on the second), tiles of the second
    axis that end inside a 16-sample block, 60 decades of dynamic range in one plane
    (the drain's safe-range test must send those groups to the plain divisions)"""
    rs = np.random.RandomState(67)
    if big(backend):
        cases = [((3, 64, 100), (2, 2)), ((2, 48, 132), (3, 3)), ((1, 512, 96), (54, 34)), ((2, 128, 260), (21, 25)),
                 ((1, 32, 4096), (10, 8)), ((2, 80, 72), (43, 17)), ((1, 16, 40), (5, 12)), ((1, 256, 68), (32, 9))]
    else:
        cases = [((2, 16, 40), (2, 2)), ((1, 32, 36), (3, 3)), ((1, 16, 20), (9, 7)), ((2, 48, 24), (14, 5))]
    for shape, (r0, r1) in cases:
        sig = np.array((_sigma_for_radius(r0), _sigma_for_radius(r1)))
        d = (rs.uniform(size=shape) * 10 ** rs.uniform(-2, 2, shape)).astype(np.float32)
        d[0, : shape[1] // 3] *= (10.0 ** rs.uniform(-30, 28, (shape[1] // 3, shape[2]))).astype(np.float32)
        fl = rs.uniform(size=shape) < 0.35
        fl[:, shape[1] // 2:, : shape[2] // 3] = True
        o = np.zeros_like(d)
        G.masked_gaussian_filter(d, fl, sig, o)
        for p in range(shape[0]):
            o2 = np.zeros_like(d[p])
            oracle.masked_gaussian_filter(d[p], fl[p], sig, o2)
            assert_same(o[p], o2, "thread-per-line filter %s r=(%d, %d) plane %d" % (shape, r0, r1, p))
        # residual form through the background loop
        ce = np.linspace(0, shape[2], 4).astype(int)
        bg = G._get_background2d(d, fl, 2, np.array((r0 * 0.6 + 1, r1 * 0.6 + 1)), 2.0, ce)
        for p in range(shape[0]):
            want = oracle._get_background2d(d[p], fl[p], 2, np.array((r0 * 0.6 + 1, r1 * 0.6 + 1)), 2.0, ce)
            assert_same(bg[p], want, "background %s plane %d" % (shape, p))


def test_masked_gaussian_filter_dynamic_range(backend):
    """samples spanning 60 decades (incl. float32 denormals and exact zeros): the
    exact-division shortcut and the integer-pipe widening must still agree bit for bit"""
    rs = np.random.RandomState(67)
    shapes = [(64, 128), (32, 256)] if big(backend) else [(16, 24), (32, 16)]
    for i, shape in enumerate(shapes):
        for r0, r1 in ((4, 9), (10, 5), (17, 12)):
            sig = np.array((_sigma_for_radius(r0), _sigma_for_radius(r1)))
            d = (rs.uniform(0.5, 1.0, size=shape) * 10.0 ** rs.uniform(-33, 30, shape)).astype(np.float32)
            d[rs.uniform(size=shape) < 0.05] = 0
            d[rs.uniform(size=shape) < 0.05] = np.float32(1e-41)     # denormal
            d[rs.uniform(size=shape) < 0.05] *= -1
            fl = rs.uniform(size=shape) < 0.2
            o = np.zeros_like(d)
            G.masked_gaussian_filter(d, fl, sig, o)
            o2 = np.zeros_like(d)
            oracle.masked_gaussian_filter(d, fl, sig, o2)
            assert_same(o, o2, "masked filter, wide dynamic range, radii (%d, %d) shape %s" % (r0, r1, shape))


def test_get_background2d(backend):
    rs = np.random.RandomState(7)
    cases = [((95, 86), 1, (10., 10.), [0, 86]), ((45, 86), 3, (2.5, 2.5), [0, 40, 86]),
             ((1, 200), 5, (0., 10.), list(np.linspace(0, 200, 11).astype(int))),
             ((30, 120), 2, (3., 0.), [0, 60, 120])]
    if big(backend):
        cases += [((64, 800), 5, (12.5, 10.), list(np.linspace(0, 800, 11).astype(int))),
                  ((128, 1024), 5, (6.5, 64.), list(np.linspace(0, 1024, 11).astype(int)))]
    else:
        cases += [((24, 160), 5, (12.5, 10.), list(np.linspace(0, 160, 11).astype(int)))]
    for shape, it, sw, ce in cases:
        d = (7.5 + rs.standard_normal(shape) * 0.1).astype(np.float32)
        d[shape[0] // 3:shape[0] // 2 + 1, 30:80] += 15
        fl = rs.uniform(size=shape) < 0.05
        fl[:, 10:14] = True
        b = G._get_background2d(d, fl, it, np.array(sw), 2.0, np.array(ce))
        b2 = oracle._get_background2d(d, fl, it, np.array(sw), 2.0, np.array(ce))
        assert_same(b, b2, "background %s it=%d" % (shape, it))
    # corner cases of tricolour/tests/test_flagging.py:335-421
    shape = (95, 86) if big(backend) else (40, 50)
    data = np.ones(shape, np.float32) * 7.5
    bgd = G._get_background2d(data, np.zeros(shape, np.uint8), 1, np.array((10., 10.)), 2.0, np.array([0, shape[1]]))
    np.testing.assert_allclose(data, bgd, rtol=1e-5)
    bgd = G._get_background2d(data, np.ones(shape, np.uint8), 1, np.array((10., 10.)), 2.0, np.array([0, shape[1]]))
    assert_same(bgd, np.zeros(shape, np.float32))
    data[::3] = 20.0
    fl = np.zeros(shape, np.uint8)
    fl[::3] = True
    bgd = G._get_background2d(data, fl, 1, np.array((10., 10.)), 2.0, np.array([0, shape[1]]))
    np.testing.assert_allclose(np.full(shape, 7.5, np.float32), bgd, rtol=1e-5)


# ---------------------------------------------------------------- SumThreshold
def test_sum_threshold(backend):
    rs = np.random.RandomState(8)
    shape = (100, 90) if big(backend) else (40, 60)
    for axis in (0, 1):
        d = rs.standard_normal(shape).astype(np.float32) * 3
        d[10, 20] = 100
        d[:, 40] += 90
        d[20:25, 30:37] += 17
        fl = rs.uniform(size=d.shape) < 0.1
        n = shape[axis]
        for ch in (None, np.array([0, n // 3, 2 * n // 3 + 1, n])):
            for wins in ([1, 2, 4, 8], [3, 5, 32]):
                o = G._sum_threshold(d, fl, axis, np.array(wins), 4.5, 1.3, ch)
                o2 = oracle._sum_threshold(d, fl, axis, np.array(wins), 4.5, 1.3, ch)
                assert_same(o, o2, "sum_threshold axis=%d" % axis)
    # all flagged -> nothing flagged (tests/test_flagging.py:436-442)
    o = G._sum_threshold(d, np.ones_like(fl), 0, np.array([1, 2, 4]), 4.5, 1.3)
    assert not o.any()
    # existing flags (tests/test_flagging.py:477-501)
    data = np.random.RandomState(seed=1).standard_normal((100, 90)).astype(np.float32) * 3.0
    in_flags = np.zeros(data.shape, np.bool_)
    data[:48] += 1000.0
    in_flags[:48] = True
    data[70, 0], data[70, 1], data[70, 2], data[70, 3] = 12.5, -12.5, 20.0, -20.0
    out = G._sum_threshold(data, in_flags, 0, np.array([1, 2, 4, 8]), 5, 1.3)
    assert_same(out, oracle._sum_threshold(data, in_flags, 0, np.array([1, 2, 4, 8]), 5, 1.3))
    np.testing.assert_array_equal([False, False, True, True], out[70, :4])
    with pytest.raises(ValueError):
        G._sum_threshold(data, in_flags, 2, np.array([1]), 5, 1.3)
    with pytest.raises(ValueError):
        G._sum_threshold(data, in_flags, 0, np.array([], np.int64), 5, 1.3)


def test_combine_and_unaverage(backend):
    rs = np.random.RandomState(9)
    T, Fa = 30, 40
    sf = rs.uniform(size=(1, Fa)) < 0.1
    tf = rs.uniform(size=(T, Fa)) < 0.05
    ff = rs.uniform(size=(T, Fa)) < 0.05
    tf[5, :36] = True
    ff[:25, 7] = True
    for te, fe, af, F0 in ((3, 3, 1, 40), (2, 4, 2, 80), (4, 3, 4, 157), (1, 1, 1, 40), (7, 5, 3, 118), (0, 0, 1, 40)):
        tmp = np.zeros((T, Fa), bool)
        oracle._combine_flags(sf, tf, ff, oracle._as_min_dtype(te), tmp)
        want = np.zeros((T, F0), bool)
        oracle._unaverage_freq(tmp, fe, af, 0.6, 0.8, want)
        got = G._combine_and_unaverage(sf, tf, ff, te, fe, af, 0.6, 0.8, F0)
        assert_same(got, want, "combine te=%d fe=%d af=%d" % (te, fe, af))
    # 16-flags-per-thread kernels (average_freq == 1, F % 16 == 0), incl. the row / column rules
    for T2, F2, te, fe, ft, ff2 in ((30, 48, 3, 3, 0.6, 0.8), (7, 96, 4, 5, 0.3, 0.2), (40, 32, 1, 1, 0.1, 0.5),
                                    (300, 64, 2, 16, 0.05, 0.3), (9, 16, 0, 0, 0.6, 0.8), (5, 80, 6, 9, 0.9, 0.1)):
        sf = rs.uniform(size=(1, F2)) < 0.1
        tf = rs.uniform(size=(T2, F2)) < 0.05
        ff = rs.uniform(size=(T2, F2)) < 0.05
        tf[T2 // 2, :F2 - 3] = True
        ff[:T2 - 2, 7] = True
        tmp = np.zeros((T2, F2), bool)
        oracle._combine_flags(sf, tf, ff, oracle._as_min_dtype(te), tmp)
        want = np.zeros((T2, F2), bool)
        oracle._unaverage_freq(tmp, fe, 1, ft, ff2, want)
        got = G._combine_and_unaverage(sf, tf, ff, te, fe, 1, ft, ff2, F2)
        assert_same(got, want, "combine v16 T=%d F=%d te=%d fe=%d" % (T2, F2, te, fe))


def _st_cases(backend):
    kw = common.DEFAULT_STRATEGY_KW
    if big(backend):
        shape = (2, 2, 64, 512)
        cases = [dict(), dict(kw["background_flags"], num_major_iterations=2), kw["final_st_very_broad"],
                 kw["final_st_broad"], kw["final_st_narrow"],
                 dict(average_freq=2, windows_freq=[2, 4, 8, 16], num_major_iterations=1),
                 dict(average_freq=3, windows_freq=[3, 6, 13], freq_chunks=4, time_extend=4, freq_extend=5,
                      num_major_iterations=2),
                 dict(freq_chunks=1, num_major_iterations=1)]
    else:
        shape = (1, 2, 20, 96)
        cases = [dict(num_major_iterations=2), dict(kw["final_st_narrow"]),
                 dict(average_freq=2, windows_freq=[2, 4, 8, 16], num_major_iterations=1, freq_chunks=3)]
    return shape, cases


def test_sum_threshold_flagger(backend):
    shape, cases = _st_cases(backend)
    vis, flags = common.make_windows(*shape, seed=21)
    vis0, flags0 = vis.copy(), flags.copy()
    for kw in cases:
        got = tb.sum_threshold_flagger(vis, flags, **kw)
        want = oracle.sum_threshold_flagger(vis, flags, **kw)
        assert_same(got, want, "sum_threshold_flagger %s" % (kw,))
    # inputs are borrowed (tests/test_flagging.py:562-567)
    assert_same(vis, vis0)
    assert_same(flags, flags0)
    f8 = flags.astype(np.uint8)
    assert_same(tb.sum_threshold_flagger(vis, f8, num_major_iterations=1),
                oracle.sum_threshold_flagger(vis, f8, num_major_iterations=1), "uint8 flags")
    with pytest.raises(ValueError):
        tb.sum_threshold_flagger(vis, flags, average_freq=2)   # window 0, as in the reference


def test_sum_threshold_flagger_aligned(backend):
    """shapes with T % 16 == 0 and F % 16 == 0 take the vectorised / thread-per-line kernels"""
    shape = (2, 2, 64, 256) if big(backend) else (1, 2, 32, 64)
    vis, flags = common.make_windows(*shape, seed=27)
    cases = [dict(num_major_iterations=1, background_iterations=2),
             dict(num_major_iterations=1, background_iterations=1, spike_width_time=2, spike_width_freq=4.0,
                  time_extend=4, freq_extend=5, flag_all_time_frac=0.3, flag_all_freq_frac=0.4)]
    if big(backend):
        cases += [dict(common.DEFAULT_STRATEGY_KW["background_flags"], num_major_iterations=2),
                  dict(common.DEFAULT_STRATEGY_KW["final_st_very_broad"])]
    for kw in cases:
        got = tb.sum_threshold_flagger(vis, flags, **kw)
        want = oracle.sum_threshold_flagger(vis, flags, **kw)
        assert_same(got, want, "sum_threshold_flagger aligned %s" % (kw,))


def test_sum_threshold_flagger_class(backend):
    shape = (3, 64, 345) if big(backend) else (1, 20, 90)
    rs = np.random.RandomState(11)
    vis, flags = common.make_windows(1, shape[0], shape[1], shape[2], seed=22)
    data = np.abs(vis[0])
    fl = flags[0]
    configs = [dict(), dict(average_freq=2), dict(freq_chunks=1), dict(average_freq=4, freq_chunks=15)]
    for kw in configs if big(backend) else configs[:2]:
        got = G.SumThresholdFlagger(**kw).get_flags(data, fl)
        want = oracle.SumThresholdFlagger(**kw).get_flags(data, fl)
        assert got.dtype == np.bool_
        assert_same(got, want, "class %s" % (kw,))
    # all flagged -> zeros (tests/test_flagging.py:619-630)
    z = G.SumThresholdFlagger(average_freq=4).get_flags(np.zeros((2, 30, 40), np.float32), np.ones((2, 30, 40), bool))
    assert not z.any()
    with pytest.raises(ValueError):
        G.SumThresholdFlagger().get_flags(data, fl[:, :3])
    with pytest.raises(ValueError):
        G.SumThresholdFlagger().get_flags(data[0], fl[0])


def test_uvcontsub_flagger(backend):
    shape = (2, 2, 40, 300) if big(backend) else (1, 2, 16, 64)
    vis, flags = common.make_windows(*shape, seed=23)
    flags[0, 1] = True  # fully flagged plane is skipped (flagging.py:1034-1035)
    for kw in (dict(), dict(major_cycles=7, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0),
               dict(major_cycles=3, or_original_from_cycle=0, taylor_degrees=25, sigma=13.0)):
        got = tb.uvcontsub_flagger(vis, flags, **kw)
        want = oracle.uvcontsub_flagger(vis.copy(), flags, **kw)
        assert got.shape == want.shape and got.dtype == want.dtype
        # numpy's float32 FFT is not reproduced bit for bit: equal up to threshold ties
        assert (got != want).mean() <= 1e-4, kw
        assert got[0, 1].all()
    with pytest.raises(ValueError):
        tb.uvcontsub_flagger(vis, flags[:, :1])


def test_sum_threshold_flagger_complex128_input(backend):
    """complex128 visibilities with average_freq == 1: the amplitude is taken in float64 and rounded once to
    float32, as the reference does (flagging.py:856-871) -- same flags as handing that float32 amplitude over;
    with frequency averaging the dtype is refused"""
    rs = np.random.RandomState(71)
    shape = (2, 2, 32, 64)
    v = (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)) * 3 + 10
    v[0, 0, 3, 5] = np.nan
    v[1, 1, 7, :] *= 8
    fl = rs.uniform(size=shape) < 0.05
    kw = dict(outlier_nsigma=6, background_iterations=2, num_major_iterations=2)
    got = G.sum_threshold_flagger(v.astype(np.complex128), fl, **kw)
    amp = np.abs(v).astype(np.float32)
    want = oracle.sum_threshold_flagger(amp, fl, **kw)
    assert_same(got, want, "complex128 input")
    with pytest.raises(TypeError):
        G.sum_threshold_flagger(v.astype(np.complex128), fl, average_freq=2, **kw)



# ---------------------------------------------------------------- stokes ------
@pytest.mark.parametrize('stokes', [['YX', 'XX', 'XY', 'YY'], ['XX', 'XY', 'YX', 'YY'],
                                    ['RR', 'RL', 'LR', 'LL'], ['RL', 'RR', 'LL', 'LR']])
def test_stokes(backend, stokes):
    ct = [tb.STOKES_TYPES[s] for s in stokes]
    smap = tb.stokes_corr_map(ct)
    assert smap == oracle.stokes_corr_map(ct)
    pol = tuple(v for k, v in smap.items() if k != 'I')
    unpol = tuple(v for k, v in smap.items() if k == 'I')
    # closed form of tricolour/tests/test_stokes.py
    vis = np.asarray([[[1 + 1j, 2 + 2j, 3 + 3j, 4 + 4j]]], np.complex128)
    p = sum(np.abs(a * (s1 * vis[0, 0, c1] + s2 * vis[0, 0, c2])) ** 2 for c1, c2, a, s1, s2 in pol)
    u = sum(np.abs(a * (s1 * vis[0, 0, c1] + s2 * vis[0, 0, c2])) for c1, c2, a, s1, s2 in unpol)
    assert np.allclose(tb.polarised_intensity(vis, pol), np.sqrt(p))
    assert np.allclose(tb.unpolarised_intensity(vis, unpol, pol), u - np.sqrt(p))
    rs = np.random.RandomState(12)
    shape = (500, 64, 4) if big(backend) else (20, 7, 4)
    v = (rs.standard_normal(shape) * 10 ** rs.uniform(-2, 2, shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    got, want = tb.polarised_intensity(v, pol), oracle.polarised_intensity(v, pol)
    # bit-exact: the kernel restates glibc's hypot operation for operation
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (got.imag == 0).all() and got.shape == (shape[0], shape[1], 1)
    got, want = tb.unpolarised_intensity(v, unpol, pol), oracle.unpolarised_intensity(v, unpol, pol)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # operands far apart in magnitude, huge and tiny ones (the scaled branches of hypot)
    w = v.copy()
    w[::3] *= np.float32(1e-30)
    w[1::3, :, ::2] *= np.float32(1e25)
    w[2::3, :, 1] = 0
    got, want = tb.polarised_intensity(w, pol), oracle.polarised_intensity(w, pol)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # results that sit exactly on (or within a few float64 ulps of) a float32 rounding midpoint: the
    # interval shortcut of the kernel must notice and replay the exact sequence.  With XY = YX = 0 the
    # polarised intensity is |Q| = |XX - YY| / 2 (or whatever single term the set keeps): operands are
    # chosen so that it is 2^24 + 1 + k ulp(float64), k = -3..3.
    if len(stokes) == 4:
        m = np.zeros((16, 1, 4), np.complex64)
        m[:, 0, 0] = np.float32(2.0 ** 25)
        m[:, 0, 3] = -np.float32(2.0) * (1 + np.arange(16, dtype=np.float32) * np.float32(2.0 ** -23))
        m[8:, 0, 0] *= np.float32(2.0 ** -40)
        m[8:, 0, 3] *= np.float32(2.0 ** -40)
        m[3, 0, :] = 0                       # all-zero sample
        m[5, 0, 1] = np.nan                  # non-finite term
        for fn_t, fn_o, args in ((tb.polarised_intensity, oracle.polarised_intensity, (pol,)),
                                 (tb.unpolarised_intensity, oracle.unpolarised_intensity, (unpol, pol))):
            got, want = fn_t(m, *args), fn_o(m, *args)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(ValueError):
        tb.unpolarised_intensity(v, (), pol)
    with pytest.raises(ValueError):
        tb.unpolarised_intensity(v, unpol, ())


# ------------------------------------------------------- pack / unpack / stats
@pytest.mark.parametrize("nchan", [16, 13])
def test_pack_unpack_roundtrip(backend, nchan):
    """tricolour/tests/test_packing.py:28-109 without the dask glue"""
    rs = np.random.RandomState(13)
    na, ntime, ncorr = 7, 10, 4
    a1, a2 = (a.astype(np.int32) for a in np.triu_indices(na, 1))
    nbl = a1.size
    A1, A2 = np.tile(a1, ntime), np.tile(a2, ntime)
    tinv = np.repeat(np.arange(ntime), nbl)
    nrow = A1.size
    vis = (rs.standard_normal((nrow, nchan, ncorr)) + 1j * rs.standard_normal((nrow, nchan, ncorr))).astype(np.complex64)
    flag = rs.randint(0, 2, (nrow, nchan, ncorr)).astype(bool)
    dele = rs.randint(nrow, size=15)
    A1, A2, tinv = np.delete(A1, dele), np.delete(A2, dele), np.delete(tinv, dele)
    vis, flag = np.delete(vis, dele, 0), np.delete(flag, dele, 0)
    ubl = tb.unique_baselines(A1, A2).view(np.int32).reshape(-1, 2)
    ubl = np.concatenate([np.arange(ubl.shape[0], dtype=np.int32)[:, None], ubl], axis=1)
    vw, fw = tb.pack_data(tinv, ubl, A1, A2, vis, flag, ntime)
    vw2, fw2 = oracle.pack_data(tinv, ubl, A1, A2, vis, flag, ntime)
    assert_same(vw, vw2, "pack vis")
    assert_same(fw, fw2, "pack flags")
    assert_same(tb.unpack_data(A1, A2, tinv, ubl, fw), flag, "round trip flags")
    assert_same(tb.unpack_data(A1, A2, tinv, ubl, vw), vis, "round trip vis")
    eq = tb.packing.unpack_flags_equalised(A1, A2, tinv, ubl, fw)
    assert_same(eq, np.broadcast_to(flag.any(axis=2, keepdims=True), flag.shape))
    # duplicate (baseline, time) rows: the last one wins, like the reference's loops
    A1d, A2d, td = np.append(A1, A1[:3]), np.append(A2, A2[:3]), np.append(tinv, tinv[:3])
    vd = np.concatenate([vis, vis[:3] + 1])
    fd = np.concatenate([flag, ~flag[:3]])
    vw3, fw3 = tb.pack_data(td, ubl, A1d, A2d, vd, fd, ntime)
    vw4, fw4 = oracle.pack_data(td, ubl, A1d, A2d, vd, fd, ntime)
    assert_same(vw3, vw4, "duplicates vis")
    assert_same(fw3, fw4, "duplicates flags")
    # a baseline chunk whose indices do not start at zero (packing.py:399-404)
    sub = ubl[5:12]
    got = tb.unpack_data(A1, A2, tinv, sub, fw[5:12])
    want = oracle.unpack_data(A1, A2, tinv, sub, fw[5:12])
    assert_same(got, want, "chunked unpack")
    with pytest.raises(TypeError):
        tb.unique_baselines(A1.astype(np.int64), A2)


@pytest.mark.parametrize("nchan,ncorr", [(32, 4), (48, 2), (21, 4)])
def test_polarised_pack_and_broadcast_unpack(backend, nchan, ncorr):
    """N2: Stokes + any(corr) fused into the pack (app.py:415-432 + packing.py:243-278)
    and the one-correlation window broadcast back to row order (app.py:479-480)"""
    rs = np.random.RandomState(17)
    na, ntime = 5, 9
    a1, a2 = (a.astype(np.int32) for a in np.triu_indices(na, 0))
    nbl = a1.size
    A1, A2 = np.tile(a1, ntime), np.tile(a2, ntime)
    tinv = np.repeat(np.arange(ntime), nbl)
    keep = np.ones(A1.size, bool)
    keep[rs.choice(A1.size, 11, replace=False)] = False
    A1, A2, tinv = A1[keep], A2[keep], tinv[keep]
    nrow = A1.size
    vis = (rs.standard_normal((nrow, nchan, ncorr)) + 1j * rs.standard_normal((nrow, nchan, ncorr))).astype(np.complex64)
    flag = rs.uniform(size=(nrow, nchan, ncorr)) < 0.1
    ubl = tb.unique_baselines(A1, A2).view(np.int32).reshape(-1, 2)
    ubl = np.concatenate([np.arange(ubl.shape[0], dtype=np.int32)[:, None], ubl], axis=1)
    if ncorr == 4:
        smap = tb.stokes_corr_map([9, 10, 11, 12])
    else:
        smap = tb.stokes_corr_map([9, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    unpol = tuple(v for k, v in smap.items() if k == 'I')
    vw, fw = tb.packing.pack_polarised(tinv, ubl, A1, A2, vis, flag, ntime, pol)
    pi = oracle.polarised_intensity(vis, pol)
    vw2, fw2 = oracle.pack_data(tinv, ubl, A1, A2, pi, flag.any(axis=2, keepdims=True), ntime)
    assert vw.shape == (nbl, 1, ntime, nchan)
    assert_same(vw, vw2, "polarised pack vis")
    assert_same(fw, fw2, "polarised pack flags")
    vw, fw = tb.packing.pack_polarised(tinv, ubl, A1, A2, vis, flag, ntime, pol, unpol)
    ui = oracle.unpolarised_intensity(vis, unpol, pol)
    vw2, _ = oracle.pack_data(tinv, ubl, A1, A2, ui, flag.any(axis=2, keepdims=True), ntime)
    assert_same(vw, vw2, "unpolarised pack vis")
    # flag the window some more, then back to rows of ncorr correlations
    fw = fw | (rs.uniform(size=fw.shape) < 0.2)
    got = tb.packing.unpack_flags_equalised(A1, A2, tinv, ubl, fw, ncorr_out=ncorr)
    want = np.broadcast_to(oracle.unpack_data(A1, A2, tinv, ubl, fw), (nrow, nchan, ncorr))
    assert got.shape == (nrow, nchan, ncorr)
    assert_same(got, want, "broadcast unpack")
    # uint8 windows holding values other than 0/1 still come back as booleans
    got8 = tb.packing.unpack_flags_equalised(A1, A2, tinv, ubl, fw.astype(np.uint8) * 7, ncorr_out=ncorr)
    assert_same(got8, want)
    with pytest.raises(ValueError):
        tb.packing.pack_polarised(tinv, ubl, A1, A2, vis, flag, ntime, ())


@pytest.mark.parametrize("shape", [(5, 4, 70, 64), (3, 2, 300, 48), (2, 1, 7, 4112), (4, 4, 9, 20)])
def test_window_counts_shapes(backend, shape):
    """the 16-byte counting kernel (F % 16 == 0; more rows than one 256-row segment;
    more than one 4096-channel tile) and the general one, bool and valued uint8 windows"""
    rs = np.random.RandomState(sum(shape))
    nbl, ncorr, T, F = shape
    ubl = common.baselines(8)[:nbl]
    cf, _ = common.channels(F)
    names = ["m%03d" % i for i in range(8)]
    for fw in (rs.uniform(size=shape) < 0.3, (rs.uniform(size=shape) < 0.3).astype(np.uint8) * rs.randint(1, 256, shape).astype(np.uint8)):
        blc, chc = tb.window_statistics._counts(fw)
        assert blc.tolist() == fw.reshape(nbl, -1).sum(axis=1, dtype=np.uint64).tolist()
        assert chc.tolist() == fw.sum(axis=(0, 1, 2), dtype=np.uint64).tolist()
        st = tb.window_stats(fw, ubl, cf, names, 0, "f", 0)
        want = oracle.window_counts(fw, ubl, cf, 8)
        assert [int(st._counts_per_ant[n]) for n in names] == [int(x) for x in want[0]]
        assert np.array_equal(st._counts_per_ddid[0], want[6].astype(np.uint64))


def test_window_stats(backend):
    g = golden("packing.npz")
    ubl, fw, cf = g["ubl"], g["flag_win"], g["chan_freqs"]
    names = ["A%d" % i for i in range(6)]
    st = tb.window_statistics._window_stats([[[fw]]], [ubl], [cf], names, 3, "M87", 0, 10)
    assert [int(st._counts_per_ant[n]) for n in names] == g["counts_per_ant"].tolist()
    assert [int(st._size_per_ant[n]) for n in names] == g["size_per_ant"].tolist()
    assert [int(st._counts_per_bl["%s&%s" % (names[b[1]], names[b[2]])]) for b in ubl] == g["counts_per_bl"].tolist()
    assert int(st._counts_per_field["M87"]) == int(g["counts_field"])
    assert int(st._size_per_scan[3]) == int(g["size_scan"])
    assert np.array_equal(st._counts_per_ddid[0], g["bins"])
    assert np.array_equal(st._bins_per_ddid[0], g["bin_edges"])
    assert st._counts_per_ddid[0][-1] == 0   # nchanbins edges -> last bin always empty
    s2 = tb.window_stats(fw, ubl, cf, names, 4, "M87", 0, prev_stats=st)
    assert int(s2._counts_per_field["M87"]) == 2 * int(g["counts_field"])
    assert set(s2._counts_per_scan.keys()) == {3, 4}
    comb = tb.combine_window_stats([st, s2])
    assert isinstance(comb, tb.WindowStatistics)
    assert "BEGINNING OF FLAG SUMMARY" in "\n".join(tb.summarise_stats(s2, comb))


# ------------------------------------------------------------ golden fixtures -
def test_golden_sum_threshold_flagger(backend):
    """outputs of the unmodified reference (tests/golden/make_golden.py)"""
    g = golden("sum_threshold_flagger.npz")
    vis, flags = g["vis"], g["flags"]
    names = ["background_flags", "final_st_very_broad", "final_st_narrow", "defaults", "avg2"]
    if not big(backend):
        vis, flags = vis[:1, :1], flags[:1, :1]
        names = ["final_st_narrow"]
    for name in names:
        if name == "defaults":
            kw = {}
        elif name == "avg2":
            kw = dict(average_freq=2, windows_freq=[2, 4, 8, 16], num_major_iterations=1)
        else:
            kw = dict(common.DEFAULT_STRATEGY_KW[name])
            if name == "background_flags":
                kw["num_major_iterations"] = 2
        got = tb.sum_threshold_flagger(vis, flags, **kw)
        want = g[name][:vis.shape[0], :vis.shape[1]]
        assert_same(got, want, "golden %s" % name)


def test_golden_stages(backend):
    g = golden("stages.npz")
    data, fl, ce = g["data"], g["flags"], g["chunk_ends"]
    if big(backend):
        assert_same(G._get_background2d(data, fl, 5, np.array((12.5, 10.0)), 2.0, ce), g["background"], "background")
    mf = np.zeros_like(data)
    G.masked_gaussian_filter(data, fl, np.array((12.5, 10.0)), mf)
    assert_same(mf, g["masked_filter"], "masked filter")
    tm, tmf = G._time_median(data, fl)
    assert_same(tm, g["time_median"])
    assert_same(tmf, g["time_median_flags"])
    res = data - g["background"]
    assert_same(G._sum_threshold(res, fl, 0, np.array([1, 2, 4, 8]), 10, 1.3), g["st_time"])
    assert_same(G._sum_threshold(res, fl, 1, np.array([1, 2, 4, 8]), 10, 1.3, ce), g["st_freq"])


def test_golden_companions(backend):
    g = golden("companions.npz")
    st = golden("sum_threshold_flagger.npz")
    assert_same(tb.flag_nans_and_zeros(st["vis"], st["flags"]), g["nanzero"])
    v = g["rowvis"]
    for tag, ct in (("lin", [9, 10, 11, 12]), ("circ", [5, 6, 7, 8]), ("mixed", [11, 9, 10, 12])):
        m = tb.stokes_corr_map(ct)
        pol = tuple(x for k, x in m.items() if k != 'I')
        unpol = tuple(x for k, x in m.items() if k == 'I')
        assert_same(tb.polarised_intensity(v, pol), g["pol_" + tag])
        assert_same(tb.unpolarised_intensity(v, unpol, pol), g["unpol_" + tag])
    p = golden("packing.npz")
    vw, fw = tb.pack_data(p["time_inv"], p["ubl"], p["ant1"], p["ant2"], p["vis"], p["flags"], int(p["ntime"]))
    assert_same(vw, p["vis_win"])
    assert_same(fw, p["flag_win"])
    assert_same(tb.unpack_data(p["ant1"], p["ant2"], p["time_inv"], p["ubl"], fw), p["unpacked"])


def test_golden_uvcontsub(backend):
    g = golden("uvcontsub.npz")
    vis, flags = g["vis"], g["flags"]
    if not big(backend):
        vis, flags = vis[:1, :1], flags[:1, :1]
    got = tb.uvcontsub_flagger(vis, flags, major_cycles=7, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0)
    assert (got != g["cycles7"][:vis.shape[0], :vis.shape[1]]).mean() <= 1e-4
    got = tb.uvcontsub_flagger(vis, flags, major_cycles=3, or_original_from_cycle=0, taylor_degrees=25, sigma=13.0)
    assert (got != g["cycles3_or0"][:vis.shape[0], :vis.shape[1]]).mean() <= 1e-4


def test_strategy_chain(backend):
    """strat_executor.py combine rules with the per-function numpy API"""
    nant = 2
    ubl = common.baselines(nant)
    ants = common.antenna_layout(nant)
    T, F = (48, 256) if big(backend) else (16, 64)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    vis, flags = common.make_windows(ubl.shape[0], 2, T, F, seed=31, ubl=ubl)
    strategies = common.default_strategies()
    if not big(backend):
        for s in strategies:           # keep the emulated run short
            if s["task"] == "sum_threshold":
                s["kwargs"].update(num_major_iterations=1, background_iterations=1)
            if s["task"] == "uvcontsub_flagger":
                s["kwargs"].update(major_cycles=2)
    got = common.run_strategies(tb, strategies, vis, flags, ubl, ants, masks, cf, cw)
    want = common.run_strategies(oracle, strategies, vis, flags, ubl, ants, masks, cf, cw)
    assert (got != want).mean() <= 1e-6


_MISSED_BRACKET_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import conftest, oracle
from tricolour_b200 import _cabi
if %(emu)r:
    _cabi._set_library_for_testing(_cabi.load(conftest.EMU_LIB))
from tricolour_b200 import flagging as G
rs = np.random.RandomState(91)
bad = 0
for shape, ce in [((2, 80, 500), [0, 500]), ((1, 96, 520), [0, 130, 520]), ((1, 41, 257), [0, 257])]:
    d = (rs.standard_normal(shape) * 10 ** rs.uniform(-2, 1, shape)).astype(np.float32)
    d[0, 3, :9] = 0.25
    fl = rs.uniform(size=shape) < 0.3
    for odd in (0, 1):
        fl[0, 0, 0] = bool(odd)
        got = np.atleast_2d(G._median_abs(d, fl, ce))
        for p in range(shape[0]):
            for k in range(len(ce) - 1):
                want = oracle._median_abs(d[p][:, ce[k]:ce[k + 1]], fl[p][:, ce[k]:ce[k + 1]])
                bad += int(got[p, k] != want)
vis = (rs.standard_normal((2, 2, 48, 256)) + 1j * rs.standard_normal((2, 2, 48, 256))).astype(np.complex64)
vis[:, :, :, 40] *= 8
fl = rs.uniform(size=vis.shape) < 0.05
kw = dict(outlier_nsigma=10, background_iterations=3, num_major_iterations=1)
bad += int((G.sum_threshold_flagger(vis, fl, **kw) != oracle.sum_threshold_flagger(vis, fl, **kw)).sum())
# background loop on ranges longer than the 4096-key sample: the thresholds of such ranges are applied by the
# block that settles the range (or by k_sel_update when TC_BRK_UPDATE_IN_TAIL=0)
for shape, it, sw, ce in [((95, 86), 2, (10., 10.), [0, 86]), ((70, 160), 3, (2.5, 4.), [0, 70, 160])]:
    d = (7.5 + rs.standard_normal(shape) * 0.1).astype(np.float32)
    d[shape[0] // 3:shape[0] // 2 + 1, 30:80] += 15
    fl = rs.uniform(size=shape) < 0.05
    b = G._get_background2d(d, fl, it, np.array(sw), 2.0, np.array(ce))
    b2 = oracle._get_background2d(d, fl, it, np.array(sw), 2.0, np.array(ce))
    bad += int((b.view(np.uint32) != b2.view(np.uint32)).sum())
print("MISSED_BRACKET_DIFF", bad)
'''


@pytest.mark.parametrize("brk_k,update_in_tail", [("0.0", "1"), ("0.0", "0"), ("1.75", "0")])
def test_select_missed_bracket_is_redone_by_the_sweep_tail(backend, brk_k, update_in_tail):
    """TC_BRK_K=0 narrows the sample bracket to a few ranks, so that nearly every range
    misses it: the last collecting block of the range must then find the exact median
    itself (block_range_median in the sweep's tail) and, by default, apply the range's
    threshold too; TC_BRK_UPDATE_IN_TAIL=0 keeps the separate k_sel_update launch.  The
    knobs are read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TC_BRK_K=brk_k, TC_BRK_UPDATE_IN_TAIL=update_in_tail)
    script = _MISSED_BRACKET_SCRIPT % {"root": root, "tests": os.path.join(root, "tests"), "emu": backend == "emu"}
    r = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MISSED_BRACKET_DIFF 0" in r.stdout, r.stdout[-500:]
