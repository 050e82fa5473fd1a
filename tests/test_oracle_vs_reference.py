# -*- coding: utf-8 -*-
"""Oracle against the live reference (imported from /root/reference through
oracle/ref_loader.py).  Only runs where the reference checkout exists, i.e. in
the build container; skipped on the GPU box."""
import numpy as np
import pytest

import oracle
from oracle import ref_loader
import common

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not available")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(a, b, equal_nan=a.dtype.kind in "fc")


def test_stages(ref):
    F = ref[0]
    rs = np.random.RandomState(31)
    d = (rs.standard_normal((2, 9, 37)) * 10 ** rs.uniform(-2, 2, (2, 9, 37)) + 1j * rs.standard_normal((2, 9, 37))).astype(np.complex64)
    fl = rs.uniform(size=d.shape) < 0.2
    for fac in (1, 2, 5):
        a, b = F._average_freq(d, fl, F._as_min_dtype(fac))
        a2, b2 = oracle._average_freq(d, fl, oracle._as_min_dtype(fac))
        same(a, a2)
        same(b, b2)
    x = rs.standard_normal((45, 120)).astype(np.float32)
    fx = rs.uniform(size=x.shape) < 0.25
    same(F._time_median(x, fx)[0], oracle._time_median(x, fx)[0])
    for sig in ((8, 2.3), (0, 30.), (12.5, 10.)):
        o, o2 = np.zeros_like(x), np.zeros_like(x)
        F.masked_gaussian_filter(x, fx, np.array(sig), o)
        oracle.masked_gaussian_filter(x, fx, np.array(sig), o2)
        same(o, o2)
    ce = np.linspace(0, 120, 5).astype(int)
    same(F._get_background2d(x, fx, 3, np.array((5., 7.)), 2.0, ce), oracle._get_background2d(x, fx, 3, np.array((5., 7.)), 2.0, ce))
    for axis, ch in ((0, None), (1, ce)):
        same(F._sum_threshold(x * 3, fx, axis, np.array([1, 2, 4, 8]), 4.5, 1.3, ch),
             oracle._sum_threshold(x * 3, fx, axis, np.array([1, 2, 4, 8]), 4.5, 1.3, ch))


def test_flaggers(ref):
    F = ref[0]
    vis, flags = common.make_windows(1, 2, 32, 200, seed=41)
    for kw in (dict(num_major_iterations=2), dict(common.DEFAULT_STRATEGY_KW["final_st_very_broad"])):
        same(F.sum_threshold_flagger(vis, flags, **kw), oracle.sum_threshold_flagger(vis, flags, **kw))
    same(F.uvcontsub_flagger(vis.copy(), flags, major_cycles=3, sigma=13.0),
         oracle.uvcontsub_flagger(vis.copy(), flags, major_cycles=3, sigma=13.0))
    same(F.flag_nans_and_zeros(vis, flags), oracle.flag_nans_and_zeros(vis, flags))
