# -*- coding: utf-8 -*-
"""Host-side logic of the drop-in layer (no kernels involved)."""
import numpy as np
import pytest

import oracle
import tricolour_b200 as tb
from tricolour_b200 import flagging as G, packing as P
from tricolour_b200.util import casa_style_range
import common


def test_casa_style_range():
    assert casa_style_range("") == (0, np.inf)
    assert casa_style_range(" * ") == (0, np.inf)
    assert casa_style_range("0~550") == [0.0, 550.0]
    assert casa_style_range("1.5e2~2e3m") == [150.0, 2000.0]
    for bad in ("abc", "1~", "~2", "1-2"):
        with pytest.raises(ValueError):
            casa_style_range(bad)
    with pytest.raises(ValueError):
        casa_style_range(3)


def test_as_min_dtype_and_radii():
    assert G._as_min_dtype(3).dtype == np.uint8
    assert G._as_min_dtype(300).dtype == np.uint16
    assert G._as_min_dtype(70000).dtype == np.uint32
    assert G._as_min_dtype(-1).dtype == np.int64
    # SURVEY.md section 7 table
    for sig, r in (((62.5, 50), (54, 43)), ((12.5, 10), (10, 8)), ((32.5, 320), (28, 277)),
                   ((6.5, 64), (5, 55)), ((10, 50), (8, 43)), ((2, 10), (1, 8))):
        assert tuple(G._box_radii(sig)) == r
    rad = G._background_radii(5, (12.5, 10.0))
    assert rad.shape == (6, 2) and tuple(rad[0]) == (54, 43) and tuple(rad[-1]) == tuple(rad[-2]) == (10, 8)
    assert np.array_equal(rad, oracle.background_radii(5, (12.5, 10.0)))


def test_plan_matches_reference_conditioning():
    plan = G._StPlan(10, [1, 2, 4, 8], [32, 48, 64, 128], 2.0, 5, 6.5, 64.0, G._as_min_dtype(3),
                     G._as_min_dtype(3), np.linspace(0, 4096, 11).astype(np.int_), G._as_min_dtype(1),
                     0.6, 0.8, 1.3, 1)
    assert plan.ce.tolist() == [0, 409, 819, 1228, 1638, 2048, 2457, 2867, 3276, 3686, 4096]
    assert np.array_equal(plan.tff, oracle.threshold_factors([32, 48, 64, 128], 1.3))
    assert plan.scf.dtype == np.float32 and plan.scf[1] == np.float32(1.0 / 48)
    assert plan.params.nwin_freq == 4 and plan.params.background_iterations == 5
    assert plan.r2[0].tolist() == [28, 277]


def test_unique_baselines_and_row_slots():
    a1 = np.array([0, 0, 1, 0, 1, 2, 0], np.int32)
    a2 = np.array([1, 2, 2, 1, 1, 2, 0], np.int32)
    ub = P.unique_baselines(a1, a2).view(np.int32).reshape(-1, 2)
    # int64 order: antenna2 is the high word
    assert ub.tolist() == [[0, 0], [0, 1], [1, 1], [0, 2], [1, 2], [2, 2]]
    ubl = np.concatenate([np.arange(6, dtype=np.int32)[:, None], ub], 1)
    tinv = np.array([0, 0, 0, 0, 1, 1, 1])
    slot, t = P._row_slots(ubl, a1, a2, tinv, last_wins=True)
    # rows 0 and 3 collide on (baseline 0-1, t=0): the later row wins
    assert slot.tolist() == [-1, 3, 4, 1, 2, 5, 0]
    slot, _ = P._row_slots(ubl[:3], a1, a2, tinv, last_wins=False)
    assert slot.tolist() == [1, -1, -1, 1, 2, -1, 0]


def test_stokes_corr_map():
    m = tb.stokes_corr_map([9, 10, 11, 12])
    assert m == {'I': (0, 3, 0.5 + 0j, 1, 1), 'Q': (0, 3, 0.5 + 0j, 1, -1),
                 'U': (1, 2, 0.5 + 0j, 1, 1), 'V': (1, 2, -0.5j, 1, -1)}
    assert tb.stokes_corr_map([9, 12]) .keys() == {'I', 'Q'}
    assert tb.stokes_corr_map([5, 6, 7, 8])['U'] == (1, 2, -0.5j, 1, -1)


def test_default_strategy_matches_yaml():
    """tests/common.py restates tricolour/conf/default.yaml; check it against the
    file when the reference checkout is present"""
    import os
    path = "/root/reference/tricolour/conf/default.yaml"
    if not os.path.exists(path):
        pytest.skip("reference checkout not available")
    assert tb.strategy.load_strategies(path) == common.default_strategies()


def test_window_statistics_container():
    a, b = tb.WindowStatistics(4), tb.WindowStatistics(4)
    a._counts_per_ant["x"] += 2
    a._size_per_ant["x"] += 10
    b._counts_per_ant["x"] += 3
    b._size_per_ant["x"] += 10
    b._counts_per_ddid[0] += np.arange(4, dtype=np.uint64)
    b._bins_per_ddid[0] = np.arange(4.0)
    c = tb.combine_window_stats([a, b])
    assert c._counts_per_ant["x"] == 5 and c._size_per_ant["x"] == 20
    assert c._counts_per_ddid[0].tolist() == [0, 1, 2, 3]
    assert a._counts_per_ant["x"] == 2


def test_hoisted_reciprocal_division_is_exact_over_its_guard_range():
    """B2Div (csrc/k_filter2.cuh): x / d by q0 = x * rinv, rem = fma(-q0, d, x), q = fma(rem, rinv, q0)
    with rinv = the refined reciprocal of d = float32(2r + 1) ** 4.  The kernels use it for
    |x| in [2^-87, 2^123) (and zero) and the plain division elsewhere; here the sequence is
    replayed in exact rational arithmetic (every fma rounded once to float32) and compared
    with the correctly rounded quotient, over the whole guarded range incl. its ends."""
    import random
    from fractions import Fraction as Fr

    def f32(fr):
        if fr == 0:
            return Fr(0)
        s = 1 if fr > 0 else -1
        a = abs(fr)
        e = a.numerator.bit_length() - a.denominator.bit_length()
        if Fr(2) ** e > a:
            e -= 1
        if Fr(2) ** (e + 1) <= a:
            e += 1
        e = max(e, -126)
        q = Fr(2) ** (e - 23)
        m = a / q
        fl = m.numerator // m.denominator
        rem = m - fl
        if rem > Fr(1, 2) or (rem == Fr(1, 2) and fl % 2 == 1):
            fl += 1
        return s * fl * q

    def fma(a, b, c):
        return f32(a * b + c)

    rnd = random.Random(11)
    for r in (4, 8, 43, 127, 277):
        d32 = np.float32(2 * r + 1)
        d32 = np.float32(d32 * d32)
        d32 = np.float32(d32 * d32)
        d = Fr(float(d32))
        r0 = f32(1 / d)
        rinv = fma(r0, fma(-d, r0, Fr(1)), r0)
        for i in range(1500):
            ex = (-87, 122, rnd.randint(-87, -30), rnd.randint(-30, 60), rnd.randint(60, 122))[i % 5]
            x = Fr(rnd.randint(2 ** 23, 2 ** 24 - 1)) * Fr(2) ** (ex - 23)
            q0 = f32(x * rinv)
            q = fma(fma(-q0, d, x), rinv, q0)
            assert q == f32(x / d), (r, ex)


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm the driver launches beside ours) prints ONE JSON
    line with the keys of the bench contract, runs without a GPU, and does nothing on ranks
    other than 0 when it is launched under torchrun."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--ntime", "32", "--nchan", "256", "--cpu-planes", "2"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GVis/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["config_index"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GVis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True,
                        timeout=600)
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]
