# -*- coding: utf-8 -*-
"""GPU-only tests of the pieces around the kernels: the device-resident strategy
executor, torch-tensor (device pointer) entry, thread re-entrancy, the
BASELINE.json configurations at reduced baseline counts, and size-independent
properties at full block size."""
from multiprocessing.pool import ThreadPool

import numpy as np
import pytest

import oracle
import tricolour_b200 as tb
import common

pytestmark = pytest.mark.gpu


def _setup(nant, F):
    ubl = common.baselines(nant)
    ants = common.antenna_layout(nant)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    return ubl, ants, cf, cw, masks


def test_default_strategy_matches_oracle_chain(cuda_lib):
    """config 2 in miniature: the 12 default.yaml tasks with the reference's
    combine rules, device resident, against the oracle run task by task"""
    nant, T, F = 3, 64, 512
    ubl, ants, cf, cw, masks = _setup(nant, F)
    vis, flags = common.make_windows(ubl.shape[0], 2, T, F, seed=101, ubl=ubl)
    strategies = common.default_strategies()
    want = common.run_strategies(oracle, strategies, vis, flags, ubl, ants, masks, cf, cw)
    ex = tb.StrategyExecutor(ants, ubl, cf, cw, masks, strategies)
    got = ex.apply_strategies(flags, vis)
    assert got.dtype == flags.dtype and got.shape == flags.shape
    assert (got != want).mean() <= 1e-6, int((got != want).sum())
    # the per-function numpy API chained on the host gives the same thing
    got2 = common.run_strategies(tb, strategies, vis, flags, ubl, ants, masks, cf, cw)
    assert np.array_equal(got, got2)
    # autos are fully flagged, input flags survive (tasks 11 and 12)
    assert got[ubl[:, 1] == ubl[:, 2]].all()
    assert got[flags].all()


def test_device_tensor_api(cuda_lib):
    import torch
    vis, flags = common.make_windows(2, 2, 48, 256, seed=102)
    dv, df = torch.from_numpy(vis).cuda(), torch.from_numpy(flags).cuda()
    kw = dict(common.DEFAULT_STRATEGY_KW["final_st_broad"])
    out = tb.sum_threshold_flagger(dv, df, **kw)
    assert out.is_cuda and out.dtype == torch.bool
    assert np.array_equal(out.cpu().numpy(), tb.sum_threshold_flagger(vis, flags, **kw))
    assert np.array_equal(tb.flag_nans_and_zeros(dv, df).cpu().numpy(), oracle.flag_nans_and_zeros(vis, flags))
    uv = tb.uvcontsub_flagger(dv, df, major_cycles=2, sigma=15.0)
    assert np.array_equal(uv.cpu().numpy(), tb.uvcontsub_flagger(vis, flags, major_cycles=2, sigma=15.0))
    # inputs untouched
    assert np.array_equal(dv.cpu().numpy(), vis, equal_nan=True) and np.array_equal(df.cpu().numpy(), flags)
    # a non-default stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out2 = tb.sum_threshold_flagger(dv, df, **kw)
    s.synchronize()
    assert torch.equal(out, out2)


def test_thread_reentrancy(cuda_lib):
    """dask calls the functions from a ThreadPool (app.py:266-271)"""
    blocks = [common.make_windows(1, 2, 40, 200, seed=200 + i) for i in range(6)]
    kw = dict(outlier_nsigma=10, background_iterations=3, num_major_iterations=2)
    serial = [tb.sum_threshold_flagger(v, f, **kw) for v, f in blocks]
    with ThreadPool(4) as pool:
        par = pool.starmap(lambda v, f: tb.sum_threshold_flagger(v, f, **kw), blocks)
    for a, b in zip(serial, par):
        assert np.array_equal(a, b)
    assert np.array_equal(serial[0], oracle.sum_threshold_flagger(*blocks[0], **kw))


def test_config3_polarised_wideband(cuda_lib):
    """config 3 in miniature: rows (T*nbl, 32768 chan, 4 corr) -> Q,U,V polarised
    intensity -> windows (nbl, 1, T, 32768) -> sum_threshold"""
    nant, T, F = 2, 16, 32768
    ubl = common.baselines(nant)
    nbl = ubl.shape[0]
    rs = np.random.RandomState(5)
    a1 = np.tile(ubl[:, 1], T).astype(np.int32)
    a2 = np.tile(ubl[:, 2], T).astype(np.int32)
    tinv = np.repeat(np.arange(T), nbl)
    rows = (rs.standard_normal((T * nbl, F, 4)) + 1j * rs.standard_normal((T * nbl, F, 4))).astype(np.complex64)
    rows[:, 1000:1003] += 40
    smap = tb.stokes_corr_map([9, 10, 11, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    pi = tb.polarised_intensity(rows, pol)
    np.testing.assert_allclose(pi.real, oracle.polarised_intensity(rows, pol).real, rtol=1.2e-7)
    flags = np.zeros(pi.shape, bool)
    vw, fw = tb.pack_data(tinv, ubl, a1, a2, pi, flags, T)
    assert vw.shape == (nbl, 1, T, F)
    kw = dict(common.DEFAULT_STRATEGY_KW["final_st_very_broad"])
    got = tb.sum_threshold_flagger(vw, fw, **kw)
    want = oracle.sum_threshold_flagger(vw, fw, nthreads=4, **kw)
    assert np.array_equal(got, want)
    assert got[:, :, :, 1000:1003].mean() > 0.9


def test_config4_uvcontsub_then_sumthreshold(cuda_lib):
    """config 4 in miniature: default.yaml tasks 4..7 on 1024 dumps, cross baselines"""
    nant, T, F = 3, 1024, 256
    ubl = common.baselines(nant, autos=False)
    ants = common.antenna_layout(nant)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    vis, flags = common.make_windows(ubl.shape[0], 2, T, F, seed=104, ubl=ubl)
    strategies = common.default_strategies()[3:7]
    want = common.run_strategies(oracle, strategies, vis, flags, ubl, ants, masks, cf, cw)
    got = tb.StrategyExecutor(ants, ubl, cf, cw, masks, strategies).apply_strategies(flags, vis)
    assert (got != want).mean() <= 1e-6


def test_config5_rows_to_rows(cuda_lib):
    """config 5 in miniature: MS row order -> pack -> strategy -> unpack (+corr
    equalisation) -> window statistics, baselines split over two 'ranks'"""
    nant, T, F, ncorr = 4, 32, 128, 4
    ubl, ants, cf, cw, masks = _setup(nant, F)
    nbl = ubl.shape[0]
    vis_w, flag_w = common.make_windows(nbl, ncorr, T, F, seed=105, ubl=ubl)
    vis_w = np.nan_to_num(vis_w, nan=0.0)
    # row order: time-major, one row per (t, baseline)
    a1 = np.tile(ubl[:, 1], T).astype(np.int32)
    a2 = np.tile(ubl[:, 2], T).astype(np.int32)
    tinv = np.repeat(np.arange(T), nbl)
    rows = np.ascontiguousarray(vis_w.transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr))
    rflags = np.ascontiguousarray(flag_w.transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr))
    strategies = common.default_strategies()[:3] + common.default_strategies()[10:]
    names = ["m%03d" % i for i in range(nant)]
    outs, stats = [], []
    for lo, hi in ((0, nbl // 2), (nbl // 2, nbl)):       # two baseline shards
        sub = ubl[lo:hi]
        vw, fw = tb.pack_data(tinv, np.concatenate([np.arange(hi - lo, dtype=np.int32)[:, None], sub[:, 1:]], 1),
                              a1, a2, rows, rflags, T)
        assert np.array_equal(vw, vis_w[lo:hi]) and np.array_equal(fw, flag_w[lo:hi])
        su = sub.copy()
        su[:, 0] -= su[0, 0]
        fl = tb.StrategyExecutor(ants, su, cf, cw, masks, strategies).apply_strategies(fw, vw)
        want = common.run_strategies(oracle, strategies, vw, fw, su, ants, masks, cf, cw)
        assert np.array_equal(fl, want)
        outs.append((su, fl))
        stats.append(tb.window_stats(fl, sub, cf, names, 0, "f", 0))
    total = tb.combine_window_stats(stats)
    full = np.concatenate([o[1] for o in outs])
    assert int(total._counts_per_field["f"]) == int(full.sum())
    # unpack every shard back to row order and OR them (each row belongs to one shard)
    unp = np.zeros(rflags.shape, bool)
    for su, fl in outs:
        sel = np.concatenate([np.arange(su.shape[0], dtype=np.int32)[:, None], su[:, 1:]], 1)
        unp |= tb.packing.unpack_flags_equalised(a1, a2, tinv, sel, fl)
    expect = full.transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr)
    expect = np.broadcast_to(expect.any(axis=2, keepdims=True), expect.shape)
    assert np.array_equal(unp, expect)


def test_full_block_properties(cuda_lib):
    """one dask block of config 2 at full size (8 baselines to keep it short):
    no oracle, size-independent properties only"""
    nbl, ncorr, T, F = 8, 4, 512, 4096
    ubl = common.baselines(64)[:nbl]
    vis, flags = common.make_windows(nbl, ncorr, T, F, seed=106, ubl=ubl)
    nz = tb.flag_nans_and_zeros(vis, flags)
    assert np.array_equal(nz, flags | (vis == 0) | np.isnan(vis))
    assert np.array_equal(tb.flag_nans_and_zeros(vis, nz), nz)            # idempotent
    kw = dict(common.DEFAULT_STRATEGY_KW["final_st_broad"])
    a = tb.sum_threshold_flagger(vis, nz, **kw)
    b = tb.sum_threshold_flagger(vis, nz, **kw)
    assert np.array_equal(a, b)                                            # deterministic
    # planes are independent: a sub-block gives the same flags as inside the batch
    c = tb.sum_threshold_flagger(vis[2:4], nz[2:4], **kw)
    assert np.array_equal(a[2:4], c)
    assert a[np.isnan(vis)].all()                                          # isnan OR (flagging.py:776-781)
    # one plane against the oracle at full size
    d = oracle.sum_threshold_flagger(vis[5:6, 1:2], nz[5:6, 1:2], **kw)
    assert np.array_equal(a[5:6, 1:2], d)
    st = tb.window_stats(a, ubl, common.channels(F)[0], ["m%03d" % i for i in range(64)], 0, "f", 0)
    assert int(st._counts_per_field["f"]) == int(a.sum())


@pytest.mark.gpu
def test_pipelined_executor_matches_block_by_block(cuda_lib):
    """apply_strategies_pipelined (transfers overlapped with flagging) returns, block
    for block, what apply_strategies returns"""
    import tricolour_b200 as tb
    nbl, T, F = 3, 32, 128
    ubl = common.baselines(8)[:nbl].copy()
    ants = common.antenna_layout(8)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    ex = tb.StrategyExecutor(ants, ubl, cf, cw, masks, common.default_strategies())
    blocks = [common.make_windows(nbl, 2, T, F, seed=90 + k, ubl=ubl) for k in range(4)]
    want = [ex.apply_strategies(f, v) for v, f in blocks]
    got = list(ex.apply_strategies_pipelined([(f, v) for v, f in blocks]))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.dtype == w.dtype and g.shape == w.shape
        assert np.array_equal(g, w)
    assert list(ex.apply_strategies_pipelined([])) == []


@pytest.mark.gpu
def test_pipelined_executor_with_packing_hooks(cuda_lib):
    """MS rows through the pipelined executor: pre = pack_data, post = unpack + correlation
    equalisation on the flagging stream; block for block the result of the same calls made
    one after the other on resident tensors"""
    import torch
    nbl, ncorr, T, F = 3, 4, 32, 128
    ubl = common.baselines(8)[:nbl].copy()
    ubl[:, 0] = np.arange(nbl)
    ants = common.antenna_layout(8)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    ex = tb.StrategyExecutor(ants, ubl, cf, cw, masks, common.default_strategies()[:4])
    a1 = np.tile(ubl[:, 1], T).astype(np.int32)
    a2 = np.tile(ubl[:, 2], T).astype(np.int32)
    tinv = np.repeat(np.arange(T), nbl)

    def rows_of(w):          # (bl, corr, T, F) -> (row = t * nbl + bl, chan, corr)
        return np.ascontiguousarray(w.transpose(2, 0, 3, 1).reshape(T * nbl, F, ncorr))

    blocks = []
    for k in range(4):
        v, f = common.make_windows(nbl, ncorr, T, F, seed=290 + k, ubl=ubl)
        blocks.append((rows_of(f), rows_of(v)))
    seen = []

    def pre(rfl, rows):
        vw, fw = tb.pack_data(tinv, ubl, a1, a2, rows, rfl, T)
        seen.append(tuple(fw.shape))
        return fw, vw

    def post(out):
        return tb.packing.unpack_flags_equalised(a1, a2, tinv, ubl, out)

    want = []
    for rf, rv in blocks:
        out = ex.apply_strategies(*pre(torch.from_numpy(rf).cuda(), torch.from_numpy(rv).cuda()))
        want.append(post(out).cpu().numpy())
    got = list(ex.apply_strategies_pipelined(blocks, pre=pre, post=post))
    assert seen[0] == (nbl, ncorr, T, F)
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.shape == (T * nbl, F, ncorr) and g.dtype == np.bool_
        assert np.array_equal(g, w), "block %d" % k
    # the same executor goes back to plain windows (download buffers follow the result shape)
    v, f = common.make_windows(nbl, 2, T, F, seed=300, ubl=ubl)
    plain = list(ex.apply_strategies_pipelined([(f, v)]))
    assert np.array_equal(plain[0], ex.apply_strategies(f, v))


@pytest.mark.gpu
def test_plane_batching_is_invisible(cuda_lib, monkeypatch):
    """tc_sum_threshold cuts the planes into batches that fit TC_WORKSPACE_MB; the
    flags must not depend on where the cuts fall"""
    vis, flags = common.make_windows(12, 4, 64, 1024, seed=77)
    kw = dict(common.DEFAULT_STRATEGY_KW["final_st_broad"])
    whole = tb.sum_threshold_flagger(vis, flags, **kw)
    monkeypatch.setenv("TC_WORKSPACE_MB", "64")          # 48 planes x 64 Ki samples x ~50 B: several batches
    cut = tb.sum_threshold_flagger(vis, flags, **kw)
    assert np.array_equal(whole, cut)
    sub = oracle.sum_threshold_flagger(vis[:1], flags[:1], **kw)
    assert np.array_equal(whole[:1], sub)


@pytest.mark.gpu
def test_pipelined_executor_single_refilled_pinned_buffer(cuda_lib):
    """the usage INTEGRATION.md recommends: ONE page-locked (flags, vis) pair that the
    generator refills in place for every block.  The executor must have finished reading
    a block before it asks the generator for the next one."""
    from tricolour_b200 import _cabi
    nbl, T, F = 4, 64, 512
    ubl = common.baselines(8)[:nbl].copy()
    ants = common.antenna_layout(8)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    strategies = common.default_strategies()[:3]
    ex = tb.StrategyExecutor(ants, ubl, cf, cw, masks, strategies)
    blocks = [common.make_windows(nbl, 2, T, F, seed=190 + k, ubl=ubl) for k in range(5)]
    want = [ex.apply_strategies(f, v) for v, f in blocks]
    hv = _cabi.pinned_empty((nbl, 2, T, F), np.complex64)
    hf = _cabi.pinned_empty((nbl, 2, T, F), np.bool_)

    def gen():
        for v, f in blocks:
            hv[...] = v            # overwrites what the previous block was uploaded from
            hf[...] = f
            yield hf, hv

    got = list(ex.apply_strategies_pipelined(gen()))
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), "block %d" % k
    _cabi.free_pinned(hv)
    _cabi.free_pinned(hf)


@pytest.mark.gpu
def test_mixed_host_and_device_arguments_are_rejected(cuda_lib):
    import torch
    vis, flags = common.make_windows(1, 1, 16, 64, seed=5)
    dv = torch.from_numpy(vis).cuda()
    with pytest.raises(TypeError):
        tb.sum_threshold_flagger(dv, flags)
    with pytest.raises(TypeError):
        tb.uvcontsub_flagger(vis, torch.from_numpy(flags).cuda())
    with pytest.raises(TypeError):
        tb.flag_nans_and_zeros(dv, flags)


@pytest.mark.gpu
def test_workspace_budget_is_shared_between_thread_contexts(cuda_lib, monkeypatch):
    """every worker thread owns a context; together they must stay inside
    TC_WORKSPACE_MB instead of each growing to the whole budget"""
    import threading
    from tricolour_b200 import _cabi
    monkeypatch.setenv("TC_WORKSPACE_MB", "2048")
    vis, flags = common.make_windows(4, 4, 64, 1024, seed=78)
    kw = dict(common.DEFAULT_STRATEGY_KW["final_st_broad"])
    want = tb.sum_threshold_flagger(vis, flags, **kw)
    held, outs, shares = [], [], []
    gate = threading.Barrier(4)

    def work():
        out = tb.sum_threshold_flagger(vis, flags, **kw)
        gate.wait()                       # all four arenas alive at the same time
        ctx = _cabi.get_context()
        held.append(ctx.workspace_held())
        shares.append(ctx.workspace_share())
        out2 = tb.sum_threshold_flagger(vis, flags, **kw)
        outs.append((out, out2))
        _cabi.release_contexts()

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(outs) == 4
    for a, b in outs:
        assert np.array_equal(a, want) and np.array_equal(b, want)
    assert sum(held) <= 2048 << 20
    assert max(shares) <= (2048 << 20) // 4


_TMA_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import oracle
from tricolour_b200 import flagging as G
rs = np.random.RandomState(77)
worst = 0
for shape, (r0, r1) in [((2, 64, 256), (5, 8)), ((1, 128, 132), (10, 12)), ((3, 32, 512), (3, 4)), ((1, 48, 96), (11, 1))]:
    sig = np.array([np.sqrt(((2 * r + 1) ** 2 - 1) / 3.0) + 1e-6 for r in (r0, r1)])
    for p in range(shape[0]):
        d = (rs.uniform(size=shape[1:]) * 10 ** rs.uniform(-2, 2, shape[1:])).astype(np.float32)
        fl = rs.uniform(size=shape[1:]) < 0.3
        fl[shape[1] // 4:shape[1] // 2, 2:shape[2] // 2] = True
        o = np.zeros_like(d); o2 = np.zeros_like(d)
        G.masked_gaussian_filter(d, fl, sig, o)
        oracle.masked_gaussian_filter(d, fl, sig, o2)
        worst += int((o.view(np.uint32) != o2.view(np.uint32)).sum() - (np.isnan(o) & np.isnan(o2)).sum())
print("TMA_DIFF", worst)
'''


def test_tma_form_of_the_second_axis_filter(cuda_lib):
    """the TMA-staged second-axis filter (k_filter5t.cuh) is opt-in since the plain-load
    form overtook it: run it in a process of its own (the switch is read once) and
    compare with the oracle bit for bit"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TC_FILTER_TMA="1", TC_FILTER_TRACE="1")
    script = _TMA_SCRIPT % {"root": root, "tests": os.path.join(root, "tests")}
    r = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "TMA_DIFF 0" in r.stdout, r.stdout[-500:]
    assert "b5t filter" in r.stderr, "the TMA form did not run:\n" + r.stderr[-1000:]
