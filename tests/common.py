# -*- coding: utf-8 -*-
"""Seeded synthetic MeerKAT-like inputs shared by the tests and bench.py
(smooth bandpass, slow gain drift, complex Gaussian noise,
persistent / broadband / blob RFI, exact zeros, NaNs, pre-flagged band)."""
import numpy as np

DEFAULT_STRATEGY_KW = {
    "background_flags": dict(outlier_nsigma=10, windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                             background_reject=2.0, background_iterations=5, spike_width_time=12.5,
                             spike_width_freq=10.0, time_extend=3, freq_extend=3, freq_chunks=10,
                             average_freq=1, flag_all_time_frac=0.6, flag_all_freq_frac=0.8, rho=1.3,
                             num_major_iterations=5),
    "final_st_very_broad": dict(outlier_nsigma=10, windows_time=[1, 2, 4, 8], windows_freq=[32, 48, 64, 128],
                                background_reject=2.0, background_iterations=5, spike_width_time=6.5,
                                spike_width_freq=64.0, time_extend=3, freq_extend=3, freq_chunks=10,
                                average_freq=1, flag_all_time_frac=0.6, flag_all_freq_frac=0.8, rho=1.3,
                                num_major_iterations=1),
    "final_st_broad": dict(outlier_nsigma=10, windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                           background_reject=2.0, background_iterations=5, spike_width_time=6.5,
                           spike_width_freq=10.0, time_extend=3, freq_extend=3, freq_chunks=10,
                           average_freq=1, flag_all_time_frac=0.6, flag_all_freq_frac=0.8, rho=1.3,
                           num_major_iterations=1),
    "final_st_narrow": dict(outlier_nsigma=10, windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                            background_reject=2.0, background_iterations=5, spike_width_time=2,
                            spike_width_freq=10.0, time_extend=3, freq_extend=3, freq_chunks=10,
                            average_freq=1, flag_all_time_frac=0.6, flag_all_freq_frac=0.8, rho=1.3,
                            num_major_iterations=1),
}


def default_strategies():
    """The 12 tasks of tricolour/conf/default.yaml:1-126, as the YAML loader
    yields them (ints stay ints)."""
    kw = DEFAULT_STRATEGY_KW
    return [
        dict(name="nan_dropouts_flag", task="flag_nans_zeros"),
        dict(name="background_static_mask", task="apply_static_mask",
             kwargs=dict(accumulation_mode="or", uvrange="")),
        dict(name="background_flags", task="sum_threshold", kwargs=dict(kw["background_flags"])),
        dict(name="residual_flag_initial", task="uvcontsub_flagger",
             kwargs=dict(major_cycles=7, or_original_from_cycle=1, taylor_degrees=20, sigma=15.0)),
        dict(name="nan_dropouts_reflag", task="flag_nans_zeros"),
        dict(name="uvrange_static_mask", task="apply_static_mask",
             kwargs=dict(accumulation_mode="or", uvrange="0~550")),
        dict(name="final_st_very_broad", task="sum_threshold", kwargs=dict(kw["final_st_very_broad"])),
        dict(name="final_st_broad", task="sum_threshold", kwargs=dict(kw["final_st_broad"])),
        dict(name="final_st_narrow", task="sum_threshold", kwargs=dict(kw["final_st_narrow"])),
        dict(name="residual_flag_final", task="uvcontsub_flagger",
             kwargs=dict(major_cycles=10, or_original_from_cycle=0, taylor_degrees=25, sigma=13.0)),
        dict(name="flag_autos", task="flag_autos"),
        dict(name="combine_with_input_flags", task="combine_with_input_flags"),
    ]


def antenna_layout(nant=64, seed=20261018):
    """ECEF-like metres: uniform disc of 4 km radius plus a dense core."""
    rs = np.random.RandomState(seed)
    r = 4000.0 * np.sqrt(rs.uniform(size=nant))
    r[:max(nant // 8, 1)] *= 0.1
    th = rs.uniform(0, 2 * np.pi, nant)
    return np.stack([5109000.0 + r * np.cos(th), 2006000.0 + r * np.sin(th),
                     np.full(nant, -3239000.0)], axis=1)


def baselines(nant=64, autos=True):
    """(nbl, 3) int32 (index, a1, a2) in unique_baselines order (a2-major)."""
    a1, a2 = np.triu_indices(nant, 0 if autos else 1)
    pairs = np.stack([a1, a2], axis=1).astype(np.int32)
    key = np.ascontiguousarray(pairs).view(np.int64).ravel()
    pairs = pairs[np.argsort(key, kind="stable")]
    idx = np.arange(pairs.shape[0], dtype=np.int32)[:, None]
    return np.concatenate([idx, pairs], axis=1)


def channels(nchan=4096):
    width = 856e6 / nchan
    return 856e6 + (np.arange(nchan) + 0.5) * width, np.full(nchan, width)


def synthetic_static_mask(chan_freqs, seed=3):
    """A few masked frequency ranges (Hz), shape (n, 1) like load_mask output."""
    rs = np.random.RandomState(seed)
    lo, hi = chan_freqs.min(), chan_freqs.max()
    out = []
    for _ in range(6):
        c = rs.uniform(lo, hi)
        w = rs.uniform(0.002, 0.01) * (hi - lo)
        out.append(np.arange(c, c + w, (hi - lo) / (4 * chan_freqs.size)))
    return [np.concatenate(out)[:, None]]


def make_windows(nbl, ncorr, T, F, seed=20261019, ubl=None):
    """(vis complex64, flags bool) windows of shape (nbl, ncorr, T, F)."""
    rs = np.random.RandomState(seed)
    x = np.linspace(0.0, 1.0, F)
    knots = np.linspace(0, 1, 10)
    kv = np.full(10, 2.34) + rs.uniform(0, 0.1, 10)
    kv[0] = kv[-1] = 0.1
    bp = np.interp(x, knots, kv).astype(np.float32)
    drift = (1.0 + 0.02 * np.sin(2 * np.pi * np.arange(T) / max(T, 2) * rs.uniform(0.5, 2)))
    amp = bp[None, None, None, :] * drift[None, None, :, None].astype(np.float32)
    amp = np.broadcast_to(amp, (nbl, ncorr, T, F)).copy()
    if ubl is not None:
        autos = ubl[:, 1] == ubl[:, 2]
        amp[autos] *= 50.0
    sigma = 0.1 * bp[None, None, None, :]
    noise = (rs.standard_normal((nbl, ncorr, T, F)) + 1j * rs.standard_normal((nbl, ncorr, T, F)))
    phase = np.exp(1j * rs.uniform(0, 2 * np.pi, (nbl, ncorr, 1, 1)))
    vis = (amp * phase + sigma * noise / np.sqrt(2)).astype(np.complex64)
    # persistent channels, broadband dumps, blobs, a faint persistent channel
    # persistent RFI: a few narrow bands (real RFI is clustered, not spread evenly)
    nband = max(F // 1000, 1)
    for f in rs.choice(max(F - 4, 1), nband, replace=False):
        wband = rs.randint(1, 4)
        vis[:, :, :, f:f + wband] += (rs.uniform(5, 40) * sigma[..., f:f + wband]).astype(np.complex64)
    for t in rs.choice(T, max(T // 200, 1), replace=False):
        vis[:, :, t, :] += (rs.uniform(5, 20) * sigma[0, 0, 0, :]).astype(np.complex64)
    for _ in range(20):
        b, c = rs.randint(nbl), rs.randint(ncorr)
        if rs.uniform() < 0.5:
            h, w = min(5, T), min(70, F)
        else:
            h, w = min(50, T), min(3, F)
        t0, f0 = rs.randint(0, T - h + 1), rs.randint(0, F - w + 1)
        vis[b, c, t0:t0 + h, f0:f0 + w] += (rs.uniform(5, 30) * 0.1 * 2.34)
    vis[:, :, :, F // 2 + 3] += np.complex64(0.2 * 0.234 * np.sqrt(T) / max(T, 1) * 10)
    n = vis.size
    flat = vis.reshape(-1)
    flat[rs.choice(n, max(n // 1000, 1), replace=False)] = 0
    flat[rs.choice(n, max(n // 1000, 1), replace=False)] = np.nan + 1j * np.nan
    flags = rs.uniform(size=vis.shape) < 0.02
    # missing (bl, t) rows look like pack_data's defaults: NaN + flagged
    miss = rs.uniform(size=(nbl, T)) < 0.01
    vis[miss[:, None, :, None] & np.ones((1, ncorr, 1, F), bool)] = np.nan + 1j * np.nan
    flags |= miss[:, None, :, None]
    b0 = min(185 * F // 345, F - 1)
    flags[:, :, :, b0:min(b0 + max(F // 70, 1), F)] = True
    return vis, flags


def run_strategies(mod, strategies, vis, flags, ubl, antspos, masks, chan_freq, chan_width):
    """The combine rules of tricolour/apps/tricolour/strat_executor.py:29-83
    applied with module ``mod`` (the oracle or tricolour_b200) on numpy arrays."""
    original = flags.copy()
    for s in strategies:
        task = s["task"]
        kw = s.get("kwargs", {}) or {}
        if task == "sum_threshold":
            flags = np.logical_or(mod.sum_threshold_flagger(vis, flags, **kw), flags)
        elif task == "uvcontsub_flagger":
            flags = mod.uvcontsub_flagger(vis, flags, **kw)
        elif task == "flag_autos":
            flags = np.logical_or(mod.flag_autos(flags, [ubl]), flags)
        elif task == "combine_with_input_flags":
            flags = np.logical_or(flags, original)
        elif task == "unflag":
            flags = np.zeros_like(flags)
        elif task == "flag_nans_zeros":
            flags = mod.flag_nans_and_zeros(vis, flags)
        elif task == "apply_static_mask":
            new = mod.apply_static_mask(flags, ubl, antspos, masks, chan_freq, chan_width, **kw)
            flags = np.logical_or(new, flags) if kw["accumulation_mode"].strip() == "or" else new
        else:
            raise ValueError(task)
    return flags


def run_strategies_planes(mod, strategies, vis, flags, ubl, antspos, masks, chan_freq, chan_width,
                          threads=None):
    """``run_strategies`` plane by plane in a ThreadPool (the reference's execution
    model, app.py:266-271).  Every (baseline, correlation) plane is independent in
    every task of the path, so this returns what one call on the whole block returns;
    it is how the full-size parity tests and bench.py's parity check keep the CPU
    oracle's run time down to seconds per plane."""
    import os
    from multiprocessing.pool import ThreadPool
    nbl, ncorr = vis.shape[:2]
    out = np.empty(flags.shape, flags.dtype)
    jobs = [(b, c) for b in range(nbl) for c in range(ncorr)]

    def work(bc):
        b, c = bc
        out[b:b + 1, c:c + 1] = run_strategies(mod, strategies, vis[b:b + 1, c:c + 1], flags[b:b + 1, c:c + 1],
                                               ubl[b:b + 1], antspos, masks, chan_freq, chan_width)

    n = max(1, min(len(jobs), threads or (os.cpu_count() or 1)))
    if n == 1:
        for j in jobs:
            work(j)
    else:
        with ThreadPool(n) as pool:
            pool.map(work, jobs)
    return out


def pick_baselines(ubl, antspos, n):
    """``n`` baseline rows of ``ubl`` that cover the cases the strategy treats
    differently: an auto-correlation (flag_autos), the shortest and the longest
    cross baseline (uvrange of the second static mask), then evenly spread ones."""
    ubl = np.asarray(ubl)
    d = np.sqrt(0.5 * ((antspos[ubl[:, 1]] - antspos[ubl[:, 2]]) ** 2).sum(axis=1))
    auto = np.flatnonzero(ubl[:, 1] == ubl[:, 2])
    cross = np.flatnonzero(ubl[:, 1] != ubl[:, 2])
    picks = []
    if auto.size:
        picks.append(int(auto[0]))
    if cross.size:
        picks.append(int(cross[np.argmax(d[cross])]))
        picks.append(int(cross[np.argmin(d[cross])]))
    for i in np.linspace(0, ubl.shape[0] - 1, max(n, 1)).astype(int):
        picks.append(int(i))
    out = []
    for p in picks:
        if p not in out:
            out.append(p)
    return out[:n]
