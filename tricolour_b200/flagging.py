# -*- coding: utf-8 -*-
"""
Drop-in replacements for the numpy-level functions of ``tricolour.flagging``
(reference: tricolour/flagging.py), executed by the sm_100a CUDA library.

Same names, positional order, defaults and exceptions as the reference; inputs
are borrowed (never modified) and outputs are freshly allocated arrays of the
input flag dtype.  Arrays may be numpy arrays (staged through the device by the
library) or torch CUDA tensors (used in place, asynchronously on torch's
current stream; results are torch tensors).

Host-side work is limited to what the reference also does in Python before its
numba kernels start: parameter conditioning (flagging.py:1160-1179) and the
O(nbl) / O(nchan) selector tables of ``apply_static_mask`` (137-160).
"""
import ctypes
import math

import numpy as np

from . import _cabi
from ._cabi import check, ptr, context_for
from .util import casa_style_range

MAD_NORMAL = 1.4826
"""Ratio between median absolute deviation and the standard deviation of a
Gaussian distribution (tricolour/flagging.py:22)."""


# ---------------------------------------------------------------------------
# array helpers
# ---------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _as_u8_flags(flags):
    """(byte view usable by the library, restore(out_u8) -> caller's dtype)"""
    if _cabi.is_device_array(flags):
        torch = _torch()
        f = flags.contiguous()
        if f.dtype == torch.bool:
            return f.view(torch.uint8), lambda o: o.view(torch.bool)
        if f.dtype == torch.uint8:
            return f, lambda o: o
        dt = f.dtype
        return (f != 0).view(torch.uint8), lambda o: o.to(dt)
    f = np.asarray(flags)
    if f.dtype == np.bool_:
        return np.ascontiguousarray(f).view(np.uint8), lambda o: o.view(np.bool_)
    if f.dtype == np.uint8:
        return np.ascontiguousarray(f), lambda o: o
    dt = f.dtype
    return np.ascontiguousarray(f != 0).view(np.uint8), lambda o: o.astype(dt)


def _empty_u8_like(a, shape=None):
    if _cabi.is_device_array(a):
        torch = _torch()
        return torch.empty(tuple(shape) if shape is not None else tuple(a.shape),
                           dtype=torch.uint8, device=a.device)
    return np.empty(shape if shape is not None else a.shape, np.uint8)


def _as_c64(vis):
    if _cabi.is_device_array(vis):
        torch = _torch()
        if vis.dtype != torch.complex64:
            raise TypeError("visibilities on the device must be complex64, got %s" % vis.dtype)
        return vis.contiguous()
    v = np.asarray(vis)
    if v.dtype != np.complex64:
        if not np.iscomplexobj(v):
            raise TypeError("visibilities must be complex, got %s" % v.dtype)
        v = v.astype(np.complex64)
    return np.ascontiguousarray(v)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _hp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# F1 flag_nans_and_zeros (tricolour/flagging.py:29-62)
# ---------------------------------------------------------------------------
def flag_nans_and_zeros(vis_windows, flag_windows):
    """
    Flag nan and zero visibilities.

    Parameters
    ----------
    vis_windows : array
        Visibilities of shape :code:`(bl, corr, time, chan)`
    flag_windows : array
        Flags of shape :code:`(bl, corr, time, chan)`

    Returns
    -------
    array
        ``(vis == 0) | isnan(vis) | (flag != 0)`` in the dtype of ``flag_windows``
    """
    if tuple(vis_windows.shape) != tuple(flag_windows.shape):
        raise ValueError("vis_windows.shape != flag_windows.shape")
    if not _cabi.is_device_array(vis_windows) and np.asarray(vis_windows).dtype == np.complex128:
        # complex128 -> complex64 would turn tiny values into zeros; keep the
        # predicate exact by classifying the parts before narrowing
        v = np.asarray(vis_windows)
        bad = np.float32(np.nan)
        re = np.where(np.isnan(v.real), bad, np.where(v.real == 0, np.float32(0), np.float32(1)))
        im = np.where(np.isnan(v.imag), bad, np.where(v.imag == 0, np.float32(0), np.float32(1)))
        vis = np.ascontiguousarray((re + 1j * im).astype(np.complex64))
    else:
        vis = _as_c64(vis_windows)
    fl, restore = _as_u8_flags(flag_windows)
    ctx, space = context_for(vis, fl)
    out = _empty_u8_like(fl)
    check(_cabi.load().tc_flag_nans_zeros(ctx.handle, ptr(vis), ptr(fl), ptr(out),
                                          int(np.prod(fl.shape)), space))
    return restore(out)


# ---------------------------------------------------------------------------
# F2 flag_autos (tricolour/flagging.py:65-95)
# ---------------------------------------------------------------------------
def flag_autos(flags, ubl):
    """
    Flags auto-correlations

    Parameters
    ----------
    flags : array
        Flags of shape :code:`(bl, corr, time, chan)`
    ubl : list holding one :class:`numpy.ndarray`
        unique baselines (blindx, a1indx, a2indx) of shape :code:`(bl, 3)`;
        list-wrapped exactly as dask hands it to the reference (line 84)
    """
    ubl = np.asarray(ubl[0])
    if flags.shape[0] != ubl.shape[0]:
        raise ValueError("flag and ubl shape mismatch %s != %s"
                         % (flags.shape[2], ubl.shape[0]))
    fl, restore = _as_u8_flags(flags)
    sel = np.ascontiguousarray(ubl[:, 1] == ubl[:, 2]).view(np.uint8)
    ctx, space = context_for(fl)
    out = _empty_u8_like(fl)
    nbl = int(fl.shape[0])
    plane = int(np.prod(fl.shape[1:])) if nbl else 0
    check(_cabi.load().tc_flag_autos(ctx.handle, ptr(fl), _hp(sel), nbl, plane, ptr(out), space))
    return restore(out)


# ---------------------------------------------------------------------------
# F3 apply_static_mask (tricolour/flagging.py:98-172)
# ---------------------------------------------------------------------------
def _static_mask_tables(ubl, antspos, masks, chan_freqs, chan_widths, accumulation_mode, uvrange):
    """(baseline selector u8[nbl], channel mask u8[nchan], kernel mode): the host-side
    part of apply_static_mask (flagging.py:125-166), O(nbl) + O(mask frequencies x nchan)
    float64 work on metadata.  The StrategyExecutor keeps the result per task, since
    the same tables serve every block of an observation."""
    uvrange = casa_style_range(uvrange)
    ubl = np.asarray(ubl)
    chan_freqs = np.asarray(chan_freqs)
    chan_widths = np.asarray(chan_widths)
    spw_chanlb = chan_freqs - chan_widths * 0.5
    spw_chanub = chan_freqs + chan_widths * 0.5
    antspos = np.asarray(antspos)
    bl_length = antspos[ubl[:, 1]] - antspos[ubl[:, 2]]
    d2 = 0.5 * np.sum(bl_length ** 2, axis=1)
    luvrange = 0.0 if uvrange is None else min(uvrange[0], uvrange[1])
    uuvrange = np.inf if uvrange is None else max(uvrange[0], uvrange[1])
    bl_sel = np.logical_and(d2 >= luvrange ** 2, d2 <= uuvrange ** 2)

    # per mask: which channels it hits; "or" accumulates over masks, "override"
    # assigns per mask so that only the last mask survives (flagging.py:163-166)
    combined = None
    for mask in masks:
        mask = np.asarray(mask)
        if mask.ndim != 2 and mask.shape[1] != 1:
            raise ValueError("masks.shape != (dim, 1)")
        lower_mask = mask[:, :] >= spw_chanlb[None, :]
        upper_mask = mask[:, :] < spw_chanub[None, :]
        masked_channels = np.logical_and(lower_mask, upper_mask).sum(axis=0) > 0
        if accumulation_mode == "or":
            combined = masked_channels if combined is None else (combined | masked_channels)
        elif accumulation_mode == "override":
            combined = masked_channels
        else:
            raise ValueError("Invalid accumulation_mode '%s'. "
                             "Should be 'or' or 'override'" % accumulation_mode)
    if combined is None:
        # no masks: plain copy
        combined = np.zeros(chan_freqs.shape[0], np.bool_)
        mode = 0
    else:
        mode = 0 if accumulation_mode == "or" else 1
    sel = np.ascontiguousarray(bl_sel).view(np.uint8)
    cm = np.ascontiguousarray(combined).view(np.uint8)
    return sel, cm, mode


def _apply_mask_tables(flag, tables):
    sel, cm, mode = tables
    fl, restore = _as_u8_flags(flag)
    ctx, space = context_for(fl)
    out = _empty_u8_like(fl)
    nbl = int(fl.shape[0])
    nchan = int(fl.shape[-1])
    rows = int(np.prod(fl.shape[1:-1])) if nbl else 0
    check(_cabi.load().tc_apply_channel_mask(ctx.handle, ptr(fl), _hp(sel), _hp(cm), mode,
                                             nbl, rows, nchan, ptr(out), space))
    return restore(out)


def apply_static_mask(flag, ubl, antspos, masks,
                      chan_freqs, chan_widths,
                      accumulation_mode="or", uvrange=""):
    """Applies static masks, flagging channels that span frequencies included
    in a mask, for the baselines whose length lies inside ``uvrange``.

    Same arguments as the reference; ``masks`` is a list of (n, 1) float64
    arrays of masked frequencies in Hz.
    """
    ubl = np.asarray(ubl)
    if flag.shape[0] != ubl.shape[0]:
        raise ValueError("flag and ubl shape mismatch %s != %s"
                         % (flag.shape[1], ubl.shape[0]))
    tables = _static_mask_tables(ubl, antspos, masks, chan_freqs, chan_widths, accumulation_mode, uvrange)
    return _apply_mask_tables(flag, tables)


# ---------------------------------------------------------------------------
# S0 / parameter conditioning
# ---------------------------------------------------------------------------
def _as_min_dtype(value):
    """Convert a non-negative integer into a numpy scalar of the narrowest
    type that will hold it (tricolour/flagging.py:175-190)."""
    if value >= 0 and value < 2 ** 8:
        dtype = np.uint8
    elif value >= 0 and value < 2 ** 16:
        dtype = np.uint16
    elif value >= 0 and value < 2 ** 32:
        dtype = np.uint32
    else:
        dtype = np.int64
    return np.array(value, dtype)


def _box_radii(sigma, passes=4):
    """r = int(0.5 * sqrt(12 sigma^2 / passes + 1)) per axis (flagging.py:451)"""
    sigma = np.asarray(sigma, dtype=np.float64)
    return (0.5 * np.sqrt(12.0 * sigma ** 2 / passes + 1)).astype(np.int_)


def _background_radii(iterations, spike_width):
    """radii for extend_factor = iterations..1 (flagging.py:553-555) followed by
    the final filter at sigma = spike_width (576)"""
    sw = np.asarray(spike_width, dtype=np.float64)
    rows = [_box_radii(ef * sw) for ef in range(int(iterations), 0, -1)]
    rows.append(_box_radii(sw))
    return np.ascontiguousarray(np.array(rows, dtype=np.int64).reshape(-1, 2))


class _StPlan(object):
    """Host-built tc_st_params plus the arrays it points to."""

    def __init__(self, outlier_nsigma, windows_time, windows_freq,
                 background_reject, background_iterations,
                 spike_width_time, spike_width_freq, time_extend, freq_extend,
                 freq_chunk_ends, average_freq, flag_all_time_frac,
                 flag_all_freq_frac, rho, num_major_iterations):
        self.wt = _i64(windows_time)
        self.wf = _i64(windows_freq)
        # tf = pow(rho, log2(window)) evaluated by glibc in float64 (flagging.py:641)
        self.tft = np.array([math.pow(float(rho), math.log2(int(w))) if w > 0 else 1.0
                             for w in self.wt], np.float64)
        self.tff = np.array([math.pow(float(rho), math.log2(int(w))) if w > 0 else 1.0
                             for w in self.wf], np.float64)
        # rolling_scale = np.float32(1.0 / window) (flagging.py:664)
        self.sct = np.array([np.float32(1.0 / int(w)) if w > 0 else 0 for w in self.wt], np.float32)
        self.scf = np.array([np.float32(1.0 / int(w)) if w > 0 else 0 for w in self.wf], np.float32)
        self.rs = _background_radii(background_iterations, (0.0, float(spike_width_freq)))
        self.r2 = _background_radii(background_iterations,
                                    (float(spike_width_time), float(spike_width_freq)))
        self.ce = _i64(freq_chunk_ends)
        p = _cabi.StParams()
        p.outlier_nsigma = float(outlier_nsigma)
        p.nwin_time = self.wt.size
        p.nwin_freq = self.wf.size
        p.windows_time = self.wt.ctypes.data
        p.tf_time = self.tft.ctypes.data
        p.scale_time = self.sct.ctypes.data
        p.windows_freq = self.wf.ctypes.data
        p.tf_freq = self.tff.ctypes.data
        p.scale_freq = self.scf.ctypes.data
        p.background_reject = float(background_reject)
        p.background_iterations = int(background_iterations)
        p.nchunk_ends = self.ce.size
        p.radii_spec = self.rs.ctypes.data
        p.radii_2d = self.r2.ctypes.data
        p.freq_chunk_ends = self.ce.ctypes.data
        p.time_extend = int(time_extend)
        p.freq_extend = int(freq_extend)
        p.average_freq = int(average_freq)
        p.flag_all_time_frac = float(flag_all_time_frac)
        p.flag_all_freq_frac = float(flag_all_freq_frac)
        p.num_major_iterations = int(num_major_iterations)
        self.params = p


def _vis_for_st(vis, average_freq):
    """(array, vis_kind): complex64 and float32 go to the device as they are."""
    if _cabi.is_device_array(vis):
        torch = _torch()
        if vis.dtype == torch.complex64:
            return vis.contiguous(), _cabi.VIS_COMPLEX64
        if vis.dtype == torch.float32:
            return vis.contiguous(), _cabi.VIS_FLOAT32
        raise TypeError("device visibilities must be complex64 or float32, got %s" % vis.dtype)
    v = np.asarray(vis)
    if v.dtype == np.complex64:
        return np.ascontiguousarray(v), _cabi.VIS_COMPLEX64
    if v.dtype == np.float32:
        return np.ascontiguousarray(v), _cabi.VIS_FLOAT32
    if v.dtype == np.float64 and int(average_freq) == 1:
        # |x| rounded once to float32 either way (avg_data is float32, weight 1)
        return np.ascontiguousarray(v.astype(np.float32)), _cabi.VIS_FLOAT32
    if v.dtype == np.complex128 and int(average_freq) == 1:
        # the reference takes |x| in float64 (glibc hypot, as numpy's np.abs) and stores it in the
        # float32 avg_data (flagging.py:856-871, weight 1): the same single rounding on the host;
        # NaN parts give a NaN amplitude, which the kernels treat like isnan(vis)
        amp = np.abs(v)
        amp[np.isnan(v.real) | np.isnan(v.imag)] = np.nan
        return np.ascontiguousarray(amp.astype(np.float32)), _cabi.VIS_FLOAT32
    raise TypeError("sum_threshold supports complex64 / float32 visibilities "
                    "(float64 / complex128 only with average_freq == 1), got %s" % v.dtype)


def _run_sum_threshold(plan, vis3, flags3):
    vis, kind = _vis_for_st(vis3, plan.params.average_freq)
    fl, restore = _as_u8_flags(flags3)
    ctx, space = context_for(vis, fl)
    out = _empty_u8_like(fl)
    ncp, T, F = (int(s) for s in fl.shape)
    check(_cabi.load().tc_sum_threshold(ctx.handle, ctypes.byref(plan.params), ptr(vis), kind,
                                        ptr(fl), ncp, T, F, ptr(out), space))
    return out, restore


# ---------------------------------------------------------------------------
# S13 sum_threshold_flagger (tricolour/flagging.py:1076-1196)
# ---------------------------------------------------------------------------
def sum_threshold_flagger(vis, flags, outlier_nsigma=4.5,
                          windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                          background_reject=2.0, background_iterations=1,
                          spike_width_time=12.5, spike_width_freq=10.0,
                          time_extend=3, freq_extend=3,
                          freq_chunks=10, average_freq=1,
                          flag_all_time_frac=0.6, flag_all_freq_frac=0.8,
                          rho=1.3, num_major_iterations=5):
    """
    Flagger that uses the SumThreshold method
    (Offringa, A., MNRAS, 405, 155-167, 2010) to detect spikes in both the
    frequency and the time axis.  Parameters, defaults and result are those of
    the reference: per major iteration the data are frequency-averaged, a
    time-median spectrum is backgrounded and thresholded, a smooth 2-D
    background is removed, SumThreshold runs along time and frequency, the
    flags are dilated and the flag-all-time/frequency fraction rules applied.
    The flags of the LAST major iteration are returned (they do not include the
    input flags; the caller ORs them, strat_executor.py:40-43).
    """
    nbl, ncorr, ntime, nchan = vis.shape
    if tuple(flags.shape) != tuple(vis.shape):
        raise ValueError('shape mismatch')
    vis3 = vis.reshape(nbl * ncorr, ntime, nchan)
    flags3 = flags.reshape(nbl * ncorr, ntime, nchan)

    # parameter conditioning, flagging.py:1160-1179
    windows_freq = np.asarray(windows_freq, dtype=np.float32)
    windows_freq = np.ceil(windows_freq) / average_freq
    windows_freq = np.unique(windows_freq.astype(np.int_))
    time_extend = _as_min_dtype(time_extend)
    freq_extend = _as_min_dtype(freq_extend)
    average_freq = _as_min_dtype(average_freq)
    averaged_channels = (int(nchan) + int(average_freq) - 1) // int(average_freq)
    freq_chunk_ends = np.linspace(0, averaged_channels, freq_chunks + 1).astype(np.int_)
    windows_time = np.array([w for w in windows_time if w <= ntime], np.int_)
    windows_freq = np.array([w for w in windows_freq if w <= averaged_channels], np.int_)

    plan = _StPlan(outlier_nsigma, windows_time, windows_freq, background_reject,
                   background_iterations, spike_width_time, spike_width_freq,
                   time_extend, freq_extend, freq_chunk_ends, average_freq,
                   flag_all_time_frac, flag_all_freq_frac, rho, num_major_iterations)
    out, restore = _run_sum_threshold(plan, vis3, flags3)
    return restore(out).reshape(nbl, ncorr, ntime, nchan)


class SumThresholdFlagger(object):
    """Legacy class API (tricolour/flagging.py:1199-1423).  Conditioning follows
    the class (lines 1274-1289: ``ceil(w / average_freq)`` and
    ``spike_width_freq / average_freq``), which differs from the function's."""

    def __init__(self, outlier_nsigma=4.5,
                 windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                 background_reject=2.0, background_iterations=1,
                 spike_width_time=12.5, spike_width_freq=10.0,
                 time_extend=3, freq_extend=3,
                 freq_chunks=10, average_freq=1,
                 flag_all_time_frac=0.6, flag_all_freq_frac=0.8,
                 rho=1.3):
        self.outlier_nsigma = outlier_nsigma
        self.windows_time = windows_time
        windows_freq = np.ceil(
            np.array(windows_freq, dtype=np.float32) / average_freq)
        self.windows_freq = np.unique(windows_freq.astype(np.int_))
        self.background_reject = background_reject
        self.background_iterations = background_iterations
        self.spike_width_time = spike_width_time
        self.spike_width_freq = spike_width_freq / average_freq
        self.time_extend = _as_min_dtype(time_extend)
        self.freq_extend = _as_min_dtype(freq_extend)
        self.freq_chunks = freq_chunks
        self.average_freq = _as_min_dtype(average_freq)
        self.flag_all_time_frac = flag_all_time_frac
        self.flag_all_freq_frac = flag_all_freq_frac
        self.rho = rho

    def get_flags(self, data, flags, pool=None, chunk_size=None,
                  is_multiprocess=None):
        """Flags for ``data`` of shape (corrprod, time, frequency); ``pool`` and
        ``chunk_size`` are accepted for compatibility (the device processes all
        planes of a call together)."""
        if tuple(data.shape) != tuple(flags.shape):
            raise ValueError('Shape mismatch')
        if data.ndim != 3:
            raise ValueError('data has wrong number of dimensions')
        ncorrprod, ntime, nchan = data.shape
        averaged_channels = ((int(nchan) + int(self.average_freq) - 1) //
                             int(self.average_freq))
        freq_chunk_ends = np.linspace(
            0, averaged_channels, self.freq_chunks + 1).astype(np.int_)
        windows_time = np.array(
            [w for w in self.windows_time if w <= ntime], np.int_)
        windows_freq = np.array(
            [w for w in self.windows_freq if w <= averaged_channels], np.int_)
        plan = _StPlan(self.outlier_nsigma, windows_time, windows_freq,
                       self.background_reject, self.background_iterations,
                       self.spike_width_time, self.spike_width_freq,
                       self.time_extend, self.freq_extend, freq_chunk_ends,
                       self.average_freq, self.flag_all_time_frac,
                       self.flag_all_freq_frac, self.rho, 1)
        out, _ = _run_sum_threshold(plan, data, flags)
        if _cabi.is_device_array(out):
            return out.view(_torch().bool)
        return out.view(np.bool_)


# ---------------------------------------------------------------------------
# U1 uvcontsub_flagger (tricolour/flagging.py:989-1073)
# ---------------------------------------------------------------------------
def uvcontsub_flagger(vis, flags, major_cycles=5,
                      or_original_from_cycle=1, taylor_degrees=20,
                      sigma=5):
    """Iteratively fits a smooth spectrum (first ``taylor_degrees`` Fourier
    components of the time-averaged unflagged visibilities), subtracts it and
    clips the residual amplitudes at ``sigma`` times their (unscaled) MAD.
    Cycles before ``or_original_from_cycle`` replace the flags, later cycles OR
    into them.  ``flags`` are treated as booleans."""
    if tuple(vis.shape) != tuple(flags.shape):
        raise ValueError("vis and flags must have the same shape")
    nbl, ncorr, ntime, nfreq = vis.shape
    v = _as_c64(vis)
    fl, restore = _as_u8_flags(flags)
    ctx, space = context_for(v, fl)
    out = _empty_u8_like(fl)
    check(_cabi.load().tc_uvcontsub(ctx.handle, ptr(v), ptr(fl), int(nbl * ncorr), int(ntime),
                                    int(nfreq), int(major_cycles), int(or_original_from_cycle),
                                    int(taylor_degrees), float(sigma), ptr(out), space))
    return restore(out)


# ---------------------------------------------------------------------------
# stage-level access (mirrors the reference's private numba kernels so that the
# parity tests can follow tricolour/tests/test_flagging.py one to one)
# ---------------------------------------------------------------------------
def _planes(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.ndim == 2:
        return a[None], True
    return a, False


def _average_freq(in_data, in_flags, factor):
    """tricolour/flagging.py:819-875 -> (float32 data, bool flags)"""
    if tuple(in_data.shape) != tuple(in_flags.shape):
        raise ValueError('shape mismatch')
    vis, kind = _vis_for_st(in_data, factor)
    fl, _ = _as_u8_flags(in_flags)
    ncp, T, F = vis.shape
    Fa = (F + int(factor) - 1) // int(factor)
    od = np.empty((ncp, T, Fa), np.float32)
    of = np.empty((ncp, T, Fa), np.uint8)
    ctx, space = context_for(vis)
    check(_cabi.load().tc_stage_average_freq(ctx.handle, ptr(vis), kind, ptr(fl), ncp, T, F,
                                             int(factor), ptr(od), ptr(of), space))
    return od, of.view(np.bool_)


def _time_median(data, flags):
    """tricolour/flagging.py:226-264"""
    d, single = _planes(data, np.float32)
    fl, _ = _as_u8_flags(flags)
    fl = fl.reshape(d.shape)
    ncp, T, F = d.shape
    od = np.empty((ncp, 1, F), np.float32)
    of = np.empty((ncp, 1, F), np.uint8)
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_time_median(ctx.handle, ptr(d), ptr(fl), ncp, T, F, ptr(od),
                                            ptr(of), space))
    of = of.view(np.bool_)
    return (od[0], of[0]) if single else (od, of)


def _median_abs(data, flags, chunk_ends=None):
    """tricolour/flagging.py:267-279 per frequency chunk -> float64"""
    d, single = _planes(data, np.float32)
    fl, _ = _as_u8_flags(flags)
    fl = fl.reshape(d.shape)
    ncp, T, F = d.shape
    ce = _i64([0, F] if chunk_ends is None else chunk_ends)
    out = np.empty((ncp, ce.size - 1), np.float64)
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_chunk_median_abs(ctx.handle, ptr(d), ptr(fl), ncp, T, F, _hp(ce),
                                                 ce.size, ptr(out), space))
    if chunk_ends is None:
        return out[0, 0] if single else out[:, 0]
    return out[0] if single else out


def masked_gaussian_filter(data, flags, sigma, out, passes=4):
    """tricolour/flagging.py:469-513 (passes must be 4)"""
    if tuple(data.shape) != tuple(flags.shape):
        raise ValueError('shape mismatch between data and flags')
    if tuple(data.shape) != tuple(out.shape):
        raise ValueError('shape mismatch between data and out')
    if passes != 4:
        raise ValueError('only passes=4 is supported')
    if len(sigma) != 2:
        raise ValueError('sigma has wrong number of elements')
    r = _box_radii(sigma, passes)
    d, _ = _planes(data, np.float32)
    fl, _ = _as_u8_flags(flags)
    fl = fl.reshape(d.shape)
    ncp, T, F = d.shape
    o = np.empty_like(d)
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_masked_filter(ctx.handle, ptr(d), ptr(fl), ncp, T, F, int(r[0]),
                                              int(r[1]), ptr(o), space))
    out[...] = o.reshape(out.shape)


def _linearly_interpolate_nans(data):
    """tricolour/flagging.py:347-359, in place on a float32 array"""
    d, _ = _planes(data, np.float32)
    ncp, T, F = d.shape
    o = np.empty_like(d)
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_interp_nans(ctx.handle, ptr(d), ncp, T, F, ptr(o), space))
    data[...] = o.reshape(data.shape)


def _get_background2d(data, flags, iterations, spike_width, reject_threshold,
                      freq_chunk_ends):
    """tricolour/flagging.py:516-579"""
    d, single = _planes(data, np.float32)
    fl, _ = _as_u8_flags(flags)
    fl = fl.reshape(d.shape)
    ncp, T, F = d.shape
    radii = _background_radii(iterations, spike_width)
    ce = _i64(freq_chunk_ends)
    o = np.empty_like(d)
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_background2d(ctx.handle, ptr(d), ptr(fl), ncp, T, F, int(iterations),
                                             _hp(radii), float(reject_threshold), _hp(ce), ce.size,
                                             ptr(o), space))
    return o[0] if single else o


def _sum_threshold(input_data, input_flags, axis, windows, outlier_nsigma, rho,
                   chunks=None):
    """tricolour/flagging.py:684-742"""
    d, single = _planes(input_data, np.float32)
    if axis < 0 or axis >= 2:
        raise ValueError('axis is out of range' if axis < 0 or axis >= input_data.ndim
                         else 'axis must be 0 or 1')
    fl, _ = _as_u8_flags(input_flags)
    fl = fl.reshape(d.shape)
    ncp, T, F = d.shape
    w = _i64(windows)
    tf = np.array([math.pow(float(rho), math.log2(int(x))) if x > 0 else 1.0 for x in w], np.float64)
    sc = np.array([np.float32(1.0 / int(x)) if x > 0 else 0 for x in w], np.float32)
    o = np.empty(d.shape, np.uint8)
    if chunks is None:
        cp, nc = None, 0
    else:
        ce = _i64(chunks)
        cp, nc = _hp(ce), ce.size
    ctx, space = context_for(d)
    check(_cabi.load().tc_stage_sum_threshold(ctx.handle, ptr(d), ptr(fl), ncp, T, F, int(axis),
                                              _hp(w), _hp(tf), _hp(sc), w.size, float(outlier_nsigma),
                                              cp, nc, ptr(o), space))
    o = o.view(np.bool_)
    return o[0] if single else o


def _combine_and_unaverage(spec_flags, time_flags, freq_flags, time_extend,
                           freq_extend, average_freq, flag_all_time_frac,
                           flag_all_freq_frac, nchan):
    """_combine_flags followed by _unaverage_freq (flagging.py:784-816, 878-918)"""
    t, single = _planes(np.asarray(time_flags) != 0, np.uint8)
    f, _ = _planes(np.asarray(freq_flags) != 0, np.uint8)
    ncp, T, Fa = t.shape
    s = np.ascontiguousarray(np.asarray(spec_flags) != 0, dtype=np.uint8).reshape(ncp, Fa)
    o = np.empty((ncp, T, int(nchan)), np.uint8)
    ctx, space = context_for(t)
    check(_cabi.load().tc_stage_combine_unaverage(ctx.handle, ptr(s), ptr(t), ptr(f), ncp, T, Fa,
                                                  int(nchan), int(time_extend), int(freq_extend),
                                                  int(average_freq), float(flag_all_time_frac),
                                                  float(flag_all_freq_frac), ptr(o), space))
    o = o.view(np.bool_)
    return o[0] if single else o
