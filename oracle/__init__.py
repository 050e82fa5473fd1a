# -*- coding: utf-8 -*-
"""
CPU parity oracle for the tricolour flagging hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tricolour_b200/`` imports this
package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

The heavy lifting is ``tricolour_oracle.c`` (a sequential C restatement of
the reference's numba kernels, built by ``oracle/Makefile``); this module is
its ctypes binding plus numpy restatements of the reference's pure-numpy
functions.  Function names follow the reference
(``/root/reference/tricolour/{flagging,stokes,packing,window_statistics}.py``)
so that tests read like the reference's own tests.

Parity status: PINNED (see tests/test_oracle_golden.py and
tests/golden/make_golden.py).
"""
import ctypes
import math
import os
import re
import subprocess
from multiprocessing.pool import ThreadPool

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtricolour_oracle.so")
_lib = None

MAD_NORMAL = 1.4826

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_u8p = ctypes.POINTER(ctypes.c_uint8)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_f64p = ctypes.POINTER(ctypes.c_double)
_c_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_vp = ctypes.c_void_p


def build(force=False):
    """Compile the C oracle (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "tricolour_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_median.restype = _dbl
        L.orc_median.argtypes = [_c_f32p, _i64]
        L.orc_median_abs.restype = _dbl
        L.orc_median_abs.argtypes = [_c_f32p, _c_u8p, _i64, _i64, _i64, _i64]
        L.orc_params_new.restype = _vp
        L.orc_params_new.argtypes = [
            _dbl, _int, _c_i64p, _c_f64p, _int, _c_i64p, _c_f64p, _dbl, _int,
            _c_i64p, _c_i64p, _i64, _int, _i64, _int, _c_i64p, _i64, _int,
            _dbl, _dbl]
        L.orc_params_free.argtypes = [_vp]
        L.orc_flag_nans_zeros.argtypes = [_c_f32p, _c_u8p, _c_u8p, _i64]
        L.orc_average_freq.argtypes = [_c_f32p, _int, _c_u8p, _i64, _i64, _i64,
                                       _i64, _int, _c_f32p, _c_u8p]
        L.orc_time_median.argtypes = [_c_f32p, _c_u8p, _i64, _i64, _c_f32p, _c_u8p]
        L.orc_median_abs_axis0.argtypes = [_c_f32p, _c_u8p, _i64, _i64, _c_f32p]
        L.orc_interp_nans.argtypes = [_c_f32p, _i64, _i64]
        L.orc_box_filter1d.argtypes = [_c_f32p, _i64, _i64, _int, _c_f32p]
        L.orc_box_gaussian_filter.argtypes = [_c_f32p, _i64, _i64, _i64, _i64,
                                              _int, _c_f32p]
        L.orc_masked_gaussian_filter.argtypes = [_c_f32p, _c_u8p, _i64, _i64,
                                                 _i64, _i64, _int, _c_f32p]
        L.orc_get_background2d.argtypes = [_c_f32p, _c_u8p, _i64, _i64, _int,
                                           _c_i64p, _dbl, _c_i64p, _int, _c_f32p]
        L.orc_sum_threshold.argtypes = [_c_f32p, _c_u8p, _i64, _i64, _int,
                                        _c_i64p, _c_f64p, _int, _dbl, _c_i64p,
                                        _int, _c_u8p]
        L.orc_combine_flags.argtypes = [_c_u8p, _c_u8p, _c_u8p, _i64, _i64, _i64,
                                        _int, _c_u8p]
        L.orc_unaverage_freq.argtypes = [_c_u8p, _i64, _i64, _i64, _i64, _i64,
                                         _dbl, _dbl, _c_u8p]
        L.orc_get_flags_impl.argtypes = [_c_f32p, _int, _c_u8p, _i64, _i64, _i64,
                                         _vp, _c_u8p]
        L.orc_polarised_intensity.argtypes = [_c_f32p, _i64, _i64, _c_i64p,
                                              _c_f64p, _int, _c_f32p]
        L.orc_unpolarised_intensity.argtypes = [_c_f32p, _i64, _i64, _c_i64p,
                                                _c_f64p, _int, _c_i64p, _c_f64p,
                                                _int, _c_f32p]
        L.orc_pack.argtypes = [_c_i64p, _c_i32p, _i64, _c_i32p, _c_i32p, _i64,
                               _vp, _i64, _i64, _i64, _int, _vp]
        L.orc_unpack.argtypes = [_c_i64p, _c_i32p, _i64, _i64, _c_i32p, _c_i32p,
                                 _i64, _vp, _i64, _i64, _i64, _int, _vp]
        L.orc_window_counts.argtypes = [_c_u8p, _i64, _i64, _i64, _i64, _c_u64p,
                                        _c_u64p]
        _lib = L
    return _lib


def _p(a, typ):
    return a.ctypes.data_as(typ)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    a = np.asarray(a)
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a).view(np.uint8)
    return np.ascontiguousarray(a != 0).view(np.uint8)


def _i64a(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# ---------------------------------------------------------------------------
# F1-F3
# ---------------------------------------------------------------------------
def flag_nans_and_zeros(vis_windows, flag_windows):
    """tricolour/flagging.py:29-62"""
    if vis_windows.shape != flag_windows.shape:
        raise ValueError("vis_windows.shape != flag_windows.shape")
    vis = np.ascontiguousarray(vis_windows, dtype=np.complex64)
    out = np.zeros(flag_windows.shape, np.uint8)
    lib().orc_flag_nans_zeros(_p(vis.view(np.float32), _c_f32p),
                              _p(_u8(flag_windows), _c_u8p), _p(out, _c_u8p),
                              vis.size)
    return out.astype(flag_windows.dtype)


def flag_autos(flags, ubl):
    """tricolour/flagging.py:65-95 (``ubl`` is list-wrapped, line 84)"""
    ubl = ubl[0]
    if flags.shape[0] != ubl.shape[0]:
        raise ValueError("flag and ubl shape mismatch %s != %s"
                         % (flags.shape[2], ubl.shape[0]))
    out = flags.copy()
    out[ubl[:, 1] == ubl[:, 2], :, :, :] = True
    return out


def casa_style_range(val):
    """tricolour/util.py:78-95"""
    if not isinstance(val, str):
        raise ValueError("Value must be a string")
    if val.strip() == "" or val.strip() == "*":
        return (0, np.inf)
    elif re.match(r"^(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?~"
                  r"(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?[\s]*[m]?$", val):
        val = val.replace(" ", "").replace("\t", "").replace("m", "")
        return list(map(float, val.split("~")))
    raise ValueError("Value must be range or blank")


def apply_static_mask(flag, ubl, antspos, masks, chan_freqs, chan_widths,
                      accumulation_mode="or", uvrange=""):
    """tricolour/flagging.py:98-172"""
    uvrange = casa_style_range(uvrange)
    if flag.shape[0] != ubl.shape[0]:
        raise ValueError("flag and ubl shape mismatch %s != %s"
                         % (flag.shape[1], ubl.shape[0]))
    lb = chan_freqs - chan_widths * 0.5
    ub = chan_freqs + chan_widths * 0.5
    bl_length = antspos[ubl[:, 1]] - antspos[ubl[:, 2]]
    d2 = 0.5 * np.sum(bl_length ** 2, axis=1)
    lo = min(uvrange[0], uvrange[1])
    hi = max(uvrange[0], uvrange[1])
    bl_sel = np.logical_and(d2 >= lo ** 2, d2 <= hi ** 2)
    out = flag.copy()
    for mask in masks:
        if mask.ndim != 2 and mask.shape[1] != 1:
            raise ValueError("masks.shape != (dim, 1)")
        mc = np.logical_and(mask >= lb[None, :], mask < ub[None, :]).sum(axis=0) > 0
        if accumulation_mode == "or":
            out[bl_sel, :, :, :] |= mc[None, None, None, :]
        elif accumulation_mode == "override":
            out[bl_sel, :, :, :] = mc[None, None, None, :]
        else:
            raise ValueError("Invalid accumulation_mode '%s'. Should be 'or' "
                             "or 'override'" % accumulation_mode)
    return out


# ---------------------------------------------------------------------------
# SumThreshold stages (S0-S11)
# ---------------------------------------------------------------------------
def _as_min_dtype(value):
    """tricolour/flagging.py:175-190"""
    if 0 <= value < 2 ** 8:
        dtype = np.uint8
    elif 0 <= value < 2 ** 16:
        dtype = np.uint16
    elif 0 <= value < 2 ** 32:
        dtype = np.uint32
    else:
        dtype = np.int64
    return np.array(value, dtype)


def _bits(v):
    return int(np.asarray(v).dtype.itemsize) * 8


def _split_data(in_data):
    in_data = np.asarray(in_data)
    if np.iscomplexobj(in_data):
        d = np.ascontiguousarray(in_data, dtype=np.complex64).view(np.float32)
        return d, 1
    return _f32(in_data), 0


def _average_freq(in_data, in_flags, factor):
    """tricolour/flagging.py:819-875; ``factor`` is a 0-d array (dtype matters)"""
    if in_data.shape != in_flags.shape:
        raise ValueError('shape mismatch')
    ncp, T, F = in_data.shape
    fac = int(factor)
    Fa = (F + fac - 1) // fac
    d, cplx = _split_data(in_data)
    out = np.empty((ncp, T, Fa), np.float32)
    oflags = np.empty((ncp, T, Fa), np.uint8)
    lib().orc_average_freq(_p(d, _c_f32p), cplx, _p(_u8(in_flags), _c_u8p), ncp,
                           T, F, fac, _bits(factor), _p(out, _c_f32p),
                           _p(oflags, _c_u8p))
    return out, oflags.view(np.bool_)


def _time_median(data, flags):
    """tricolour/flagging.py:226-264"""
    T, F = data.shape
    d = _f32(data)
    out = np.empty((1, F), np.float32)
    of = np.empty((1, F), np.uint8)
    lib().orc_time_median(_p(d, _c_f32p), _p(_u8(flags), _c_u8p), T, F,
                          _p(out, _c_f32p), _p(of, _c_u8p))
    return out, of.view(np.bool_)


def _median_abs(data, flags):
    """tricolour/flagging.py:267-279"""
    d = _f32(data)
    d2 = d.reshape(-1, d.shape[-1]) if d.ndim > 1 else d.reshape(1, -1)
    T, F = d2.shape
    return lib().orc_median_abs(_p(d2, _c_f32p), _p(_u8(flags), _c_u8p), T, F, 0, F)


def _median_abs_axis0(data, flags):
    """tricolour/flagging.py:282-304 (2-D form)"""
    d = _f32(data)
    n, m = d.shape
    out = np.empty((1, m), np.float32)
    lib().orc_median_abs_axis0(_p(d, _c_f32p), _p(_u8(flags), _c_u8p), n, m,
                               _p(out, _c_f32p))
    return out


def _linearly_interpolate_nans(data):
    """tricolour/flagging.py:347-359 (float32, in place)"""
    assert data.dtype == np.float32 and data.flags.c_contiguous
    T, F = data.shape if data.ndim == 2 else (1, data.shape[0])
    lib().orc_interp_nans(_p(data, _c_f32p), T, F)


def _box_gaussian_filter1d(data, r, out, passes):
    """tricolour/flagging.py:362-419 for a 1-D float32 array"""
    d = _f32(data)
    o = np.empty_like(d)
    lib().orc_box_filter1d(_p(d, _c_f32p), d.shape[0], int(r), int(passes),
                           _p(o, _c_f32p))
    out[...] = o


def box_radii(sigma, passes=4):
    """tricolour/flagging.py:451"""
    sigma = np.asarray(sigma, dtype=np.float64)
    return (0.5 * np.sqrt(12.0 * sigma ** 2 / passes + 1)).astype(np.int_)


def _box_gaussian_filter(data, sigma, out, passes=4):
    """tricolour/flagging.py:422-466"""
    if len(sigma) != data.ndim:
        raise ValueError('sigma has wrong number of elements')
    r = box_radii(sigma, passes)
    d = _f32(data)
    o = np.empty_like(d)
    T, F = d.shape
    lib().orc_box_gaussian_filter(_p(d, _c_f32p), T, F, int(r[0]), int(r[1]),
                                  int(passes), _p(o, _c_f32p))
    out[...] = o


def masked_gaussian_filter(data, flags, sigma, out, passes=4):
    """tricolour/flagging.py:469-513"""
    if data.shape != flags.shape:
        raise ValueError('shape mismatch between data and flags')
    if data.shape != out.shape:
        raise ValueError('shape mismatch between data and out')
    r = box_radii(sigma, passes)
    d = _f32(data)
    o = np.empty_like(d)
    T, F = d.shape
    lib().orc_masked_gaussian_filter(_p(d, _c_f32p), _p(_u8(flags), _c_u8p), T, F,
                                     int(r[0]), int(r[1]), int(passes),
                                     _p(o, _c_f32p))
    out[...] = o


def background_radii(iterations, spike_width):
    """radii per extend_factor (flagging.py:553-555) plus the final filter (576)"""
    sw = np.asarray(spike_width, dtype=np.float64)
    rows = [box_radii(ef * sw) for ef in range(iterations, 0, -1)]
    rows.append(box_radii(sw))
    return np.ascontiguousarray(np.array(rows, dtype=np.int64).reshape(-1, 2))


def _get_background2d(data, flags, iterations, spike_width, reject_threshold,
                      freq_chunk_ends):
    """tricolour/flagging.py:516-579"""
    d = _f32(data)
    T, F = d.shape
    radii = background_radii(iterations, spike_width)
    ce = _i64a(freq_chunk_ends)
    out = np.empty_like(d)
    lib().orc_get_background2d(_p(d, _c_f32p), _p(_u8(flags), _c_u8p), T, F,
                               int(iterations), _p(radii, _c_i64p),
                               float(reject_threshold), _p(ce, _c_i64p),
                               ce.size, _p(out, _c_f32p))
    return out


def threshold_factors(windows, rho):
    """tf = pow(rho, log2(window)) -- flagging.py:641 (glibc pow/log2, float64)"""
    return np.array([math.pow(float(rho), math.log2(int(w))) for w in windows],
                    dtype=np.float64)


def _sum_threshold(input_data, input_flags, axis, windows, outlier_nsigma, rho,
                   chunks=None):
    """tricolour/flagging.py:684-742"""
    d = _f32(input_data)
    T, F = d.shape
    if axis < 0 or axis >= d.ndim:
        raise ValueError('axis is out of range')
    w = _i64a(windows)
    tf = threshold_factors(w, rho)
    out = np.empty((T, F), np.uint8)
    if chunks is None:
        cp, nc = None, 0
    else:
        ce = _i64a(chunks)
        cp, nc = _p(ce, _c_i64p), ce.size
    lib().orc_sum_threshold(_p(d, _c_f32p), _p(_u8(input_flags), _c_u8p), T, F,
                            int(axis), _p(w, _c_i64p), _p(tf, _c_f64p), w.size,
                            float(outlier_nsigma), cp, nc, _p(out, _c_u8p))
    return out.view(np.bool_)


def _combine_flags(spec_flags, time_flags, freq_flags, time_extend, out):
    """tricolour/flagging.py:784-816; ``time_extend`` is a 0-d array"""
    T, F = time_flags.shape
    o = np.empty((T, F), np.uint8)
    lib().orc_combine_flags(_p(_u8(spec_flags), _c_u8p), _p(_u8(time_flags), _c_u8p),
                            _p(_u8(freq_flags), _c_u8p), T, F, int(time_extend),
                            _bits(time_extend), _p(o, _c_u8p))
    out[...] = o


def _unaverage_freq(flags, freq_extend, average_freq, flag_all_time_frac,
                    flag_all_freq_frac, out):
    """tricolour/flagging.py:878-918"""
    T, Fa = flags.shape
    F = out.shape[-1]
    o = np.empty((T, F), np.uint8)
    lib().orc_unaverage_freq(_p(_u8(flags), _c_u8p), T, Fa, F, int(freq_extend),
                             int(average_freq), float(flag_all_time_frac),
                             float(flag_all_freq_frac), _p(o, _c_u8p))
    out[...] = o


# ---------------------------------------------------------------------------
# S12-S14
# ---------------------------------------------------------------------------
class _Params(object):
    """Owns an ``orc_params`` block built from already-conditioned values."""

    def __init__(self, outlier_nsigma, windows_time, windows_freq,
                 background_reject, background_iterations, spike_width_time,
                 spike_width_freq, time_extend, freq_extend, freq_chunk_ends,
                 average_freq, flag_all_time_frac, flag_all_freq_frac, rho):
        wt = _i64a(windows_time)
        wf = _i64a(windows_freq)
        tft = threshold_factors(wt, rho)
        tff = threshold_factors(wf, rho)
        rs = background_radii(background_iterations, (0.0, spike_width_freq))
        r2 = background_radii(background_iterations,
                              (spike_width_time, spike_width_freq))
        ce = _i64a(freq_chunk_ends)
        self.average_freq = int(average_freq)
        self._h = lib().orc_params_new(
            float(outlier_nsigma), wt.size, _p(wt, _c_i64p), _p(tft, _c_f64p),
            wf.size, _p(wf, _c_i64p), _p(tff, _c_f64p), float(background_reject),
            int(background_iterations), _p(rs, _c_i64p), _p(r2, _c_i64p),
            int(time_extend), _bits(time_extend), int(freq_extend), ce.size,
            _p(ce, _c_i64p), int(average_freq), _bits(average_freq),
            float(flag_all_time_frac), float(flag_all_freq_frac))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_params_free(self._h)
            self._h = None


def _get_flags_impl(in_data, in_flags, out_flags, params, nthreads=1):
    """tricolour/flagging.py:745-781 on (cp, T, F) arrays; threads over planes
    like the reference's ThreadPool over baseline blocks (app.py:266-271)."""
    d, cplx = _split_data(in_data)
    ncp, T, F = in_flags.shape
    fl = _u8(in_flags)
    out = np.empty((ncp, T, F), np.uint8)
    L = lib()
    esz = 2 if cplx else 1

    def run(lo, hi):
        if hi <= lo:
            return
        L.orc_get_flags_impl(
            ctypes.cast(d.ctypes.data + lo * T * F * esz * 4, _c_f32p), cplx,
            ctypes.cast(fl.ctypes.data + lo * T * F, _c_u8p), hi - lo, T, F,
            params._h, ctypes.cast(out.ctypes.data + lo * T * F, _c_u8p))

    if nthreads <= 1 or ncp <= 1:
        run(0, ncp)
    else:
        with ThreadPool(nthreads) as pool:
            pool.starmap(run, [(i, i + 1) for i in range(ncp)])
    out_flags[...] = out


def condition_sum_threshold_params(nchan, ntime, windows_time, windows_freq,
                                   freq_chunks, average_freq, time_extend,
                                   freq_extend):
    """tricolour/flagging.py:1160-1179"""
    windows_freq = np.asarray(windows_freq, dtype=np.float32)
    windows_freq = np.ceil(windows_freq) / average_freq
    windows_freq = np.unique(windows_freq.astype(np.int_))
    time_extend = _as_min_dtype(time_extend)
    freq_extend = _as_min_dtype(freq_extend)
    average_freq = _as_min_dtype(average_freq)
    averaged_channels = (int(nchan) + int(average_freq) - 1) // int(average_freq)
    freq_chunk_ends = np.linspace(0, averaged_channels,
                                  freq_chunks + 1).astype(np.int_)
    windows_time = np.array([w for w in windows_time if w <= ntime], np.int_)
    windows_freq = np.array([w for w in windows_freq if w <= averaged_channels],
                            np.int_)
    return (windows_time, windows_freq, freq_chunk_ends, average_freq,
            time_extend, freq_extend)


def sum_threshold_flagger(vis, flags, outlier_nsigma=4.5,
                          windows_time=[1, 2, 4, 8], windows_freq=[1, 2, 4, 8],
                          background_reject=2.0, background_iterations=1,
                          spike_width_time=12.5, spike_width_freq=10.0,
                          time_extend=3, freq_extend=3,
                          freq_chunks=10, average_freq=1,
                          flag_all_time_frac=0.6, flag_all_freq_frac=0.8,
                          rho=1.3, num_major_iterations=5, nthreads=1):
    """tricolour/flagging.py:1076-1196"""
    nbl, ncorr, ntime, nchan = vis.shape
    vis3 = vis.reshape(nbl * ncorr, ntime, nchan)
    flags3 = flags.reshape(nbl * ncorr, ntime, nchan)
    (wt, wf, ce, af, te, fe) = condition_sum_threshold_params(
        nchan, ntime, windows_time, windows_freq, freq_chunks, average_freq,
        time_extend, freq_extend)
    params = _Params(outlier_nsigma, wt, wf, background_reject,
                     background_iterations, spike_width_time, spike_width_freq,
                     te, fe, ce, af, flag_all_time_frac, flag_all_freq_frac, rho)
    out_flags = np.empty_like(flags3)
    iter_flags = flags3.copy()
    for _ in range(num_major_iterations):
        _get_flags_impl(vis3, iter_flags, out_flags, params, nthreads)
        iter_flags = np.logical_or(iter_flags, out_flags)
    return out_flags.reshape(nbl, ncorr, ntime, nchan)


class SumThresholdFlagger(object):
    """tricolour/flagging.py:1199-1423 (parameter conditioning of the class,
    lines 1274-1289 and 1304-1327, which differs from the function's)"""

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
        wf = np.ceil(np.array(windows_freq, dtype=np.float32) / average_freq)
        self.windows_freq = np.unique(wf.astype(np.int_))
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
        if data.shape != flags.shape:
            raise ValueError('Shape mismatch')
        if data.ndim != 3:
            raise ValueError('data has wrong number of dimensions')
        ncp, ntime, nchan = data.shape
        ac = (int(nchan) + int(self.average_freq) - 1) // int(self.average_freq)
        ce = np.linspace(0, ac, self.freq_chunks + 1).astype(np.int_)
        wt = np.array([w for w in self.windows_time if w <= ntime], np.int_)
        wf = np.array([w for w in self.windows_freq if w <= ac], np.int_)
        params = _Params(self.outlier_nsigma, wt, wf, self.background_reject,
                         self.background_iterations, self.spike_width_time,
                         self.spike_width_freq, self.time_extend,
                         self.freq_extend, ce, self.average_freq,
                         self.flag_all_time_frac, self.flag_all_freq_frac,
                         self.rho)
        out = np.empty(flags.shape, np.bool_)
        _get_flags_impl(data, flags, out, params)
        return out


# ---------------------------------------------------------------------------
# U1 uvcontsub_flagger -- tricolour/flagging.py:989-1073 (pure numpy in the
# reference; restated per plane, same numpy calls in the same order)
# ---------------------------------------------------------------------------
def uvcontsub_flagger(vis, flags, major_cycles=5, or_original_from_cycle=1,
                      taylor_degrees=20, sigma=5):
    if vis.shape != flags.shape:
        raise ValueError("vis and flags must have the same shape")
    nbl, ncorr, ntime, nfreq = vis.shape
    vis = vis.reshape(nbl * ncorr, ntime, nfreq)
    result_flags = flags.reshape(nbl * ncorr, ntime, nfreq).copy()
    import warnings
    for mi in range(major_cycles):
        for cp in range(vis.shape[0]):
            rf = result_flags[cp]
            if rf.sum() == rf.size:
                continue
            scratch = vis[cp].copy()
            scratch[rf] = np.nan
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                avgvis = np.nanmean(scratch, axis=0)
            avgvis[np.isnan(avgvis)] = 0.0
            fft = np.fft.fft(avgvis, axis=0)
            fft[np.arange(taylor_degrees, fft.shape[0])] = 0
            smoothened = np.fft.ifft(fft)
            absresidual = np.abs(vis[cp] - smoothened[None, :]).real
            fa = absresidual.copy()
            fa[rf] = np.nan
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                diff = np.abs(np.abs(fa) - np.nanmedian(np.abs(fa)))
                mad = np.nanmedian(np.abs(diff))
            newflags = absresidual > sigma * mad
            if mi >= or_original_from_cycle:
                result_flags[cp] = np.logical_or(rf, newflags)
            else:
                result_flags[cp] = newflags
    return result_flags.reshape(nbl, ncorr, ntime, nfreq)


# ---------------------------------------------------------------------------
# K1/K2 stokes -- tricolour/stokes.py
# ---------------------------------------------------------------------------
STOKES_TYPES = {'I': 1, 'Q': 2, 'U': 3, 'V': 4, 'RR': 5, 'RL': 6, 'LR': 7,
                'LL': 8, 'XX': 9, 'XY': 10, 'YX': 11, 'YY': 12}

_STOKES_DEPS = {
    'I': [('XX', 'YY', 0.5 + 0.0j, 1, 1), ('RR', 'LL', 0.5 + 0.0j, 1, 1)],
    'Q': [('XX', 'YY', 0.5 + 0.0j, 1, -1), ('RL', 'LR', 0.5 + 0.0j, 1, 1)],
    'U': [('XY', 'YX', 0.5 + 0.0j, 1, 1), ('RL', 'LR', 0.0 - 0.5j, 1, -1)],
    'V': [('XY', 'YX', 0.0 - 0.5j, 1, -1), ('RR', 'LL', 0.5 + 0.0j, 1, -1)]}


def stokes_corr_map(corr_types):
    """tricolour/stokes.py:42-76"""
    have = set(corr_types)
    out = {}
    for stokes, deps in _STOKES_DEPS.items():
        for (c1, c2, a, s1, s2) in deps:
            n1, n2 = STOKES_TYPES[c1], STOKES_TYPES[c2]
            if n1 in have and n2 in have:
                out[stokes] = (corr_types.index(n1), corr_types.index(n2), a, s1, s2)
    return out


def _terms(stokes):
    idx = np.array([[t[0], t[1]] for t in stokes], np.int64).reshape(-1, 2)
    coef = np.array([[complex(t[2]).real, complex(t[2]).imag, t[3], t[4]]
                     for t in stokes], np.float64).reshape(-1, 4)
    return np.ascontiguousarray(idx), np.ascontiguousarray(coef)


def polarised_intensity(vis, stokes_pol):
    """tricolour/stokes.py:157-209 (complex64 input)"""
    v = np.ascontiguousarray(vis, dtype=np.complex64)
    nrow, nchan, ncorr = v.shape
    out = np.empty((nrow, nchan, 1), np.complex64)
    idx, coef = _terms(stokes_pol)
    lib().orc_polarised_intensity(_p(v.view(np.float32), _c_f32p), nrow * nchan,
                                  ncorr, _p(idx, _c_i64p), _p(coef, _c_f64p),
                                  idx.shape[0], _p(out.view(np.float32), _c_f32p))
    return out


def unpolarised_intensity(vis, stokes_unpol, stokes_pol):
    """tricolour/stokes.py:79-154 (complex64 input)"""
    if not len(stokes_unpol) == 1:
        raise ValueError("There should be exactly one entry "
                         "for unpolarised stokes (stokes_unpol)")
    if not len(stokes_pol) > 0:
        raise ValueError("No entries for polarised stokes (stokes_pol)")
    v = np.ascontiguousarray(vis, dtype=np.complex64)
    nrow, nchan, ncorr = v.shape
    out = np.empty((nrow, nchan, 1), np.complex64)
    ui, uc = _terms(stokes_unpol)
    pi, pc = _terms(stokes_pol)
    lib().orc_unpolarised_intensity(_p(v.view(np.float32), _c_f32p), nrow * nchan,
                                    ncorr, _p(ui, _c_i64p), _p(uc, _c_f64p),
                                    ui.shape[0], _p(pi, _c_i64p), _p(pc, _c_f64p),
                                    pi.shape[0], _p(out.view(np.float32), _c_f32p))
    return out


# ---------------------------------------------------------------------------
# P1/P2 pack / unpack numeric kernels -- tricolour/packing.py:243-278, 369-415
# ---------------------------------------------------------------------------
def pack_data(time_inv, ubl, antenna1, antenna2, data, flags, ntime):
    """Window creation (packing.py:96-98, 116-117: NaN+NaNj / 1 defaults) plus
    _numba_pack_data for one row block covering all baselines of ``ubl``."""
    nrow, nchan, ncorr = data.shape
    nbl = ubl.shape[0]
    vis_win = np.full((nbl, ncorr, ntime, nchan), np.nan + np.nan * 1j, data.dtype)
    flag_win = np.full((nbl, ncorr, ntime, nchan), 1, flags.dtype)
    ti = _i64a(time_inv)
    u = np.ascontiguousarray(ubl, dtype=np.int32)
    a1 = np.ascontiguousarray(antenna1, dtype=np.int32)
    a2 = np.ascontiguousarray(antenna2, dtype=np.int32)
    for arr, win in ((np.ascontiguousarray(data), vis_win),
                     (np.ascontiguousarray(flags), flag_win)):
        lib().orc_pack(_p(ti, _c_i64p), _p(u, _c_i32p), nbl, _p(a1, _c_i32p),
                       _p(a2, _c_i32p), nrow, arr.ctypes.data, nchan, ncorr,
                       ntime, arr.dtype.itemsize, win.ctypes.data)
    return vis_win, flag_win


def unpack_data(antenna1, antenna2, time_inv, ubl, windows):
    """_unpack_data for a single baseline chunk (packing.py:391-415)"""
    nbl, ncorr, ntime, nchan = windows.shape
    nrow = antenna1.shape[0]
    out = np.zeros((nrow, nchan, ncorr), windows.dtype)
    ti = _i64a(time_inv)
    u = np.ascontiguousarray(ubl, dtype=np.int32)
    a1 = np.ascontiguousarray(antenna1, dtype=np.int32)
    a2 = np.ascontiguousarray(antenna2, dtype=np.int32)
    w = np.ascontiguousarray(windows)
    lib().orc_unpack(_p(ti, _c_i64p), _p(u, _c_i32p), nbl, int(u[:, 0].min()),
                     _p(a1, _c_i32p), _p(a2, _c_i32p), nrow, w.ctypes.data, nchan,
                     ncorr, ntime, w.dtype.itemsize, out.ctypes.data)
    return out


# ---------------------------------------------------------------------------
# W1 window statistics counting -- tricolour/window_statistics.py:12-66
# ---------------------------------------------------------------------------
def window_counts(flag_window, ubl, chan_freqs, nant, nchanbins=10):
    """Returns the numbers _window_stats accumulates:
    (ant_counts[nant], ant_sizes[nant], bl_counts[nbl], bl_size, total_count,
    total_size, bin_counts[nchanbins] (through uint32), bin_edges)."""
    fw = np.ascontiguousarray(flag_window)
    nbl, ncorr, T, F = fw.shape
    blc = np.zeros(nbl, np.uint64)
    chc = np.zeros(F, np.uint64)
    lib().orc_window_counts(_p(fw.view(np.uint8) if fw.dtype.itemsize == 1
                               else fw.astype(np.uint8), _c_u8p), nbl, ncorr, T,
                            F, _p(blc, _c_u64p), _p(chc, _c_u64p))
    plane = ncorr * T * F
    antc = np.zeros(nant, np.uint64)
    ants = np.zeros(nant, np.uint64)
    for ai in range(nant):
        sel = np.logical_or(ubl[:, 1] == ai, ubl[:, 2] == ai)
        antc[ai] = blc[sel].sum(dtype=np.uint64)
        ants[ai] = int(sel.sum()) * plane
    edges = np.linspace(np.min(chan_freqs), np.max(chan_freqs), nchanbins)
    bins = np.zeros(nchanbins, np.uint32)
    for i in range(nchanbins - 1):
        sel = np.logical_and(chan_freqs >= edges[i], chan_freqs < edges[i + 1])
        bins[i] = np.uint64(chc[sel].sum(dtype=np.uint64)).astype(np.uint32)
    return (antc, ants, blc, plane, int(blc.sum(dtype=np.uint64)), nbl * plane,
            bins, edges)
