# -*- coding: utf-8 -*-
"""
MS row order <-> (bl, corr, time, chan) windows (reference:
tricolour/packing.py).

The reference wraps its numba kernels in dask graph glue; dask is not part of
the hot path, so these functions are the eager numpy-level equivalents with the
reference's signatures: ``pack_data`` builds default-filled windows (vis
``NaN+NaNj``, flag ``1``; packing.py:96-98, 116-117) and scatters the rows into
them, ``unpack_data`` gathers a window back into row order.  The scatter /
gather runs on the GPU; the host only turns (antenna1, antenna2) into window
slots, which is O(row) integer work on metadata.
"""
import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, ptr, context_for

_WINDOW_SCHEMA = ("bl", "corr", "time", "chan")


def unique_baselines(ant1, ant2):
    """
    Unique baseline pairs as 64 bit ints (packing.py:36-56).  Recast with
    ``ubl.view(np.int32).reshape(-1, 2)``; the sort order is that of the int64
    values, i.e. antenna2-major on little-endian hosts.
    """
    ant1 = np.asarray(ant1)
    ant2 = np.asarray(ant2)
    if not (ant1.dtype == np.int32 and ant2.dtype == np.int32):
        raise TypeError("antenna1 '%s' and antenna2 '%s' dtypes "
                        "must both be np.int32" % (ant1.dtype, ant2.dtype))
    bl = np.ascontiguousarray(np.stack([ant1, ant2], axis=1)).view(np.int64)
    return np.unique(bl)


def _row_slots(ubl, antenna1, antenna2, time_inv, last_wins):
    """window slot (or -1) and time index of every MS row.  With ``last_wins``
    rows that lose a (baseline, time) collision to a later row are dropped, which
    is what the reference's sequential loops amount to (packing.py:262-276)."""
    ubl = np.asarray(ubl)
    a1 = np.asarray(antenna1).astype(np.int64)
    a2 = np.asarray(antenna2).astype(np.int64)
    key = (a1 << 32) | (a2 & 0xffffffff)
    ukey = (ubl[:, 1].astype(np.int64) << 32) | (ubl[:, 2].astype(np.int64) & 0xffffffff)
    order = np.argsort(ukey, kind="stable")
    pos = np.searchsorted(ukey[order], key)
    pos = np.clip(pos, 0, max(len(order) - 1, 0))
    if len(order):
        hit = ukey[order][pos] == key
        slot = np.where(hit, ubl[order[pos], 0], -1).astype(np.int32)
    else:
        slot = np.full(key.shape, -1, np.int32)
    t = np.asarray(time_inv).astype(np.int32)
    if last_wins and slot.size:
        cell = slot.astype(np.int64) * (int(t.max()) + 1 if t.size else 1) + t
        cell = np.where(slot >= 0, cell, -1 - np.arange(slot.size, dtype=np.int64))
        # keep the last occurrence of every cell
        _, first_rev = np.unique(cell[::-1], return_index=True)
        keep = np.zeros(slot.size, np.bool_)
        keep[slot.size - 1 - first_rev] = True
        slot = np.where(keep, slot, -1).astype(np.int32)
    return np.ascontiguousarray(slot), np.ascontiguousarray(t)


def _hp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pack_data(time_inv, ubl,
              antenna1, antenna2,
              data, flags, ntime,
              backend="numpy", path=None,
              return_objs=False):
    """
    Packs ``data`` / ``flags`` of shape (row, chan, corr) into windows of shape
    (bl, corr, time, chan) (packing.py:306-366 with ``_numba_pack_data``
    243-278).  ``ubl`` is (nbl, 3) int32 ``(bl, a1, a2)``.  Only the in-memory
    backend exists here: windows stay resident instead of spilling to zarr.
    """
    if backend != "numpy":
        raise ValueError("Invalid backend '%s'" % backend)
    nrow, nchan, ncorr = (int(s) for s in data.shape)
    if tuple(flags.shape) != tuple(data.shape):
        raise ValueError("vis_windows.shape != flag_windows.shape")
    ubl = np.asarray(ubl)
    nbl = int(ubl.shape[0])
    slot, t = _row_slots(ubl, antenna1, antenna2, time_inv, last_wins=True)
    if slot.size and t[slot >= 0].size and (t[slot >= 0].min() < 0 or t[slot >= 0].max() >= ntime):
        raise ValueError("time_inv out of range")
    dev = _cabi.is_device_array(data)
    if dev != _cabi.is_device_array(flags):
        raise TypeError("tricolour_b200: data and flags must both be numpy arrays or both be CUDA tensors")
    if dev:
        import torch
        vis = data.contiguous()
        if vis.dtype != torch.complex64:
            raise TypeError("device visibilities must be complex64")
        fl = flags.contiguous()
        fdt = fl.dtype
        fl8 = fl.view(torch.uint8) if fdt in (torch.bool, torch.uint8) else (fl != 0).view(torch.uint8)
        vis_win = torch.empty((nbl, ncorr, int(ntime), nchan), dtype=torch.complex64, device=vis.device)
        flag_win = torch.empty((nbl, ncorr, int(ntime), nchan), dtype=torch.uint8, device=vis.device)
    else:
        vis_in = np.asarray(data)
        vdt = vis_in.dtype
        vis = np.ascontiguousarray(vis_in, dtype=np.complex64)
        f = np.asarray(flags)
        fdt = f.dtype
        fl8 = (np.ascontiguousarray(f).view(np.uint8) if f.dtype.itemsize == 1
               else np.ascontiguousarray(f != 0).view(np.uint8))
        vis_win = np.empty((nbl, ncorr, int(ntime), nchan), np.complex64)
        flag_win = np.empty((nbl, ncorr, int(ntime), nchan), np.uint8)
    ctx, space = context_for(vis, fl8)
    check(_cabi.load().tc_pack(ctx.handle, _hp(slot), _hp(t), nrow, ptr(vis), ptr(fl8), nchan, ncorr,
                               int(ntime), nbl, ptr(vis_win), ptr(flag_win), 1, space))
    if dev:
        import torch
        flag_win = flag_win.view(torch.bool) if fdt == torch.bool else flag_win.to(fdt)
    else:
        if vdt != np.complex64:
            vis_win = vis_win.astype(vdt)
        flag_win = flag_win.view(np.bool_) if fdt == np.bool_ else flag_win.astype(fdt)
    if return_objs:
        return vis_win, flag_win, vis_win, flag_win
    return vis_win, flag_win


def unpack_data(antenna1, antenna2, time_inv, ubl, flag_windows):
    """
    Gathers windows of shape (bl, corr, time, chan) back into (row, chan, corr)
    (packing.py:391-425).  Rows whose baseline is not in ``ubl`` stay zero.
    Works for flag windows (any 1-byte dtype) and complex64 visibility windows.
    """
    ubl = np.asarray(ubl)
    w = flag_windows
    nbl, ncorr, ntime, nchan = (int(s) for s in w.shape)
    # window slots are relative to the smallest baseline index of this chunk
    u = ubl.copy()
    if u.shape[0]:
        u[:, 0] = u[:, 0] - u[:, 0].min()
    slot, t = _row_slots(u, antenna1, antenna2, time_inv, last_wins=False)
    nrow = int(slot.size)
    dev = _cabi.is_device_array(w)
    if dev:
        import torch
        wc = w.contiguous()
        if wc.dtype == torch.complex64:
            elem, raw = 8, wc
        elif wc.dtype in (torch.bool, torch.uint8):
            elem, raw = 1, wc.view(torch.uint8)
        else:
            raise TypeError("unsupported window dtype %s" % wc.dtype)
        out = torch.empty((nrow, nchan, ncorr), dtype=raw.dtype, device=wc.device)
    else:
        wa = np.asarray(w)
        if wa.dtype == np.complex64:
            elem, raw = 8, np.ascontiguousarray(wa)
        elif wa.dtype.itemsize == 1:
            elem, raw = 1, np.ascontiguousarray(wa).view(np.uint8)
        elif np.iscomplexobj(wa):
            elem, raw = 8, np.ascontiguousarray(wa, dtype=np.complex64)
        else:
            elem, raw = 1, np.ascontiguousarray(wa != 0).view(np.uint8)
        out = np.empty((nrow, nchan, ncorr), raw.dtype)
    ctx, space = context_for(raw)
    check(_cabi.load().tc_unpack(ctx.handle, _hp(slot), _hp(t), nrow, ptr(raw), elem, nchan, ncorr,
                                 ntime, nbl, ptr(out), space))
    if dev:
        import torch
        return out.view(torch.bool) if w.dtype == torch.bool else out
    wa = np.asarray(w)
    if wa.dtype == np.bool_:
        return out.view(np.bool_)
    return out if out.dtype == wa.dtype else out.astype(wa.dtype)


def unpack_flags_equalised(antenna1, antenna2, time_inv, ubl, flag_windows, ncorr_out=None):
    """``unpack_data`` fused with the app's correlation equalisation
    (tricolour/apps/tricolour/app.py:479-480): a sample flagged in any
    correlation is flagged in all ``ncorr_out`` correlations of the output rows
    (default: the window's own count; a one-correlation window of the polarised
    strategies is broadcast to the measurement set's correlations this way)."""
    ubl = np.asarray(ubl)
    w = flag_windows
    nbl, ncorr, ntime, nchan = (int(s) for s in w.shape)
    nout = ncorr if ncorr_out is None else int(ncorr_out)
    u = ubl.copy()
    if u.shape[0]:
        u[:, 0] = u[:, 0] - u[:, 0].min()
    slot, t = _row_slots(u, antenna1, antenna2, time_inv, last_wins=False)
    nrow = int(slot.size)
    if _cabi.is_device_array(w):
        import torch
        raw = w.contiguous().view(torch.uint8)
        out = torch.empty((nrow, nchan, nout), dtype=torch.uint8, device=w.device)
    else:
        wa = np.asarray(w)
        raw = (np.ascontiguousarray(wa).view(np.uint8) if wa.dtype.itemsize == 1
               else np.ascontiguousarray(wa != 0).view(np.uint8))
        out = np.empty((nrow, nchan, nout), np.uint8)
    ctx, space = context_for(raw)
    check(_cabi.load().tc_unpack_flags_broadcast(ctx.handle, _hp(slot), _hp(t), nrow, ptr(raw), nchan,
                                                 ncorr, nout, ntime, nbl, ptr(out), space))
    if _cabi.is_device_array(w):
        import torch
        return out.view(torch.bool)
    return out.view(np.bool_)


def pack_polarised(time_inv, ubl, antenna1, antenna2, data, flags, ntime, stokes_pol, stokes_unpol=None):
    """The polarised / total-power front end of the application in one pass over
    the rows (app.py:415-432 followed by ``pack_data``): Stokes intensity of the
    correlations (``polarised_intensity``, or ``unpolarised_intensity`` when
    ``stokes_unpol`` is given), ``flags.any(axis=2)`` and the scatter into windows
    of shape (bl, 1, time, chan).  Equal to
    ``pack_data(..., polarised_intensity(data, stokes_pol), flags.any(2, keepdims=True), ntime)``."""
    from .stokes import _terms
    nrow, nchan, ncorr = (int(s) for s in data.shape)
    if tuple(flags.shape) != tuple(data.shape):
        raise ValueError("vis_windows.shape != flag_windows.shape")
    if stokes_unpol is not None and not len(stokes_unpol) == 1:
        raise ValueError("There should be exactly one entry "
                         "for unpolarised stokes (stokes_unpol)")
    if not len(stokes_pol) > 0:
        raise ValueError("No entries for polarised stokes (stokes_pol)")
    ubl = np.asarray(ubl)
    nbl = int(ubl.shape[0])
    slot, t = _row_slots(ubl, antenna1, antenna2, time_inv, last_wins=True)
    if slot.size and t[slot >= 0].size and (t[slot >= 0].min() < 0 or t[slot >= 0].max() >= ntime):
        raise ValueError("time_inv out of range")
    dev = _cabi.is_device_array(data)
    if dev != _cabi.is_device_array(flags):
        raise TypeError("tricolour_b200: data and flags must both be numpy arrays or both be CUDA tensors")
    if dev:
        import torch
        vis = data.contiguous()
        if vis.dtype != torch.complex64:
            raise TypeError("device visibilities must be complex64")
        fl = flags.contiguous()
        fdt = fl.dtype
        fl8 = fl.view(torch.uint8) if fdt in (torch.bool, torch.uint8) else (fl != 0).view(torch.uint8)
        vis_win = torch.empty((nbl, 1, int(ntime), nchan), dtype=torch.complex64, device=vis.device)
        flag_win = torch.empty((nbl, 1, int(ntime), nchan), dtype=torch.uint8, device=vis.device)
    else:
        vis_in = np.asarray(data)
        vdt = vis_in.dtype
        vis = np.ascontiguousarray(vis_in, dtype=np.complex64)
        f = np.asarray(flags)
        fdt = f.dtype
        fl8 = (np.ascontiguousarray(f).view(np.uint8) if f.dtype.itemsize == 1
               else np.ascontiguousarray(f != 0).view(np.uint8))
        vis_win = np.empty((nbl, 1, int(ntime), nchan), np.complex64)
        flag_win = np.empty((nbl, 1, int(ntime), nchan), np.uint8)
    pi, pc = _terms(stokes_pol)
    if stokes_unpol is not None:
        ui, uc = _terms(stokes_unpol)
    else:
        ui, uc = np.zeros((0, 2), np.int32), np.zeros((0, 4), np.float64)
    ctx, space = context_for(vis, fl8)
    check(_cabi.load().tc_stokes_pack(ctx.handle, _hp(slot), _hp(t), nrow, ptr(vis), ptr(fl8), nchan, ncorr,
                                      int(ntime), nbl, _hp(ui), _hp(uc), ui.shape[0], _hp(pi), _hp(pc),
                                      pi.shape[0], ptr(vis_win), ptr(flag_win), 1, space))
    if dev:
        import torch
        flag_win = flag_win.view(torch.bool) if fdt == torch.bool else flag_win.to(fdt)
    else:
        if vdt != np.complex64:
            vis_win = vis_win.astype(vdt)
        flag_win = flag_win.view(np.bool_) if fdt == np.bool_ else flag_win.astype(fdt)
    return vis_win, flag_win


# ---------------------------------------------------------------------------
# block functions with the signatures of the reference's private per-block
# functions; ``tricolour_b200.install()`` binds them over
# ``tricolour.packing._fast_pack_data`` / ``_unpack_data`` (packing.py:281-292,
# 391-415), which ``pack_data`` / ``unpack_data`` look up when dask runs a block
# ---------------------------------------------------------------------------
def _fast_pack_data(time_inv, ubl, ant1, ant2, data, flag, vis_windows, flag_windows):
    """One row chunk scattered into the (shared, in-place) window objects
    (packing.py:281-292 -> ``_numba_pack_data`` 243-278).  ``ubl`` arrives as
    dask's nested list of baseline blocks, the windows as one-element lists.
    The (row, chan, corr) -> (row, corr, chan) transposition runs on the GPU;
    the host then copies whole channel runs into the windows."""
    ubl = np.concatenate([bl for bl_list in ubl for bl in bl_list])
    vis_win, flag_win = vis_windows[0], flag_windows[0]
    data = np.asarray(data)
    flag = np.asarray(flag)
    rows, chans, corrs = data.shape
    if vis_win.shape[3] != chans:
        raise ValueError("channels mismatch")
    if vis_win.shape[1] != corrs:
        raise ValueError("correlations mismatch")
    if vis_win.shape != flag_win.shape:
        raise ValueError("vis_windows.shape != flag_windows.shape")
    assert ubl.shape == (vis_win.shape[0], 3)
    slot, t = _row_slots(ubl, ant1, ant2, time_inv, last_wins=True)
    keep = np.flatnonzero(slot >= 0)
    if keep.size:
        # every row is its own "baseline" of one dump: the pack kernel then is the
        # transposition of the chunk
        ident = np.ascontiguousarray(np.arange(rows, dtype=np.int32))
        zeros = np.zeros(rows, np.int32)
        vis = np.ascontiguousarray(data, dtype=np.complex64)
        fl8 = (np.ascontiguousarray(flag).view(np.uint8) if flag.dtype.itemsize == 1
               else np.ascontiguousarray(flag != 0).view(np.uint8))
        tv = np.empty((rows, corrs, 1, chans), np.complex64)
        tf = np.empty((rows, corrs, 1, chans), np.uint8)
        ctx, space = context_for(vis, fl8)
        check(_cabi.load().tc_pack(ctx.handle, _hp(ident), _hp(zeros), rows, ptr(vis), ptr(fl8), chans, corrs,
                                   1, rows, ptr(tv), ptr(tf), 1, space))
        vis_win[slot[keep], :, t[keep], :] = tv[keep, :, 0, :]
        flag_win[slot[keep], :, t[keep], :] = tf[keep, :, 0, :]
    return np.array([[[True]]])


def _unpack_data(antenna1, antenna2, time_inv, ubl, windows):
    """All baseline chunks of a window gathered back into one row chunk
    (packing.py:391-415).  ``ubl`` and ``windows`` are dask's lists of
    one-element lists (one entry per baseline chunk)."""
    exemplar = windows[0][0]
    antenna1 = np.asarray(antenna1)
    data = np.zeros((antenna1.shape[0], exemplar.shape[3], exemplar.shape[1]), dtype=exemplar.dtype)
    for baselines, window in zip(ubl, windows):
        baselines = np.asarray(baselines[0])
        window = window[0]
        u = baselines.copy()
        u[:, 0] = u[:, 0] - u[:, 0].min()
        slot, _ = _row_slots(u, antenna1, antenna2, time_inv, last_wins=False)
        mine = slot >= 0
        if not mine.any():
            continue
        data[mine] = unpack_data(antenna1, antenna2, time_inv, baselines, window)[mine]
    return data
