# -*- coding: utf-8 -*-
"""
Flag statistics of a window (reference: tricolour/window_statistics.py).

The counting (per baseline and per channel sums of the flag bytes) is one GPU
reduction; the dictionary bookkeeping, the channel-bin edges and the text
summary stay in Python exactly like the reference.  ``allreduce_window_stats``
adds the one collective of the multi-GPU path: a single small all-reduce of the
packed count vector (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import ctypes
from collections import defaultdict
from functools import partial

import numpy as np

from . import _cabi
from ._cabi import check, ptr, context_for
from .packing import _WINDOW_SCHEMA  # noqa: F401


class WindowStatistics(object):
    """Accumulator with the reference's fields (window_statistics.py:173-231)."""

    def __init__(self, nchanbins):
        self._nchanbins = nchanbins
        self._counts_per_ant = defaultdict(lambda: 0)
        self._counts_per_field = defaultdict(lambda: 0)
        self._counts_per_scan = defaultdict(lambda: 0)
        self._counts_per_bl = defaultdict(lambda: 0)
        self._size_per_ant = defaultdict(lambda: 0)
        self._size_per_field = defaultdict(lambda: 0)
        self._size_per_scan = defaultdict(lambda: 0)
        self._size_per_bl = defaultdict(lambda: 0)
        bin_factory = partial(np.zeros, nchanbins, dtype=np.uint64)
        self._counts_per_ddid = defaultdict(bin_factory)
        self._bins_per_ddid = defaultdict(lambda: 0)
        self._size_per_ddid = defaultdict(lambda: 0)

    def update(self, other):
        for name in ("_counts_per_ant", "_counts_per_field", "_counts_per_scan",
                     "_counts_per_bl", "_size_per_ant", "_size_per_field",
                     "_size_per_scan", "_size_per_ddid", "_size_per_bl"):
            mine = getattr(self, name)
            for k, v in getattr(other, name).items():
                mine[k] += v
        for d, count in other._counts_per_ddid.items():
            self._counts_per_ddid[d] += count
        for d, bins in other._bins_per_ddid.items():
            self._bins_per_ddid[d] = bins  # frequency labels

    def copy(self):
        """ Creates a copy of the current WindowStatistics"""
        result = WindowStatistics(self._nchanbins)
        result.update(self)
        return result


def _counts(flag_window):
    """(bl_counts[nbl], chan_counts[nchan]) uint64 from the GPU"""
    w = flag_window
    nbl, ncorr, T, F = (int(s) for s in w.shape)
    if _cabi.is_device_array(w):
        import torch
        raw = w.contiguous()
        raw = raw.view(torch.uint8) if raw.dtype in (torch.bool, torch.uint8) else (raw != 0).view(torch.uint8)
    else:
        wa = np.asarray(w)
        raw = (np.ascontiguousarray(wa).view(np.uint8) if wa.dtype.itemsize == 1
               else np.ascontiguousarray(wa != 0).view(np.uint8))
    blc = np.zeros(max(nbl, 1), np.uint64)
    chc = np.zeros(max(F, 1), np.uint64)
    ctx, space = context_for(raw)
    check(_cabi.load().tc_window_counts(ctx.handle, ptr(raw), nbl, ncorr, T, F,
                                        blc.ctypes.data_as(ctypes.c_void_p),
                                        chc.ctypes.data_as(ctypes.c_void_p), space))
    return blc[:nbl], chc[:F]


def _window_stats(flag_window, ubls, chan_freqs,
                  antenna_names, scan_no, field_name, ddid, nchanbins):
    """
    Stats for one block of a flag window (window_statistics.py:12-66).  The
    arguments may be list-wrapped the way dask hands them to the reference.
    """
    while isinstance(ubls, (list, tuple)):
        ubls = ubls[0]
    while isinstance(flag_window, (list, tuple)):
        flag_window = flag_window[0]
    while isinstance(chan_freqs, (list, tuple)):
        chan_freqs = chan_freqs[0]
    ubls = np.asarray(ubls)
    chan_freqs = np.asarray(chan_freqs)
    nbl, ncorr, T, F = (int(s) for s in flag_window.shape)
    plane = ncorr * T * F
    blc, chc = _counts(flag_window)

    stats = WindowStatistics(nchanbins)
    # per antenna
    for ai, a in enumerate(antenna_names):
        sel = np.logical_or(ubls[:, 1] == ai, ubls[:, 2] == ai)
        stats._counts_per_ant[a] += np.sum(blc[sel], dtype=np.uint64)
        stats._size_per_ant[a] += int(sel.sum()) * plane
    # per baseline
    for b in np.unique(ubls[:, 0]):
        sel = ubls[:, 0] == b
        sela1 = antenna_names[ubls[sel, 1][0]]
        sela2 = antenna_names[ubls[sel, 2][0]]
        blname = "{0:s}&{1:s}".format(sela1, sela2)
        stats._counts_per_bl[blname] += np.sum(blc[sel], dtype=np.uint64)
        stats._size_per_bl[blname] += int(sel.sum()) * plane
    # per scan and field
    cnt = np.sum(blc, dtype=np.uint64)
    sz = nbl * plane
    stats._counts_per_field[field_name] += cnt
    stats._size_per_field[field_name] += sz
    stats._counts_per_scan[scan_no] += cnt
    stats._size_per_scan[scan_no] += sz
    # binned per channel: nchanbins EDGES, so the last bin stays empty
    bins_edges = np.linspace(np.min(chan_freqs), np.max(chan_freqs), nchanbins)
    bins = np.zeros(nchanbins, dtype=np.uint32)
    for ch_i in range(nchanbins - 1):
        sel = np.logical_and(chan_freqs >= bins_edges[ch_i],
                             chan_freqs < bins_edges[ch_i + 1])
        bins[ch_i] = np.sum(chc[sel], dtype=np.uint64).astype(np.uint32)
    stats._counts_per_ddid[ddid] += bins
    stats._bins_per_ddid[ddid] = bins_edges
    stats._size_per_ddid[ddid] += sz
    return stats


def window_stats(flag_window, ubls, chan_freqs,
                 antenna_names, scan_no, field_name, ddid,
                 nchanbins=10, prev_stats=None):
    """
    Stats of a ``flag_window`` of shape (bl, corr, time, chan), merged into
    ``prev_stats`` if given (window_statistics.py:81-140).  Eager: returns a
    :class:`WindowStatistics`.
    """
    stats = _window_stats(flag_window, ubls, chan_freqs, antenna_names,
                          scan_no, field_name, ddid, nchanbins)
    result = WindowStatistics(nchanbins) if prev_stats is None else prev_stats.copy()
    result.update(stats)
    return result


def combine_window_stats(window_stats):
    """Combines a list of :class:`WindowStatistics` (window_statistics.py:151-170)."""
    result = window_stats[0].copy()
    for arg in window_stats[1:]:
        result.update(arg)
    return result


def allreduce_window_stats(stats, group=None, device=None):
    """Sums a :class:`WindowStatistics` over all ranks of ``torch.distributed``:
    every dictionary is packed into one int64 vector (keys are exchanged once
    with ``all_gather_object``) and reduced with a single ``all_reduce``."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats.copy()
    names = ("_counts_per_ant", "_counts_per_field", "_counts_per_scan", "_counts_per_bl",
             "_size_per_ant", "_size_per_field", "_size_per_scan", "_size_per_bl",
             "_size_per_ddid")
    local_keys = {n: list(getattr(stats, n).keys()) for n in names}
    local_keys["_ddid"] = list(stats._counts_per_ddid.keys())
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, local_keys, group=group)
    keys = {}
    for n in list(names) + ["_ddid"]:
        seen = []
        for g in gathered:
            for k in g[n]:
                if k not in seen:
                    seen.append(k)
        keys[n] = seen
    vec = []
    for n in names:
        d = getattr(stats, n)
        vec.extend(int(d[k]) if k in d else 0 for k in keys[n])
    nb = stats._nchanbins
    for k in keys["_ddid"]:
        if k in stats._counts_per_ddid:
            vec.extend(int(x) for x in stats._counts_per_ddid[k])
        else:
            vec.extend([0] * nb)
    backend = dist.get_backend(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor(vec, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = t.cpu().tolist()
    out = WindowStatistics(nb)
    pos = 0
    for n in names:
        d = getattr(out, n)
        for k in keys[n]:
            d[k] += vals[pos]
            pos += 1
    for k in keys["_ddid"]:
        out._counts_per_ddid[k] += np.array(vals[pos:pos + nb], dtype=np.uint64)
        pos += nb
    # frequency labels are identical on every rank that has them
    bins = [None] * dist.get_world_size(group)
    dist.all_gather_object(bins, {k: np.asarray(v).tolist() for k, v in stats._bins_per_ddid.items()},
                           group=group)
    for b in bins:
        for k, v in b.items():
            out._bins_per_ddid[k] = np.asarray(v)
    return out


def summarise_stats(final, original):
    """
    Returns a list of strings summarising final and original flag percentages
    (window_statistics.py:234-294).
    """
    l = []  # noqa
    l.append("********************************")
    l.append("   BEGINNING OF FLAG SUMMARY    ")
    l.append("********************************")

    def pct(cnt, size):
        return cnt * 100.0 / size

    l.append("Per antenna:")
    for a in final._counts_per_ant:
        l.append("\t {0:s}: {1:.3f}%, original {2:.3f}%".format(
            a, pct(final._counts_per_ant[a], final._size_per_ant[a]),
            pct(original._counts_per_ant[a], original._size_per_ant[a])))
    l.append("Per scan:")
    for s in final._counts_per_scan:
        l.append("\t {0:d}: {1:.3f}%, original {2:.3f}%".format(
            s, pct(final._counts_per_scan[s], final._size_per_scan[s]),
            pct(original._counts_per_scan[s], original._size_per_scan[s])))
    l.append("Per field:")
    for f in final._counts_per_field:
        l.append("\t {0:s}: {1:.3f}%, original {2:.3f}%".format(
            f, pct(final._counts_per_field[f], final._size_per_field[f]),
            pct(original._counts_per_field[f], original._size_per_field[f])))
    l.append("Per baseline:")
    for b in final._counts_per_bl:
        l.append("\t {0:s}: {1:.3f}%, original {2:.3f}%".format(
            b, pct(final._counts_per_bl[b], final._size_per_bl[b]),
            pct(original._counts_per_bl[b], original._size_per_bl[b])))
    l.append("Per data descriptor id:")
    for d in final._counts_per_ddid:
        ratios = final._counts_per_ddid[d] * 100.0 / final._size_per_ddid[d]
        ratio_str = '\t'.join(["{0:<7.2f}".format(r) for r in ratios])
        l.append("\t {0:d}: {1:s}%".format(d, ratio_str))
        ddid_freqs = final._bins_per_ddid[d] / 1e6
        ddid_freqs_str = '\t'.join(["{0:<7.1f}".format(f) for f in ddid_freqs])
        l.append("\t    {0:s} MHz".format(ddid_freqs_str))
    l.append("********************************")
    l.append("       END OF FLAG SUMMARY      ")
    l.append("********************************")
    return l
