# -*- coding: utf-8 -*-
"""
Flag statistics of a window (reference: tricolour/window_statistics.py).

The counting (per baseline and per channel sums of the flag bytes) is one GPU
reduction; the dictionary bookkeeping, the channel-bin edges and the text
summary stay in Python exactly like the reference.  ``allreduce_window_stats``
adds the one collective of the multi-GPU path: a single small all-reduce of the
count vector in the fixed layout of ``StatsLayout`` (NCCL over NVLink on GPUs,
gloo in CPU tests).
"""
import ctypes
from collections import defaultdict
from functools import partial

import numpy as np

from . import _cabi
from ._cabi import check, ptr, context_for
from .packing import _WINDOW_SCHEMA  # noqa: F401


_SCALAR_TABLES = tuple("_%s_per_%s" % (kind, what) for what in ("ant", "field", "scan", "bl")
                       for kind in ("counts", "size")) + ("_size_per_ddid",)


class WindowStatistics(object):
    """Flag-count accumulator with the attribute names of the reference's class
    (window_statistics.py:173-231), which its ``summarise_stats`` and the
    application read: ``_counts_per_{ant,field,scan,bl}``, ``_size_per_*``,
    ``_counts_per_ddid`` (``nchanbins`` uint64 bins per data descriptor),
    ``_bins_per_ddid`` (bin frequency labels) and ``_size_per_ddid``."""

    def __init__(self, nchanbins):
        self._nchanbins = nchanbins
        for table in _SCALAR_TABLES:
            setattr(self, table, defaultdict(int))
        self._counts_per_ddid = defaultdict(partial(np.zeros, nchanbins, dtype=np.uint64))
        self._bins_per_ddid = defaultdict(int)

    def update(self, other):
        """adds ``other`` into this accumulator (the frequency labels are copied)"""
        for table in _SCALAR_TABLES:
            mine, theirs = getattr(self, table), getattr(other, table)
            for key, value in theirs.items():
                mine[key] += value
        for ddid, bins in other._counts_per_ddid.items():
            self._counts_per_ddid[ddid] += bins
        self._bins_per_ddid.update(other._bins_per_ddid)

    def copy(self):
        out = WindowStatistics(self._nchanbins)
        out.update(self)
        return out


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


class StatsLayout(object):
    """Fixed layout of a :class:`WindowStatistics` as one int64 vector
    (SURVEY.md 8(e)): ``[per-antenna counts, per-antenna sizes, per-baseline counts
    (all baselines of the observation, zero where a rank owns none of a baseline),
    per-baseline sizes, per-field count+size, per-scan count+size, per-ddid
    channel-bin counts, per-ddid size]``.  Every rank builds the same layout from
    observation metadata (antenna names, the global ``ubl`` table, field names,
    scan numbers, the channel frequencies of every data descriptor), so no keys
    have to be exchanged and the statistics of all ranks are summed by ONE
    all-reduce.  It replaces the host-side ``combine_window_stats`` of the
    reference (window_statistics.py:143-170) across processes."""

    def __init__(self, antenna_names, ubl, field_names, scan_numbers, ddid_chan_freqs, nchanbins=10):
        self.antenna_names = list(antenna_names)
        ubl = np.asarray(ubl)
        self.bl_names = []
        for a1, a2 in ubl[:, 1:3]:
            name = "{0:s}&{1:s}".format(self.antenna_names[int(a1)], self.antenna_names[int(a2)])
            if name not in self.bl_names:
                self.bl_names.append(name)
        self.field_names = list(field_names)
        self.scan_numbers = list(scan_numbers)
        self.ddids = list(ddid_chan_freqs.keys())
        self.nchanbins = int(nchanbins)
        # channel-bin labels exactly as _window_stats computes them (window_statistics.py:53-55)
        self.bin_edges = {d: np.linspace(np.min(f), np.max(f), self.nchanbins)
                          for d, f in ddid_chan_freqs.items()}
        self._fields = (("_counts_per_ant", self.antenna_names), ("_size_per_ant", self.antenna_names),
                        ("_counts_per_bl", self.bl_names), ("_size_per_bl", self.bl_names),
                        ("_counts_per_field", self.field_names), ("_size_per_field", self.field_names),
                        ("_counts_per_scan", self.scan_numbers), ("_size_per_scan", self.scan_numbers),
                        ("_size_per_ddid", self.ddids))
        self.size = sum(len(k) for _, k in self._fields) + len(self.ddids) * self.nchanbins

    def pack(self, stats):
        """int64 vector of ``stats``; a key outside the layout is an error (it
        would silently vanish from the reduced statistics)"""
        if stats._nchanbins != self.nchanbins:
            raise ValueError("statistics have %d channel bins, the layout %d" % (stats._nchanbins, self.nchanbins))
        vec = np.zeros(self.size, np.int64)
        pos = 0
        for name, keys in self._fields:
            d = getattr(stats, name)
            extra = set(d.keys()) - set(keys)
            if extra:
                raise ValueError("%s holds keys outside the layout: %s" % (name, sorted(map(str, extra))[:4]))
            for k in keys:
                if k in d:
                    vec[pos] = int(d[k])
                pos += 1
        extra = set(stats._counts_per_ddid.keys()) - set(self.ddids)
        if extra:
            raise ValueError("_counts_per_ddid holds keys outside the layout: %s" % sorted(extra)[:4])
        for k in self.ddids:
            if k in stats._counts_per_ddid:
                vec[pos:pos + self.nchanbins] = np.asarray(stats._counts_per_ddid[k]).astype(np.int64)
            pos += self.nchanbins
        return vec

    def unpack(self, vec):
        """:class:`WindowStatistics` from a (reduced) vector; keys whose size is
        zero everywhere (nothing was counted for them) are left out, like in the
        reference where a key only exists once a block contributed to it"""
        vec = np.asarray(vec).astype(np.int64)
        out = WindowStatistics(self.nchanbins)
        pos = 0
        vals = {}
        for name, keys in self._fields:
            vals[name] = vec[pos:pos + len(keys)]
            pos += len(keys)
        pairs = (("_counts_per_ant", "_size_per_ant", self.antenna_names),
                 ("_counts_per_bl", "_size_per_bl", self.bl_names),
                 ("_counts_per_field", "_size_per_field", self.field_names),
                 ("_counts_per_scan", "_size_per_scan", self.scan_numbers))
        for cn, sn, keys in pairs:
            for i, k in enumerate(keys):
                if vals[sn][i] != 0 or vals[cn][i] != 0:
                    getattr(out, cn)[k] += int(vals[cn][i])
                    getattr(out, sn)[k] += int(vals[sn][i])
        for i, k in enumerate(self.ddids):
            bins = vec[pos:pos + self.nchanbins]
            pos += self.nchanbins
            if vals["_size_per_ddid"][i] != 0 or bins.any():
                out._size_per_ddid[k] += int(vals["_size_per_ddid"][i])
                out._counts_per_ddid[k] += bins.astype(np.uint64)
                out._bins_per_ddid[k] = self.bin_edges[k]
        return out


def allreduce_window_stats(stats, layout=None, group=None, device=None):
    """Sums a :class:`WindowStatistics` over all ranks of ``torch.distributed``
    with exactly ONE collective: the statistics are packed into the fixed int64
    layout of :class:`StatsLayout` and reduced by a single ``all_reduce`` (NCCL
    over NVLink on GPUs, gloo in the CPU tests).  ``stats`` may also be a
    sequence of statistics (e.g. ``(original, final)``): they share the one
    all-reduce and a tuple is returned.  Without a process group (or with one
    rank) copies are returned and nothing is communicated."""
    import torch
    import torch.distributed as dist
    many = isinstance(stats, (list, tuple))
    items = list(stats) if many else [stats]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        outs = [s.copy() for s in items]
        return tuple(outs) if many else outs[0]
    if layout is None:
        raise ValueError("allreduce_window_stats needs a StatsLayout when more than one rank takes part: "
                         "the vector layout must be identical on every rank without exchanging keys")
    vec = np.concatenate([layout.pack(s) for s in items])
    backend = dist.get_backend(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(vec).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = t.cpu().numpy()
    outs = [layout.unpack(vals[i * layout.size:(i + 1) * layout.size]) for i in range(len(items))]
    return tuple(outs) if many else outs[0]


def summarise_stats(final, original):
    """The text summary of final against original flag percentages, line for
    line what the reference logs (window_statistics.py:234-294)."""
    stars = "*" * 32
    lines = [stars, "   BEGINNING OF FLAG SUMMARY    ", stars]
    sections = (("Per antenna:", "ant", "{0:s}"), ("Per scan:", "scan", "{0:d}"),
                ("Per field:", "field", "{0:s}"), ("Per baseline:", "bl", "{0:s}"))
    for title, what, keyfmt in sections:
        lines.append(title)
        fc, fs = getattr(final, "_counts_per_" + what), getattr(final, "_size_per_" + what)
        oc, os_ = getattr(original, "_counts_per_" + what), getattr(original, "_size_per_" + what)
        for key in fc:
            lines.append(("\t " + keyfmt + ": {1:.3f}%, original {2:.3f}%").format(
                key, fc[key] * 100.0 / fs[key], oc[key] * 100.0 / os_[key]))
    lines.append("Per data descriptor id:")
    for ddid in final._counts_per_ddid:
        ratios = final._counts_per_ddid[ddid] * 100.0 / final._size_per_ddid[ddid]
        lines.append("\t {0:d}: {1:s}%".format(ddid, '\t'.join("{0:<7.2f}".format(r) for r in ratios)))
        mhz = final._bins_per_ddid[ddid] / 1e6
        lines.append("\t    {0:s} MHz".format('\t'.join("{0:<7.1f}".format(f) for f in mhz)))
    lines += [stars, "       END OF FLAG SUMMARY      ", stars]
    return lines
