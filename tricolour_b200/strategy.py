# -*- coding: utf-8 -*-
"""
Strategy execution with device-resident windows (reference:
tricolour/apps/tricolour/strat_executor.py:29-83 and the YAML schema of
tricolour/conf/default.yaml).

The reference chains dask arrays, so every task moves the whole block through
host memory.  Here a block is uploaded once, all tasks run on the GPU against
the resident visibilities and flags with the reference's combine rule per task
(OR / replace), and only the final flags come back.
"""
import os
import sys
import time

import numpy as np

from . import _cabi
from ._cabi import check, ptr
from .flagging import (sum_threshold_flagger, uvcontsub_flagger, flag_autos,
                       flag_nans_and_zeros, _static_mask_tables, _apply_mask_tables)


def load_strategies(path):
    """``strategies:`` list of a tricolour YAML file (app.py:101-120)."""
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)["strategies"]


def _flags_or(a, b):
    """device flags a | b through the library (strat_executor.py:43, 54, 59, 76)"""
    import torch
    au8 = a.view(torch.uint8) if a.dtype == torch.bool else a
    bu8 = b.view(torch.uint8) if b.dtype == torch.bool else b
    out = torch.empty_like(au8)
    dev = a.device.index if a.device.index is not None else torch.cuda.current_device()
    ctx = _cabi.get_context(dev, _cabi.torch_stream_handle(dev))
    check(_cabi.load().tc_flags_or(ctx.handle, ptr(au8.contiguous()), ptr(bu8.contiguous()),
                                   ptr(out), int(out.numel()), _cabi.DEVICE))
    return out.view(torch.bool) if a.dtype == torch.bool else out


class StrategyExecutor(object):
    """Same constructor and ``apply_strategies`` contract as the reference's
    executor; ``flag_windows`` / ``vis_windows`` are numpy arrays (uploaded
    once) or torch CUDA tensors (used in place)."""

    def __init__(self, antenna_positions, unique_baselines,
                 chan_freq, chan_width, masked_channels, strategies):
        self.ant_pos = antenna_positions
        self.ubl = unique_baselines
        self.chan_freq = chan_freq
        self.chan_width = chan_width
        self.masked_channels = masked_channels
        self.strategies = strategies

    def __enter__(self):
        return self

    def __exit__(self, etype, evalue, etraceback):
        pass

    def apply_strategies(self, flag_windows, vis_windows, device=None):
        import torch
        host_io = not _cabi.is_device_array(flag_windows)
        if host_io:
            if not torch.cuda.is_available():
                raise RuntimeError("tricolour_b200: no CUDA device available; there is no CPU fallback")
            dev = torch.device("cuda", _cabi.default_device() if device is None else int(device))
            f_np = np.ascontiguousarray(flag_windows)
            fdt = f_np.dtype
            f8 = f_np.view(np.uint8) if f_np.dtype.itemsize == 1 else (f_np != 0).view(np.uint8)
            v_np = np.ascontiguousarray(vis_windows, dtype=np.complex64)
            with torch.cuda.device(dev):
                flags = torch.from_numpy(f8).to(dev, non_blocking=True).view(torch.bool)
                vis = torch.from_numpy(v_np).to(dev, non_blocking=True)
        else:
            dev = flag_windows.device
            flags, vis = flag_windows, vis_windows
            fdt = None
            if flags.dtype != torch.bool:
                flags = flags != 0
        with torch.cuda.device(dev):
            flags = self._run(flags, vis)
            if not host_io:
                return flags if flag_windows.dtype == torch.bool else flags.to(flag_windows.dtype)
            out = flags.view(torch.uint8).cpu().numpy()
        return out.view(np.bool_) if fdt == np.bool_ else out.astype(fdt)

    def apply_strategies_pipelined(self, blocks, device=None, depth=2, pre=None, post=None):
        """Flags a sequence of host blocks, overlapping transfers with flagging.

        ``blocks`` yields ``(flag_windows, vis_windows)`` numpy pairs (page-locked
        buffers make the uploads asynchronous); the flags of every block are
        yielded in order as numpy arrays.  The upload of block i+1 runs on a copy
        stream while block i is being flagged, the download of block i on a third
        stream while block i+1 is being flagged -- the dask scheduler of the
        reference overlaps its I/O with flagging the same way (app.py:266-271).
        Device and page-locked staging buffers for ``depth`` + 1 blocks are kept
        on the executor and reused.

        Buffer reuse: the arrays of a block are read until the iterator is asked
        for the following block (the executor waits for the block's upload before
        it calls ``next``), so a generator may refill one pair of page-locked
        buffers in place for every block it yields.

        ``pre`` / ``post`` put the packing of the reference's graph inside the same
        pipeline (app.py:415-432, 479-480): a block may then be any pair of equally
        shaped (flag, visibility) arrays, e.g. MS rows ``(row, chan, corr)``;
        ``pre(flags, vis)`` gets the uploaded device tensors and returns the
        ``(flag_windows, vis_windows)`` to flag (``packing.pack_data`` ...),
        ``post(flag_windows)`` turns the flagged windows into the device tensor that
        is downloaded (``packing.unpack_flags_equalised`` ...).  Both run on the
        flagging stream; whatever else they compute (window statistics) is theirs to keep.
        """
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("tricolour_b200: no CUDA device available; there is no CPU fallback")
        dev = torch.device("cuda", _cabi.default_device() if device is None else int(device))
        nslots = max(1, int(depth)) + 1
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            pipe = self.__dict__.setdefault("_pipe", {})
            if pipe.get("dev") != dev or pipe.get("nslots") != nslots:
                pipe.clear()
                pipe.update(dev=dev, nslots=nslots, up=torch.cuda.Stream(device=dev),
                            down=torch.cuda.Stream(device=dev), shape=None)
            up_s, down_s = pipe["up"], pipe["down"]

            def buffers(shape):
                if pipe["shape"] != tuple(shape):
                    pipe["shape"] = tuple(shape)
                    pipe["f"] = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(nslots)]
                    pipe["v"] = [torch.empty(shape, dtype=torch.complex64, device=dev) for _ in range(nslots)]
                    pipe["h"] = [None] * nslots         # page-locked download buffers, shaped like the results
                    pipe["hshape"] = None
                    pipe["free"] = [None] * nslots      # event: the slot's device buffers may be overwritten
                    if pre is None and post is None:
                        download_buffer(0, shape)       # plain windows: the results have the blocks' shape
                return pipe

            def download_buffer(slot, shape):
                # page-locked, shaped like the results; all slots at once the first time a result shape is
                # seen: page-locking memory is slow (~0.3 ms per MB) and belongs to the first block, not to
                # the steady state.  Downloads still in flight keep their own reference to the old buffers.
                shape = tuple(shape)
                if pipe.get("hshape") != shape or pipe["h"][slot] is None:
                    pipe["hshape"] = shape
                    pipe["h"] = [torch.empty(shape, dtype=torch.uint8, pin_memory=True) for _ in range(nslots)]
                return pipe["h"][slot]

            lib = _cabi.load()
            up_ctx = _cabi.get_context(dev.index, int(up_s.cuda_stream))
            down_ctx = _cabi.get_context(dev.index, int(down_s.cuda_stream))

            def upload(pair, k):
                f_np = np.ascontiguousarray(pair[0])
                f8 = f_np.view(np.uint8) if f_np.dtype.itemsize == 1 else (f_np != 0).view(np.uint8)
                v_np = np.ascontiguousarray(pair[1], dtype=np.complex64)
                b = buffers(f8.shape)
                slot = k % nslots
                if b["free"][slot] is not None:
                    up_s.wait_event(b["free"][slot])
                # raw asynchronous copies on the upload stream: with page-locked sources the host
                # does not wait (torch would treat foreign page-locked memory as pageable)
                check(lib.tc_memcpy_async(up_ctx.handle, ptr(b["f"][slot]), ptr(f8), int(f8.nbytes), 0))
                check(lib.tc_memcpy_async(up_ctx.handle, ptr(b["v"][slot]), ptr(v_np), int(v_np.nbytes), 0))
                ev = torch.cuda.Event(enable_timing=trace is not None)
                ev.record(up_s)
                if trace is not None:
                    trace.append(("upload %d enqueued (host)" % k, time.perf_counter()))
                    trace.append(("upload %d done" % k, ev))
                return slot, ev, f_np.dtype, (f8, v_np)      # keep the sources alive until the copy ran

            from concurrent.futures import ThreadPoolExecutor
            copier = pipe.setdefault("copier", ThreadPoolExecutor(max_workers=1))
            slicers = pipe.setdefault("slicers", ThreadPoolExecutor(max_workers=8))

            def prepare(shape):
                # the caller's array is allocated and its pages are touched on the helper thread
                # while the block is being flagged
                out = np.empty(shape, np.uint8)
                out.reshape(-1)[::4096] = 0
                return out

            def finish(item):
                hbuf, ev, fdt, prep = item
                out = prep.result()
                src = hbuf.numpy()                       # the page-locked buffer is reused
                flat_o, flat_s = out.reshape(-1), src.reshape(-1)
                ev.synchronize()
                # once the download has landed only a plain copy is left, cut into slices for a
                # few threads (numpy releases the GIL while copying)
                n = flat_o.size
                ncut = 8 if n >= (1 << 24) else 1
                cuts = [n * i // ncut for i in range(ncut + 1)]
                list(slicers.map(lambda i: np.copyto(flat_o[cuts[i]:cuts[i + 1]], flat_s[cuts[i]:cuts[i + 1]]),
                                 range(ncut)))
                return out.view(np.bool_) if fdt == np.bool_ else out.astype(fdt)

            # TC_PIPE_TRACE=1 prints a timeline (device events and host times) after the last block
            trace = [] if os.environ.get("TC_PIPE_TRACE") else None
            if trace is not None:
                t_host0 = time.perf_counter()
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(main)
            it = iter(blocks)
            pending = []          # futures of downloads in flight
            k = 0
            try:
                nxt = upload(next(it), k)
            except StopIteration:
                return
            while nxt is not None:
                cur = nxt
                k += 1
                # a slot's page-locked output must have been consumed before its reuse
                while len(pending) >= nslots - 1:
                    yield pending.pop(0).result()
                slot, ev_up, fdt, _src = cur
                # The copy of the block just taken from the iterator reads the caller's arrays
                # asynchronously.  The iterator may refill those very buffers as soon as it is
                # asked for the next block (one page-locked buffer per stream of blocks is the
                # usage INTEGRATION.md recommends), so the copy must have finished first.  It was
                # queued one whole block of flagging ago: this wait is normally free.
                ev_up.synchronize()
                _src = None
                try:
                    nxt = upload(next(it), k)      # queued before this block's kernels: overlaps them
                except StopIteration:
                    nxt = None
                if pre is None and post is None:
                    prep = copier.submit(prepare, tuple(pipe["f"][slot].shape))
                main.wait_event(ev_up)
                if trace is not None:
                    ev_start = torch.cuda.Event(enable_timing=True)
                    ev_start.record(main)
                    trace.append(("flag %d start" % (k - 1), ev_start))
                fw, vw = pipe["f"][slot].view(torch.bool), pipe["v"][slot]
                if pre is not None:
                    fw, vw = pre(fw, vw)
                res = self._run(fw, vw)
                if post is not None:
                    res = post(res)
                res = res.contiguous()
                if pre is not None or post is not None:
                    prep = copier.submit(prepare, tuple(res.shape))
                ev_done = torch.cuda.Event(enable_timing=trace is not None)
                ev_done.record(main)
                if trace is not None:
                    trace.append(("flag %d enqueued (host)" % (k - 1), time.perf_counter()))
                    trace.append(("flag %d done" % (k - 1), ev_done))
                pipe["free"][slot] = ev_done
                down_s.wait_event(ev_done)
                r8 = res.view(torch.uint8) if res.dtype == torch.bool else res
                hbuf = download_buffer(slot, r8.shape)
                check(lib.tc_memcpy_async(down_ctx.handle, _cabi._vp(hbuf.data_ptr()), ptr(r8), int(r8.numel()), 1))
                ev_out = torch.cuda.Event(enable_timing=trace is not None)
                ev_out.record(down_s)
                if trace is not None:
                    trace.append(("download %d done" % (k - 1), ev_out))
                res.record_stream(down_s)
                # the host-side copy out of the page-locked buffer runs on a helper thread
                pending.append(copier.submit(finish, (hbuf, ev_out, fdt, prep)))
            while pending:
                yield pending.pop(0).result()
            if trace is not None:
                trace.append(("all results handed over (host)", time.perf_counter()))
                for name, what in trace:
                    ms = (what - t_host0) * 1e3 if isinstance(what, float) else ev0.elapsed_time(what)
                    sys.stderr.write("pipe trace %9.1f ms  %s\n" % (ms, name))

    def _run(self, flag_windows, vis_windows):
        original = flag_windows.clone()
        ubl = np.asarray(self.ubl)
        for strategy in self.strategies:
            try:
                task = strategy['task']
            except KeyError:
                raise ValueError("strategy has no 'task': %s" % strategy)
            kwargs = strategy.get('kwargs', {}) or {}
            if task == "sum_threshold":
                new_flags = sum_threshold_flagger(vis_windows, flag_windows, **kwargs)
                # sum threshold builds upon any flags that came previous
                flag_windows = _flags_or(new_flags, flag_windows)
            elif task == "uvcontsub_flagger":
                # discards previous flags per its or_original_from_cycle rule
                flag_windows = uvcontsub_flagger(vis_windows, flag_windows, **kwargs)
            elif task == "flag_autos":
                new_flags = flag_autos(flag_windows, [ubl])
                flag_windows = _flags_or(new_flags, flag_windows)
            elif task == "combine_with_input_flags":
                flag_windows = _flags_or(flag_windows, original)
            elif task == "unflag":
                flag_windows = flag_windows.new_zeros(flag_windows.shape)
            elif task == "flag_nans_zeros":
                flag_windows = flag_nans_and_zeros(vis_windows, flag_windows)
            elif task == "apply_static_mask":
                if flag_windows.shape[0] != ubl.shape[0]:
                    raise ValueError("flag and ubl shape mismatch %s != %s"
                                     % (flag_windows.shape[1], ubl.shape[0]))
                # the selector tables depend on observation metadata only: built once per task
                cache = self.__dict__.setdefault("_mask_tables", {})
                key = (kwargs.get("accumulation_mode", "or"), kwargs.get("uvrange", ""))
                if key not in cache:
                    cache[key] = _static_mask_tables(ubl, self.ant_pos, self.masked_channels, self.chan_freq,
                                                     self.chan_width, *key)
                new_flags = _apply_mask_tables(flag_windows, cache[key])
                if kwargs["accumulation_mode"].strip() == "or":
                    flag_windows = _flags_or(new_flags, flag_windows)
                else:
                    flag_windows = new_flags
            else:
                raise ValueError("Task '%s' does not name a valid task", task)
        return flag_windows
