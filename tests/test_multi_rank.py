# -*- coding: utf-8 -*-
"""N > 1 path on CPU: two gloo ranks shard the baselines (no data-path
collective), flag their shard, and all-reduce only the window statistics
(SURVEY.md 8e).  Kernels run under the CPU emulator here; bench.py --gpus N
exercises the same code with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import tricolour_b200 as tb
    from tricolour_b200 import _cabi
    import common
    _cabi._set_library_for_testing(_cabi.load(os.path.join(ROOT, "tests", "_emu", "libtricolour_b200_emu.so")))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nant = 4
    ubl = common.baselines(nant)
    nbl = ubl.shape[0]
    cf, _ = common.channels(32)
    names = ["m%03d" % i for i in range(nant)]
    vis, flags = common.make_windows(nbl, 2, 8, 32, seed=77, ubl=ubl)
    # contiguous baseline ranges, ceil(nbl / world) per rank
    per = (nbl + world - 1) // world
    lo, hi = rank * per, min((rank + 1) * per, nbl)
    mine = tb.flag_nans_and_zeros(vis[lo:hi], flags[lo:hi])
    st0 = tb.window_stats(flags[lo:hi], ubl[lo:hi], cf, names, 1, "f", 0)
    st = tb.window_stats(mine, ubl[lo:hi], cf, names, 1, "f", 0)
    layout = tb.StatsLayout(names, ubl, ["f"], [1], {0: cf})
    # exactly one collective for both statistics objects: count the calls torch.distributed sees
    calls = []
    for fn in ("all_reduce", "all_gather", "all_gather_object", "broadcast", "reduce", "all_to_all",
               "gather", "scatter", "broadcast_object_list", "all_gather_into_tensor", "reduce_scatter"):
        orig = getattr(dist, fn)

        def counted(*a, _orig=orig, _fn=fn, **k):
            calls.append(_fn)
            return _orig(*a, **k)
        setattr(dist, fn, counted)
    red0, red = tb.window_statistics.allreduce_window_stats((st0, st), layout)
    ncalls = list(calls)
    if rank == 0:
        full = tb.window_stats(tb.flag_nans_and_zeros(vis, flags), ubl, cf, names, 1, "f", 0)
        ok = (dict(red._counts_per_ant) == {k: int(v) for k, v in full._counts_per_ant.items()}
              and dict(red._size_per_ant) == dict(full._size_per_ant)
              and {k: int(v) for k, v in red._counts_per_bl.items()} == {k: int(v) for k, v in full._counts_per_bl.items()}
              and int(red._counts_per_field["f"]) == int(full._counts_per_field["f"])
              and int(red._size_per_scan[1]) == int(full._size_per_scan[1])
              and np.array_equal(red._counts_per_ddid[0], full._counts_per_ddid[0])
              and np.array_equal(red._bins_per_ddid[0], full._bins_per_ddid[0])
              and list(red._counts_per_bl.keys()) == list(full._counts_per_bl.keys())
              and int(red0._counts_per_field["f"]) == int(flags.sum())
              and ncalls == ["all_reduce"])
        open(os.path.join(tmpdir, "result"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_baseline_sharding_and_stats_allreduce(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "result")).read() == "ok"


def test_stats_layout_round_trip():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tricolour_b200 as tb
    import common
    names = ["a", "b", "c"]
    ubl = common.baselines(3)
    cf, _ = common.channels(16)
    lay = tb.StatsLayout(names, ubl, ["f0", "f1"], [4, 5], {0: cf, 1: cf * 2})
    st = tb.WindowStatistics(10)
    st._counts_per_ant["b"] += 5
    st._size_per_ant["b"] += 50
    st._counts_per_bl["a&c"] += 2
    st._size_per_bl["a&c"] += 20
    st._counts_per_scan[5] += 9
    st._size_per_scan[5] += 90
    st._counts_per_field["f1"] += 9
    st._size_per_field["f1"] += 90
    st._counts_per_ddid[1] += np.arange(10, dtype=np.uint64)
    st._size_per_ddid[1] += 90
    v = lay.pack(st)
    assert v.dtype == np.int64 and v.size == lay.size == 2 * 3 + 2 * 6 + 4 + 4 + 2 + 20
    back = lay.unpack(v * 2)
    assert dict(back._counts_per_ant) == {"b": 10} and dict(back._size_per_bl) == {"a&c": 40}
    assert dict(back._counts_per_scan) == {5: 18} and dict(back._size_per_field) == {"f1": 180}
    assert np.array_equal(back._counts_per_ddid[1], 2 * np.arange(10, dtype=np.uint64))
    assert np.array_equal(back._bins_per_ddid[1], np.linspace(cf.min() * 2, cf.max() * 2, 10))
    st._counts_per_ant["zz"] += 1
    with pytest.raises(ValueError):
        lay.pack(st)


def test_allreduce_is_identity_without_process_group():
    sys.path.insert(0, ROOT)
    import tricolour_b200 as tb
    st = tb.WindowStatistics(10)
    st._counts_per_ant["a"] += 3
    out = tb.window_statistics.allreduce_window_stats(st)
    assert out._counts_per_ant["a"] == 3 and out is not st
