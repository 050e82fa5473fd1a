# -*- coding: utf-8 -*-
"""How long does the host spend inside one apply_strategies call (device tensors)
compared with the GPU time of that call?  Finds hidden host<->device syncs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, common
import tricolour_b200 as tb
from tricolour_b200 import flagging
dev = torch.device("cuda", 0)
B, T, F = 16, 512, 4096
ubl = common.baselines(64); ants = common.antenna_layout(64); cf, cw = common.channels(F)
masks = common.synthetic_static_mask(cf); strategies = common.default_strategies()
my = ubl[:B].copy()
vis, flags = bench.make_block_torch(B, 4, T, F, 0, ubl, dev, 1)
ex = tb.StrategyExecutor(ants, my, cf, cw, masks, strategies)
for _ in range(2): ex.apply_strategies(flags, vis)
torch.cuda.synchronize()
t0 = time.perf_counter(); out = ex.apply_strategies(flags, vis); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("host inside call %.1f ms, until GPU done %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
# per task
orig = flagging.sum_threshold_flagger
for name in ("sum_threshold_flagger", "uvcontsub_flagger", "flag_nans_and_zeros", "apply_static_mask", "flag_autos"):
    fn = getattr(flagging, name)
kw = strategies[2]["kwargs"]
torch.cuda.synchronize(); t0 = time.perf_counter(); r = tb.sum_threshold_flagger(vis, flags, **kw); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("sum_threshold: host %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
kw = strategies[3]["kwargs"]
torch.cuda.synchronize(); t0 = time.perf_counter(); r = tb.uvcontsub_flagger(vis, flags, **kw); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("uvcontsub: host %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
# host time of every task inside the executor
import tricolour_b200.strategy as S
def wrap(name):
    fn = getattr(S, name)
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); dt = (time.perf_counter() - t0) * 1e3
        print("   %-24s host %.2f ms" % (name, dt))
        return r
    setattr(S, name, w)
for n in ("sum_threshold_flagger", "uvcontsub_flagger", "flag_autos", "flag_nans_and_zeros", "apply_static_mask", "_flags_or"):
    wrap(n)
torch.cuda.synchronize()
t0 = time.perf_counter(); out = ex.apply_strategies(flags, vis); t1 = time.perf_counter()
print("total host %.1f ms" % ((t1 - t0) * 1e3))
