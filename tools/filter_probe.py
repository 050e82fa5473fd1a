#!/usr/bin/env python
"""Per-launch times of the masked box filter forms, by radius.

Calls the stage entry tc_stage_masked_filter (device pointers) on one block of
planes for each (r0, r1) of default.yaml's passes and prints the library's own
CUDA-event profile (first axis / second axis kernels separately).  Environment
knobs select the kernel form; they are read once per process, so run one process
per variant:   TC_FILTER_TPL=1 python tools/filter_probe.py
usage: filter_probe.py [nbl=16] [T=512] [F=4096] [r0,r1 ...]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tricolour_b200 import _cabi
from tricolour_b200._cabi import check, ptr

args = [a for a in sys.argv[1:]]
nums = [int(a) for a in args if "," not in a]
nbl = nums[0] if len(nums) > 0 else 16
T = nums[1] if len(nums) > 1 else 512
F = nums[2] if len(nums) > 2 else 4096
pairs = [tuple(int(x) for x in a.split(",")) for a in args if "," in a]
if not pairs:
    pairs = [(54, 43), (43, 34), (32, 25), (21, 17), (10, 8), (55, 277), (27, 138), (6, 34), (5, 8), (0, 8), (0, 43)]
ncp = nbl * 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
data = torch.rand((ncp, T, F), device=dev, generator=g, dtype=torch.float32) + 2.0
flags = (torch.rand((ncp, T, F), device=dev, generator=g) < float(os.environ.get("PROBE_FLAG_FRAC", "0.1"))).to(torch.uint8)
out = torch.empty_like(data)
ctx = _cabi.get_context(0, _cabi.torch_stream_handle(0))
lib = _cabi.load()
res = []
for r0, r1 in pairs:
    for rep in range(3):
        if rep == 1:
            ctx.profile(True); ctx.profile_reset()
        check(lib.tc_stage_masked_filter(ctx.handle, ptr(data), ptr(flags), ncp, T, F, r0, r1, ptr(out), 1))
        ctx.synchronize()
    prof = ctx.profile_read(); ctx.profile(False)
    row = {"r0": r0, "r1": r1}
    for k, (ms, n) in prof.items():
        if k.startswith("box") and n:
            row[k] = round(ms / n, 3)
    res.append(row)
    print(json.dumps(row), flush=True)
vis = ncp * T * F
print(json.dumps({"summary": "filter_probe", "block": [nbl, 4, T, F], "chain_steps_per_axis_launch": vis * 8,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("TC_")}, "rows": res}))
