# -*- coding: utf-8 -*-
"""Times the masked box-Gaussian filter launches alone (device resident block,
CUDA-event profile of the library) for a list of (r0, r1) radii.
usage: python tools/bench_filter.py [nplanes T F]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tricolour_b200 import _cabi  # noqa: E402

P, T, F = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 512, 4096)
dev = torch.device("cuda", 0)
torch.manual_seed(1)
d = (torch.rand(P, T, F, device=dev) + 0.5).float()
fl = (torch.rand(P, T, F, device=dev) < 0.3).to(torch.uint8)
out = torch.empty_like(d)
ctx = _cabi.get_context(0, _cabi.torch_stream_handle(0))
if os.environ.get("TC_LIB"):
    _cabi._set_library_for_testing(_cabi.load(os.environ["TC_LIB"]))
lib = _cabi.load()
radii = [(54, 43), (43, 34), (32, 25), (21, 17), (10, 8), (28, 277), (22, 221), (16, 166), (11, 110), (5, 55),
         (8, 43), (6, 34), (5, 25), (3, 17), (1, 8), (0, 43)]
if os.environ.get("RADII"):
    radii = [tuple(int(v) for v in p.split(",")) for p in os.environ["RADII"].split(";")]
for r0, r1 in radii:
    for it in range(2):
        ctx.profile(True)
        ctx.profile_reset()
        _cabi.check(lib.tc_stage_masked_filter(ctx.handle, d.data_ptr(), fl.data_ptr(), P, T, F, r0, r1,
                                               out.data_ptr(), 1))
        torch.cuda.synchronize()
        prof = ctx.profile_read()
    nsamp = P * T * F
    msg = "r=(%3d,%3d)" % (r0, r1)
    for k in ("box_filter_axis0", "box_filter"):
        if k in prof and prof[k][1]:
            ms = prof[k][0]
            msg += "  %s %.3f ms (%d launches, %.1f Gsample/s)" % (k, ms, prof[k][1], nsamp / ms / 1e6)
    print(msg, flush=True)
