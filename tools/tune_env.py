# -*- coding: utf-8 -*-
"""A/B of a runtime knob of the library (an environment variable read at launch
time) on the bench workload: one process, one resident block, for every value
one warm-up step and `--steps` timed steps of the full default strategy plus the
per-family CUDA-event profile of one more step.

    python tools/tune_env.py --env TC_BRK_SLICE --values 16384,32768,65536
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", required=True)
    ap.add_argument("--values", required=True, help="comma separated; an empty entry unsets the variable")
    ap.add_argument("--baselines", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--families", default="chunk_select,box_filter,box_filter_axis0,line_median,st_scan")
    args = ap.parse_args()
    import torch
    import bench
    import common
    import tricolour_b200 as tb
    from tricolour_b200 import _cabi
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    T, F, B = bench.NTIME, bench.NCHAN, args.baselines
    ubl = common.baselines(bench.NANT)
    cf, cw = common.channels(F)
    my_ubl = ubl[:B].copy()
    vis, flags = bench.make_block_torch(B, bench.NCORR, T, F, 0, ubl, dev, 20261019)
    ex = tb.StrategyExecutor(common.antenna_layout(bench.NANT), my_ubl, cf, cw,
                             common.synthetic_static_mask(cf), common.default_strategies())
    ctx = _cabi.get_context(0, _cabi.torch_stream_handle(0))
    ref = None
    for val in args.values.split(","):
        if val == "":
            os.environ.pop(args.env, None)
        else:
            os.environ[args.env] = val
        out = ex.apply_strategies(flags, vis)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = ex.apply_strategies(flags, vis)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        ctx.profile(True)
        ctx.profile_reset()
        out = ex.apply_strategies(flags, vis)
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        same = True if ref is None else bool(torch.equal(ref, out))
        if ref is None:
            ref = out.clone()
        fam = {k: round(prof[k][0], 2) for k in args.families.split(",") if k in prof}
        print(json.dumps({"env": args.env, "value": val, "ms_per_step": round(ms, 2),
                          "same_flags_as_first": same, "family_ms": fam}), flush=True)


if __name__ == "__main__":
    main()
