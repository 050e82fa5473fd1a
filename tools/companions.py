# -*- coding: utf-8 -*-
"""Achieved HBM bandwidth of the HBM-bound companions of the SumThreshold core
(SURVEY 8a rows F1-F3, K2, P1, P2, W1 and the fused N2 forms) on ALGORITHMIC bytes:
every kernel is called through the Python boundary on device-resident arrays of
configs[1] block size and timed with CUDA events (best of `--reps`).  Run it under
`ncu --set full -k regex:...` for the DRAM counters of the same launches.

    python tools/companions.py [--baselines 32] [--reps 5] > gpurun_out/companions.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--baselines", type=int, default=32)
    ap.add_argument("--ntime", type=int, default=512)
    ap.add_argument("--nchan", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import tricolour_b200 as tb
    import common
    dev = torch.device("cuda", 0)
    B, T, F, C = args.baselines, args.ntime, args.nchan, 4
    peak = 6537.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    ubl = common.baselines(64)[:B].copy()
    ubl[:, 0] = np.arange(B)
    ants = common.antenna_layout(64)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    n = B * C * T * F
    vis = torch.view_as_complex(torch.randn((B, C, T, F, 2), generator=g, device=dev, dtype=torch.float32))
    flags = torch.rand((B, C, T, F), generator=g, device=dev) < 0.1
    rows = vis.permute(2, 0, 3, 1).reshape(T * B, F, C).contiguous()
    rflags = flags.permute(2, 0, 3, 1).reshape(T * B, F, C).contiguous()
    a1 = np.tile(ubl[:, 1], T).astype(np.int32)
    a2 = np.tile(ubl[:, 2], T).astype(np.int32)
    tinv = np.repeat(np.arange(T), B)
    smap = tb.stokes_corr_map([9, 10, 11, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    win1 = torch.rand((B, 1, T, F), generator=g, device=dev) < 0.1

    def timeit(fn):
        best = None
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
            del r
        return best

    cases = [
        ("F1 flag_nans_and_zeros", lambda: tb.flag_nans_and_zeros(vis, flags), n * 10),
        ("F2 flag_autos", lambda: tb.flag_autos(flags, [ubl]), n * 2),
        ("F3 apply_static_mask", lambda: tb.apply_static_mask(flags, ubl, ants, masks, cf, cw, "or", "0~550"), n * 2),
        ("flags_or", lambda: tb.strategy._flags_or(flags, flags), n * 3),
        ("K2 polarised_intensity", lambda: tb.polarised_intensity(rows, pol), (n // C) * (8 * C + 8)),
        ("P1 pack_data", lambda: tb.pack_data(tinv, ubl, a1, a2, rows, rflags, T), n * 18),
        ("N2 pack_polarised (K2 + any + P1)", lambda: tb.packing.pack_polarised(tinv, ubl, a1, a2, rows, rflags, T, pol),
         (n // C) * (9 * C + 9)),
        ("P2 unpack_data (flags)", lambda: tb.unpack_data(a1, a2, tinv, ubl, flags), n * 2),
        ("P2 unpack_data (vis)", lambda: tb.unpack_data(a1, a2, tinv, ubl, vis), n * 16),
        ("N2 unpack_flags_equalised", lambda: tb.packing.unpack_flags_equalised(a1, a2, tinv, ubl, flags), n * 2),
        ("N2 unpack 1 -> 4 corr", lambda: tb.packing.unpack_flags_equalised(a1, a2, tinv, ubl, win1, ncorr_out=4),
         (n // C) * 5),
        ("W1 window_stats counts", lambda: tb.window_statistics._counts(flags), n),
    ]
    from tricolour_b200 import _cabi
    ctx = _cabi.get_context(0, _cabi.torch_stream_handle(0))
    out = []
    for name, fn, nbytes in cases:
        fn()                     # warm-up (arena growth, first launch)
        ms = timeit(fn)
        # the kernels alone: the library's own CUDA events around every launch of the call
        ctx.profile(True)
        ctx.profile_reset()
        for _ in range(args.reps):
            r = fn()
            del r
        ctx.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        kms = sum(v[0] for v in prof.values()) / args.reps
        nk = sum(v[1] for v in prof.values()) // args.reps
        gbs = nbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        rec = {"kernel": name, "kernel_ms": round(kms, 4), "launches": int(nk), "algorithmic_bytes": int(nbytes),
               "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
               "api_ms": round(ms, 4), "api_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    print(json.dumps({"summary": "companions", "block": [B, C, T, F], "hbm_peak_gbs": peak,
                      "note": "kernel_ms: the library's CUDA events around the launches of one call (device time "
                              "of the kernels); api_ms: CUDA events around the Python call on device-resident "
                              "arrays, best of reps (host-side table work of pack / unpack / static mask included)",
                      "results": out}))


if __name__ == "__main__":
    main()
