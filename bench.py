# -*- coding: utf-8 -*-
"""
bench.py -- throughput of the tricolour flagging hot path on B200.

A "step" is one pass of the FULL default strategy (tricolour/conf/default.yaml:
2x nan/zero flags, 2x static mask, 4 sum_threshold tasks = 8 SumThreshold passes,
2 uvcontsub tasks = 17 cycles, flag_autos, combine_with_input_flags) over one
block of synthetic MeerKAT-shaped windows: `--baselines` baselines (default 64:
the reference's `--baseline-chunks` option, app.py:189, whose default is 16;
batches are sized to the GPU) x 4 correlations x 512 dumps x 4096
channels of BASELINE.json's configs[1].  Every rank owns a different block of
the 2080 baselines (weak scaling, no data-path collective; only the window
statistics are all-reduced once after the timed region).

Printed JSON (one line, rank 0):
  value      GVis/s with the block already resident in HBM (CUDA events)
  e2e        the same metric through StrategyExecutor.apply_strategies with
             pinned HOST buffers: H2D of vis+flags and D2H of the flags inside
             the timed region, every step
  roofline   the dominant kernel family (fused four-pass box filter) against the
             measured HBM peak, plus the whole-strategy figure
  cpu_baseline  the oracle port of the reference's CPU path on a bounded sample,
             ThreadPool over baselines like the reference's dask pool

`--impl reference` times only that CPU arm (no GPU work).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_VIS = 10.0  # 8 B complex64 read + 1 B flag read + 1 B flag written
NANT, NCORR, NTIME, NCHAN = 64, 4, 512, 4096
NBL_TOTAL = 2080


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--baselines", type=int, default=64, help="baselines per step and rank (tricolour --baseline-chunks)")
    ap.add_argument("--ntime", type=int, default=NTIME)
    ap.add_argument("--nchan", type=int, default=NCHAN)
    ap.add_argument("--cpu-baselines", type=int, default=0, help="baselines of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("TC_BENCH_STREAMS", "1")),
                    help="independent blocks flagged concurrently per GPU, one CUDA stream each "
                         "(the reference's dask ThreadPool flags several blocks at once, app.py:266-271)")
    return ap.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel family from the committed ncu
    --set full capture (profiles/rNN_traffic.json, written by profiles/summarize.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), os.path.basename(files[-1])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------- inputs ----
def make_block_torch(nbl, ncorr, T, F, bl0, ubl, device, seed):
    """Synthetic windows generated on the device (same ingredients as
    tests/common.py:make_windows): bandpass x drift x (autos x50) + complex
    noise, persistent / broadband / blob RFI, zeros, NaNs, missing rows, flags."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    x = torch.linspace(0, 1, F, device=device)
    bp = (2.34 - 2.24 * (2 * x - 1) ** 8).to(torch.float32)
    t = torch.arange(T, device=device, dtype=torch.float32)
    drift = 1.0 + 0.02 * torch.sin(2 * np.pi * t / max(T, 2) * 1.3)
    amp = bp[None, None, None, :] * drift[None, None, :, None]
    auto = torch.from_numpy((ubl[bl0:bl0 + nbl, 1] == ubl[bl0:bl0 + nbl, 2])).to(device)
    scale = torch.where(auto, torch.tensor(50.0, device=device), torch.tensor(1.0, device=device))
    amp = amp * scale[:, None, None, None]
    shape = (nbl, ncorr, T, F)
    noise = torch.randn(shape + (2,), generator=g, device=device, dtype=torch.float32) * (0.1 / np.sqrt(2))
    ph = torch.rand((nbl, ncorr, 1, 1), generator=g, device=device) * (2 * np.pi)
    re = amp * torch.cos(ph) + noise[..., 0] * bp
    im = amp * torch.sin(ph) + noise[..., 1] * bp
    del noise
    rs = np.random.RandomState(seed)
    for f in rs.choice(max(F - 4, 1), max(F // 1000, 1), replace=False):
        wband = rs.randint(1, 4)   # persistent RFI: a few narrow bands
        re[:, :, :, f:f + wband] += float(rs.uniform(5, 40)) * 0.1 * bp[f:f + wband]
    for tt in rs.choice(T, max(T // 200, 1), replace=False):
        re[:, :, tt, :] += float(rs.uniform(5, 20)) * 0.1 * bp
    for _ in range(20 * nbl):
        b, c = rs.randint(nbl), rs.randint(ncorr)
        h, w = (min(5, T), min(70, F)) if rs.uniform() < 0.5 else (min(50, T), min(3, F))
        t0, f0 = rs.randint(0, T - h + 1), rs.randint(0, F - w + 1)
        re[b, c, t0:t0 + h, f0:f0 + w] += float(rs.uniform(5, 30)) * 0.234
    re[:, :, :, F // 2 + 3] += 0.2 * 0.234 / np.sqrt(T) * 10
    u = torch.rand(shape, generator=g, device=device)
    zero = u < 0.001
    nan = (u >= 0.001) & (u < 0.002)
    re[zero] = 0
    im[zero] = 0
    re[nan] = float("nan")
    im[nan] = float("nan")
    flags = (u > 0.98)
    miss = torch.rand((nbl, 1, T, 1), generator=g, device=device) < 0.01
    re = torch.where(miss, torch.tensor(float("nan"), device=device), re)
    im = torch.where(miss, torch.tensor(float("nan"), device=device), im)
    flags = flags | miss
    b0 = min(185 * F // 345, F - 1)
    flags[:, :, :, b0:min(b0 + max(F // 70, 1), F)] = True
    vis = torch.complex(re, im)
    return vis.contiguous(), flags.contiguous()


# --------------------------------------------------------------- clocks ------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag.is_set():
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if bit and (r & bit):
                        self.reasons.add(nm)
                time.sleep(0.2)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ----------------------------------------------------------- CPU baseline ----
def cpu_reference_run(nbl, T, F, threads, seed=5):
    """Full default strategy with the oracle port, one task per baseline in a
    ThreadPool (the reference's execution model, app.py:266-271)."""
    from multiprocessing.pool import ThreadPool
    import oracle
    import common
    ubl = common.baselines(NANT)
    ants = common.antenna_layout(NANT)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    strategies = common.default_strategies()
    sel = np.linspace(0, ubl.shape[0] - 1, nbl).astype(int)  # mix of short and long baselines
    blocks = []
    for i, b in enumerate(sel):
        u = ubl[b:b + 1].copy()
        vis, flags = common.make_windows(1, NCORR, T, F, seed=seed + i, ubl=u)
        blocks.append((vis, flags, u))
    oracle.lib()

    def work(blk):
        vis, flags, u = blk
        return common.run_strategies(oracle, strategies, vis, flags, u, ants, masks, cf, cw)

    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPool(threads) as pool:
            outs = pool.map(work, blocks)
    else:
        outs = [work(b) for b in blocks]
    dt = time.perf_counter() - t0
    nvis = nbl * NCORR * T * F
    return nvis / dt / 1e9, dt, float(np.mean([o.mean() for o in outs]))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 32)
    nbl = args.cpu_baselines or threads
    T, F = args.ntime, args.nchan
    vals = []
    for i in range(args.warmup + args.steps):
        # every step is a bounded sample of the workload; warm-up steps use a tiny one
        if i < args.warmup:
            cpu_reference_run(1, 32, 256, 1)
            continue
        v, dt, frac = cpu_reference_run(nbl, T, F, threads)
        vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    dt = float(np.mean([x[1] for x in vals]))
    sample = "%d baselines x %d corr x %d dumps x %d chans (%.1f MVis) per step" % (
        nbl, NCORR, T, F, nbl * NCORR * T * F / 1e6)
    print(json.dumps({
        "impl": "reference", "metric": "visibilities flagged/sec, full default strategy", "value": v,
        "unit": "GVis/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/f64 (complex64 in, u8 flags)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "GVis/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "GVis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": "MeerKAT 64-antenna L-band %d-chan, %d dumps, 4 corr, full default strategy "
                        "(configs[1]); %d of 2080 baselines per step per GPU" % (args.nchan, args.ntime, args.baselines),
            "baselines_per_step": args.baselines, "ncorr": NCORR, "ntime": args.ntime, "nchan": args.nchan,
            "strategy": "default.yaml (12 tasks)", "sharding": "baselines x%d" % world,
            "concurrent_blocks_per_gpu": max(1, args.streams),
            "cache": "inputs (%.0f MiB per step) larger than L2" % (args.baselines * NCORR * args.ntime * args.nchan * 9 / 2 ** 20)}


# ------------------------------------------------------------------- ours ----
def ours(args):
    import torch
    import torch.distributed as dist
    import tricolour_b200 as tb
    from tricolour_b200 import _cabi
    import common

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, F, B = args.ntime, args.nchan, args.baselines
    ubl = common.baselines(NANT)
    ants = common.antenna_layout(NANT)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    strategies = common.default_strategies()
    bl0 = (rank * B) % max(NBL_TOTAL - B, 1)
    my_ubl = ubl[bl0:bl0 + B].copy()
    my_ubl[:, 0] -= my_ubl[0, 0]
    vis, flags = make_block_torch(B, NCORR, T, F, bl0, ubl, dev, 20261019 + rank)
    ex = tb.StrategyExecutor(ants, my_ubl, cf, cw, masks, strategies)
    nvis = B * NCORR * T * F
    ctx = _cabi.get_context(local, _cabi.torch_stream_handle(local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    # extra concurrent blocks (own stream, own library context and workspace)
    S = max(1, args.streams)
    blocks = [(vis, flags, ex)]
    for k in range(1, S):
        blk = (bl0 + k * B) % max(NBL_TOTAL - B, 1)
        v2, f2 = make_block_torch(B, NCORR, T, F, blk, ubl, dev, 20261019 + rank + 1000 * k)
        u2 = ubl[blk:blk + B].copy()
        u2[:, 0] -= u2[0, 0]
        blocks.append((v2, f2, tb.StrategyExecutor(ants, u2, cf, cw, masks, strategies)))
    side = [torch.cuda.Stream(device=dev) for _ in range(S - 1)]

    def run_step():
        """one step: every block once, each on its own stream, enqueued by its own host thread
        (one thread cannot keep two streams fed: its launches block once a stream's queue is full)"""
        cur = torch.cuda.current_stream(dev)
        outs = [None] * S

        def work(k):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(side[k - 1]):
                outs[k] = blocks[k][2].apply_strategies(blocks[k][1], blocks[k][0])

        threads = []
        for k in range(1, S):
            side[k - 1].wait_stream(cur)
            th = threading.Thread(target=work, args=(k,))
            th.start()
            threads.append(th)
        outs[0] = ex.apply_strategies(flags, vis)
        for th in threads:
            th.join()
        for k in range(1, S):
            cur.wait_stream(side[k - 1])
        return outs[0]

    out = None
    for _ in range(args.warmup):
        out = run_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = run_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (ctx.launch_count() - l0) * S
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    # per-kernel-family times: one more step of block 0 alone with the library's event profile on
    ctx.profile(True)
    ctx.profile_reset()
    out = ex.apply_strategies(flags, vis)
    barrier()
    prof = ctx.profile_read()
    ctx.profile(False)
    prof_steps = 1
    nvis_step = nvis * S
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * nvis_step * args.steps / (ms * 1e-3) / 1e9
    flag_frac = float(out.float().mean().item())

    # ---- end to end through the host API (pinned buffers)
    e2e = None
    if not args.no_e2e:
        hv = _cabi.pinned_empty((B, NCORR, T, F), np.complex64)
        hf = _cabi.pinned_empty((B, NCORR, T, F), np.bool_)
        hv[...] = vis.cpu().numpy()
        hf[...] = flags.cpu().numpy()
        for _ in range(1):
            ex.apply_strategies(hf, hv, device=local)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = ex.apply_strategies(hf, hv, device=local)     # one block at a time, nothing overlapped
        torch.cuda.synchronize()
        dt_serial = time.perf_counter() - t0
        for res in ex.apply_strategies_pipelined(((hf, hv) for _ in range(2)), device=local):
            pass                                                # staging buffers exist from here on
        barrier()
        t0 = time.perf_counter()
        res = None
        done_at = []
        for res in ex.apply_strategies_pipelined(((hf, hv) for _ in range(args.steps)), device=local):
            done_at.append(round(time.perf_counter() - t0, 4))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert res.shape == (B, NCORR, T, F)
        e2e = {"value": world * nvis * args.steps / dt / 1e9, "unit": "GVis/s",
               "h2d_bytes_per_step": int(nvis * 9), "d2h_bytes_per_step": int(nvis),
               "serial_value": nvis * args.steps / dt_serial / 1e9, "block_done_s": done_at,
               "api": "tricolour_b200.StrategyExecutor.apply_strategies_pipelined(numpy (flags, vis) blocks): "
                      "apply_strategies of tricolour.apps.tricolour.strat_executor over a sequence of blocks, "
                      "pinned host buffers; every block is uploaded, flagged and downloaded inside the timed "
                      "region, the transfers of neighbouring blocks overlap the flagging"}

    # ---- one small collective: window statistics of the final flags
    st = tb.window_stats(out, my_ubl, cf, ["m%03d" % i for i in range(NANT)], 0, "synthetic", 0)
    if world > 1:
        st = tb.window_statistics.allreduce_window_stats(st)

    if rank == 0:
        peak, peak_src = measured_peaks()
        fam = max(prof.items(), key=lambda kv: kv[1][0])
        fam_ms, fam_n = fam[1]
        total_prof = sum(v[0] for v in prof.values())
        per_launch_ms = fam_ms / max(fam_n, 1)
        achieved = nvis * BYTES_PER_VIS / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic()
        if traffic is not None and B != 16:
            traffic = traffic * B / 16.0     # the capture was taken on a 16-baseline block
        roofline = {
            "bound": "hbm", "kernel": fam[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": nvis * BYTES_PER_VIS, "peak_source": peak_src,
            "launches_per_step": fam_n / prof_steps, "avg_launch_ms": per_launch_ms,
            "share_of_step": fam_ms / max(total_prof, 1e-9),
            "strategy": {"achieved": value / world * BYTES_PER_VIS, "frac": value / world * BYTES_PER_VIS / peak,
                         "note": "whole 12-task strategy: GVis/s x 10 B / measured HBM peak; the chain is "
                                 "FP64/convert-issue bound, not HBM bound (DESIGN.md)"},
            "kernel_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items() if v[1]},
            "kernel_ms_note": "CUDA-event profile of one extra step of one block run alone after the timed region",
        }
        line = {
            "metric": "visibilities flagged/sec, full default strategy", "value": value, "unit": "GVis/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32/f64 (complex64 in, u8 flags)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": sampler.result(),
            "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e,
            "flag_fraction": flag_frac,
        }
        if not args.no_cpu_baseline and world == 1:     # reported at N = 1 only
            cores = os.cpu_count() or 1
            threads = min(cores, 32)
            nblc = args.cpu_baselines or threads
            v, dt, _ = cpu_reference_run(nblc, T, F, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "GVis/s", "cores": threads, "kind": "port",
                "sample": "%d baselines x %d corr x %d dumps x %d chans, one ThreadPool task per baseline, %.1f s"
                          % (nblc, NCORR, T, F, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
