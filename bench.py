# -*- coding: utf-8 -*-
"""
bench.py -- throughput of the tricolour flagging hot path on B200.

`--config N` selects one of BASELINE.json's configurations (default 1, the one the
metric is quoted on); a "step" is one pass of that configuration's path over one
block of synthetic MeerKAT-shaped input:

  0  sum_threshold_flagger, default.yaml step 3, on ONE window (1, 4, 64, 4096)
  1  the FULL default strategy (default.yaml: 2x nan/zero, 2x static mask, 4
     sum_threshold tasks = 8 SumThreshold passes, 2 uvcontsub tasks = 17 cycles,
     flag_autos, combine_with_input_flags) on `--baselines` (64) x 4 x 512 x 4096
  2  32768-channel mode, 256 dumps, polarised: rows -> Stokes Q,U,V intensity +
     any(corr) flags -> windows (bl, 1, 256, 32768) -> the full default strategy
  3  default.yaml tasks 4 -> 7 (uvcontsub, nan/zero reflag, static mask 0~550,
     final_st_very_broad) on (bl, 4, 1024, 4096), cross baselines only
  4  MS row order -> pack_data -> window_stats -> full default strategy ->
     window_stats -> unpack_data + correlation equalisation, baselines sharded over
     the ranks, ONE all-reduce of both statistics inside the end-to-end figure

Every rank owns a different block of baselines (weak scaling, no data-path
collective).  Printed JSON (one line, rank 0):
  value         GVis/s with the block already resident in HBM (CUDA events)
  e2e           the same metric through the host-facing API with pinned HOST buffers:
                H2D of the inputs and D2H of the flags inside the timed region
  roofline      the dominant kernel family against the measured HBM peak, plus the
                whole-strategy figure and the per-family CUDA-event profile
  parity_check  after the timed region: planes of the step's output (an auto, the
                longest and the shortest cross baseline, ...) against the CPU oracle
  cpu_baseline  the oracle port of the reference's CPU path on a bounded sample
  extra         a lightly flagged workload (config 1) beside the default one

`--impl reference` times only the CPU arm (no GPU work): one (baseline,
correlation) plane of the configuration per core and step.
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
NANT, NCORR = 64, 4
NBL_TOTAL = 2080

CONFIGS = {
    0: dict(name="configs[0]", ntime=64, nchan=4096, baselines=1, autos=True,
            what="sum_threshold_flagger (default.yaml step 3) on a single window"),
    1: dict(name="configs[1]", ntime=512, nchan=4096, baselines=64, autos=True,
            what="full default strategy (12 tasks)"),
    2: dict(name="configs[2]", ntime=256, nchan=32768, baselines=32, autos=True,
            what="rows -> polarised intensity (Q,U,V) + any(corr) -> windows (bl,1,T,F) -> full default strategy"),
    3: dict(name="configs[3]", ntime=1024, nchan=4096, baselines=32, autos=False,
            what="default.yaml tasks 4->7: uvcontsub (7 cycles), nan/zero reflag, static mask 0~550, final_st_very_broad"),
    4: dict(name="configs[4]", ntime=512, nchan=4096, baselines=32, autos=True,
            what="MS rows -> pack_data -> window_stats -> full default strategy -> window_stats -> "
                 "unpack_data + corr equalisation; stats all-reduce"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--baselines", type=int, default=0,
                    help="baselines per step and rank (tricolour --baseline-chunks); 0 = the configuration's default")
    ap.add_argument("--ntime", type=int, default=0)
    ap.add_argument("--nchan", type=int, default=0)
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes of the CPU sample (0 = one per core)")
    ap.add_argument("--parity-planes", type=int, default=-1, help="planes checked against the oracle (-1 = auto, 0 = off)")
    ap.add_argument("--clock-interval", type=float, default=0.2, help="seconds between NVML clock samples")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-light", action="store_true", help="skip the lightly flagged workload (config 1 only)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.baselines = args.baselines or cfg["baselines"]
    args.ntime = args.ntime or cfg["ntime"]
    args.nchan = args.nchan or cfg["nchan"]
    return args


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel family from the committed ncu
    --set full capture (profiles/rNN_traffic.json, written by profiles/summarize.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        d = json.load(f)
    d["source"] = os.path.basename(files[-1])
    return d


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------- inputs ----
def make_block_torch(nbl, ncorr, T, F, sub_ubl, device, seed, light=False):
    """Synthetic windows generated on the device (same ingredients as
    tests/common.py:make_windows): bandpass x drift x (autos x50) + complex
    noise, persistent / broadband / blob RFI, zeros, NaNs, missing rows, flags.
    `light`: sparse RFI and input flags only (a quiet band on a quiet day)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    x = torch.linspace(0, 1, F, device=device)
    if light:
        # a band-limited ripple: the 20 / 25 Fourier terms of uvcontsub_flagger can follow it,
        # a steep band edge they cannot (the residual at the edges then gets flagged)
        bp = (2.34 * (1.0 + 0.02 * torch.sin(2 * np.pi * 3 * x))).to(torch.float32)
    else:
        bp = (2.34 - 2.24 * (2 * x - 1) ** 8).to(torch.float32)
    t = torch.arange(T, device=device, dtype=torch.float32)
    drift = 1.0 + 0.02 * torch.sin(2 * np.pi * t / max(T, 2) * 1.3)
    amp = bp[None, None, None, :] * drift[None, None, :, None]
    auto = torch.from_numpy((sub_ubl[:, 1] == sub_ubl[:, 2])).to(device)
    scale = torch.where(auto, torch.tensor(50.0, device=device), torch.tensor(1.0, device=device))
    amp = amp * scale[:, None, None, None]
    shape = (nbl, ncorr, T, F)
    noise = torch.randn(shape + (2,), generator=g, device=device, dtype=torch.float32) * (0.1 / np.sqrt(2))
    ph = torch.rand((nbl, ncorr, 1, 1), generator=g, device=device) * (2 * np.pi)
    re = amp * torch.cos(ph) + noise[..., 0] * bp
    im = amp * torch.sin(ph) + noise[..., 1] * bp
    del noise, amp
    rs = np.random.RandomState(seed)
    nband = 1 if light else max(F // 1000, 1)
    for f in rs.choice(max(F - 4, 1), nband, replace=False):
        wband = rs.randint(1, 4)   # persistent RFI: a few narrow bands
        re[:, :, :, f:f + wband] += float(rs.uniform(5, 40)) * 0.1 * bp[f:f + wband]
    if not light:
        for tt in rs.choice(T, max(T // 200, 1), replace=False):
            re[:, :, tt, :] += float(rs.uniform(5, 20)) * 0.1 * bp
    for _ in range((2 if light else 20) * nbl):
        b, c = rs.randint(nbl), rs.randint(ncorr)
        h, w = (min(5, T), min(70, F)) if rs.uniform() < 0.5 else (min(50, T), min(3, F))
        t0, f0 = rs.randint(0, T - h + 1), rs.randint(0, F - w + 1)
        re[b, c, t0:t0 + h, f0:f0 + w] += float(rs.uniform(5, 30)) * 0.234
    if not light:
        re[:, :, :, F // 2 + 3] += 0.2 * 0.234 / np.sqrt(T) * 10
    u = torch.rand(shape, generator=g, device=device)
    pz = 0.0001 if light else 0.001
    zero = u < pz
    nan = (u >= pz) & (u < 2 * pz)
    re[zero] = 0
    im[zero] = 0
    re[nan] = float("nan")
    im[nan] = float("nan")
    flags = (u > (0.999 if light else 0.98))
    del u, zero, nan
    if not light:
        miss = torch.rand((nbl, 1, T, 1), generator=g, device=device) < 0.01
        re = torch.where(miss, torch.tensor(float("nan"), device=device), re)
        im = torch.where(miss, torch.tensor(float("nan"), device=device), im)
        flags = flags | miss
        b0 = min(185 * F // 345, F - 1)
        flags[:, :, :, b0:min(b0 + max(F // 70, 1), F)] = True
    vis = torch.complex(re, im)
    return vis.contiguous(), flags.contiguous()


def windows_to_rows(win):
    """(bl, corr, T, F) -> MS row order (row = t * nbl + bl, chan, corr)"""
    nbl, ncorr, T, F = win.shape
    return win.permute(2, 0, 3, 1).reshape(T * nbl, F, ncorr).contiguous()


# --------------------------------------------------------------- clocks ------
class ClockSampler(threading.Thread):
    def __init__(self, index, interval=0.2):
        super().__init__(daemon=True)
        self.index = index
        self.interval = interval
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
                self.stop_flag.wait(self.interval)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ------------------------------------------------------------ the workloads --
class Workload(object):
    """One BASELINE.json configuration: device-resident step, host-facing e2e call,
    the oracle's version of a plane (parity check and CPU arm)."""

    def __init__(self, args, rank=0, world=1):
        import common
        self.args, self.rank, self.world = args, rank, world
        self.cfg = CONFIGS[args.config]
        self.idx = args.config
        self.T, self.F, self.B = args.ntime, args.nchan, args.baselines
        self.ubl_all = common.baselines(NANT, autos=self.cfg["autos"])
        self.ants = common.antenna_layout(NANT)
        self.cf, self.cw = common.channels(self.F)
        self.masks = common.synthetic_static_mask(self.cf)
        all_s = common.default_strategies()
        self.kw3 = dict(common.DEFAULT_STRATEGY_KW["background_flags"])
        self.strategies = {0: None, 1: all_s, 2: all_s, 3: all_s[3:7], 4: all_s}[self.idx]
        nbl_total = self.ubl_all.shape[0]
        self.bl0 = (rank * self.B) % max(nbl_total - self.B, 1)
        self.sub = self.ubl_all[self.bl0:self.bl0 + self.B].copy()
        self.sub[:, 0] = np.arange(self.sub.shape[0])
        # visibilities per step and rank: every input correlation counts
        self.nvis = self.B * NCORR * self.T * self.F
        self.names = ["m%03d" % i for i in range(NANT)]

    # -- description
    def config(self):
        a = self.args
        in_mib = self.nvis * 9 / 2 ** 20
        return {"workload": "%s: MeerKAT 64-antenna L-band %d-chan, %d dumps, 4 corr; %s; %d baselines per step "
                            "per GPU" % (self.cfg["name"], a.nchan, a.ntime, self.cfg["what"], self.B),
                "config_index": self.idx, "baselines_per_step": self.B, "ncorr": NCORR, "ntime": a.ntime,
                "nchan": a.nchan, "sharding": "baselines x%d" % self.world,
                "cache": "inputs (%.0f MiB per step) %s L2; the working set of a step (%.1f GB) is far larger"
                         % (in_mib, "larger than" if in_mib > 126 else "SMALLER than", self.nvis * 50 / 1e9)}

    # -- device-resident
    def setup_device(self, dev, light=False):
        import torch
        import tricolour_b200 as tb
        self.dev = dev
        seed = 20261018 + self.idx + 100 * self.rank + (7 if light else 0)
        ncorr = NCORR
        vis, flags = make_block_torch(self.B, ncorr, self.T, self.F, self.sub, dev, seed, light=light)
        self.h2d = self.nvis * 9
        self.d2h = self.nvis
        if self.idx == 0:
            self.vis, self.flags = vis, flags
        elif self.idx in (1, 3):
            self.vis, self.flags = vis, flags
            masks = [self.masks[0][::3]] if light else self.masks
            self.ex = tb.StrategyExecutor(self.ants, self.sub, self.cf, self.cw, masks, self.strategies)
        else:
            # MS row order (time-major), the windows themselves are dropped
            self.rows = windows_to_rows(vis)
            self.rflags = windows_to_rows(flags)
            del vis, flags
            torch.cuda.empty_cache()
            self.a1 = np.tile(self.sub[:, 1], self.T).astype(np.int32)
            self.a2 = np.tile(self.sub[:, 2], self.T).astype(np.int32)
            self.tinv = np.repeat(np.arange(self.T), self.B)
            self.ex = tb.StrategyExecutor(self.ants, self.sub, self.cf, self.cw, self.masks, self.strategies)
            smap = tb.stokes_corr_map([9, 10, 11, 12])
            self.pol = tuple(v for k, v in smap.items() if k != 'I')
            if self.idx == 2:
                self.d2h = self.nvis // NCORR          # one-correlation flag windows come back
            self.layout = tb.StatsLayout(self.names, self.ubl_all, ["synthetic"], [0], {0: self.cf})

    def step_device(self):
        """one step on resident inputs; returns the step's flag output (device)"""
        import tricolour_b200 as tb
        if self.idx == 0:
            return tb.sum_threshold_flagger(self.vis, self.flags, **self.kw3)
        if self.idx in (1, 3):
            return self.ex.apply_strategies(self.flags, self.vis)
        if self.idx == 2:
            vw, fw = tb.packing.pack_polarised(self.tinv, self.sub, self.a1, self.a2, self.rows, self.rflags,
                                               self.T, self.pol)
            return self.ex.apply_strategies(fw, vw)
        # 4: rows -> windows -> stats -> strategy -> stats -> rows
        vw, fw = tb.pack_data(self.tinv, self.sub, self.a1, self.a2, self.rows, self.rflags, self.T)
        st0 = tb.window_stats(fw, self.ubl_all[self.bl0:self.bl0 + self.B], self.cf, self.names, 0, "synthetic", 0)
        out = self.ex.apply_strategies(fw, vw)
        st1 = tb.window_stats(out, self.ubl_all[self.bl0:self.bl0 + self.B], self.cf, self.names, 0, "synthetic", 0)
        self.last_stats = (st0, st1)
        self.last_windows = out
        return tb.packing.unpack_flags_equalised(self.a1, self.a2, self.tinv, self.sub, out)

    # -- end to end through the host-facing API
    def setup_host(self):
        from tricolour_b200 import _cabi
        if self.idx in (0, 1, 3):
            self.hv = _cabi.pinned_empty(tuple(self.vis.shape), np.complex64)
            self.hf = _cabi.pinned_empty(tuple(self.flags.shape), np.bool_)
            self.hv[...] = self.vis.cpu().numpy()
            self.hf[...] = self.flags.cpu().numpy()
        else:
            self.hv = _cabi.pinned_empty(tuple(self.rows.shape), np.complex64)
            self.hf = _cabi.pinned_empty(tuple(self.rflags.shape), np.bool_)
            self.hv[...] = self.rows.cpu().numpy()
            self.hf[...] = self.rflags.cpu().numpy()

    def e2e_api(self):
        return {0: "tricolour_b200.sum_threshold_flagger(numpy vis, numpy flags, **step-3 kwargs): the library stages "
                   "the block to the device and the flags back, per call",
                1: "tricolour_b200.StrategyExecutor.apply_strategies_pipelined(numpy (flags, vis) blocks): "
                   "apply_strategies of tricolour.apps.tricolour.strat_executor over a sequence of blocks, pinned host "
                   "buffers; every block is uploaded, flagged and downloaded inside the timed region, the transfers "
                   "of neighbouring blocks overlap the flagging",
                2: "StrategyExecutor.apply_strategies_pipelined(numpy (row flags, rows) blocks, pre=packing.pack_polarised): "
                   "host rows -> device, Stokes + pack, strategy on the resident windows, flag windows -> host; "
                   "the transfers of neighbouring blocks overlap the flagging",
                3: "StrategyExecutor.apply_strategies_pipelined(numpy (flags, vis) blocks), tasks 4->7",
                4: "StrategyExecutor.apply_strategies_pipelined(numpy (row flags, rows) blocks, pre=pack_data + "
                   "window_stats, post=window_stats + unpack_flags_equalised): host rows -> device -> host row flags, "
                   "allreduce_window_stats((original, final)) per block; transfers of neighbouring blocks overlap "
                   "the flagging",
                }[self.idx]

    def run_e2e(self, steps):
        """`steps` blocks from pinned HOST buffers to HOST results; returns seconds"""
        import torch
        import tricolour_b200 as tb
        dev = self.dev
        res = None
        t0 = time.perf_counter()
        if self.idx == 0:
            for _ in range(steps):
                res = tb.sum_threshold_flagger(self.hv, self.hf, **self.kw3)
        elif self.idx in (1, 3):
            for res in self.ex.apply_strategies_pipelined(((self.hf, self.hv) for _ in range(steps)), device=dev.index):
                pass
        else:
            # rows -> windows and windows -> rows run inside the executor's pipeline (pre / post hooks on the
            # flagging stream): the rows of block i+1 are uploaded and the row flags of block i-1 downloaded
            # while block i is being flagged
            ub = self.ubl_all[self.bl0:self.bl0 + self.B]
            stats = []

            def pre(rfl, rows):
                if self.idx == 2:
                    vw, fw = tb.packing.pack_polarised(self.tinv, self.sub, self.a1, self.a2, rows, rfl, self.T, self.pol)
                    return fw, vw
                vw, fw = tb.pack_data(self.tinv, self.sub, self.a1, self.a2, rows, rfl, self.T)
                stats.append([tb.window_stats(fw, ub, self.cf, self.names, 0, "synthetic", 0)])
                return fw, vw

            def post(out):
                stats[-1].append(tb.window_stats(out, ub, self.cf, self.names, 0, "synthetic", 0))
                return tb.packing.unpack_flags_equalised(self.a1, self.a2, self.tinv, self.sub, out)

            blocks = ((self.hf, self.hv) for _ in range(steps))
            for res in self.ex.apply_strategies_pipelined(blocks, device=dev.index, pre=pre,
                                                          post=post if self.idx == 4 else None):
                if self.idx == 4:
                    self.reduced = tb.allreduce_window_stats(tuple(stats.pop(0)), self.layout)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        self.e2e_result_shape = tuple(res.shape)
        return dt

    # -- the oracle's version of some planes of the step (parity check)
    def parity(self, out, nplanes):
        """compares `nplanes` planes of the step's output with the CPU oracle"""
        import torch
        import oracle
        import common
        t0 = time.perf_counter()
        if self.idx == 0:
            vis, flags = self.vis.cpu().numpy(), self.flags.cpu().numpy()
            nb = max(1, min(self.B, (nplanes + NCORR - 1) // NCORR))
            want = oracle.sum_threshold_flagger(vis[:nb], flags[:nb], nthreads=os.cpu_count() or 1, **self.kw3)
            got = out[:nb].cpu().numpy()
            nd, n, picks = int((got != want).sum()), int(want.size), list(range(nb))
            nfl = int(want.sum())
        else:
            picks = common.pick_baselines(self.sub, self.ants, max(1, min(self.B, nplanes)))
            if self.idx in (2, 4) and len(picks) == 1:
                # one baseline only: an auto-correlation ends fully flagged (flag_autos), which checks
                # nothing; take the longest cross baseline of the block instead
                picks = common.pick_baselines(self.sub, self.ants, min(self.B, 2))[-1:]
            corr_of = {b: (k % NCORR) for k, b in enumerate(picks)}
            if nplanes >= 2 * len(picks):
                corr_all = True
            else:
                corr_all = False
            nd = n = nfl = 0
            jobs = []
            for b in picks:
                su = self.sub[b:b + 1].copy()
                su[:, 0] = 0
                if self.idx in (1, 3):
                    cs = list(range(NCORR)) if corr_all else [corr_of[b]]
                    for c in cs:
                        jobs.append((b, c, su, self.vis[b:b + 1, c:c + 1].cpu().numpy(),
                                     self.flags[b:b + 1, c:c + 1].cpu().numpy()))
                else:
                    rows = self.rows.view(self.T, self.B, self.F, NCORR)[:, b].contiguous().cpu().numpy()
                    rfl = self.rflags.view(self.T, self.B, self.F, NCORR)[:, b].contiguous().cpu().numpy()
                    jobs.append((b, None, su, rows, rfl))

            def work(job):
                b, c, su, v, f = job
                if self.idx in (1, 3):
                    return common.run_strategies(oracle, self.strategies, v, f, su, self.ants, self.masks,
                                                 self.cf, self.cw)
                a1 = np.full(self.T, su[0, 1], np.int32)
                a2 = np.full(self.T, su[0, 2], np.int32)
                tinv = np.arange(self.T)
                if self.idx == 2:
                    pi = oracle.polarised_intensity(v, self.pol)
                    vw, fw = oracle.pack_data(tinv, su, a1, a2, pi, f.any(axis=2, keepdims=True), self.T)
                    return common.run_strategies(oracle, self.strategies, vw, fw, su, self.ants, self.masks,
                                                 self.cf, self.cw)
                vw, fw = oracle.pack_data(tinv, su, a1, a2, v, f, self.T)
                fl = common.run_strategies_planes(oracle, self.strategies, vw, fw, su, self.ants, self.masks,
                                                  self.cf, self.cw, threads=1)
                rows = oracle.unpack_data(a1, a2, tinv, su, fl)
                return np.broadcast_to(rows.any(axis=2, keepdims=True), rows.shape)

            from multiprocessing.pool import ThreadPool
            with ThreadPool(max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
                wants = pool.map(work, jobs)
            for (b, c, su, v, f), want in zip(jobs, wants):
                if self.idx in (1, 3):
                    got = out[b:b + 1, c:c + 1].cpu().numpy()
                elif self.idx == 2:
                    got = out[b:b + 1].cpu().numpy()
                else:
                    got = out.view(self.T, self.B, self.F, NCORR)[:, b].cpu().numpy()
                nd += int((got != want).sum())
                n += int(want.size)
                nfl += int(want.sum())
        return {"planes": (len(picks) * NCORR if self.idx == 0 else len(jobs)), "baselines": [int(p) for p in picks],
                "samples": n, "ndiff": nd, "nflags": nfl, "oracle_s": round(time.perf_counter() - t0, 1),
                "what": "the step's output against oracle/ (CPU restatement of the reference) on the same inputs"}


# ----------------------------------------------------------- CPU arm ---------
def cpu_blocks(args, nplanes, seed=5):
    """host inputs of `nplanes` independent (baseline, correlation) planes of the
    configuration, mixed short / long baselines; built ONCE, outside the timed steps"""
    import common
    cfg = CONFIGS[args.config]
    T, F = args.ntime, args.nchan
    ubl = common.baselines(NANT, autos=cfg["autos"])
    sel = np.linspace(0, ubl.shape[0] - 1, nplanes).astype(int)
    blocks = []
    for i, b in enumerate(sel):
        u = ubl[b:b + 1].copy()
        u[:, 0] = 0
        if args.config == 0:
            vis, flags = common.make_windows(1, NCORR, T, F, seed=seed + i, ubl=u)
        elif args.config in (1, 3):
            vis, flags = common.make_windows(1, 1, T, F, seed=seed + i, ubl=u)
        else:
            v4, f4 = common.make_windows(1, NCORR, T, F, seed=seed + i, ubl=u)
            vis = np.ascontiguousarray(v4[0].transpose(1, 2, 0))       # rows of one baseline: (T, F, corr)
            flags = np.ascontiguousarray(f4[0].transpose(1, 2, 0))
        blocks.append((vis, flags, u))
    return blocks


def cpu_run(args, blocks, threads):
    """one step of the CPU arm: every block through the configuration's path with the
    oracle port, one ThreadPool task per block (the reference's execution model,
    app.py:266-271).  Returns (GVis/s, seconds, flagged fraction)."""
    from multiprocessing.pool import ThreadPool
    import oracle
    import common
    T, F = args.ntime, args.nchan
    ants = common.antenna_layout(NANT)
    cf, cw = common.channels(F)
    masks = common.synthetic_static_mask(cf)
    all_s = common.default_strategies()
    strategies = {0: None, 1: all_s, 2: all_s, 3: all_s[3:7], 4: all_s}[args.config]
    kw3 = dict(common.DEFAULT_STRATEGY_KW["background_flags"])
    smap = oracle.stokes_corr_map([9, 10, 11, 12])
    pol = tuple(v for k, v in smap.items() if k != 'I')
    oracle.lib()

    def work(blk):
        vis, flags, u = blk
        if args.config == 0:
            return oracle.sum_threshold_flagger(vis, flags, **kw3)
        if args.config in (1, 3):
            return common.run_strategies(oracle, strategies, vis, flags, u, ants, masks, cf, cw)
        a1 = np.full(T, u[0, 1], np.int32)
        a2 = np.full(T, u[0, 2], np.int32)
        tinv = np.arange(T)
        if args.config == 2:
            pi = oracle.polarised_intensity(vis, pol)
            vw, fw = oracle.pack_data(tinv, u, a1, a2, pi, flags.any(axis=2, keepdims=True), T)
            return common.run_strategies(oracle, strategies, vw, fw, u, ants, masks, cf, cw)
        vw, fw = oracle.pack_data(tinv, u, a1, a2, vis, flags, T)
        oracle.window_counts(fw, u, cf, NANT)
        fl = common.run_strategies(oracle, strategies, vw, fw, u, ants, masks, cf, cw)
        oracle.window_counts(fl, u, cf, NANT)
        rows = oracle.unpack_data(a1, a2, tinv, u, fl)
        return np.broadcast_to(rows.any(axis=2, keepdims=True), rows.shape)

    t0 = time.perf_counter()
    if threads > 1 and len(blocks) > 1:
        with ThreadPool(min(threads, len(blocks))) as pool:
            outs = pool.map(work, blocks)
    else:
        outs = [work(b) for b in blocks]
    dt = time.perf_counter() - t0
    nvis = sum(int(b[0].size) for b in blocks)
    return nvis / dt / 1e9, dt, float(np.mean([o.mean() for o in outs]))


def cpu_sample_desc(args, blocks, dt):
    per = blocks[0][0].shape
    what = {0: "windows (1, 4, %d, %d)" % (args.ntime, args.nchan),
            1: "(baseline, correlation) planes of %d x %d" % (args.ntime, args.nchan),
            2: "baselines of (%d dumps, %d chans, 4 corr) rows" % (args.ntime, args.nchan),
            3: "(baseline, correlation) planes of %d x %d" % (args.ntime, args.nchan),
            4: "baselines of (%d dumps, %d chans, 4 corr) rows" % (args.ntime, args.nchan)}[args.config]
    return "%d %s (%.1f MVis), one ThreadPool task each, %.1f s per step" % (
        len(blocks), what, sum(int(b[0].size) for b in blocks) / 1e6, dt), per


def reference_arm(args):
    """The reference's CPU implementation of the path (the oracle port: numba-compiled
    code cannot travel to the GPU box; port and numba were timed side by side on
    (1,4,64,4096) step 3 in the build container: 1.84 s vs 1.90 s, identical flags) on
    all host cores: every step flags one plane of the configuration per core.  The
    inputs are built once, before the steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nplanes = args.cpu_planes or cores
    t_in = time.perf_counter()
    blocks = cpu_blocks(args, nplanes)
    t_in = time.perf_counter() - t_in
    # warm-up: a tiny case (library load, first-touch of the numpy paths), not the full sample
    small = argparse.Namespace(**vars(args))
    small.ntime, small.nchan = 32, 256
    tiny = cpu_blocks(small, 1)
    for _ in range(max(args.warmup, 1)):
        cpu_run(small, tiny, 1)
    vals = []
    budget_s = float(os.environ.get("TC_REFERENCE_BUDGET_S", "420"))
    t_start = time.perf_counter()
    for i in range(args.steps):
        v, dt, frac = cpu_run(args, blocks, cores)
        vals.append((v, dt))
        # the run must end within a few minutes whatever K is: later steps repeat the same
        # sample, so stop measuring once the budget is spent and report the steps done
        if time.perf_counter() - t_start + dt > budget_s and i + 1 < args.steps:
            break
    v = float(np.mean([x[0] for x in vals]))
    dt = float(np.mean([x[1] for x in vals]))
    sample, _ = cpu_sample_desc(args, blocks, dt)
    w = Workload(args, 0, 1)
    cfgd = w.config()        # the workload the sample is drawn from: identical to the GPU arm's
    print(json.dumps({
        "impl": "reference", "metric": "visibilities flagged/sec, " + CONFIGS[args.config]["what"], "value": v,
        "unit": "GVis/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "steps_measured": len(vals), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64 (complex64 in, u8 flags)", "data": "synthetic",
        "config": cfgd, "reference_sample_per_step": sample, "input_build_s": round(t_in, 1),
        "cpu_baseline": {"value": v, "unit": "GVis/s", "cores": min(cores, nplanes), "kind": "port", "sample": sample,
                         "equivalence": "oracle port vs numba reference on (1,4,64,4096) step 3: 1.84 s vs 1.90 s, "
                                        "identical flags (VERDICT r01; tests/test_oracle_vs_reference.py)"},
        "e2e": {"value": v, "unit": "GVis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------- ours ----
def ours(args):
    import torch
    import torch.distributed as dist
    from tricolour_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = Workload(args, rank, world)
    w.setup_device(dev)
    ctx = _cabi.get_context(local, _cabi.torch_stream_handle(local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(work, steps, warmup):
        out = None
        for _ in range(warmup):
            out = work()
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = work()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count() - l0, out

    # ---- device-resident throughput
    sampler = ClockSampler(local, args.clock_interval)
    # the sampler covers the timed region (warm-up excluded as far as a thread can tell)
    for _ in range(args.warmup):
        w.step_device()
    barrier()
    sampler.start()
    ms, launches, out = timed(w.step_device, args.steps, 0)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    value = world * w.nvis * args.steps / (ms * 1e-3) / 1e9
    flag_frac = float(out.float().mean().item())

    # per-kernel-family times: one more step alone with the library's event profile on
    ctx.profile(True)
    ctx.profile_reset()
    out = w.step_device()
    barrier()
    prof = ctx.profile_read()
    ctx.profile(False)

    # ---- parity of the step's output against the oracle (rank 0, after the timed region)
    parity = None
    npar = args.parity_planes if args.parity_planes >= 0 else {0: 4, 1: 4, 2: 1, 3: 2, 4: 1}[args.config]
    if npar > 0 and rank == 0:
        parity = w.parity(out, npar)
    del out

    # ---- end to end through the host API (pinned buffers)
    e2e = None
    if not args.no_e2e:
        w.setup_host()
        w.run_e2e(2 if args.config in (1, 3) else 1)           # staging buffers exist from here on
        barrier()
        dt = w.run_e2e(args.steps)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * w.nvis * args.steps / dt / 1e9, "unit": "GVis/s",
               "h2d_bytes_per_step": int(w.h2d), "d2h_bytes_per_step": int(w.d2h),
               "result_shape": list(w.e2e_result_shape), "api": w.e2e_api()}
        if args.config == 4:
            red0, red1 = w.reduced
            e2e["stats_allreduce"] = {"collectives_per_block": 1 if world > 1 else 0,
                                      "vector_int64": 2 * w.layout.size,
                                      "flagged_original": int(red0._counts_per_field["synthetic"]),
                                      "flagged_final": int(red1._counts_per_field["synthetic"]),
                                      "size": int(red1._size_per_field["synthetic"])}

    extra = {}
    # ---- config 0: the same call from a pool of host threads, one window each (how dask's
    # ThreadPool drives the np_* functions, app.py:266-271): every thread has its own context and
    # stream, the latency-bound launch chains of different windows overlap on the device
    if args.config == 0 and not args.no_e2e and world == 1:
        from concurrent.futures import ThreadPoolExecutor
        import tricolour_b200 as tb
        nthreads = 16
        rounds = max(2, min(args.steps, 4))
        wins = [(w.hv.copy(), w.hf.copy()) for _ in range(nthreads)]

        def one(i):
            return tb.sum_threshold_flagger(wins[i][0], wins[i][1], **w.kw3)

        with ThreadPoolExecutor(nthreads) as pool:
            ref_out = list(pool.map(one, range(nthreads)))             # warm-up: contexts, arenas
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(rounds):
                outs = list(pool.map(one, range(nthreads)))
            torch.cuda.synchronize()
            dtc = time.perf_counter() - t0
        same = all(np.array_equal(o, ref_out[0]) for o in outs)
        extra["concurrent_windows"] = {
            "value": nthreads * rounds * w.nvis / dtc / 1e9, "unit": "GVis/s", "host_threads": nthreads,
            "windows": nthreads * rounds, "identical_outputs": bool(same),
            "what": "sum_threshold_flagger(numpy window) called from %d host threads at once, host buffers in and "
                    "out (the way dask's ThreadPool calls it); the single-window figures above are latency-bound"
                    % nthreads}

    # ---- a lightly flagged workload beside the default one (config 1)
    if args.config == 1 and not args.no_light:
        del w.vis, w.flags
        torch.cuda.empty_cache()
        wl = Workload(args, rank, world)
        wl.sub = wl.ubl_all[(wl.ubl_all[:, 1] != wl.ubl_all[:, 2])][rank * wl.B:(rank + 1) * wl.B].copy()
        wl.sub[:, 0] = np.arange(wl.sub.shape[0])
        wl.setup_device(dev, light=True)
        in_frac = float(wl.flags.float().mean().item())
        lms, _, lout = timed(wl.step_device, max(1, min(args.steps, 3)), 1)
        extra["light_workload"] = {
            "value": world * wl.nvis * max(1, min(args.steps, 3)) / (lms * 1e-3) / 1e9, "unit": "GVis/s",
            "flag_fraction_in": in_frac, "flag_fraction_out": float(lout.float().mean().item()),
            "what": "same configuration on a quiet sky: cross baselines only, a band-limited bandpass ripple, "
                    "sparse RFI, 0.1 % input flags"}
        del lout, wl

    if rank == 0:
        peak, peak_src = measured_peaks()
        fam = max(prof.items(), key=lambda kv: kv[1][0])
        fam_ms, fam_n = fam[1]
        total_prof = sum(v[0] for v in prof.values())
        per_launch_ms = fam_ms / max(fam_n, 1)
        achieved = w.nvis * BYTES_PER_VIS / (per_launch_ms * 1e-3) / 1e9
        tr = measured_traffic()
        traffic = None
        if tr is not None and tr.get("dram_bytes_per_launch") is not None:
            # the capture's block size is recorded with it; traffic scales with the planes of a launch
            traffic = tr["dram_bytes_per_launch"] * (w.nvis / float(tr.get("nvis_per_launch", 16 * 4 * 512 * 4096)))
        roofline = {
            "bound": "hbm", "kernel": fam[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": tr.get("source") if tr else None,
            "algorithmic_bytes_per_launch": w.nvis * BYTES_PER_VIS, "peak_source": peak_src,
            "launches_per_step": fam_n, "avg_launch_ms": per_launch_ms,
            "share_of_step": fam_ms / max(total_prof, 1e-9),
            "strategy": {"achieved": value / world * BYTES_PER_VIS, "frac": value / world * BYTES_PER_VIS / peak,
                         "note": "whole path of the step: GVis/s x 10 B / measured HBM peak; the SumThreshold chain "
                                 "is FP64/convert-issue bound, not HBM bound (DESIGN.md)"},
            "kernel_ms_per_step": {k: v[0] for k, v in prof.items() if v[1]},
            "kernel_launches_per_step": {k: v[1] for k, v in prof.items() if v[1]},
            "kernel_ms_note": "CUDA-event profile of one extra step run alone after the timed region",
        }
        line = {
            "metric": "visibilities flagged/sec, " + CONFIGS[args.config]["what"], "value": value, "unit": "GVis/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32/f64 (complex64 in, u8 flags)", "data": "synthetic",
            "config": w.config(), "clocks": sampler.result(),
            "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e,
            "flag_fraction": flag_frac, "parity_check": parity, "extra": extra,
        }
        if not args.no_cpu_baseline and world == 1:     # reported at N = 1 only
            cores = os.cpu_count() or 1
            nplanes = args.cpu_planes or cores
            blocks = cpu_blocks(args, nplanes)
            v, dtc, _ = cpu_run(args, blocks, cores)
            sample, _ = cpu_sample_desc(args, blocks, dtc)
            line["cpu_baseline"] = {"value": v, "unit": "GVis/s", "cores": min(cores, nplanes), "kind": "port",
                                    "sample": sample}
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
