# -*- coding: utf-8 -*-
"""
Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/tricolour, imported through oracle/ref_loader.py) on small
seeded inputs.  Run in the build container (the GPU box has no reference):

    python tests/golden/make_golden.py

Every file stores the inputs next to the reference outputs so that the
fixtures stay valid even if the generators in tests/common.py change.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import common  # noqa: E402


def main():
    F, S, P, W = ref_loader.load()
    rs = np.random.RandomState(20261018)

    # --- sum_threshold_flagger, three parameter sets of default.yaml (steps 3, 7, 9)
    vis, flags = common.make_windows(2, 2, 48, 320, seed=11)
    out = {"vis": vis, "flags": flags}
    for name in ("background_flags", "final_st_very_broad", "final_st_narrow"):
        kw = dict(common.DEFAULT_STRATEGY_KW[name])
        if name == "background_flags":
            kw["num_major_iterations"] = 2
        out[name] = F.sum_threshold_flagger(vis, flags, **kw)
    out["defaults"] = F.sum_threshold_flagger(vis, flags)
    out["avg2"] = F.sum_threshold_flagger(vis, flags, average_freq=2, windows_freq=[2, 4, 8, 16],
                                          num_major_iterations=1)
    np.savez_compressed(os.path.join(HERE, "sum_threshold_flagger.npz"), **out)

    # --- stages of one plane (intermediate backgrounds for the 1e-5 check)
    data = np.abs(vis[0, 0]).astype(np.float32)
    data[np.isnan(data)] = 0
    fl = flags[0, 0]
    ce = np.linspace(0, data.shape[1], 11).astype(np.int_)
    bg = F._get_background2d(data, fl, 5, np.array((12.5, 10.0)), 2.0, ce)
    mf = np.zeros_like(data)
    F.masked_gaussian_filter(data, fl, np.array((12.5, 10.0)), mf)
    tm, tmf = F._time_median(data, fl)
    st0 = F._sum_threshold(data - bg, fl, 0, np.array([1, 2, 4, 8]), 10, 1.3)
    st1 = F._sum_threshold(data - bg, fl, 1, np.array([1, 2, 4, 8]), 10, 1.3, ce)
    np.savez_compressed(os.path.join(HERE, "stages.npz"), data=data, flags=fl, chunk_ends=ce,
                        background=bg, masked_filter=mf, time_median=tm, time_median_flags=tmf,
                        st_time=st0, st_freq=st1)

    # --- uvcontsub_flagger
    uv = {"vis": vis, "flags": flags}
    uv["cycles7"] = F.uvcontsub_flagger(vis.copy(), flags, major_cycles=7, or_original_from_cycle=1,
                                        taylor_degrees=20, sigma=15.0)
    uv["cycles3_or0"] = F.uvcontsub_flagger(vis.copy(), flags, major_cycles=3, or_original_from_cycle=0,
                                            taylor_degrees=25, sigma=13.0)
    np.savez_compressed(os.path.join(HERE, "uvcontsub.npz"), **uv)

    # --- companions
    rowvis = (rs.standard_normal((40, 24, 4)) + 1j * rs.standard_normal((40, 24, 4))).astype(np.complex64)
    comp = {"rowvis": rowvis, "nanzero": F.flag_nans_and_zeros(vis, flags)}
    for tag, ct in (("lin", [9, 10, 11, 12]), ("circ", [5, 6, 7, 8]), ("mixed", [11, 9, 10, 12])):
        m = S.stokes_corr_map(ct)
        pol = tuple(v for k, v in m.items() if k != 'I')
        unpol = tuple(v for k, v in m.items() if k == 'I')
        comp["pol_" + tag] = S.polarised_intensity(rowvis, pol)
        comp["unpol_" + tag] = S.unpolarised_intensity(rowvis, unpol, pol)
    np.savez_compressed(os.path.join(HERE, "companions.npz"), **comp)

    # --- pack / unpack / window stats
    na, ntime, nchan, ncorr = 6, 8, 12, 4
    a1, a2 = (a.astype(np.int32) for a in np.triu_indices(na, 0))
    nbl = a1.size
    A1, A2 = np.tile(a1, ntime), np.tile(a2, ntime)
    tinv = np.repeat(np.arange(ntime), nbl)
    nrow = A1.size
    dv = (rs.standard_normal((nrow, nchan, ncorr)) + 1j * rs.standard_normal((nrow, nchan, ncorr))).astype(np.complex64)
    df = rs.randint(0, 2, (nrow, nchan, ncorr)).astype(bool)
    dele = rs.randint(nrow, size=11)
    A1, A2, tinv = np.delete(A1, dele), np.delete(A2, dele), np.delete(tinv, dele)
    dv, df = np.delete(dv, dele, 0), np.delete(df, dele, 0)
    ubl = np.unique(np.stack([A1, A2], 1).view(np.int64)).view(np.int32).reshape(-1, 2)
    ubl = np.concatenate([np.arange(ubl.shape[0], dtype=np.int32)[:, None], ubl], 1)
    vw = np.full((nbl, ncorr, ntime, nchan), np.nan + np.nan * 1j, np.complex64)
    fw = np.full((nbl, ncorr, ntime, nchan), 1, bool)
    P._numba_pack_data(tinv, ubl, A1, A2, dv, df, vw, fw)
    un = P._unpack_data(A1, A2, tinv, [[ubl]], [[fw]])
    names = ["A%d" % i for i in range(na)]
    cf = np.linspace(.856e9, 2 * .856e9, nchan)
    st = W._window_stats([[[fw]]], [ubl], [cf], names, 3, "M87", 0, 10)
    np.savez_compressed(
        os.path.join(HERE, "packing.npz"), ant1=A1, ant2=A2, time_inv=tinv, vis=dv, flags=df, ubl=ubl,
        ntime=ntime, vis_win=vw, flag_win=fw, unpacked=un, chan_freqs=cf,
        counts_per_ant=np.array([int(st._counts_per_ant[n]) for n in names], np.uint64),
        size_per_ant=np.array([int(st._size_per_ant[n]) for n in names], np.uint64),
        counts_per_bl=np.array([int(st._counts_per_bl["%s&%s" % (names[b[1]], names[b[2]])]) for b in ubl], np.uint64),
        counts_field=np.uint64(st._counts_per_field["M87"]), size_scan=np.uint64(st._size_per_scan[3]),
        bins=st._counts_per_ddid[0], bin_edges=st._bins_per_ddid[0])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
