// st_driver.cuh -- host-side sequencing of the SumThreshold stages on the
// device (reference: _get_baseline_flags flagging.py:921-976, _get_flags_impl
// 745-781, _get_background2d 516-579).
//
// Two layouts of every per-plane array are used: "TF" = (plane, time, chan)
// and "FT" = (plane, chan, time).  Whatever walks the time axis sequentially
// (time box passes, time SumThreshold scan) runs on TF with one thread per
// channel; whatever walks the frequency axis (frequency box passes, frequency
// SumThreshold scan, NaN interpolation) runs on FT with one thread per dump.
// Either way a warp touches 32 consecutive addresses per step.  Frequency
// chunks are contiguous in FT, which makes the per-chunk medians 1-D ranges.
#pragma once
#include "k_elementwise.cuh"
#include "k_filter.cuh"
#include "k_filter2.cuh"
#include "k_filter3.cuh"
#include "k_filter5.cuh"
#include "k_filter5t.cuh"
#include "k_select.cuh"
#include "k_sumthreshold.cuh"

#define MAD_NORMAL 1.4826  // tricolour/flagging.py:22

static int dev_upload_i64(tc_context *c, const int64_t *h, size_t n, int64_t **d)
{
    TC_TRY(tc_alloc(c, n, d));
    TC_TRY(tc_upload_small(c, h, n * sizeof(int64_t), *d));
    return TC_OK;
}

// contiguous FT ranges of every (plane, chunk)
static int dev_make_ranges(tc_context *c, int64_t np, int64_t T, int64_t Fa, const int64_t *ce,
                           int nce, int64_t **lo, int64_t **hi, int64_t *max_range)
{
    int nch = nce - 1;
    std::vector<int64_t> hlo((size_t)np * nch), hhi((size_t)np * nch);
    int64_t mr = 0;
    for (int64_t p = 0; p < np; p++)
        for (int k = 0; k < nch; k++) {
            hlo[p * nch + k] = p * T * Fa + ce[k] * T;
            hhi[p * nch + k] = p * T * Fa + ce[k + 1] * T;
            if ((ce[k + 1] - ce[k]) * T > mr) mr = (ce[k + 1] - ce[k]) * T;
        }
    TC_TRY(dev_upload_i64(c, hlo.data(), hlo.size(), lo));
    TC_TRY(dev_upload_i64(c, hhi.data(), hhi.size(), hi));
    *max_range = mr;
    return TC_OK;
}

// device tables that depend only on the batch shape and the chunk boundaries: built once per plane
// batch (tc_sum_threshold) instead of once per pass
struct PassTables {
    int64_t np; int T, Fa, nce;
    int64_t *slo, *shi, smax;      // (plane, chunk) ranges of the spectra (T = 1)
    int64_t *rlo, *rhi, rmax;      // (plane, chunk) ranges of the planes in (F,T) order
    int64_t *d_ce, *d_tce;         // frequency chunk ends; {0, T}
};

struct BgWork {
    u8 *fl_FT;            // working copy of the flags (gets modified), (plane, chan, time)
    float *v_FT, *w_FT;   // time-filtered value / weight, already in (plane, chan, time)
};

static int dev_bg_work_alloc(tc_context *c, int64_t N, bool two_axes, BgWork *w)
{
    TC_TRY(tc_alloc(c, N, &w->fl_FT));
    w->v_FT = w->w_FT = nullptr;
    if (two_axes) {
        // one allocation: the TMA form of the second-axis filter addresses the pair as one 3-D tensor
        TC_TRY(tc_alloc(c, 2 * N, &w->v_FT));
        w->w_FT = w->v_FT + N;
    }
    return TC_OK;
}

__global__ void __launch_bounds__(256)
k_abs_sub(const float *a, const float *b, float *out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fabsf(a[i] - b[i]);
}

// one masked_gaussian_filter (flagging.py:469-513) from (data, work flags) to
// out_FT; resid != 0 stores |data - background| instead (flagging.py:561-566).
// The lean kernels (k_filter2.cuh) want their input line-contiguous: the time
// axis reads the (F,T) copies, the frequency axis the (T,F) ones; whichever
// kernel runs first writes the intermediate (value, weight) pair in the layout
// the second one reads.
static int dev_masked_filter(tc_context *c, int64_t np, int T, int Fa, const float *data_TF,
                             const float *data_FT, BgWork &w, int64_t r0, int64_t r1, int resid,
                             float *out_FT, int want_TF = 0, int *got_TF = nullptr)
{
    if (got_TF) *got_TF = 0;
    int64_t N = np * (int64_t)T * Fa;
    FilterArgs a;
    memset(&a, 0, sizeof(a));
    FilterArgs probe;
    memset(&probe, 0, sizeof(probe));
    probe.n = T; probe.r = (int)r0;
    const bool lean0 = r0 > 0 && b2_supported(c, probe);
    probe.n = Fa; probe.r = (int)r1;
    const bool lean1 = r1 > 0 && b2_supported(c, probe) && (r0 > 0 || T == 1);
    // the B5 forms (k_filter5.cuh) replace the lane-per-chain kernels of k_filter2.cuh
    // wherever they apply; TC_B5_A_MINR / TC_B5_B_MINR keep an axis on the older forms
    // below a radius (the thread-per-line kernel of small first-axis radii)
    // (measured on B200: first axis from r = 37, where the thread-per-line weight chains stop; second axis
    // at every radius up to r ~ 200, beyond which the longer rings cost occupancy -- until the residual's
    // samples were prefetched B5 lost to k_box4 from r = 17 up)
    static const int b5_a_minr = tpl_env_int("TC_B5_A_MINR", 37), b5_b_minr = tpl_env_int("TC_B5_B_MINR", 1);
    static const int b5_b_maxr = tpl_env_int("TC_B5_B_MAXR", 200);
    probe.n = T; probe.r = (int)r0; probe.data = data_FT; probe.flags = w.fl_FT;
    const bool b5_0 = r0 >= b5_a_minr && r0 > 0 && b5_supported(c, probe);
    probe.n = Fa; probe.r = (int)r1; probe.data = w.v_FT; probe.win = w.w_FT; probe.flags = nullptr;
    const bool b5_1 = r1 >= b5_b_minr && r1 <= b5_b_maxr && r1 > 0 && b5_supported(c, probe) && (r0 > 0 || T == 1);
    if (r0 > 0 && r1 > 0) {
        // second axis first: which kernel takes it decides the layout of the intermediate pair
        FilterArgs b;
        memset(&b, 0, sizeof(b));
        b.n = Fa; b.nj = T; b.nlines = np * T; b.r = (int)r1;
        b.mode_in = FIN_PAIR; b.mode_out = resid ? FOUT_RESID : FOUT_BG;
        b.data = w.v_FT; b.win = w.w_FT; b.vout = out_FT;
        const bool tplb = !b5_1 && tpl_b_supported(c, b);
        // time axis: flags are read from the (F,T) layout; the pair goes to (T,F)
        // when a kernel that reads line-contiguous input follows, else to (F,T)
        a.n = T; a.nj = Fa; a.nlines = np * Fa; a.r = (int)r0;
        a.mode_in = FIN_MASKED; a.mode_out = FOUT_PAIR;
        a.flags = w.fl_FT; a.flags_transposed = 1; a.out_transposed = (b5_1 || lean1 || tplb) ? 0 : 1;
        a.vout = w.v_FT; a.wout = w.w_FT;
        a.data = data_TF;
        if (b5_0) { a.data = data_FT; TC_TRY(launch_box_filter5(c, a)); }
        else if (tpl_a_supported(c, a)) TC_TRY(launch_box_tpl_a(c, a));
        else if (t4a_supported(c, a)) TC_TRY(launch_box_t4a(c, a));
        else if (lean0 && t4a_weights_supported(c, a)) {
            // values: lane-per-chain kernel; weights: thread-per-line integer chains
            a.role = 2;
            TC_TRY(launch_box_t4a(c, a));
            a.role = 1; a.data = data_FT;
            TC_TRY(launch_box_filter2(c, a));
        } else if (lean0) { a.data = data_FT; TC_TRY(launch_box_filter2(c, a)); }
        else TC_TRY(launch_box_filter(c, a));
        if (b5_1 || (!tplb && lean1)) {
            // the lane-per-chain kernels can leave their output line-contiguous, i.e. in (T,F)
            b.data2 = data_TF;
            if (want_TF && got_TF) { b.out_transposed = 1; *got_TF = 1; }
            if (b5_1 && b5t_supported(c, b)) TC_TRY(launch_box_filter5t(c, b));
            else if (b5_1) TC_TRY(launch_box_filter5(c, b));
            else TC_TRY(launch_box_filter2(c, b));
        } else if (tplb) {
            // thread per line: the output is sample-major for this axis, i.e. (F,T)
            b.data2 = data_FT;
            TC_TRY(launch_box_tpl_b(c, b));
        } else { b.data2 = data_FT; TC_TRY(launch_box_filter(c, b)); }
    } else if (r0 > 0) {
        a.n = T; a.nj = Fa; a.nlines = np * Fa; a.r = (int)r0; a.single_axis = 1;
        a.mode_in = FIN_MASKED; a.mode_out = FOUT_BG;
        a.flags = w.fl_FT; a.flags_transposed = 1; a.out_transposed = 1;
        a.vout = out_FT;
        if (b5_0) { a.data = data_FT; TC_TRY(launch_box_filter5(c, a)); }
        else if (lean0) { a.data = data_FT; TC_TRY(launch_box_filter2(c, a)); }
        else { a.data = data_TF; TC_TRY(launch_box_filter(c, a)); }
        if (resid) {
            TC_LAUNCH_NOSYNC(k_abs_sub, tc_blocks_for(N, 256), 256, 0, c->stream, data_FT, out_FT, out_FT, N);
            c->launches++;
        }
    } else if (r1 > 0) {
        a.n = Fa; a.nj = T; a.nlines = np * T; a.r = (int)r1; a.single_axis = 1;
        a.mode_in = FIN_MASKED; a.mode_out = resid ? FOUT_RESID : FOUT_BG;
        a.data = data_FT; a.flags = w.fl_FT; a.vout = out_FT; a.data2 = data_FT;
        // T == 1: both layouts coincide and the lines are contiguous
        if (b5_1) TC_TRY(launch_box_filter5(c, a));
        else if (lean1) TC_TRY(launch_box_filter2(c, a));
        else TC_TRY(launch_box_filter(c, a));
    } else {
        TC_LAUNCH_NOSYNC(k_masked_copy, tc_blocks_for(N, 256), 256, 0, c->stream, data_FT, w.fl_FT, N,
                         resid ? FOUT_RESID : FOUT_BG, data_FT, out_FT);
        c->launches++;
    }
    TC_KERNEL_CHECK();
    return TC_OK;
}

// _get_background2d (flagging.py:516-579).  flags_* are the caller's flags
// (not modified).  work_FT receives the filter outputs; the interpolated
// background lands in out_TF, or minuend_TF - background when minuend_TF is
// given (the `data -= background` of flagging.py:950/962 fused in).
static int dev_background2d(tc_context *c, int64_t np, int T, int Fa, const float *data_TF,
                            const float *data_FT, const u8 *flags_TF, const u8 *flags_FT,
                            int iterations, const int64_t *radii, double reject,
                            const int64_t *range_lo, const int64_t *range_hi, int nchunks,
                            int64_t max_range, float *work_FT, float *out_TF,
                            const float *minuend_TF)
{
    int64_t N = np * (int64_t)T * Fa;
    bool two_axes = false;
    for (int k = 0; k <= iterations; k++) if (radii[2 * k] > 0 && radii[2 * k + 1] > 0) two_axes = true;
    if (T == 1) two_axes = false;
    (void)flags_TF;
    BgWork w;
    tc_mark mark = tc_arena_mark(c);
    TC_TRY(dev_bg_work_alloc(c, N, two_axes, &w));
    TC_TRY(tc_copy_d2d(c, w.fl_FT, flags_FT, N));
    int bg_is_TF = 0;
    for (int it = 0; it <= iterations; it++) {
        int64_t r0 = T == 1 ? 0 : radii[2 * it], r1 = radii[2 * it + 1];
        bool final_pass = it == iterations;
        TC_TRY(dev_masked_filter(c, np, T, Fa, data_TF, data_FT, w, r0, r1, final_pass ? 0 : 1, work_FT,
                                 final_pass ? 1 : 0, &bg_is_TF));
        if (final_pass) break;
        ChunkSelectArgs s;
        memset(&s, 0, sizeof(s));
        s.resid = work_FT; s.flags = w.fl_FT; s.range_lo = range_lo; s.range_hi = range_hi;
        s.thr_mult = MAD_NORMAL * reject; s.mode = CS_BACKGROUND; s.take_abs = 0; s.medians = nullptr;
        TC_TRY(launch_chunk_select(c, s, np * nchunks, max_range));
    }
    // _linearly_interpolate_nans along frequency for every (plane, dump): one
    // warp per contiguous (T,F) row
    const float *bg_TF = work_FT;
    if (T != 1 && !bg_is_TF) {
        float *tmp;
        TC_TRY(tc_alloc(c, (size_t)N, &tmp));
        TC_TRY(launch_transpose<float>(c, work_FT, tmp, np, Fa, T));
        bg_TF = tmp;
    }
    int *rv;
    TC_TRY(tc_alloc(c, (size_t)N, &rv));
    tc_prof_begin(c, TCP_INTERP);
    TC_LAUNCH(k_interp_nans_rows, tc_blocks_for(np * (int64_t)T * 32, 128), 128, 0, c->stream, bg_TF, minuend_TF,
              out_TF, rv, np * (int64_t)T, Fa);
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    tc_arena_release(c, mark);
    return TC_OK;
}

// _sum_threshold (flagging.py:684-742) for np planes.
// axis 0: scans time (TF layout for the scan, FT for the medians)
// axis 1: scans frequency (FT for the scan, TF for the medians)
// chunk_ends (host, nce entries) partition the scanned axis.
static int dev_sum_threshold(tc_context *c, int64_t np, int T, int Fa, int axis, const float *d_TF,
                             const float *d_FT, const u8 *fl_TF, const u8 *fl_FT,
                             const u8 *fl2_TF, const int64_t *windows, const double *tf,
                             const float *scale, int nwin, double nsigma, const int64_t *ce,
                             int nce, u8 *out /* TF for axis 0, FT for axis 1 */, int64_t *d_ce_cached = nullptr)
{
    TC_REQUIRE(nwin > 0, "zero-size array to reduction operation maximum which has no identity");
    TC_REQUIRE(nwin <= TC_MAX_WINDOWS, "at most %d SumThreshold windows are supported", TC_MAX_WINDOWS);
    int n = axis == 0 ? T : Fa;
    int nchunks = nce - 1;
    int64_t maxw = 0;
    for (int k = 0; k < nwin; k++) {
        TC_REQUIRE(windows[k] >= 1, "unable to broadcast argument 1 to output array (window %lld < 1)",
                   (long long)windows[k]);
        if (windows[k] > maxw) maxw = windows[k];
    }
    int maxlen = 0, mpad = 0;
    for (int k = 0; k < nchunks; k++) {
        TC_REQUIRE(ce[k + 1] >= ce[k] && ce[k] >= 0 && ce[k + 1] <= n, "bad chunk boundaries");
        int len = (int)(ce[k + 1] - ce[k]);
        if (len > maxlen) maxlen = len;
        int64_t p0 = ce[k] - maxw + 1; if (p0 < 0) p0 = 0;
        int64_t p1 = ce[k + 1] + maxw - 1; if (p1 > n) p1 = n;
        if ((int)(p1 - p0) > mpad) mpad = (int)(p1 - p0);
    }
    if (nchunks <= 0 || np == 0) return TC_OK;
    tc_mark mark = tc_arena_mark(c);
    int64_t *d_ce = d_ce_cached;
    if (!d_ce) TC_TRY(dev_upload_i64(c, ce, (size_t)nce, &d_ce));
    int64_t ninner = axis == 0 ? Fa : T;  // lines per plane
    int64_t nlines = np * ninner;
    float *thr = nullptr;
    TC_TRY(tc_alloc(c, (size_t)nlines * nchunks, &thr));
    LineMedianArgs m;
    memset(&m, 0, sizeof(m));
    m.nlines = nlines; m.ninner = ninner; m.outer_stride = (int64_t)T * Fa;
    m.elem_stride = 1; m.mode = LM_ST_THRESHOLD; m.use_abs = 1;
    m.thr_scale = nsigma * MAD_NORMAL; m.out = thr;
    if (axis == 0) {  // a time line of channel f is contiguous in FT
        m.data = d_FT; m.flags = fl_FT; m.flags2 = nullptr; m.inner_stride = T;
    } else {          // a frequency line of dump t is contiguous in TF
        m.data = d_TF; m.flags = fl_TF; m.flags2 = fl2_TF; m.inner_stride = Fa;
    }
    m.seg_ends = d_ce; m.nseg = nchunks; m.n = n;
    TC_TRY(launch_line_median(c, m, maxlen));

    StScanArgs s;
    memset(&s, 0, sizeof(s));
    s.data = axis == 0 ? d_TF : d_FT;
    s.thr = thr; s.out = out; s.nlines = nlines; s.ninner = ninner;
    s.outer_stride = (int64_t)T * Fa; s.estride = ninner; s.n = n; s.nchunks = nchunks;
    s.chunk_ends = d_ce; s.nwin = nwin; s.maxw = (int)maxw; s.mpad = mpad;
    for (int k = 0; k < nwin; k++) { s.windows[k] = windows[k]; s.tf[k] = tf[k]; s.scale[k] = scale[k]; }
    s.fused1248 = (nwin == 4 && windows[0] == 1 && windows[1] == 2 && windows[2] == 4 && windows[3] == 8 &&
                   !TC_ENV_FLAG("TC_ST_UNFUSED")) ? (TC_ENV_FLAG("TC_ST_V1") ? 1 : 2) : 0;
    bool need_cum = false;
    for (int k = 0; k < nwin; k++)
        if (windows[k] != 1 && windows[k] != 2 && windows[k] != 4 && windows[k] != 8) need_cum = true;
    s.cum = nullptr;
    if (need_cum) TC_TRY(tc_alloc(c, (size_t)np * nchunks * (mpad + 1) * ninner, &s.cum));
    s.pn = s.pn2 = nullptr;
    if (!s.fused1248) {
        TC_TRY(tc_alloc(c, (size_t)np * nchunks * (mpad > 0 ? mpad : 1) * ninner, &s.pn));
        TC_TRY(tc_alloc(c, (size_t)np * nchunks * (mpad > 0 ? mpad : 1) * ninner, &s.pn2));
    }
    tc_prof_begin(c, TCP_ST_SCAN);
    TC_LAUNCH_NOSYNC(k_st_scan, tc_blocks_for(nlines * nchunks, 128), 128, 0, c->stream, s);
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    tc_arena_release(c, mark);
    return TC_OK;
}

// _combine_flags + _unaverage_freq + the final isnan OR (flagging.py:784-816,
// 878-918, 776-781) for np planes; c1 is N bytes of scratch
static int dev_combine_flags(tc_context *c, int64_t np, int T, int Fa, int F, int avg, int te, int fe,
                             double frac_t, double frac_f, const u8 *spec_out, const u8 *time_TF,
                             const u8 *freq_TF, u8 *c1, const void *vis, int vis_kind, u8 *out_flags,
                             u8 *iter_flags_accum)
{
    const int64_t N = np * (int64_t)T * Fa, NF = np * (int64_t)T * F;
    tc_prof_begin(c, TCP_COMBINE);
    u8 *dflags;
    int *rowcnt, *colcnt;
    TC_TRY(tc_alloc(c, NF, &dflags));
    TC_TRY(tc_alloc(c, np * (int64_t)T, &rowcnt));
    TC_TRY(tc_alloc(c, np * (int64_t)F, &colcnt));
    const bool vec16 = avg == 1 && (F & 15) == 0 && fe >= 0 && fe <= 16 && ((uintptr_t)vis & 15) == 0 &&
                       ((uintptr_t)out_flags & 3) == 0 && ((uintptr_t)iter_flags_accum & 3) == 0 &&
                       ((uintptr_t)spec_out & 15) == 0 && ((uintptr_t)time_TF & 15) == 0 &&
                       ((uintptr_t)freq_TF & 15) == 0 && T <= 65535 && np <= 65535 && F + 32 <= 48 * 1024 &&
                       np * (int64_t)T < ((int64_t)1 << 31) && !TC_ENV_FLAG("TC_COMBINE_SCALAR");
    if (vec16) {
        const int F16 = F / 16, F4 = F / 4;
        TC_LAUNCH_NOSYNC(k_combine_time_v16, dim3(tc_blocks_for(F16, 256), (unsigned)T, (unsigned)np), 256, 0, c->stream,
                         (const uint4 *)spec_out,
                         (const uint4 *)time_TF, (const uint4 *)freq_TF, N / 16, T, F16, -(te / 2), te, (uint4 *)c1);
        c->launches++;
        TC_LAUNCH(k_dilate_rows_v16, (unsigned)(np * T), 256, (size_t)(F + 32), c->stream, (const uint4 *)c1, F16,
                  -(fe / 2), fe, (uint4 *)dflags, rowcnt);
        c->launches++;
        TC_LAUNCH_NOSYNC(k_colcnt_v4, dim3(tc_blocks_for(F4, 128), (unsigned)np), 128, 0, c->stream,
                         (const unsigned *)dflags, T, F4, np * (int64_t)F4, colcnt);
        c->launches++;
        TC_LAUNCH_NOSYNC(k_finalize_flags_v4, dim3(tc_blocks_for(F4, 256), (unsigned)T, (unsigned)np), 256, 0, c->stream,
                         (const unsigned *)dflags,
                         rowcnt, (const int4 *)colcnt, vis, vis_kind, NF / 4, T, F4, frac_f * (double)F,
                         (double)T * frac_t, (unsigned *)out_flags, (unsigned *)iter_flags_accum);
        c->launches++;
    } else {
        TC_LAUNCH_NOSYNC(k_combine_time, tc_blocks_for(N, 256), 256, 0, c->stream, spec_out, time_TF, freq_TF, N, T,
                         Fa, -(te / 2), te, c1);
        c->launches++;
        TC_CUDA(cudaMemsetAsync(colcnt, 0, sizeof(int) * np * (size_t)F, c->stream));
        TC_LAUNCH(k_unaverage_rows, (unsigned)(np * T), 256, 0, c->stream, c1, T, Fa, F, -(fe / 2), fe, avg, dflags,
                  rowcnt, colcnt);
        c->launches++;
        TC_LAUNCH_NOSYNC(k_finalize_flags, tc_blocks_for(NF, 256), 256, 0, c->stream, dflags, rowcnt, colcnt, vis,
                         vis_kind, NF, T, F, frac_f * (double)F, (double)T * frac_t, out_flags, iter_flags_accum);
        c->launches++;
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    return TC_OK;
}

static int dev_pass_tables(tc_context *c, const tc_st_params *p, int64_t np, int T, int F, PassTables *t)
{
    const int avg = (int)p->average_freq;
    t->np = np; t->T = T; t->Fa = (F + avg - 1) / avg; t->nce = p->nchunk_ends;
    TC_TRY(dev_make_ranges(c, np, 1, t->Fa, p->freq_chunk_ends, t->nce, &t->slo, &t->shi, &t->smax));
    TC_TRY(dev_make_ranges(c, np, T, t->Fa, p->freq_chunk_ends, t->nce, &t->rlo, &t->rhi, &t->rmax));
    TC_TRY(dev_upload_i64(c, p->freq_chunk_ends, (size_t)t->nce, &t->d_ce));
    const int64_t tce[2] = {0, T};
    TC_TRY(dev_upload_i64(c, tce, 2, &t->d_tce));
    return TC_OK;
}

// ----------------------------------------------------------------------------
// one _get_flags_impl pass over np planes
// ----------------------------------------------------------------------------
static int dev_get_flags_pass(tc_context *c, const tc_st_params *p, const void *vis, int vis_kind,
                              const u8 *in_flags, int64_t np, int T, int F, u8 *out_flags,
                              u8 *iter_flags_accum, const PassTables *tab = nullptr)
{
    const int avg = (int)p->average_freq;
    const int Fa = (F + avg - 1) / avg;
    const int64_t N = np * (int64_t)T * Fa;
    const int nce = p->nchunk_ends, nchunks = nce - 1;
    const int iters = p->background_iterations;
    TC_REQUIRE(nchunks >= 0, "freq_chunk_ends must not be empty");
    tc_mark pass_mark = tc_arena_mark(c);

    float *data_TF, *data_FT;
    u8 *fl_TF, *fl_FT;
    TC_TRY(tc_alloc(c, N, &data_TF)); TC_TRY(tc_alloc(c, N, &data_FT));
    TC_TRY(tc_alloc(c, N, &fl_TF)); TC_TRY(tc_alloc(c, N, &fl_FT));
    // S1 _average_freq
    tc_prof_begin(c, TCP_PREP);
    const bool vec_ok = avg == 1 && vis_kind == TC_VIS_COMPLEX64 && (N & 3) == 0 && ((uintptr_t)vis & 15) == 0 &&
                        ((uintptr_t)in_flags & 3) == 0;
    if (vec_ok && (T & 31) == 0 && (F & 31) == 0 && np <= 65535 && T / 32 <= 65535) {
        // amplitudes and flags in both layouts from one kernel
        TC_LAUNCH(k_prep_c64_tile, dim3((unsigned)(F / 32), (unsigned)(T / 32), (unsigned)np), 256, 0, c->stream,
                  (const float4 *)vis, (const unsigned *)in_flags, T, F, (float4 *)data_TF, (unsigned *)fl_TF,
                  (float4 *)data_FT, (unsigned *)fl_FT);
        tc_prof_end(c);
        c->launches++;
        TC_KERNEL_CHECK();
    } else {
        if (vec_ok)
            TC_LAUNCH_NOSYNC(k_prep_c64_v4, tc_blocks_for(N / 4, 256), 256, 0, c->stream, (const float4 *)vis,
                             (const unsigned *)in_flags, N / 4, (float4 *)data_TF, (unsigned *)fl_TF);
        else
            TC_LAUNCH_NOSYNC(k_prep, tc_blocks_for(N, 256), 256, 0, c->stream, vis, vis_kind, in_flags, N, F, Fa, avg,
                             data_TF, fl_TF);
        tc_prof_end(c);
        c->launches++;
        TC_KERNEL_CHECK();
        TC_TRY(launch_transpose<float>(c, data_TF, data_FT, np, T, Fa));
        TC_TRY(launch_transpose<u8>(c, fl_TF, fl_FT, np, T, Fa));
    }

    // S2 _time_median -> (np, Fa)
    float *spec_data, *spec_bg;
    u8 *spec_fl, *spec_out;
    TC_TRY(tc_alloc(c, np * (int64_t)Fa, &spec_data)); TC_TRY(tc_alloc(c, np * (int64_t)Fa, &spec_bg));
    TC_TRY(tc_alloc(c, np * (int64_t)Fa, &spec_fl)); TC_TRY(tc_alloc(c, np * (int64_t)Fa, &spec_out));
    {
        LineMedianArgs m;
        memset(&m, 0, sizeof(m));
        m.data = data_FT; m.flags = fl_FT; m.nlines = np * Fa; m.ninner = Fa;
        m.outer_stride = (int64_t)T * Fa; m.inner_stride = T; m.elem_stride = 1; m.n = T;
        m.mode = LM_TIME_MEDIAN; m.use_abs = 0; m.out = spec_data; m.out_flags = spec_fl;
        TC_TRY(launch_line_median(c, m, T));
    }
    // spectrum background (flagging.py:945-949) and its SumThreshold (950-952)
    PassTables local;
    if (!tab) {
        TC_TRY(dev_pass_tables(c, p, np, T, F, &local));
        tab = &local;
    }
    int64_t *slo = tab->slo, *shi = tab->shi, smax = tab->smax;
    float *spec_res;
    TC_TRY(tc_alloc(c, np * (int64_t)Fa, &spec_res));
    TC_TRY(dev_background2d(c, np, 1, Fa, spec_data, spec_data, spec_fl, spec_fl, iters, p->radii_spec,
                            p->background_reject, slo, shi, nchunks, smax, spec_bg, spec_res, spec_data));
    TC_TRY(dev_sum_threshold(c, np, 1, Fa, 1, spec_res, spec_res, spec_fl, spec_fl, nullptr,
                             p->windows_freq, p->tf_freq, p->scale_freq, p->nwin_freq, p->outlier_nsigma,
                             p->freq_chunk_ends, nce, spec_out, tab->d_ce));
    // flags |= spec_flags (flagging.py:954), both layouts
    if ((Fa & 15) == 0 && (T & 15) == 0 && T <= 65535 && np <= 65535) {
        // flag bytes are 0/1 (k_prep wrote them): OR whole vectors, in both layouts
        TC_LAUNCH_NOSYNC(k_or_spec_tf16, dim3(tc_blocks_for(Fa / 16, 256), (unsigned)T, (unsigned)np), 256, 0, c->stream,
                         (uint4 *)fl_TF,
                         (const uint4 *)spec_out, N / 16, T, Fa / 16);
        TC_LAUNCH_NOSYNC(k_or_spec_ft16, tc_blocks_for(N / 16, 256), 256, 0, c->stream, (uint4 *)fl_FT, spec_out,
                         N / 16, T / 16);
        c->launches += 2;
        TC_KERNEL_CHECK();
    } else {
        TC_LAUNCH_NOSYNC(k_or_spec, tc_blocks_for(N, 256), 256, 0, c->stream, fl_TF, spec_out, N, T, Fa);
        c->launches++;
        TC_KERNEL_CHECK();
        TC_TRY(launch_transpose<u8>(c, fl_TF, fl_FT, np, T, Fa));
    }

    // 2-D background (flagging.py:957-961)
    float *bg_FT, *dres_TF;
    TC_TRY(tc_alloc(c, N, &bg_FT)); TC_TRY(tc_alloc(c, N, &dres_TF));
    int64_t *rlo = tab->rlo, *rhi = tab->rhi, rmax = tab->rmax;
    // background and data -= background (flagging.py:957-962), in both layouts
    TC_TRY(dev_background2d(c, np, T, Fa, data_TF, data_FT, fl_TF, fl_FT, iters, p->radii_2d,
                            p->background_reject, rlo, rhi, nchunks, rmax, bg_FT, dres_TF, data_TF));
    float *dres_FT = bg_FT;
    TC_TRY(launch_transpose<float>(c, dres_TF, dres_FT, np, T, Fa));

    // SumThreshold along time (flagging.py:964-965)
    u8 *time_TF, *freq_FT, *freq_TF;
    TC_TRY(tc_alloc(c, N, &time_TF)); TC_TRY(tc_alloc(c, N, &freq_FT)); TC_TRY(tc_alloc(c, N, &freq_TF));
    int64_t tce[2] = {0, T};
    TC_TRY(dev_sum_threshold(c, np, T, Fa, 0, dres_TF, dres_FT, fl_TF, fl_FT, nullptr, p->windows_time,
                             p->tf_time, p->scale_time, p->nwin_time, p->outlier_nsigma, tce, 2, time_TF, tab->d_tce));
    // flags |= time_flags; SumThreshold along frequency (flagging.py:967-969)
    TC_TRY(dev_sum_threshold(c, np, T, Fa, 1, dres_TF, dres_FT, fl_TF, fl_FT, time_TF, p->windows_freq,
                             p->tf_freq, p->scale_freq, p->nwin_freq, p->outlier_nsigma,
                             p->freq_chunk_ends, nce, freq_FT, tab->d_ce));
    TC_TRY(launch_transpose<u8>(c, freq_FT, freq_TF, np, Fa, T));

    // _combine_flags + _unaverage_freq + final isnan OR
    u8 *c1 = fl_FT;  // fl_FT is no longer needed
    int te = (int)p->time_extend, fe = (int)p->freq_extend;
    TC_TRY(dev_combine_flags(c, np, T, Fa, F, avg, te, fe, p->flag_all_time_frac, p->flag_all_freq_frac, spec_out,
                             time_TF, freq_TF, c1, vis, vis_kind, out_flags, iter_flags_accum));
    tc_arena_release(c, pass_mark);
    return TC_OK;
}
