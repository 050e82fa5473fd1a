// k_filter.cuh -- masked box-Gaussian filter (reference: _box_gaussian_filter1d
// flagging.py:362-419, _box_gaussian_filter 422-466, masked_gaussian_filter
// 469-513).
//
// The reference makes K=4 sequential box passes over a zero-padded copy of
// every line, each pass a float64 running sum whose outputs are rounded to
// float32.  Here all four passes of a line are fused into ONE streaming loop:
// pass p+1 consumes what pass p emitted 2r samples earlier, so a line is read
// once and written once, and the only state is four float64 accumulators and
// three 2r-deep float32 delay lines per array.  The order of floating point
// operations on every accumulator is exactly the reference's (add the entering
// sample, round+emit, subtract the leaving sample), so results are bit-exact.
//
// One thread owns one line and filters the `value` and the `weight` array of
// the masked filter together.  Lines are addressed so that neighbouring threads
// touch neighbouring addresses: element i of line j of plane p lives at
// p*n*nj + i*nj + j.  Along time that is the (T,F) layout with j = channel;
// along frequency it is the transposed (F,T) layout with j = time.
//
// The kernel is FP64-add / convert issue bound and runs at low occupancy (the
// delay lines fill shared memory), so the loop is arranged for ILP instead:
//   * the four passes are software pipelined -- pass p+1 works on what pass p
//     produced one tick earlier -- which leaves eight independent dependency
//     chains (4 passes x 2 arrays) per thread and tick;
//   * global loads run one unrolled group (4 ticks) ahead of their use.
//
// Delay lines live in shared memory as [line][slot][thread] (bank-conflict
// free); when 2r is too deep for shared memory they fall back to an L2-resident
// global ring with the same indexing.
#pragma once
#include "tc_common.cuh"

enum { FIN_MASKED = 0, FIN_PAIR = 1 };
enum { FOUT_PAIR = 0, FOUT_BG = 1, FOUT_RESID = 2 };

struct FilterArgs {
    int n;            // samples per line
    int nj;           // lines per plane
    int64_t nlines;   // nplanes * nj
    int r;            // box radius (> 0)
    float div;        // float32(d)**4 by repeated squaring (flagging.py:419)
    int mode_in, mode_out;
    const float *data;  // FIN_MASKED: samples;            FIN_PAIR: filtered values
    const u8 *flags;    // FIN_MASKED: flags
    const float *win;   // FIN_PAIR: filtered weights
    float *vout;        // FOUT_PAIR: values; FOUT_BG / FOUT_RESID: background / |data2-bg|
    float *wout;        // FOUT_PAIR: weights
    const float *data2; // FOUT_RESID: the unfiltered samples in the output layout
    float *gring;       // global delay-line scratch (when not in shared memory)
    int64_t gring_stride;
};

#define TC_FILT_U 4  // ticks per unrolled group / prefetch distance

template <bool SMEM_RING, int MODE_IN, int MODE_OUT>
__global__ void k_box_filter(FilterArgs a)
{
    TC_DYN_SMEM(float, sring);
    const int L = 2 * a.r;
    float *ring;
    int64_t rstride, rtid;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    if (SMEM_RING) { ring = sring; rstride = blockDim.x; rtid = threadIdx.x; }
    else { ring = a.gring; rstride = a.gring_stride; rtid = gtid; }
    const int n = a.n, r2 = 2 * a.r, r4 = 4 * a.r;
    const int64_t nj = a.nj;
    const int64_t lstride = (int64_t)L * rstride;
    const int nticks = n + r4 + 3;  // pass p lags pass 1 by p-1 ticks

    for (int64_t line = gtid; line < a.nlines; line += gthreads) {
        const int64_t plane = line / nj;
        const int64_t base = plane * (int64_t)n * nj + (line - plane * nj);
        double s1v = 0, s1w = 0, s2v = 0, s2w = 0, s3v = 0, s3w = 0, s4v = 0, s4w = 0;
        float y1v = 0.f, y1w = 0.f, y2v = 0.f, y2w = 0.f, y3v = 0.f, y3w = 0.f;
        int slot = 0;

        // raw prefetch registers: entering sample (tick m) and leaving sample (m - 2r)
        float ea[TC_FILT_U], eb[TC_FILT_U], la[TC_FILT_U], lb[TC_FILT_U];
        float nea[TC_FILT_U], neb[TC_FILT_U], nla[TC_FILT_U], nlb[TC_FILT_U];

#define TC_FILT_LOAD(m, A, B)                                                        \
        do {                                                                         \
            A = 0.f; B = 0.f;                                                        \
            if ((m) >= 0 && (m) < n) {                                               \
                int64_t idx_ = base + (int64_t)(m) * nj;                             \
                if (MODE_IN == FIN_MASKED) { A = a.data[idx_]; B = a.flags[idx_] ? 0.f : 1.f; } \
                else { A = a.data[idx_]; B = a.win[idx_]; }                          \
            }                                                                        \
        } while (0)

#pragma unroll
        for (int k = 0; k < TC_FILT_U; k++) {
            TC_FILT_LOAD(k, ea[k], eb[k]);
            TC_FILT_LOAD(k - r2, la[k], lb[k]);
        }

        for (int t0 = 0; t0 < nticks; t0 += TC_FILT_U) {
            // issue the loads of the next group first
#pragma unroll
            for (int k = 0; k < TC_FILT_U; k++) {
                const int m = t0 + TC_FILT_U + k;
                TC_FILT_LOAD(m, nea[k], neb[k]);
                TC_FILT_LOAD(m - r2, nla[k], nlb[k]);
            }
#pragma unroll
            for (int k = 0; k < TC_FILT_U; k++) {
                const int tick = t0 + k;
                float *rp = ring + ((int64_t)slot * rstride + rtid);
                // ---- pass 4 (local tick tick-3): entering x3 = y3 of the previous tick
                {
                    const int m = tick - 3;
                    const bool warm = m >= r2;
                    if (warm) { s4v += (double)y3v; s4w += (double)y3w; }
                    const float y4v = (float)s4v, y4w = (float)s4w;
                    if (m >= 0) {
                        const float ov = warm ? rp[4 * lstride] : 0.f, ow = warm ? rp[5 * lstride] : 0.f;
                        rp[4 * lstride] = warm ? y3v : 0.f;
                        rp[5 * lstride] = warm ? y3w : 0.f;
                        s4v -= (double)ov; s4w -= (double)ow;
                    }
                    if (m >= r4 && m < n + r4) {
                        const int64_t idx = base + (int64_t)(m - r4) * nj;
                        const float fv = y4v / a.div, fw = y4w / a.div;
                        if (MODE_OUT == FOUT_PAIR) {
                            a.vout[idx] = fv;
                            a.wout[idx] = fw;
                        } else {
                            float bg = (fw == 0.f) ? NAN : fv / fw;
                            if (MODE_OUT == FOUT_RESID) bg = fabsf(a.data2[idx] - bg);
                            a.vout[idx] = bg;
                        }
                    }
                }
                // ---- pass 3 (local tick tick-2): entering x2 = y2 of the previous tick
                {
                    const int m = tick - 2;
                    if (m >= 0) {
                        s3v += (double)y2v; s3w += (double)y2w;
                        y3v = (float)s3v; y3w = (float)s3w;
                        const bool warm = m >= r2;
                        const float ov = warm ? rp[2 * lstride] : 0.f, ow = warm ? rp[3 * lstride] : 0.f;
                        rp[2 * lstride] = y2v; rp[3 * lstride] = y2w;
                        s3v -= (double)ov; s3w -= (double)ow;
                    }
                }
                // ---- pass 2 (local tick tick-1): entering x1 = y1 of the previous tick
                {
                    const int m = tick - 1;
                    if (m >= 0) {
                        const bool valid = m < n + r2;  // x1 runs off the padded array afterwards
                        if (valid) { s2v += (double)y1v; s2w += (double)y1w; }
                        y2v = (float)s2v; y2w = (float)s2w;
                        const bool warm = m >= r2;
                        const float ov = warm ? rp[0] : 0.f, ow = warm ? rp[lstride] : 0.f;
                        rp[0] = valid ? y1v : 0.f; rp[lstride] = valid ? y1w : 0.f;
                        s2v -= (double)ov; s2w -= (double)ow;
                    }
                }
                // ---- pass 1 (local tick tick): entering line[m], leaving line[m-2r]
                {
                    const int m = tick;
                    if (m < n + r2) {
                        float ev = ea[k], ew = eb[k];
                        if (MODE_IN == FIN_MASKED) ev = (ew != 0.f) ? ev : 0.f;
                        if (m < n) { s1v += (double)ev; s1w += (double)ew; }
                        y1v = (float)s1v; y1w = (float)s1w;
                        if (m >= r2) {
                            float lv = la[k], lw = lb[k];
                            if (MODE_IN == FIN_MASKED) lv = (lw != 0.f) ? lv : 0.f;
                            s1v -= (double)lv; s1w -= (double)lw;
                        }
                    }
                }
                slot++;
                if (slot == L) slot = 0;
            }
#pragma unroll
            for (int k = 0; k < TC_FILT_U; k++) { ea[k] = nea[k]; eb[k] = neb[k]; la[k] = nla[k]; lb[k] = nlb[k]; }
        }
#undef TC_FILT_LOAD
    }
}

// both radii zero (flagging.py:465-466): the "filter" is a copy, so the
// background is data/1 where unflagged and NaN elsewhere
__global__ void __launch_bounds__(256)
k_masked_copy(const float *__restrict__ data, const u8 *__restrict__ flags, int64_t n,
              int mode_out, const float *__restrict__ data2, float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool fl = flags[i] != 0;
    float bg = fl ? NAN : data[i] / 1.0f;
    if (mode_out == FOUT_RESID) bg = fabsf(data2[i] - bg);
    out[i] = bg;
}

static float tc_f32_pow4(int64_t d)
{
    // numba's float32 ** int: square-and-multiply in float32
    volatile float a = (float)d;
    volatile float a2 = a * a;
    volatile float a4 = a2 * a2;
    return a4;
}

template <bool SMEM_RING>
static int launch_box_filter_mode(tc_context *c, const FilterArgs &a, unsigned grid, int bd, size_t smem)
{
#define TC_FILT_CASE(MI, MO)                                                                          \
    if (a.mode_in == MI && a.mode_out == MO) {                                                        \
        if (smem > 48 * 1024)                                                                         \
            TC_CUDA(cudaFuncSetAttribute(k_box_filter<SMEM_RING, MI, MO>,                              \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        TC_LAUNCH_NOSYNC((k_box_filter<SMEM_RING, MI, MO>), grid, bd, smem, c->stream, a);            \
        return TC_OK;                                                                                 \
    }
    TC_FILT_CASE(FIN_MASKED, FOUT_PAIR)
    TC_FILT_CASE(FIN_MASKED, FOUT_BG)
    TC_FILT_CASE(FIN_MASKED, FOUT_RESID)
    TC_FILT_CASE(FIN_PAIR, FOUT_PAIR)
    TC_FILT_CASE(FIN_PAIR, FOUT_BG)
    TC_FILT_CASE(FIN_PAIR, FOUT_RESID)
#undef TC_FILT_CASE
    return tc_fail(TC_ERR_VALUE, "bad filter mode");
}

static int launch_box_filter(tc_context *c, FilterArgs a)
{
    if (a.nlines == 0 || a.n == 0) return TC_OK;
    a.div = tc_f32_pow4(2 * (int64_t)a.r + 1);
    tc_prof_begin(c, TCP_BOX_FILTER);
    const size_t per_thread = (size_t)6 * 2 * a.r * sizeof(float);
    const size_t smem_cap = (size_t)c->smem_optin - 1024;
    int bd = (int)(smem_cap / per_thread);
    bd = bd / 32 * 32;
    if (bd > 128) bd = 128;
    if (bd >= 32) {
        // keep several blocks per SM when the delay lines are short
        while (bd > 32 && (int64_t)tc_blocks_for(a.nlines, bd) < 2 * (int64_t)c->sm_count) bd -= 32;
        size_t smem = per_thread * bd;
        TC_TRY(launch_box_filter_mode<true>(c, a, tc_blocks_for(a.nlines, bd), bd, smem));
    } else {
        // delay lines too deep for shared memory: persistent threads with an
        // L2-resident global ring
        int bd2 = 64;
        int64_t blocks = (int64_t)c->sm_count * 2;
        if (blocks > (int64_t)tc_blocks_for(a.nlines, bd2)) blocks = tc_blocks_for(a.nlines, bd2);
        int64_t threads = blocks * bd2;
        float *g = nullptr;
        TC_TRY(tc_alloc(c, (size_t)threads * 6 * 2 * a.r, &g));
        a.gring = g;
        a.gring_stride = threads;
        TC_TRY(launch_box_filter_mode<false>(c, a, (unsigned)blocks, bd2, 0));
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
