// k_filter.cuh -- masked box-Gaussian filter (reference: _box_gaussian_filter1d
// flagging.py:362-419, _box_gaussian_filter 422-466, masked_gaussian_filter
// 469-513).
//
// The reference makes K=4 sequential box passes over a zero-padded copy of
// every line, each pass a float64 running sum whose outputs are rounded to
// float32.  Here all four passes of a line are fused into ONE streaming loop:
// pass p+1 consumes what pass p emitted 2r ticks earlier, so a line is read
// once and written once, and the only state is four float64 accumulators and
// three 2r-deep float32 delay lines per array.  The order of floating point
// operations on every accumulator is exactly the reference's (add the entering
// sample, round+emit, subtract the leaving sample), so results are bit-exact.
//
// One thread owns one line and filters the `value` and the `weight` array of
// the masked filter together (two independent dependency chains per pass).
// Lines are addressed so that neighbouring threads touch neighbouring
// addresses: element i of line j of plane p lives at p*n*nj + i*nj + j.  Along
// time that is the (T,F) layout with j = channel; along frequency it is the
// transposed (F,T) layout with j = time.
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

template <bool SMEM_RING>
__global__ void k_box_filter(FilterArgs a)
{
    TC_DYN_SMEM(float, sring);
    const int L = 2 * a.r;
    float *ring;
    int64_t rstride, rtid;
    int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    if (SMEM_RING) { ring = sring; rstride = blockDim.x; rtid = threadIdx.x; }
    else { ring = a.gring; rstride = a.gring_stride; rtid = gtid; }
    const int n = a.n, r2 = 2 * a.r, r4 = 4 * a.r;
    const int64_t nj = a.nj;

    for (int64_t line = gtid; line < a.nlines; line += gthreads) {
        int64_t plane = line / nj;
        int64_t base = plane * (int64_t)n * nj + (line - plane * nj);
        double s1v = 0, s1w = 0, s2v = 0, s2w = 0, s3v = 0, s3w = 0, s4v = 0, s4w = 0;
        int slot = 0;
        for (int m = 0; m < n + r4; m++) {
            float y1v = 0.f, y1w = 0.f;
            const bool u1 = m < n + r2;
            if (u1) {
                // pass 1: entering sample x0[4r+m] = line[m]
                if (m < n) {
                    int64_t idx = base + (int64_t)m * nj;
                    float v, w;
                    if (a.mode_in == FIN_MASKED) {
                        bool fl = a.flags[idx] != 0;
                        w = fl ? 0.f : 1.f;
                        v = fl ? 0.f : a.data[idx];
                    } else {
                        v = a.data[idx];
                        w = a.win[idx];
                    }
                    s1v += (double)v;
                    s1w += (double)w;
                }
                y1v = (float)s1v;
                y1w = (float)s1w;
                if (m >= r2) {
                    // leaving sample x0[2r+m] = line[m-2r]
                    int64_t idx = base + (int64_t)(m - r2) * nj;
                    float v, w;
                    if (a.mode_in == FIN_MASKED) {
                        bool fl = a.flags[idx] != 0;
                        w = fl ? 0.f : 1.f;
                        v = fl ? 0.f : a.data[idx];
                    } else {
                        v = a.data[idx];
                        w = a.win[idx];
                    }
                    s1v -= (double)v;
                    s1w -= (double)w;
                }
                // pass 2: entering x1[2r+m]
                s2v += (double)y1v;
                s2w += (double)y1w;
            }
            const bool warm = m >= r2;  // delay lines hold real samples from tick 2r on
            float *rp = ring + ((int64_t)slot * rstride + rtid);
            const int64_t lstride = (int64_t)L * rstride;
            float y2v = (float)s2v, y2w = (float)s2w;
            {
                float ov = warm ? rp[0] : 0.f, ow = warm ? rp[lstride] : 0.f;
                rp[0] = y1v; rp[lstride] = y1w;  // zero once x1 runs off the array
                s2v -= (double)ov; s2w -= (double)ow;
            }
            // pass 3: entering x2[m]
            s3v += (double)y2v; s3w += (double)y2w;
            float y3v = (float)s3v, y3w = (float)s3w;
            {
                float ov = warm ? rp[2 * lstride] : 0.f, ow = warm ? rp[3 * lstride] : 0.f;
                rp[2 * lstride] = y2v; rp[3 * lstride] = y2w;
                s3v -= (double)ov; s3w -= (double)ow;
            }
            // pass 4: entering x3[m-2r] (exists from tick 2r on)
            if (warm) { s4v += (double)y3v; s4w += (double)y3w; }
            float y4v = (float)s4v, y4w = (float)s4w;
            {
                float ov = warm ? rp[4 * lstride] : 0.f, ow = warm ? rp[5 * lstride] : 0.f;
                rp[4 * lstride] = warm ? y3v : 0.f; rp[5 * lstride] = warm ? y3w : 0.f;
                s4v -= (double)ov; s4w -= (double)ow;
            }
            if (m >= r4) {
                int64_t idx = base + (int64_t)(m - r4) * nj;
                float fv = y4v / a.div, fw = y4w / a.div;
                if (a.mode_out == FOUT_PAIR) {
                    a.vout[idx] = fv;
                    a.wout[idx] = fw;
                } else {
                    float bg = (fw == 0.f) ? NAN : fv / fw;
                    if (a.mode_out == FOUT_RESID) bg = fabsf(a.data2[idx] - bg);
                    a.vout[idx] = bg;
                }
            }
            slot++;
            if (slot == L) slot = 0;
        }
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
        if (smem > 48 * 1024)
            TC_CUDA(cudaFuncSetAttribute(k_box_filter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        TC_LAUNCH_NOSYNC(k_box_filter<true>, tc_blocks_for(a.nlines, bd), bd, smem, c->stream, a);
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
        TC_LAUNCH_NOSYNC(k_box_filter<false>, (unsigned)blocks, bd2, 0, c->stream, a);
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
