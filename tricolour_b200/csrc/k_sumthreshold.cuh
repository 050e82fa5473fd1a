// k_sumthreshold.cuh -- the remaining stages of _get_baseline_flags
// (flagging.py:921-976): amplitude/averaging prep, NaN interpolation, the 1-D
// SumThreshold scans and the flag combination / dilation / fraction rules.
#pragma once
#include "tc_common.cuh"

// ----------------------------------------------------------------------------
// S1 _average_freq (flagging.py:819-875)
// |complex64| is numba's hypotf: (float) sqrt((double)re*re + (double)im*im)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float tc_abs_c64(float re, float im)
{
    if (isinf(re) || isinf(im)) return INFINITY;
    double dr = (double)re, di = (double)im;
    // both products are exact in float64, so one rounding happens in the add
    return (float)__dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(di, di)));
}

// one thread per averaged sample (cp, t, fa); output in (cp, T, Fa) layout
__global__ void __launch_bounds__(256)
k_prep(const void *__restrict__ vis, int vis_kind, const u8 *__restrict__ flags,
       int64_t total_out, int F, int Fa, int factor, float *__restrict__ out_data,
       u8 *__restrict__ out_flags)
{
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total_out) return;
    int64_t row = o / Fa;
    int fa = (int)(o - row * Fa);
    int f0 = fa * factor;
    int f1 = f0 + factor < F ? f0 + factor : F;
    float sum = 0.0f;
    int cnt = 0;
    for (int f = f0; f < f1; f++) {
        int64_t i = row * F + f;
        float a;
        if (vis_kind == TC_VIS_COMPLEX64) {
            float2 v = ((const float2 *)vis)[i];
            a = tc_abs_c64(v.x, v.y);
        } else {
            a = fabsf(((const float *)vis)[i]);
        }
        if (!flags[i] && !(a != a)) { sum = __fadd_rn(sum, a); cnt++; }
    }
    if (cnt == 0) { out_data[o] = 0.0f; out_flags[o] = 1; }
    else { out_data[o] = __fdiv_rn(sum, (float)cnt); out_flags[o] = 0; }
}

// factor == 1, complex64: four samples per thread
__global__ void __launch_bounds__(256)
k_prep_c64_v4(const float4 *__restrict__ vis, const unsigned *__restrict__ flags, int64_t total4,
              float4 *__restrict__ out_data, unsigned *__restrict__ out_flags)
{
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total4) return;
    const float4 a = vis[2 * o], b = vis[2 * o + 1];
    const unsigned f = flags[o];
    float amp[4] = {tc_abs_c64(a.x, a.y), tc_abs_c64(a.z, a.w), tc_abs_c64(b.x, b.y), tc_abs_c64(b.z, b.w)};
    unsigned of = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool bad = ((f >> (8 * k)) & 0xffu) || amp[k] != amp[k];
        // a lone sample: sum = 0 + a, average = sum / 1
        amp[k] = bad ? 0.0f : __fdiv_rn(__fadd_rn(0.0f, amp[k]), 1.0f);
        of |= bad ? (1u << (8 * k)) : 0u;
    }
    out_data[o] = make_float4(amp[0], amp[1], amp[2], amp[3]);
    out_flags[o] = of;
}

// The same with both layouts written at once: a block owns a tile of 32 dumps x
// 32 channels, writes it to (cp, T, F) directly and to (cp, F, T) through a
// shared-memory transpose.  grid (F / 32, T / 32, cp), 256 threads.
__global__ void __launch_bounds__(256)
k_prep_c64_tile(const float4 *__restrict__ vis, const unsigned *__restrict__ flags, int T, int F,
                float4 *__restrict__ d_TF, unsigned *__restrict__ f_TF, float4 *__restrict__ d_FT,
                unsigned *__restrict__ f_FT)
{
    __shared__ float td[32][33];
    __shared__ u8 tf[32][36];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int64_t cp = blockIdx.z;
    const int t0 = blockIdx.y * 32, f0 = blockIdx.x * 32;
    {
        const int64_t idx = ((cp * T + t0 + ty) * (int64_t)F + f0 + tx * 4) >> 2;   // in units of 4 samples
        const float4 a = vis[2 * idx], b = vis[2 * idx + 1];
        const unsigned f = flags[idx];
        float amp[4] = {tc_abs_c64(a.x, a.y), tc_abs_c64(a.z, a.w), tc_abs_c64(b.x, b.y), tc_abs_c64(b.z, b.w)};
        unsigned of = 0u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool bad = ((f >> (8 * k)) & 0xffu) || amp[k] != amp[k];
            amp[k] = bad ? 0.0f : __fdiv_rn(__fadd_rn(0.0f, amp[k]), 1.0f);
            of |= bad ? (1u << (8 * k)) : 0u;
            td[tx * 4 + k][ty] = amp[k];
            tf[tx * 4 + k][ty] = bad ? 1 : 0;
        }
        d_TF[idx] = make_float4(amp[0], amp[1], amp[2], amp[3]);
        f_TF[idx] = of;
    }
    __syncthreads();
    {
        // row ty of the transposed tile = channel f0 + ty, dumps t0 + 4 tx .. + 3
        const int64_t idx = ((cp * F + f0 + ty) * (int64_t)T + t0 + tx * 4) >> 2;
        d_FT[idx] = make_float4(td[ty][tx * 4], td[ty][tx * 4 + 1], td[ty][tx * 4 + 2], td[ty][tx * 4 + 3]);
        f_FT[idx] = (unsigned)tf[ty][tx * 4] | ((unsigned)tf[ty][tx * 4 + 1] << 8) |
                    ((unsigned)tf[ty][tx * 4 + 2] << 16) | ((unsigned)tf[ty][tx * 4 + 3] << 24);
    }
}

// flags[(cp,t,f)] |= spec[(cp,f)]   (flagging.py:954)
__global__ void __launch_bounds__(256)
k_or_spec(u8 *__restrict__ flags, const u8 *__restrict__ spec, int64_t total, int T, int Fa)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t row = i / Fa;
    int f = (int)(i - row * Fa);
    int64_t cp = row / T;
    if (spec[cp * Fa + f]) flags[i] = 1;
}

// 16 flags per thread; layout (cp, T, Fa) with Fa % 16 == 0
__global__ void __launch_bounds__(256)
k_or_spec_tf16(uint4 *__restrict__ flags, const uint4 *__restrict__ spec, int64_t total16, int T, int F16)
{
    // grid (ceil(F16 / 256), T, planes): no index divisions
    (void)total16;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F16) return;
    const int64_t cp = blockIdx.z;
    const int64_t i = (cp * T + blockIdx.y) * (int64_t)F16 + f;
    const uint4 sp = spec[cp * F16 + f];
    if (sp.x | sp.y | sp.z | sp.w) {
        uint4 v = flags[i];
        v.x |= sp.x; v.y |= sp.y; v.z |= sp.z; v.w |= sp.w;
        flags[i] = v;
    }
}

// the same on the transposed layout (cp, Fa, T) with T % 16 == 0: one spectrum flag per row of T
__global__ void __launch_bounds__(256)
k_or_spec_ft16(uint4 *__restrict__ flags, const u8 *__restrict__ spec, int64_t total16, int T16)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total16) return;
    if (spec[i / T16]) flags[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
}

// out = a - b elementwise
__global__ void __launch_bounds__(256)
k_sub(const float *a, const float *b, float *out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = a[i] - b[i];
}

// out = a | b for byte flags
__global__ void __launch_bounds__(256)
k_or_bytes(const u8 *__restrict__ a, const u8 *__restrict__ b, u8 *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (u8)((a[i] | b[i]) ? 1 : 0);
}

// ----------------------------------------------------------------------------
// S5 _linearly_interpolate_nans (flagging.py:307-359): one thread per line,
// line element i at base + i*stride (lines are the frequency axis).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_interp_nans(float *__restrict__ d, int64_t nlines, int64_t ninner, int64_t outer_stride,
              int64_t inner_stride, int n, int64_t stride)
{
    int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nlines) return;
    int64_t outer = line / ninner, inner = line - outer * ninner;
    float *x = d + outer * outer_stride + inner * inner_stride;
#define X(i) x[(int64_t)(i) * stride]
    int p = 0;
    while (p < n && X(p) != X(p)) p++;
    if (p == n) {
        for (int i = 0; i < n; i++) X(i) = 0.0f;
        return;
    }
    float first = X(p);
    for (int i = 0; i < p; i++) X(i) = first;  // extrapolate backwards
    p += 1;
    while (p < n) {
        float cur = X(p);
        if (cur != cur) {
            int q = p + 1;
            while (q < n && X(q) != X(q)) q++;
            float start = X(p - 1);
            if (q == n) {
                for (int i = p; i < n; i++) X(i) = start;  // extrapolate forwards
            } else {
                // float32 difference, true division by an int64 -> float64
                double grad = __ddiv_rn((double)(X(q) - start), (double)(q - (p - 1)));
                for (int i = p; i < q; i++)
                    X(i) = (float)__dadd_rn((double)start, __dmul_rn((double)(i - (p - 1)), grad));
            }
            p = q;
        } else {
            p += 1;
        }
    }
#undef X
}

// Warp-per-line form for contiguous lines: two tile sweeps with ballots find,
// for every NaN, the nearest valid sample on either side (`rv` is scratch for the
// right neighbours), after which every sample is independent.  Writes
// out = interp(bg), or out = minuend - interp(bg) when `minuend` is given
// (flagging.py:962 fused).
__global__ void __launch_bounds__(128)
k_interp_nans_rows(const float *__restrict__ bg, const float *__restrict__ minuend,
                   float *__restrict__ out, int *__restrict__ rv, int64_t nlines, int n)
{
    const int lane = threadIdx.x & 31;
    const int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (line >= nlines) return;
    const float *x = bg + line * (int64_t)n;
    int *r = rv + line * (int64_t)n;
    const int ntiles = (n + 31) / 32;
    // Both sweeps walk the row in tiles of 32 with a carried index; the samples of
    // four tiles are loaded up front so that one memory latency covers four steps.
    // backward sweep: index of the next valid sample at or after i (n if none)
    int carry = n;
    int nnan = 0;
    for (int t4 = ntiles - 1; t4 >= 0; t4 -= 4) {
        float xs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = (t4 - u) * 32 + lane;
            xs[u] = (t4 - u >= 0 && i < n) ? x[i] : NAN;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = t4 - u;
            if (t < 0) break;
            const int i = t * 32 + lane;
            const bool valid = i < n && !(xs[u] != xs[u]);
            const unsigned m = __ballot_sync(TC_FULL_MASK, valid);
            const unsigned mm = m >> lane;
            if (i < n && !valid) r[i] = mm ? i + __ffs((int)mm) - 1 : carry;   // only NaN samples look it up
            if (m) carry = t * 32 + __ffs((int)m) - 1;
            nnan += (i < n && !valid) ? 1 : 0;
        }
    }
    __syncwarp();
    // forward sweep: previous valid sample at or before i (-1 if none), then fill
    int carl = -1;
    for (int t4 = 0; t4 < ntiles; t4 += 4) {
        float xs[4], ms[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = (t4 + u) * 32 + lane;
            const bool in = t4 + u < ntiles && i < n;
            xs[u] = in ? x[i] : 0.f;
            ms[u] = (in && minuend) ? minuend[line * (int64_t)n + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = t4 + u;
            if (t >= ntiles) break;
            const int i = t * 32 + lane;
            const float xi = xs[u];
            const bool valid = i < n && !(xi != xi);
            const unsigned m = __ballot_sync(TC_FULL_MASK, valid);
            const unsigned below = m & (0xffffffffu >> (31 - lane));
            const int lv = below ? t * 32 + 31 - __clz((int)below) : carl;
            if (m) carl = t * 32 + 31 - __clz((int)m);
            if (i < n) {
                float val = xi;
                if (!valid) {
                    const int rr = r[i];
                    if (lv < 0 && rr >= n) val = 0.0f;
                    else if (lv < 0) val = x[rr];
                    else if (rr >= n) val = x[lv];
                    else {
                        const float start = x[lv];
                        // float32 difference, true division by an int64 -> float64
                        const double grad = __ddiv_rn((double)(x[rr] - start), (double)(rr - lv));
                        val = (float)__dadd_rn((double)start, __dmul_rn((double)(i - lv), grad));
                    }
                }
                out[line * (int64_t)n + i] = minuend ? ms[u] - val : val;
            }
        }
    }
    (void)nnan;
}

// ----------------------------------------------------------------------------
// S9 _sum_threshold1d (flagging.py:610-681) with _convolve_flags (582-607).
// One thread scans one (line, chunk): for every window, a sequential float64
// prefix sum of the clamped samples over the padded chunk, window sums as
// prefix differences, and smearing of every flagged window over its samples.
// Samples of a line: data[base + i*estride]; neighbouring threads own
// neighbouring lines (stride 1), so every access is coalesced.
// Scratch per thread: a float64 prefix ring and one byte of pos/neg state per
// padded sample, both laid out [chunk][sample][line] like the data.
// ----------------------------------------------------------------------------
#define TC_MAX_WINDOWS 16
#ifndef TC_ST_MINBLOCKS
#define TC_ST_MINBLOCKS 5   // 96 registers: measured best of 3..5 blocks of 128 threads per SM
#endif
struct StScanArgs {
    const float *data;
    const float *thr;         // [line*nchunks + chunk] float32 thresholds (inf = none)
    u8 *out;                  // same layout as data
    int64_t nlines;           // total lines (all planes)
    int64_t ninner;           // lines per plane (contiguous)
    int64_t outer_stride;     // plane stride in samples
    int64_t estride;          // distance between consecutive samples of a line (= ninner)
    int n;                    // line length
    int nchunks;
    const int64_t *chunk_ends;  // device [nchunks+1]
    int nwin;
    int maxw;
    int64_t windows[TC_MAX_WINDOWS];
    double tf[TC_MAX_WINDOWS];
    float scale[TC_MAX_WINDOWS];
    int mpad;                 // max padded chunk length
    double *cum;              // [nplanes][nchunks][mpad+1][ninner]
    u8 *pn;                   // [nplanes][nchunks][mpad][ninner] pos/neg state (ping)
    u8 *pn2;                  // same size (pong)
    int fused1248;            // windows are exactly [1, 2, 4, 8]: single-sweep kernel, no scratch
};

// One SumThreshold window over a padded chunk.  The pos/neg state of the
// previous windows is read from `pin` and the updated state written to `pout`
// (two separate scratch planes, so that loads can run ahead of the stores; a
// pass never reads a byte it has already rewritten).  The prefix values the
// window sums need (cum[i+1-w]) stay in a w-deep register ring when w is 1, 2,
// 4 or 8 (every default.yaml time window and all but one frequency list); other
// widths spill the prefix to the coalesced global scratch.
template <int W>
__device__ __forceinline__ void st_window_reg(const float *__restrict__ d, const u8 *__restrict__ pin,
                                              u8 *__restrict__ pout, int m, int64_t es, int64_t ss,
                                              double limit, double sc, double nsc)
{
    double h[W];
    u8 sr[W];
#pragma unroll
    for (int k = 0; k < W; k++) { h[k] = 0.0; sr[k] = 0; }
    double c = 0.0;
    int lastpos = -(1 << 30), lastneg = -(1 << 30);
    for (int i0 = 0; i0 < m; i0 += 8) {
        float xv[8];
        u8 sv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + k;
            xv[k] = i < m ? d[(int64_t)i * es] : 0.f;
            sv[k] = (pin && i < m) ? pin[(int64_t)i * ss] : (u8)0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + k;
            if (i < m) {
                double x = (double)xv[k];
                const u8 st = sv[k];
                if ((st & 1) && x > limit) x = limit;
                else if ((st & 2) && x < -limit) x = -limit;
                c = c + x;
                const int j = i + 1 - W;
                const double cj = h[k % W];   // cum[j] (0 for j == 0)
                const u8 sj = W == 1 ? st : sr[(k + 1) % W];  // state of sample j, stored W-1 steps ago
                h[k % W] = c;
                sr[k % W] = st;
                if (j >= 0) {
                    const double avg = c - cj;
                    if (avg * sc > limit) lastpos = j;
                    if (avg * nsc > limit) lastneg = j;
                    const u8 add = (u8)(((j - lastpos < W) ? 1 : 0) | ((j - lastneg < W) ? 2 : 0));
                    pout[(int64_t)j * ss] = (u8)(sj | add);
                }
            }
        }
    }
    int jt = m - W + 1; if (jt < 0) jt = 0;
    for (int j = jt; j < m; j++) {
        const u8 add = (u8)(((j - lastpos < W) ? 1 : 0) | ((j - lastneg < W) ? 2 : 0));
        const u8 sj = pin ? pin[(int64_t)j * ss] : (u8)0;
        pout[(int64_t)j * ss] = (u8)(sj | add);
    }
}

__device__ __forceinline__ void st_window_mem(const float *__restrict__ d, const u8 *__restrict__ pin,
                                              u8 *__restrict__ pout, double *cum, int m, int w, int64_t es,
                                              int64_t ss, double limit, double sc, double nsc)
{
    double c = 0.0;
    cum[0] = 0.0;
    int lastpos = -(1 << 30), lastneg = -(1 << 30);
    int i = 0;
    // Blocks of four samples with all their loads issued up front (one memory latency
    // per block instead of per sample).  cum[j .. j + 3] of a block were written by
    // earlier blocks as soon as w >= 4, which holds for every width that gets here
    // except 3 (widths 1, 2, 4, 8 use the register rings).
    if (w >= 4) {
        for (; i + 4 <= m; i += 4) {
            float xs[4];
            u8 si[4], sj[4];
            double cj[4];
            const int j0 = i + 1 - w;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                xs[k] = d[(int64_t)(i + k) * es];
                si[k] = pin ? pin[(int64_t)(i + k) * ss] : (u8)0;
                const int j = j0 + k;
                cj[k] = j >= 0 ? cum[(int64_t)j * ss] : 0.0;
                sj[k] = (pin && j >= 0) ? pin[(int64_t)j * ss] : (u8)0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double x = (double)xs[k];
                const u8 st = si[k];
                if ((st & 1) && x > limit) x = limit;
                else if ((st & 2) && x < -limit) x = -limit;
                c = c + x;
                cum[(int64_t)(i + k + 1) * ss] = c;
                const int j = j0 + k;
                if (j >= 0) {
                    const double avg = c - cj[k];
                    if (avg * sc > limit) lastpos = j;
                    if (avg * nsc > limit) lastneg = j;
                    const u8 add = (u8)(((j - lastpos < w) ? 1 : 0) | ((j - lastneg < w) ? 2 : 0));
                    pout[(int64_t)j * ss] = (u8)(sj[k] | add);
                }
            }
        }
    }
    for (; i < m; i++) {
        double x = (double)d[(int64_t)i * es];
        const u8 st = pin ? pin[(int64_t)i * ss] : (u8)0;
        if ((st & 1) && x > limit) x = limit;
        else if ((st & 2) && x < -limit) x = -limit;
        c = c + x;
        cum[(int64_t)(i + 1) * ss] = c;
        const int j = i + 1 - w;
        if (j >= 0) {
            const double avg = c - cum[(int64_t)j * ss];
            if (avg * sc > limit) lastpos = j;
            if (avg * nsc > limit) lastneg = j;
            const u8 add = (u8)(((j - lastpos < w) ? 1 : 0) | ((j - lastneg < w) ? 2 : 0));
            const u8 sj = pin ? pin[(int64_t)j * ss] : (u8)0;
            pout[(int64_t)j * ss] = (u8)(sj | add);
        }
    }
    int jt = m - w + 1; if (jt < 0) jt = 0;
    for (int j = jt; j < m; j++) {
        const u8 add = (u8)(((j - lastpos < w) ? 1 : 0) | ((j - lastneg < w) ? 2 : 0));
        const u8 sj = pin ? pin[(int64_t)j * ss] : (u8)0;
        pout[(int64_t)j * ss] = (u8)(sj | add);
    }
}

// All four windows of the default lists [1, 2, 4, 8] in ONE sweep.  Window k
// only needs the flags of window k-1 at the sample it is processing, and those
// are final W(k-1)-1 samples after window k-1 passed it, so the windows can
// follow each other at fixed lags (0, 0, 1, 4 samples): the line is read once,
// the pos/neg state lives in small register rings and never touches memory,
// and the four float64 prefix chains give the scheduler independent work.
// Per window the arithmetic is exactly st_window_reg's (sequential prefix,
// prefix differences, smear), so the flags are identical.
struct StW {
    double c;
    int lastpos, lastneg;
};

template <int W>
__device__ __forceinline__ u8 st_step(StW &w, double *h, int slot, bool have, float xf, u8 st, int i,
                                      double limit, double sc, double nsc)
{
    // processes sample i (if `have`), evaluates the window starting at j = i+1-W and
    // returns the pos/neg bits that window coverage adds to sample j
    const int j = i + 1 - W;
    if (have) {
        double x = (double)xf;
        if ((st & 1) && x > limit) x = limit;
        else if ((st & 2) && x < -limit) x = -limit;
        w.c = w.c + x;
        const double cj = h[slot];
        h[slot] = w.c;
        if (j >= 0) {
            const double avg = w.c - cj;
            if (avg * sc > limit) w.lastpos = j;
            if (avg * nsc > limit) w.lastneg = j;
        }
    }
    return (u8)(((j - w.lastpos < W) ? 1 : 0) | ((j - w.lastneg < W) ? 2 : 0));
}

__device__ __forceinline__ void st_fused_1248(const float *__restrict__ d, u8 *__restrict__ out, int m, int rel,
                                              int nout, int64_t es, float thr, const double *tf,
                                              const float *scale)
{
    const double l0 = (double)thr / tf[0], l1 = (double)thr / tf[1], l2 = (double)thr / tf[2], l3 = (double)thr / tf[3];
    const double s0 = (double)scale[0], s1 = (double)scale[1], s2 = (double)scale[2], s3 = (double)scale[3];
    const double n0 = (double)(-scale[0]), n1 = (double)(-scale[1]), n2 = (double)(-scale[2]), n3 = (double)(-scale[3]);
    StW w0 = {0.0, -(1 << 30), -(1 << 30)}, w1 = w0, w2 = w0, w3 = w0;
    double h0[1] = {0.0}, h1[2] = {0.0, 0.0}, h2[4] = {0.0, 0.0, 0.0, 0.0}, h3[8];
    float xr[8];      // x[s - q] lives in xr[(s - q) & 7]
    u8 p0prev = 0;    // pn_0(s-1)
    u8 p1r[4];        // pn_1(j) in p1r[j & 3]
    u8 p2r[8];        // pn_2(j) in p2r[j & 7]
#pragma unroll
    for (int k = 0; k < 8; k++) { h3[k] = 0.0; xr[k] = 0.f; p2r[k] = 0; }
#pragma unroll
    for (int k = 0; k < 4; k++) p1r[k] = 0;
    // window positions at step s: i0 = i1 = s, i2 = s - 1, i3 = s - 4; the last
    // sample (m-1) leaves window 3 at step m - 1 + 7 + 4
    for (int s0i = 0; s0i < m + 11; s0i += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int s = s0i + k;
            const float xs = s < m ? d[(int64_t)s * es] : 0.f;
            xr[k] = xs;                                            // (s & 7) == k
            // window 0 (W=1) at sample s, finalises pn_0(s)
            u8 p0 = 0;
            if (s < m) p0 = st_step<1>(w0, h0, 0, true, xs, 0, s, l0, s0, n0);
            // window 1 (W=2) at sample s with pn_0(s), finalises pn_1(s-1)
            {
                const u8 add = st_step<2>(w1, h1, k & 1, s < m, xs, p0, s, l1, s1, n1);
                const int j = s - 1;
                if (j >= 0 && j < m) p1r[(k + 3) & 3] = (u8)(p0prev | add);
            }
            p0prev = p0;
            // window 2 (W=4) at sample s-1 with pn_1(s-1), finalises pn_2(s-4)
            {
                const int i = s - 1;
                const bool have = i >= 0 && i < m;
                const u8 add = st_step<4>(w2, h2, (k + 3) & 3, have, xr[(k + 7) & 7], p1r[(k + 3) & 3], i, l2, s2, n2);
                const int j = i - 3;
                if (j >= 0 && j < m) p2r[(k + 4) & 7] = (u8)(p1r[k & 3] | add);   // (j & 3) == (k & 3), (j & 7) == ((k+4) & 7)
            }
            // window 3 (W=8) at sample s-4 with pn_2(s-4), finalises pn_3(s-11)
            {
                const int i = s - 4;
                const bool have = i >= 0 && i < m;
                const u8 add = st_step<8>(w3, h3, (k + 4) & 7, have, xr[(k + 4) & 7], p2r[(k + 4) & 7], i, l3, s3, n3);
                const int j = i - 7;
                if (j >= 0 && j < m) {
                    const u8 fin = (u8)(p2r[(k + 5) & 7] | add);               // (j & 7) == ((k + 5) & 7)
                    const int o = j - rel;
                    if (o >= 0 && o < nout) out[(int64_t)o * es] = fin ? 1 : 0;
                }
            }
        }
    }
}

// Branch-free form of the single-sweep scan (same per-window arithmetic, so
// the flags are identical).  Differences from st_fused_1248:
//   * the line is cut into blocks of 8 steps; interior blocks (every window
//     inside the line, every output inside the chunk) carry no predicates;
//   * "x > limit" on a float32 sample is decided in float32 against the
//     smallest float32 above the float64 limit (exactly the same predicate);
//   * avg * (-scale) > limit is evaluated as avg * scale < -limit (the product
//     only changes sign), one multiply per window instead of two;
//   * the pos / neg smear is a 2-bit-per-step shift register instead of "last
//     flagged start" indices;
//   * loads and stores walk pointers instead of recomputing 64-bit indices.
struct St2Win {
    double c;        // running prefix sum (cum[i + 1])
    unsigned h;      // bit 2q: the window that started q steps ago exceeded +limit; bit 2q+1: -limit
    double limit, sc;
    float hif;       // smallest float32 > limit (NaN when there is none)
};

__device__ __forceinline__ float st2_float_above(double limit)
{
    // smallest float32 f with (double)f > limit
    if (!(limit < 3.0e38)) return NAN;              // inf / NaN limits never clamp
    float f = __double2float_rn(limit);
    if (!((double)f > limit)) f = __uint_as_float(__float_as_uint(f) + 1u);   // limit >= 0: next float up
    if (!((double)f > limit)) f = __uint_as_float(__float_as_uint(f) + 1u);
    return f;
}

template <int W, bool INTERIOR>
__device__ __forceinline__ unsigned st2_step(St2Win &w, double *ring, int slot, float xf, unsigned st, int j, int nj)
{
    // processes sample i = j + W - 1 and returns the pos/neg bits that window
    // coverage adds to sample j.  nj = number of window starts (m - W + 1)
    const double xd = (double)xf;
    double x1 = xd;
    if (W > 1) {
        const bool ph = (st & 1u) && xf >= w.hif;
        const bool pl = (st & 2u) && xf <= -w.hif;
        x1 = ph ? w.limit : (pl ? -w.limit : xd);
    }
    w.c = w.c + x1;
    const double cj = ring[slot];
    ring[slot] = w.c;
    const double avg = w.c - cj;
    const double t = W == 1 ? avg : avg * w.sc;
    const bool ok = INTERIOR || ((unsigned)j < (unsigned)nj);
    const bool pos = ok && t > w.limit;
    const bool neg = ok && t < -w.limit;
    w.h = (w.h << 2) | (pos ? 1u : 0u) | (neg ? 2u : 0u);
    const unsigned MP = W == 1 ? 0x1u : (W == 2 ? 0x5u : (W == 4 ? 0x55u : 0x5555u));
    return ((w.h & MP) ? 1u : 0u) | ((w.h & (MP << 1)) ? 2u : 0u);
}

// Window k+1 trails window k by one step more than it strictly has to (sample
// positions s, s-1, s-3, s-7; outputs at s-14), so that within a step no window
// waits for another: the four float64 chains are independent.  The pos/neg
// state of the samples in flight lives in 2-bit-per-sample shift registers.
__device__ __forceinline__ void st_fused_1248_v2(const float *__restrict__ d, u8 *__restrict__ out, int m, int rel,
                                                 int nout, int64_t es, float thr, const double *tf,
                                                 const float *scale)
{
    St2Win w0, w1, w2, w3;
    w0.limit = (double)thr / tf[0]; w1.limit = (double)thr / tf[1];
    w2.limit = (double)thr / tf[2]; w3.limit = (double)thr / tf[3];
    w0.sc = (double)scale[0]; w1.sc = (double)scale[1]; w2.sc = (double)scale[2]; w3.sc = (double)scale[3];
    w0.hif = st2_float_above(w0.limit); w1.hif = st2_float_above(w1.limit);
    w2.hif = st2_float_above(w2.limit); w3.hif = st2_float_above(w3.limit);
    w0.c = w1.c = w2.c = w3.c = 0.0;
    w0.h = w1.h = w2.h = w3.h = 0u;
    const int nj0 = m, nj1 = m - 1 > 0 ? m - 1 : 0, nj2 = m - 3 > 0 ? m - 3 : 0, nj3 = m - 7 > 0 ? m - 7 : 0;
    double h0[1] = {0.0}, h1[2] = {0.0, 0.0}, h2[4] = {0.0, 0.0, 0.0, 0.0}, h3[8];
    float xr[8];             // x[s - q] in xr[(s - q) & 7]
    // state histories, newest sample in bits 1:0 (before this step's push):
    unsigned p0h = 0u;       // pn_0(s - 1), pn_0(s - 2), ...
    unsigned p1h = 0u;       // pn_1(s - 3), pn_1(s - 4), ...
    unsigned p2h = 0u;       // pn_2(s - 7), pn_2(s - 8), ...
#pragma unroll
    for (int k = 0; k < 8; k++) { h3[k] = 0.0; xr[k] = 0.f; }
    const float *pd = d;                              // sample s
    u8 *po = out + (int64_t)(-14 - rel) * es;         // output of sample j = s - 14 (only dereferenced in range)
    const int nsteps = m + 14;
    // window starts at step s: j0 = s, j1 = s - 2, j2 = s - 6, j3 = s - 14
    for (int s0 = 0; s0 < nsteps; s0 += 8) {
        // interior: all starts valid for the whole block, all 8 outputs inside the chunk
        const bool interior = s0 >= 14 + rel && s0 + 7 < nj3 && s0 + 7 - 14 - rel < nout;
        if (interior) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float xs = *pd;
                pd += es;
                xr[k] = xs;
                const unsigned a0 = st2_step<1, true>(w0, h0, 0, xs, 0u, 0, 0);
                const unsigned a1 = st2_step<2, true>(w1, h1, (k + 1) & 1, xr[(k + 7) & 7], p0h & 3u, 0, 0);
                const unsigned a2 = st2_step<4, true>(w2, h2, (k + 5) & 3, xr[(k + 5) & 7], p1h & 3u, 0, 0);
                const unsigned a3 = st2_step<8, true>(w3, h3, (k + 1) & 7, xr[(k + 1) & 7], p2h & 3u, 0, 0);
                *po = (((p2h >> 14) & 3u) | a3) ? 1 : 0;
                po += es;
                const unsigned n1 = ((p0h >> 2) & 3u) | a1, n2 = ((p1h >> 6) & 3u) | a2;
                p0h = (p0h << 2) | a0; p1h = (p1h << 2) | n1; p2h = (p2h << 2) | n2;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int s = s0 + k;
                const float xs = s < m ? *pd : 0.f;
                pd += es;
                xr[k] = xs;
                const unsigned a0 = st2_step<1, false>(w0, h0, 0, xs, 0u, s, nj0);
                const unsigned a1 = st2_step<2, false>(w1, h1, (k + 1) & 1, xr[(k + 7) & 7], p0h & 3u, s - 2, nj1);
                const unsigned a2 = st2_step<4, false>(w2, h2, (k + 5) & 3, xr[(k + 5) & 7], p1h & 3u, s - 6, nj2);
                const unsigned a3 = st2_step<8, false>(w3, h3, (k + 1) & 7, xr[(k + 1) & 7], p2h & 3u, s - 14, nj3);
                const int o = s - 14 - rel;
                if ((unsigned)o < (unsigned)nout && s - 14 < m) *po = (((p2h >> 14) & 3u) | a3) ? 1 : 0;
                po += es;
                const unsigned n1 = ((p0h >> 2) & 3u) | a1, n2 = ((p1h >> 6) & 3u) | a2;
                p0h = (p0h << 2) | a0; p1h = (p1h << 2) | n1; p2h = (p2h << 2) | n2;
            }
        }
    }
}

__global__ void __launch_bounds__(128, TC_ST_MINBLOCKS)
k_st_scan(StScanArgs a)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.nlines * a.nchunks) return;
    // thread order: (plane, chunk, inner) so that a warp shares plane and chunk
    int64_t per_plane = (int64_t)a.nchunks * a.ninner;
    int64_t plane = g / per_plane;
    int64_t rem = g - plane * per_plane;
    int chunk = (int)(rem / a.ninner);
    int64_t inner = rem - (int64_t)chunk * a.ninner;
    int64_t line = plane * a.ninner + inner;

    int c0 = (int)a.chunk_ends[chunk], c1 = (int)a.chunk_ends[chunk + 1];
    int p0 = c0 - a.maxw + 1; if (p0 < 0) p0 = 0;
    int p1 = c1 + a.maxw - 1; if (p1 > a.n) p1 = a.n;
    int m = p1 - p0;
    const float *d = a.data + plane * a.outer_stride + inner + (int64_t)p0 * a.estride;
    u8 *o = a.out + plane * a.outer_stride + inner + (int64_t)c0 * a.estride;
    int64_t sbase = ((plane * a.nchunks + chunk) * (int64_t)(a.mpad + 1)) * a.ninner + inner;
    double *cum = a.cum ? a.cum + sbase : nullptr;
    int64_t pbase = ((plane * a.nchunks + chunk) * (int64_t)a.mpad) * a.ninner + inner;
    u8 *pa = a.pn ? a.pn + pbase : nullptr, *pb = a.pn2 ? a.pn2 + pbase : nullptr;
    const int64_t es = a.estride, ss = a.ninner;
    float thr = a.thr[line * a.nchunks + chunk];

    if (a.fused1248 == 2) {
        st_fused_1248_v2(d, o, m, c0 - p0, c1 - c0, es, thr, a.tf, a.scale);
        return;
    }
    if (a.fused1248) {
        st_fused_1248(d, o, m, c0 - p0, c1 - c0, es, thr, a.tf, a.scale);
        return;
    }
    const u8 *pin = nullptr;   // no state before the first window
    u8 *pout = pa;
    for (int wi = 0; wi < a.nwin; wi++) {
        const int w = (int)a.windows[wi];
        const double limit = (double)thr / a.tf[wi];
        const double sc = (double)a.scale[wi];
        const double nsc = (double)(-a.scale[wi]);
        if (w == 1) st_window_reg<1>(d, pin, pout, m, es, ss, limit, sc, nsc);
        else if (w == 2) st_window_reg<2>(d, pin, pout, m, es, ss, limit, sc, nsc);
        else if (w == 4) st_window_reg<4>(d, pin, pout, m, es, ss, limit, sc, nsc);
        else if (w == 8) st_window_reg<8>(d, pin, pout, m, es, ss, limit, sc, nsc);
        else st_window_mem(d, pin, pout, cum, m, w, es, ss, limit, sc, nsc);
        pin = pout;
        pout = (pout == pa) ? pb : pa;
    }
    const u8 *pn = pin;
    int rel = c0 - p0;
    for (int i = 0; i < c1 - c0; i++) o[(int64_t)i * es] = (pn && pn[(int64_t)(rel + i) * ss]) ? 1 : 0;
}

// ----------------------------------------------------------------------------
// S10 _combine_flags (flagging.py:784-816) + S11 _unaverage_freq (878-918)
// ----------------------------------------------------------------------------
// c1[t,fa] = any over the time window of (spec | time | freq); all (cp,T,Fa)
__global__ void __launch_bounds__(256)
k_combine_time(const u8 *__restrict__ spec, const u8 *__restrict__ time_f,
               const u8 *__restrict__ freq_f, int64_t total, int T, int Fa, int lo, int ext,
               u8 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t row = i / Fa;
    int f = (int)(i - row * Fa);
    int64_t cp = row / T;
    int t = (int)(row - cp * T);
    if (spec[cp * Fa + f]) { out[i] = ext > 0 ? 1 : 0; return; }
    int t0 = t + lo < 0 ? 0 : t + lo;
    int t1 = t + lo + ext > T ? T : t + lo + ext;
    u8 any = 0;
    for (int tt = t0; tt < t1; tt++) {
        int64_t k = (cp * T + tt) * (int64_t)Fa + f;
        any |= (u8)(time_f[k] | freq_f[k]);
    }
    out[i] = any ? 1 : 0;
}

// d[t,f] = any over the frequency window of c1[t, f'/avg]; per-row and
// per-column totals of d (the counts taken BEFORE the row extension,
// flagging.py:902-911).  One block per (cp, t) row.
__global__ void __launch_bounds__(256)
k_unaverage_rows(const u8 *__restrict__ c1, int T, int Fa, int F, int lo, int ext, int avg,
                 u8 *__restrict__ d, int *__restrict__ rowcnt, int *__restrict__ colcnt)
{
    __shared__ int s_cnt;
    int64_t row = blockIdx.x;  // cp*T + t
    int64_t cp = row / T;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const u8 *src = c1 + row * Fa;
    int local = 0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        int f0 = f + lo < 0 ? 0 : f + lo;
        int f1 = f + lo + ext > F ? F : f + lo + ext;
        u8 any = 0;
        for (int ff = f0; ff < f1; ff++) any |= src[ff / avg];
        any = any ? 1 : 0;
        d[row * F + f] = any;
        if (any) { local++; atomicAdd(&colcnt[cp * F + f], 1); }
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) rowcnt[row] = s_cnt;
}

// out = d | row rule | column rule | isnan(input)   (flagging.py:910-918, 776-781)
// optionally accumulates iter_flags |= out for the next major iteration
__global__ void __launch_bounds__(256)
k_finalize_flags(const u8 *__restrict__ d, const int *__restrict__ rowcnt,
                 const int *__restrict__ colcnt, const void *__restrict__ vis, int vis_kind,
                 int64_t total, int T, int F, double row_limit, double col_limit,
                 u8 *__restrict__ out, u8 *__restrict__ iter_flags)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t row = i / F;
    int f = (int)(i - row * F);
    int64_t cp = row / T;
    bool fl = d[i] != 0;
    if ((double)rowcnt[row] > row_limit) fl = true;
    if ((double)colcnt[cp * F + f] > col_limit) fl = true;
    if (vis) {
        if (vis_kind == TC_VIS_COMPLEX64) {
            float2 v = ((const float2 *)vis)[i];
            if (v.x != v.x || v.y != v.y) fl = true;
        } else {
            float v = ((const float *)vis)[i];
            if (v != v) fl = true;
        }
    }
    out[i] = fl ? 1 : 0;
    if (iter_flags && fl) iter_flags[i] = 1;
}

// ----------------------------------------------------------------------------
// 16-bytes-per-thread forms of the three kernels above for average_freq == 1 and
// F % 16 == 0 (flag bytes are 0/1 throughout, so bytewise OR is the logical OR
// and a popcount of a word is the number of set flags in it).
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 tc_or4(uint4 a, uint4 b)
{
    return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
}

// c1 = ext > 0 ? spec | OR over the time window of (time | freq) : 0
__global__ void __launch_bounds__(256)
k_combine_time_v16(const uint4 *__restrict__ spec, const uint4 *__restrict__ time_f,
                   const uint4 *__restrict__ freq_f, int64_t total16, int T, int F16, int lo, int ext,
                   uint4 *__restrict__ out)
{
    // grid (ceil(F16 / 256), T, planes): no index divisions
    (void)total16;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F16) return;
    const int64_t cp = blockIdx.z;
    const int t = blockIdx.y;
    const int64_t i = (cp * T + t) * (int64_t)F16 + f;
    uint4 any = make_uint4(0u, 0u, 0u, 0u);
    if (ext > 0) {
        any = spec[cp * F16 + f];
        int t0 = t + lo < 0 ? 0 : t + lo;
        int t1 = t + lo + ext > T ? T : t + lo + ext;
        for (int tt = t0; tt < t1; tt++) {
            int64_t k = (cp * T + tt) * (int64_t)F16 + f;
            any = tc_or4(any, tc_or4(time_f[k], freq_f[k]));
        }
    }
    out[i] = any;
}

// d[t, f] = OR over [f + lo, f + lo + ext) of c1[t, .]; rowcnt[row] = number of
// set d in the row (before the row rule).  One block per row, the row (plus 16
// zero bytes on either side) staged in shared memory as words.
__global__ void __launch_bounds__(256)
k_dilate_rows_v16(const uint4 *__restrict__ c1, int F16, int lo, int ext, uint4 *__restrict__ d,
                  int *__restrict__ rowcnt)
{
    TC_DYN_SMEM(unsigned, row);          // [4 pad][F/4][4 pad] words
    __shared__ int s_cnt;
    const int64_t r = blockIdx.x;
    const int nw = F16 * 4;
    if (threadIdx.x == 0) s_cnt = 0;
    if (threadIdx.x < 4) { row[threadIdx.x] = 0u; row[4 + nw + threadIdx.x] = 0u; }
    for (int v = threadIdx.x; v < F16; v += blockDim.x) {
        uint4 q = c1[r * F16 + v];
        row[4 + 4 * v] = q.x; row[5 + 4 * v] = q.y; row[6 + 4 * v] = q.z; row[7 + 4 * v] = q.w;
    }
    __syncthreads();
    int local = 0;
    for (int v = threadIdx.x; v < F16; v += blockDim.x) {
        unsigned o[4] = {0u, 0u, 0u, 0u};
        for (int sft = lo; sft < lo + ext; sft++) {
            // bytes [16 v + sft, 16 v + sft + 16) of the row; |sft| < 16 keeps it inside the padding
            const int b0 = 16 * v + sft + 16;          // byte offset in the padded buffer
            const int w0 = b0 >> 2, sh = (b0 & 3) * 8;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned a = row[w0 + q], b = row[w0 + q + 1];
                o[q] |= sh ? ((a >> sh) | (b << (32 - sh))) : a;
            }
        }
        d[r * F16 + v] = make_uint4(o[0], o[1], o[2], o[3]);
        local += __popc(o[0]) + __popc(o[1]) + __popc(o[2]) + __popc(o[3]);
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(TC_FULL_MASK, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) rowcnt[r] = s_cnt;
}

// colcnt[cp, f] = number of set d[cp, :, f]; one thread per word (4 channels)
__global__ void __launch_bounds__(128)
k_colcnt_v4(const unsigned *__restrict__ d, int T, int F4, int64_t total4, int *__restrict__ colcnt)
{
    // grid (ceil(F4 / 128), planes): no index divisions
    (void)total4;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F4) return;
    const int64_t cp = blockIdx.y;
    const int64_t i = cp * F4 + f;
    const unsigned *p = d + cp * (int64_t)T * F4 + f;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int t0 = 0; t0 < T; t0 += 128) {
        unsigned acc = 0u;                     // four byte counters, at most 128 each
        const int t1 = t0 + 128 < T ? t0 + 128 : T;
        for (int t = t0; t < t1; t++) acc += p[(int64_t)t * F4];
        c0 += acc & 0xffu; c1 += (acc >> 8) & 0xffu; c2 += (acc >> 16) & 0xffu; c3 += acc >> 24;
    }
    reinterpret_cast<int4 *>(colcnt)[i] = make_int4((int)c0, (int)c1, (int)c2, (int)c3);
}

// out = d | row rule | column rule | isnan(input), four samples per thread
__global__ void __launch_bounds__(256)
k_finalize_flags_v4(const unsigned *__restrict__ d, const int *__restrict__ rowcnt,
                    const int4 *__restrict__ colcnt, const void *__restrict__ vis, int vis_kind,
                    int64_t total4, int T, int F4, double row_limit, double col_limit,
                    unsigned *__restrict__ out, unsigned *__restrict__ iter_flags)
{
    // grid (ceil(F4 / 256), T, planes): no index divisions
    (void)total4;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F4) return;
    const int64_t cp = blockIdx.z;
    const int64_t row = cp * T + blockIdx.y;
    const int64_t i = row * F4 + f;
    unsigned w = d[i];
    if ((double)rowcnt[row] > row_limit) w = 0x01010101u;
    const int4 cc = colcnt[cp * F4 + f];
    if ((double)cc.x > col_limit) w |= 0x1u;
    if ((double)cc.y > col_limit) w |= 0x100u;
    if ((double)cc.z > col_limit) w |= 0x10000u;
    if ((double)cc.w > col_limit) w |= 0x1000000u;
    if (vis) {
        if (vis_kind == TC_VIS_COMPLEX64) {
            const float4 a = ((const float4 *)vis)[2 * i], b = ((const float4 *)vis)[2 * i + 1];
            if (a.x != a.x || a.y != a.y) w |= 0x1u;
            if (a.z != a.z || a.w != a.w) w |= 0x100u;
            if (b.x != b.x || b.y != b.y) w |= 0x10000u;
            if (b.z != b.z || b.w != b.w) w |= 0x1000000u;
        } else {
            const float4 a = ((const float4 *)vis)[i];
            if (a.x != a.x) w |= 0x1u;
            if (a.y != a.y) w |= 0x100u;
            if (a.z != a.z) w |= 0x10000u;
            if (a.w != a.w) w |= 0x1000000u;
        }
    }
    out[i] = w;
    if (iter_flags && w) iter_flags[i] |= w;
}

// normalise arbitrary non-zero flag bytes to 0/1
__global__ void __launch_bounds__(256)
k_norm_flags(const u8 *__restrict__ in, u8 *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[i] ? 1 : 0;
}

// the same, 16 flags per thread (both pointers 16-byte aligned, n16 = n / 16)
__global__ void __launch_bounds__(256)
k_norm_flags_v16(const uint4 *__restrict__ in, uint4 *__restrict__ out, int64_t n16)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    uint4 v = in[i];
    unsigned *w = reinterpret_cast<unsigned *>(&v);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // every non-zero byte -> 1: fold the bits of each byte into its lowest bit
        unsigned x = w[q];
        x |= x >> 4; x |= x >> 2; x |= x >> 1;
        w[q] = x & 0x01010101u;
    }
    out[i] = v;
}

static int launch_norm_flags(tc_context *c, const u8 *in, u8 *out, int64_t n)
{
    if (n <= 0) return TC_OK;
    const int64_t n16 = ((((uintptr_t)in | (uintptr_t)out) & 15) == 0) ? n / 16 : 0;
    if (n16) TC_LAUNCH_NOSYNC(k_norm_flags_v16, tc_blocks_for(n16, 256), 256, 0, c->stream, (const uint4 *)in, (uint4 *)out, n16);
    if (n - 16 * n16)
        TC_LAUNCH_NOSYNC(k_norm_flags, tc_blocks_for(n - 16 * n16, 256), 256, 0, c->stream, in + 16 * n16, out + 16 * n16,
                         n - 16 * n16);
    return TC_OK;
}
