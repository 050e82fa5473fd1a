/*
 * tricolour_oracle.c -- CPU restatement of tricolour's flagging hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * implementation in tricolour_b200/csrc.  It may be used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product path never links, loads or calls it.
 *
 * Every function restates, in plain sequential C, the numba/numpy function
 * of the reference it cites (paths relative to /root/reference/).  Parity
 * status: PINNED -- tests/test_oracle_golden.py checks this file bit-for-bit
 * against fixtures produced by running the reference itself in the build
 * container (tests/golden/make_golden.py) and against the known-answer
 * vectors of the reference's own unit tests.
 *
 * Numerics follow what numba 0.65 / glibc 2.39 do for the reference:
 *   |complex64|       = (float) sqrt((double)re*re + (double)im*im)   (glibc hypotf)
 *   np.median         = exact order statistics; even count -> (double)(float)(a+b) / 2
 *   box filter        = float64 running sum, float32 store after every pass,
 *                       divisor = float32 power by squaring
 *   thresholds        = float64 (background) / float32->float64 (SumThreshold)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAD_NORMAL 1.4826 /* tricolour/flagging.py:22 */

typedef uint8_t u8;

/* ------------------------------------------------------------------------- */
/* exact order statistics (stands in for numba/np/arraymath.py:1573-1635)    */
/* ------------------------------------------------------------------------- */

static void select_kth(float *a, int64_t n, int64_t k)
{
    /* iterative quickselect with median-of-3; on return a[k] is the k-th
     * smallest, a[0..k) <= a[k] <= a(k..n) */
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        float t;
        if (a[mid] < a[lo]) { t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
        if (a[hi] < a[mid]) { t = a[hi]; a[hi] = a[mid]; a[mid] = t; }
        if (a[mid] < a[lo]) { t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
        float pivot = a[mid];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (a[i] < pivot) i++;
            while (pivot < a[j]) j--;
            if (i <= j) {
                t = a[i]; a[i] = a[j]; a[j] = t;
                i++; j--;
            }
        }
        if (k <= j) hi = j;
        else if (k >= i) lo = i;
        else break;
    }
}

/* np.median of a float32 scratch array (destroyed); returns float64 exactly
 * as numba's _median_inner does (arraymath.py:1621-1635). n > 0. */
static double median_f32(float *a, int64_t n)
{
    int64_t half = n >> 1;
    select_kth(a, n, half);
    if (n & 1) return (double)a[half];
    float lower = a[0];
    for (int64_t i = 1; i < half; i++)
        if (a[i] > lower) lower = a[i];
    float s = lower + a[half]; /* float32 add */
    return (double)s / 2.0;
}

double orc_median(const float *x, int64_t n)
{
    if (n == 0) return NAN;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(tmp, x, sizeof(float) * (size_t)n);
    double m = median_f32(tmp, n);
    free(tmp);
    return m;
}

/* ------------------------------------------------------------------------- */
/* F1 flag_nans_and_zeros  -- tricolour/flagging.py:29-62                     */
/* ------------------------------------------------------------------------- */
void orc_flag_nans_zeros(const float *vis /* (n,2) re,im */, const u8 *flags,
                         u8 *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float re = vis[2 * i], im = vis[2 * i + 1];
        int flag = (re == 0.0f && im == 0.0f) || isnan(re) || isnan(im);
        out[i] = (u8)(flag || flags[i] != 0);
    }
}

/* ------------------------------------------------------------------------- */
/* S1 _average_freq -- tricolour/flagging.py:819-875                          */
/* ------------------------------------------------------------------------- */
static inline float abs_c64(float re, float im)
{
    /* numba lowers np.abs(complex64) to hypotf; glibc >= 2.35 evaluates it in
     * double and rounds once more to float. */
    if (isinf(re) || isinf(im)) return INFINITY;
    return (float)sqrt((double)re * (double)re + (double)im * (double)im);
}

/* is_complex: 1 -> in_data is (cp,T,F,2) float32 pairs, 0 -> (cp,T,F) float32.
 * weight_bits selects the accumulator width of avg_weight (factor.dtype,
 * flagging.py:848): 8, 16, 32 or 64. */
void orc_average_freq(const float *in_data, int is_complex, const u8 *in_flags,
                      int64_t ncp, int64_t T, int64_t F, int64_t factor,
                      int weight_bits, float *avg_data, u8 *avg_flags)
{
    int64_t Fa = (F + factor - 1) / factor;
    uint64_t wmask = weight_bits >= 64 ? ~(uint64_t)0
                                       : (((uint64_t)1 << weight_bits) - 1);
    uint64_t *w = (uint64_t *)calloc((size_t)Fa, sizeof(uint64_t));
    for (int64_t cp = 0; cp < ncp; cp++)
        for (int64_t t = 0; t < T; t++) {
            const float *row = in_data + (size_t)(cp * T + t) * F * (is_complex ? 2 : 1);
            const u8 *frow = in_flags + (size_t)(cp * T + t) * F;
            float *orow = avg_data + (size_t)(cp * T + t) * Fa;
            u8 *oflag = avg_flags + (size_t)(cp * T + t) * Fa;
            for (int64_t f = 0; f < Fa; f++) { orow[f] = 0.0f; w[f] = 0; }
            for (int64_t f = 0; f < F; f++) {
                float a = is_complex ? abs_c64(row[2 * f], row[2 * f + 1])
                                     : fabsf(row[f]);
                if (!frow[f] && !isnan(a)) {
                    orow[f / factor] += a;
                    w[f / factor] = (w[f / factor] + 1) & wmask;
                }
            }
            for (int64_t f = 0; f < Fa; f++) {
                if (w[f] == 0) { orow[f] = 0.0f; oflag[f] = 1; }
                else { orow[f] = orow[f] / (float)w[f]; oflag[f] = 0; }
            }
        }
    free(w);
}

/* ------------------------------------------------------------------------- */
/* S2 _time_median -- tricolour/flagging.py:226-264                           */
/* ------------------------------------------------------------------------- */
void orc_time_median(const float *data, const u8 *flags, int64_t T, int64_t F,
                     float *out_data, u8 *out_flags)
{
    float *tmp = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
    for (int64_t f = 0; f < F; f++) {
        int64_t n = 0;
        for (int64_t t = 0; t < T; t++)
            if (!flags[t * F + f]) tmp[n++] = data[t * F + f];
        if (n == 0) { out_data[f] = 0.0f; out_flags[f] = 1; }
        else { out_data[f] = (float)median_f32(tmp, n); out_flags[f] = 0; }
    }
    free(tmp);
}

/* S3 _median_abs on a (T, f0:f1) sub-block of a (T,F) array -- flagging.py:267-279 */
double orc_median_abs(const float *data, const u8 *flags, int64_t T, int64_t F,
                      int64_t f0, int64_t f1)
{
    int64_t cap = T * (f1 - f0);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)(cap > 0 ? cap : 1));
    int64_t n = 0;
    for (int64_t t = 0; t < T; t++)
        for (int64_t f = f0; f < f1; f++)
            if (!flags[t * F + f]) tmp[n++] = fabsf(data[t * F + f]);
    double m = n ? median_f32(tmp, n) : NAN;
    free(tmp);
    return m;
}

/* S4 _median_abs_axis0 on a strided line -- flagging.py:282-304; returns the
 * float32-rounded value the reference stores in out_data. */
static float median_abs_line(const float *data, const u8 *flags, int64_t n,
                             int64_t stride, float *tmp)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++)
        if (!flags[i * stride]) tmp[m++] = fabsf(data[i * stride]);
    if (m == 0) return NAN;
    return (float)median_f32(tmp, m);
}

/* exported 2-D form used by the unit tests: out[j] for j in [0,ncols) */
void orc_median_abs_axis0(const float *data, const u8 *flags, int64_t nrows,
                          int64_t ncols, float *out)
{
    float *tmp = (float *)malloc(sizeof(float) * (size_t)(nrows > 0 ? nrows : 1));
    for (int64_t j = 0; j < ncols; j++)
        out[j] = median_abs_line(data + j, flags + j, nrows, ncols, tmp);
    free(tmp);
}

/* ------------------------------------------------------------------------- */
/* S5 _linearly_interpolate_nans1d -- flagging.py:307-345                     */
/* ------------------------------------------------------------------------- */
void orc_interp_nans1d(float *d, int64_t n)
{
    int64_t p = 0;
    while (p < n && isnan(d[p])) p++;
    if (p == n) { for (int64_t i = 0; i < n; i++) d[i] = 0.0f; return; }
    for (int64_t i = 0; i < p; i++) d[i] = d[p];
    p += 1;
    while (p < n) {
        if (isnan(d[p])) {
            int64_t q = p + 1;
            while (q < n && isnan(d[q])) q++;
            if (q == n) {
                for (int64_t i = p; i < n; i++) d[i] = d[p - 1];
            } else {
                float start = d[p - 1];
                /* float32 difference, then true division by an int64 -> float64 */
                double grad = (double)(d[q] - start) / (double)(q - (p - 1));
                for (int64_t i = p; i < q; i++)
                    d[i] = (float)((double)start + (double)(i - (p - 1)) * grad);
            }
            p = q;
        } else {
            p += 1;
        }
    }
}

void orc_interp_nans(float *d, int64_t T, int64_t F)
{
    for (int64_t t = 0; t < T; t++) orc_interp_nans1d(d + t * F, F);
}

/* ------------------------------------------------------------------------- */
/* S6 _box_gaussian_filter1d -- flagging.py:362-419                           */
/* one line of length n with element stride `stride`; `padded` is scratch of */
/* n + r*K floats.  In-place (out == data) is allowed.                       */
/* ------------------------------------------------------------------------- */
static float f32_int_power(float a, int64_t b)
{
    /* numba's float32 ** int64: multiply-and-square in float32 */
    float r = 1.0f;
    int64_t e = b < 0 ? -b : b;
    while (e != 0) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    return b < 0 ? 1.0f / r : r;
}

static void box_filter_line(const float *data, int64_t n, int64_t stride,
                            int64_t r, int K, float *out, int64_t ostride,
                            float *padded)
{
    if (n == 0 || K == 0) {
        for (int64_t i = 0; i < n; i++) out[i * ostride] = data[i * stride];
        return;
    }
    int64_t d = 2 * r + 1;
    int64_t padding = r * K;
    int64_t plen = n + padding;
    for (int64_t i = 0; i < padding; i++) padded[i] = 0.0f;
    for (int64_t i = 0; i < n; i++) padded[padding + i] = data[i * stride];
    int64_t prev_start = padding;
    for (int p = 1; p <= K; p++) {
        double s = 0.0;
        int64_t start = padding - 2 * r * p;
        int64_t stop = start + n + 2 * padding;
        if (start < 0) start = 0;
        if (stop > plen) stop = plen;
        int64_t tail = stop < plen - 2 * r ? stop : plen - 2 * r;
        int64_t lim = start + 2 * r < plen ? start + 2 * r : plen;
        for (int64_t i = prev_start; i < lim; i++) s += (double)padded[i];
        for (int64_t i = start; i < tail; i++) {
            s += (double)padded[i + 2 * r];
            float prev = padded[i];
            padded[i] = (float)s;
            s -= (double)prev;
        }
        for (int64_t i = tail; i < stop; i++) {
            float prev = padded[i];
            padded[i] = (float)s;
            s -= (double)prev;
        }
        prev_start = start;
    }
    float div = f32_int_power((float)d, K);
    for (int64_t i = 0; i < n; i++) out[i * ostride] = padded[i] / div;
}

void orc_box_filter1d(const float *data, int64_t n, int64_t r, int K, float *out)
{
    float *padded = (float *)malloc(sizeof(float) * (size_t)(n + r * K + 1));
    box_filter_line(data, n, 1, r, K, out, 1, padded);
    free(padded);
}

/* _box_gaussian_filter -- flagging.py:422-466.  r0/r1 are the radii the
 * reference derives at line 451 (computed by the caller in float64). */
void orc_box_gaussian_filter(const float *data, int64_t T, int64_t F, int64_t r0,
                             int64_t r1, int K, float *out)
{
    int64_t maxn = T > F ? T : F;
    int64_t maxr = r0 > r1 ? r0 : r1;
    float *padded = (float *)malloc(sizeof(float) * (size_t)(maxn + maxr * K + 1));
    const float *src = data;
    int need_copy = 1;
    if (r0 > 0) {
        for (int64_t f = 0; f < F; f++)
            box_filter_line(src + f, T, F, r0, K, out + f, F, padded);
        src = out;
        need_copy = 0;
    }
    if (r1 > 0) {
        for (int64_t t = 0; t < T; t++)
            box_filter_line(src + t * F, F, 1, r1, K, out + t * F, 1, padded);
        need_copy = 0;
    }
    if (need_copy && out != data) memcpy(out, data, sizeof(float) * (size_t)(T * F));
    free(padded);
}

/* S7 masked_gaussian_filter -- flagging.py:469-513 */
void orc_masked_gaussian_filter(const float *data, const u8 *flags, int64_t T,
                                int64_t F, int64_t r0, int64_t r1, int K,
                                float *out)
{
    int64_t N = T * F;
    float *weight = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; i++) {
        weight[i] = flags[i] ? 0.0f : 1.0f;
        out[i] = flags[i] ? 0.0f : data[i];
    }
    orc_box_gaussian_filter(weight, T, F, r0, r1, K, weight);
    orc_box_gaussian_filter(out, T, F, r0, r1, K, out);
    for (int64_t i = 0; i < N; i++) {
        if (weight[i] == 0.0f) out[i] = NAN;
        else out[i] = out[i] / weight[i];
    }
    free(weight);
}

/* ------------------------------------------------------------------------- */
/* S8 _get_background2d -- flagging.py:516-579                                */
/* radii: (iterations+1) x 2 int64, row k for extend_factor = iterations-k,   */
/* last row for the final filter (sigma = spike_width).                      */
/* ------------------------------------------------------------------------- */
void orc_get_background2d(const float *data, const u8 *flags_in, int64_t T,
                          int64_t F, int iterations, const int64_t *radii,
                          double reject_threshold, const int64_t *chunk_ends,
                          int nchunk_ends, float *background)
{
    int64_t N = T * F;
    u8 *flags = (u8 *)malloc((size_t)(N > 0 ? N : 1));
    memcpy(flags, flags_in, (size_t)N);
    for (int it = 0; it < iterations; it++) {
        orc_masked_gaussian_filter(data, flags, T, F, radii[2 * it],
                                   radii[2 * it + 1], 4, background);
        for (int c = 0; c + 1 < nchunk_ends; c++) {
            int64_t f0 = chunk_ends[c], f1 = chunk_ends[c + 1];
            for (int64_t t = 0; t < T; t++)
                for (int64_t f = f0; f < f1; f++)
                    background[t * F + f] = fabsf(data[t * F + f] - background[t * F + f]);
            double threshold = orc_median_abs(background, flags, T, F, f0, f1);
            threshold *= MAD_NORMAL * reject_threshold;
            for (int64_t t = 0; t < T; t++)
                for (int64_t f = f0; f < f1; f++)
                    if ((double)background[t * F + f] > threshold) flags[t * F + f] = 1;
        }
    }
    orc_masked_gaussian_filter(data, flags, T, F, radii[2 * iterations],
                               radii[2 * iterations + 1], 4, background);
    orc_interp_nans(background, T, F);
    free(flags);
}

/* ------------------------------------------------------------------------- */
/* S9 _convolve_flags / _sum_threshold1d / _sum_threshold                     */
/*    flagging.py:582-607, 610-681, 684-742                                   */
/* one line (stride `stride`) of length n; chunks over the line.              */
/* tf[w] = pow(rho, log2(window)) is supplied by the caller (glibc, float64)  */
/* ------------------------------------------------------------------------- */
static void sum_threshold_line(const float *data, const u8 *in_flags, u8 *out_flags,
                               int64_t n, int64_t stride, const int64_t *windows,
                               const double *tf, int nwin, double outlier_nsigma,
                               const int64_t *chunks, int nchunk_ends, float *ftmp,
                               double *cum, u8 *pos, u8 *neg, uint32_t *fcum)
{
    int64_t maxw = 0;
    for (int w = 0; w < nwin; w++) if (windows[w] > maxw) maxw = windows[w];
    for (int ci = 0; ci + 1 < nchunk_ends; ci++) {
        int64_t c0 = chunks[ci], c1 = chunks[ci + 1];
        float thr = median_abs_line(data + c0 * stride, in_flags + c0 * stride,
                                    c1 - c0, stride, ftmp);
        double threshold_scale = outlier_nsigma * MAD_NORMAL;
        if (isnan(thr)) thr = INFINITY;
        else thr = (float)((double)thr * threshold_scale);
        int64_t p0 = c0 - maxw + 1; if (p0 < 0) p0 = 0;
        int64_t p1 = c1 + maxw - 1; if (p1 > n) p1 = n;
        int64_t m = p1 - p0;
        const float *pd = data + p0 * stride;
        memset(pos, 0, (size_t)m);
        memset(neg, 0, (size_t)m);
        for (int wi = 0; wi < nwin; wi++) {
            int64_t window = windows[wi];
            double limit = (double)thr / tf[wi];
            cum[0] = 0.0;
            for (int64_t i = 0; i < m; i++) {
                double clamped = (double)pd[i * stride];
                if (pos[i] && clamped > limit) clamped = limit;
                else if (neg[i] && clamped < -limit) clamped = -limit;
                cum[i + 1] = cum[i] + clamped;
            }
            float rolling_scale = (float)(1.0 / (double)window);
            int64_t navg = m + 1 - window; /* len(cum[window:] - cum[:-window]) */
            if (navg < 0) navg = 0;
            for (int pol = 0; pol < 2; pol++) {
                double scale = pol == 0 ? (double)rolling_scale : (double)(-rolling_scale);
                u8 *of = pol == 0 ? pos : neg;
                /* _convolve_flags (flagging.py:582-607) */
                int64_t cum_size = navg + 2 * window - 1;
                for (int64_t i = 0; i < window; i++) fcum[i] = 0;
                for (int64_t i = 0; i < navg; i++) {
                    double avg = cum[i + window] - cum[i];
                    uint32_t flag = (avg * scale > limit) ? 1u : 0u;
                    fcum[window + i] = fcum[window + i - 1] + flag;
                }
                for (int64_t i = cum_size - (window - 1); i < cum_size; i++)
                    fcum[i] = fcum[cum_size - window];
                for (int64_t i = 0; i < m; i++)
                    of[i] |= (u8)(fcum[i + window] - fcum[i] != 0);
            }
        }
        int64_t rel = c0 - p0;
        for (int64_t i = 0; i < c1 - c0; i++)
            out_flags[(c0 + i) * stride] = (u8)(pos[rel + i] | neg[rel + i]);
    }
}

void orc_sum_threshold(const float *data, const u8 *in_flags, int64_t T, int64_t F,
                       int axis, const int64_t *windows, const double *tf, int nwin,
                       double outlier_nsigma, const int64_t *chunks,
                       int nchunk_ends, u8 *out_flags)
{
    int64_t n = axis == 0 ? T : F;
    int64_t nlines = axis == 0 ? F : T;
    int64_t maxw = 1;
    for (int w = 0; w < nwin; w++) if (windows[w] > maxw) maxw = windows[w];
    float *ftmp = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    double *cum = (double *)malloc(sizeof(double) * (size_t)(n + 2));
    u8 *pos = (u8 *)malloc((size_t)(n + 1));
    u8 *neg = (u8 *)malloc((size_t)(n + 1));
    uint32_t *fcum = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 2 * maxw + 2));
    int64_t default_chunks[2] = {0, n};
    if (chunks == NULL) { chunks = default_chunks; nchunk_ends = 2; }
    for (int64_t l = 0; l < nlines; l++) {
        int64_t off = axis == 0 ? l : l * F;
        int64_t stride = axis == 0 ? F : 1;
        sum_threshold_line(data + off, in_flags + off, out_flags + off, n, stride,
                           windows, tf, nwin, outlier_nsigma, chunks, nchunk_ends,
                           ftmp, cum, pos, neg, fcum);
    }
    free(ftmp); free(cum); free(pos); free(neg); free(fcum);
}

/* ------------------------------------------------------------------------- */
/* S10 _combine_flags -- flagging.py:784-816.  The cumulative sum lives in    */
/* time_extend.dtype (extend_bits wide) and wraps like the reference's.       */
/* ------------------------------------------------------------------------- */
void orc_combine_flags(const u8 *spec_flags, const u8 *time_flags,
                       const u8 *freq_flags, int64_t T, int64_t F,
                       int64_t time_extend, int extend_bits, u8 *out)
{
    uint64_t mask = extend_bits >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << extend_bits) - 1);
    uint64_t *fs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)((T + 1) * F + 1));
    for (int64_t f = 0; f < F; f++) fs[f] = 0;
    for (int64_t t = 0; t < T; t++)
        for (int64_t f = 0; f < F; f++) {
            uint64_t flag = (spec_flags[f] || time_flags[t * F + f] || freq_flags[t * F + f]) ? 1 : 0;
            fs[(t + 1) * F + f] = (fs[t * F + f] + flag) & mask;
        }
    int64_t lo = -(time_extend / 2);
    int64_t hi = lo + time_extend;
    for (int64_t t = 0; t < T; t++) {
        int64_t t0 = t + lo < 0 ? 0 : t + lo;
        int64_t t1 = t + hi > T ? T : t + hi;
        for (int64_t f = 0; f < F; f++)
            out[t * F + f] = (u8)(fs[t0 * F + f] != fs[t1 * F + f]);
    }
    free(fs);
}

/* ------------------------------------------------------------------------- */
/* S11 _unaverage_freq -- flagging.py:878-918                                 */
/* flags: (T, Fa) averaged flags; out: (T, F)                                 */
/* ------------------------------------------------------------------------- */
void orc_unaverage_freq(const u8 *flags, int64_t T, int64_t Fa, int64_t F,
                        int64_t freq_extend, int64_t average_freq,
                        double flag_all_time_frac, double flag_all_freq_frac, u8 *out)
{
    int32_t *fs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F + 1));
    int32_t *fst = (int32_t *)calloc((size_t)(F > 0 ? F : 1), sizeof(int32_t));
    int64_t lo = -(freq_extend / 2);
    int64_t hi = lo + freq_extend;
    for (int64_t t = 0; t < T; t++) {
        fs[0] = 0;
        for (int64_t f = 0; f < F; f++)
            fs[f + 1] = fs[f] + (flags[t * Fa + f / average_freq] ? 1 : 0);
        int64_t tot = 0;
        for (int64_t f = 0; f < F; f++) {
            int64_t f0 = f + lo < 0 ? 0 : f + lo;
            int64_t f1 = f + hi > F ? F : f + hi;
            int flag = fs[f1] != fs[f0];
            out[t * F + f] = (u8)flag;
            tot += flag;
            fst[f] += flag;
        }
        if ((double)tot > flag_all_freq_frac * (double)F)
            for (int64_t f = 0; f < F; f++) out[t * F + f] = 1;
    }
    for (int64_t f = 0; f < F; f++)
        if ((double)fst[f] > (double)T * flag_all_time_frac)
            for (int64_t t = 0; t < T; t++) out[t * F + f] = 1;
    free(fs); free(fst);
}

/* ------------------------------------------------------------------------- */
/* parameter block shared by S12/S13                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
    double outlier_nsigma;
    int nwin_time;  const int64_t *windows_time;  const double *tf_time;
    int nwin_freq;  const int64_t *windows_freq;  const double *tf_freq;
    double background_reject;
    int background_iterations;
    const int64_t *radii_spec; /* (iterations+1) x 2, sigma = ef*(0, spike_f)   */
    const int64_t *radii_2d;   /* (iterations+1) x 2, sigma = ef*(spike_t, spike_f) */
    int64_t time_extend; int time_extend_bits;
    int64_t freq_extend;
    int nchunk_ends; const int64_t *freq_chunk_ends;
    int64_t average_freq; int average_freq_bits;
    double flag_all_time_frac, flag_all_freq_frac;
} orc_params;

/* S12 _get_baseline_flags -- flagging.py:921-976.  data/flags are the (T,Fa)
 * outputs of _average_freq and are modified in place like the reference's. */
void orc_get_baseline_flags(float *data, u8 *flags, int64_t T, int64_t Fa, int64_t F,
                            const orc_params *p, u8 *out_flags)
{
    int64_t N = T * Fa;
    float *spec_data = (float *)malloc(sizeof(float) * (size_t)(Fa + 1));
    float *spec_bg = (float *)malloc(sizeof(float) * (size_t)(Fa + 1));
    u8 *spec_flags = (u8 *)malloc((size_t)(Fa + 1));
    u8 *spec_out = (u8 *)malloc((size_t)(Fa + 1));
    float *background = (float *)malloc(sizeof(float) * (size_t)(N + 1));
    u8 *time_flags = (u8 *)malloc((size_t)(N + 1));
    u8 *freq_flags = (u8 *)malloc((size_t)(N + 1));

    orc_time_median(data, flags, T, Fa, spec_data, spec_flags);
    orc_get_background2d(spec_data, spec_flags, 1, Fa, p->background_iterations,
                         p->radii_spec, p->background_reject, p->freq_chunk_ends,
                         p->nchunk_ends, spec_bg);
    for (int64_t f = 0; f < Fa; f++) spec_data[f] -= spec_bg[f];
    orc_sum_threshold(spec_data, spec_flags, 1, Fa, 1, p->windows_freq, p->tf_freq,
                      p->nwin_freq, p->outlier_nsigma, p->freq_chunk_ends,
                      p->nchunk_ends, spec_out);
    for (int64_t t = 0; t < T; t++)
        for (int64_t f = 0; f < Fa; f++) flags[t * Fa + f] |= spec_out[f];

    orc_get_background2d(data, flags, T, Fa, p->background_iterations, p->radii_2d,
                         p->background_reject, p->freq_chunk_ends, p->nchunk_ends,
                         background);
    for (int64_t i = 0; i < N; i++) data[i] -= background[i];
    orc_sum_threshold(data, flags, T, Fa, 0, p->windows_time, p->tf_time,
                      p->nwin_time, p->outlier_nsigma, NULL, 0, time_flags);
    for (int64_t i = 0; i < N; i++) flags[i] |= time_flags[i];
    orc_sum_threshold(data, flags, T, Fa, 1, p->windows_freq, p->tf_freq,
                      p->nwin_freq, p->outlier_nsigma, p->freq_chunk_ends,
                      p->nchunk_ends, freq_flags);
    orc_combine_flags(spec_out, time_flags, freq_flags, T, Fa, p->time_extend,
                      p->time_extend_bits, flags);
    orc_unaverage_freq(flags, T, Fa, F, p->freq_extend, p->average_freq,
                       p->flag_all_time_frac, p->flag_all_freq_frac, out_flags);
    free(spec_data); free(spec_bg); free(spec_flags); free(spec_out);
    free(background); free(time_flags); free(freq_flags);
}

/* _get_flags_impl -- flagging.py:745-781.  Re-entrant: the Python wrapper
 * threads over plane ranges with a ThreadPool, the way the reference's dask
 * ThreadPool threads over baseline blocks (ctypes drops the GIL). */
void orc_get_flags_impl(const float *in_data, int is_complex, const u8 *in_flags,
                        int64_t ncp, int64_t T, int64_t F, const orc_params *p,
                        u8 *out_flags)
{
    int64_t Fa = (F + p->average_freq - 1) / p->average_freq;
    for (int64_t cp = 0; cp < ncp; cp++) {
        float *data = (float *)malloc(sizeof(float) * (size_t)(T * Fa + 1));
        u8 *flags = (u8 *)malloc((size_t)(T * Fa + 1));
        const float *pin = in_data + (size_t)cp * T * F * (is_complex ? 2 : 1);
        orc_average_freq(pin, is_complex, in_flags + (size_t)cp * T * F, 1, T, F,
                         p->average_freq, p->average_freq_bits, data, flags);
        u8 *o = out_flags + (size_t)cp * T * F;
        orc_get_baseline_flags(data, flags, T, Fa, F, p, o);
        for (int64_t i = 0; i < T * F; i++) {
            int nan = is_complex ? (isnan(pin[2 * i]) || isnan(pin[2 * i + 1]))
                                 : isnan(pin[i]);
            o[i] = (u8)(o[i] || nan);
        }
        free(data); free(flags);
    }
}

/* S13 sum_threshold_flagger major-iteration loop -- flagging.py:1181-1196 */
void orc_sum_threshold_flagger(const float *vis, int is_complex, const u8 *flags,
                               int64_t ncp, int64_t T, int64_t F,
                               const orc_params *p, int num_major_iterations,
                               u8 *out_flags)
{
    int64_t N = ncp * T * F;
    u8 *iter_flags = (u8 *)malloc((size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; i++) iter_flags[i] = flags[i] != 0;
    for (int it = 0; it < num_major_iterations; it++) {
        orc_get_flags_impl(vis, is_complex, iter_flags, ncp, T, F, p, out_flags);
        for (int64_t i = 0; i < N; i++) iter_flags[i] |= out_flags[i];
    }
    free(iter_flags);
}

/* constructor helper so ctypes callers need not mirror the struct layout */
orc_params *orc_params_new(double outlier_nsigma, int nwin_time,
                           const int64_t *windows_time, const double *tf_time,
                           int nwin_freq, const int64_t *windows_freq,
                           const double *tf_freq, double background_reject,
                           int background_iterations, const int64_t *radii_spec,
                           const int64_t *radii_2d, int64_t time_extend,
                           int time_extend_bits, int64_t freq_extend,
                           int nchunk_ends, const int64_t *freq_chunk_ends,
                           int64_t average_freq, int average_freq_bits,
                           double flag_all_time_frac, double flag_all_freq_frac)
{
    orc_params *p = (orc_params *)calloc(1, sizeof(orc_params));
    size_t nr = (size_t)(background_iterations + 1) * 2;
    int64_t *wt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nwin_time + 1));
    int64_t *wf = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nwin_freq + 1));
    double *tt = (double *)malloc(sizeof(double) * (size_t)(nwin_time + 1));
    double *tfq = (double *)malloc(sizeof(double) * (size_t)(nwin_freq + 1));
    int64_t *rs = (int64_t *)malloc(sizeof(int64_t) * nr);
    int64_t *r2 = (int64_t *)malloc(sizeof(int64_t) * nr);
    int64_t *ce = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nchunk_ends + 1));
    memcpy(wt, windows_time, sizeof(int64_t) * (size_t)nwin_time);
    memcpy(wf, windows_freq, sizeof(int64_t) * (size_t)nwin_freq);
    memcpy(tt, tf_time, sizeof(double) * (size_t)nwin_time);
    memcpy(tfq, tf_freq, sizeof(double) * (size_t)nwin_freq);
    memcpy(rs, radii_spec, sizeof(int64_t) * nr);
    memcpy(r2, radii_2d, sizeof(int64_t) * nr);
    memcpy(ce, freq_chunk_ends, sizeof(int64_t) * (size_t)nchunk_ends);
    p->outlier_nsigma = outlier_nsigma;
    p->nwin_time = nwin_time; p->windows_time = wt; p->tf_time = tt;
    p->nwin_freq = nwin_freq; p->windows_freq = wf; p->tf_freq = tfq;
    p->background_reject = background_reject;
    p->background_iterations = background_iterations;
    p->radii_spec = rs; p->radii_2d = r2;
    p->time_extend = time_extend; p->time_extend_bits = time_extend_bits;
    p->freq_extend = freq_extend;
    p->nchunk_ends = nchunk_ends; p->freq_chunk_ends = ce;
    p->average_freq = average_freq; p->average_freq_bits = average_freq_bits;
    p->flag_all_time_frac = flag_all_time_frac;
    p->flag_all_freq_frac = flag_all_freq_frac;
    return p;
}

void orc_params_free(orc_params *p)
{
    if (!p) return;
    free((void *)p->windows_time); free((void *)p->windows_freq);
    free((void *)p->tf_time); free((void *)p->tf_freq);
    free((void *)p->radii_spec); free((void *)p->radii_2d);
    free((void *)p->freq_chunk_ends);
    free(p);
}

/* ------------------------------------------------------------------------- */
/* K2 polarised_intensity / unpolarised_intensity -- tricolour/stokes.py:79-209 */
/* vis (row, chan, ncorr) complex64; terms: nterm x (c1, c2) ints,            */
/* coef: nterm x (a_re, a_im, s1, s2) doubles.  Arithmetic in complex128.     */
/* ------------------------------------------------------------------------- */
static inline double stokes_abs(const float *v, int64_t c1, int64_t c2,
                                const double *coef)
{
    double s1 = coef[2], s2 = coef[3];
    double re = s1 * (double)v[2 * c1] + s2 * (double)v[2 * c2];
    double im = s1 * (double)v[2 * c1 + 1] + s2 * (double)v[2 * c2 + 1];
    double vr = coef[0] * re - coef[1] * im;
    double vi = coef[0] * im + coef[1] * re;
    return hypot(vr, vi);
}

void orc_polarised_intensity(const float *vis, int64_t nrowchan, int64_t ncorr,
                             const int64_t *pol_idx, const double *pol_coef,
                             int npol, float *out /* (nrowchan,2) */)
{
    for (int64_t i = 0; i < nrowchan; i++) {
        const float *v = vis + (size_t)i * ncorr * 2;
        double pol = 0.0;
        for (int k = 0; k < npol; k++) {
            double a = stokes_abs(v, pol_idx[2 * k], pol_idx[2 * k + 1], pol_coef + 4 * k);
            pol += a * a;
        }
        out[2 * i] = (float)sqrt(pol);
        out[2 * i + 1] = 0.0f;
    }
}

void orc_unpolarised_intensity(const float *vis, int64_t nrowchan, int64_t ncorr,
                               const int64_t *unpol_idx, const double *unpol_coef,
                               int nunpol, const int64_t *pol_idx,
                               const double *pol_coef, int npol, float *out)
{
    for (int64_t i = 0; i < nrowchan; i++) {
        const float *v = vis + (size_t)i * ncorr * 2;
        double pol = 0.0, unpol = 0.0;
        for (int k = 0; k < npol; k++) {
            double a = stokes_abs(v, pol_idx[2 * k], pol_idx[2 * k + 1], pol_coef + 4 * k);
            pol += a * a;
        }
        for (int k = 0; k < nunpol; k++)
            unpol += stokes_abs(v, unpol_idx[2 * k], unpol_idx[2 * k + 1], unpol_coef + 4 * k);
        out[2 * i] = (float)(unpol - sqrt(pol));
        out[2 * i + 1] = 0.0f;
    }
}

/* ------------------------------------------------------------------------- */
/* P1 _numba_pack_data -- tricolour/packing.py:243-278                        */
/* P2 _unpack_data / _numpy_unpack_transpose -- packing.py:369-415            */
/* elem = bytes per element (8 for complex64 vis, 1 for flags)                */
/* ------------------------------------------------------------------------- */
void orc_pack(const int64_t *time_inv, const int32_t *ubl /* (nbl,3) */, int64_t nbl,
              const int32_t *ant1, const int32_t *ant2, int64_t nrow,
              const void *data, int64_t nchan, int64_t ncorr, int64_t ntime,
              int elem, void *window /* (nbl_total,ncorr,ntime,nchan) */)
{
    const char *src = (const char *)data;
    char *dst = (char *)window;
    for (int64_t b = 0; b < nbl; b++) {
        int64_t bl = ubl[3 * b];
        int32_t a1 = ubl[3 * b + 1], a2 = ubl[3 * b + 2];
        for (int64_t r = 0; r < nrow; r++) {
            if (ant1[r] != a1 || ant2[r] != a2) continue;
            int64_t t = time_inv[r];
            for (int64_t f = 0; f < nchan; f++)
                for (int64_t c = 0; c < ncorr; c++)
                    memcpy(dst + (size_t)(((bl * ncorr + c) * ntime + t) * nchan + f) * elem,
                           src + (size_t)((r * nchan + f) * ncorr + c) * elem, (size_t)elem);
        }
    }
}

void orc_unpack(const int64_t *time_inv, const int32_t *ubl, int64_t nbl,
                int64_t bl_min, const int32_t *ant1, const int32_t *ant2,
                int64_t nrow, const void *window, int64_t nchan, int64_t ncorr,
                int64_t ntime, int elem, void *data /* (nrow,nchan,ncorr), pre-zeroed */)
{
    const char *src = (const char *)window;
    char *dst = (char *)data;
    for (int64_t b = 0; b < nbl; b++) {
        int64_t bl = ubl[3 * b] - bl_min;
        int32_t a1 = ubl[3 * b + 1], a2 = ubl[3 * b + 2];
        for (int64_t r = 0; r < nrow; r++) {
            if (ant1[r] != a1 || ant2[r] != a2) continue;
            int64_t t = time_inv[r];
            for (int64_t f = 0; f < nchan; f++)
                for (int64_t c = 0; c < ncorr; c++)
                    memcpy(dst + (size_t)((r * nchan + f) * ncorr + c) * elem,
                           src + (size_t)(((bl * ncorr + c) * ntime + t) * nchan + f) * elem,
                           (size_t)elem);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* W1 _window_stats counting part -- tricolour/window_statistics.py:12-66     */
/* flag window (nbl,ncorr,T,F) uint8.  Outputs:                               */
/*   bl_counts[nbl], chan_counts[F] (uint64); the per-antenna / field / scan  */
/*   / channel-bin numbers are sums of these and are formed by the caller.    */
/* ------------------------------------------------------------------------- */
void orc_window_counts(const u8 *flags, int64_t nbl, int64_t ncorr, int64_t T,
                       int64_t F, uint64_t *bl_counts, uint64_t *chan_counts)
{
    for (int64_t f = 0; f < F; f++) chan_counts[f] = 0;
    for (int64_t b = 0; b < nbl; b++) {
        uint64_t cnt = 0;
        const u8 *p = flags + (size_t)b * ncorr * T * F;
        for (int64_t i = 0; i < ncorr * T; i++)
            for (int64_t f = 0; f < F; f++) {
                uint64_t v = p[i * F + f]; /* np.sum of the values */
                cnt += v;
                chan_counts[f] += v;
            }
        bl_counts[b] = cnt;
    }
}
