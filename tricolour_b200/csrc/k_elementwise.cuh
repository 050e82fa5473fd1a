// k_elementwise.cuh -- HBM-bound companions of the SumThreshold core:
// flag_nans_and_zeros, flag_autos, apply_static_mask, OR, Stokes intensities,
// tiled transposes.  One pass over the data, 16-byte accesses where the
// pointers allow it.
#pragma once
#include "tc_common.cuh"

// ----------------------------------------------------------------------------
// F1 flag_nans_and_zeros (tricolour/flagging.py:29-62)
// out = (vis == 0) | isnan(vis) | (flag != 0)
// ----------------------------------------------------------------------------
__device__ __forceinline__ u8 nz_flag(float re, float im, u8 f)
{
    bool z = (re == 0.0f) && (im == 0.0f);
    bool n = (re != re) || (im != im);
    return (u8)((z || n || f != 0) ? 1 : 0);
}

// 4 samples per thread: 2 x 16 B of visibilities, 4 B of flags in, 4 B out
__global__ void __launch_bounds__(256)
k_flag_nans_zeros_v4(const float4 *__restrict__ vis, const uint32_t *__restrict__ flags,
                     uint32_t *__restrict__ out, int64_t n4)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = vis[2 * i], b = vis[2 * i + 1];
    uint32_t f = flags[i];
    uint32_t o = (uint32_t)nz_flag(a.x, a.y, (u8)(f & 0xff)) |
                 ((uint32_t)nz_flag(a.z, a.w, (u8)((f >> 8) & 0xff)) << 8) |
                 ((uint32_t)nz_flag(b.x, b.y, (u8)((f >> 16) & 0xff)) << 16) |
                 ((uint32_t)nz_flag(b.z, b.w, (u8)((f >> 24) & 0xff)) << 24);
    out[i] = o;
}

__global__ void __launch_bounds__(256)
k_flag_nans_zeros(const float2 *__restrict__ vis, const u8 *__restrict__ flags,
                  u8 *__restrict__ out, int64_t n, int64_t start)
{
    int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v = vis[i];
    out[i] = nz_flag(v.x, v.y, flags[i]);
}

static int launch_flag_nans_zeros(tc_context *c, const void *vis, const u8 *flags, u8 *out, int64_t n)
{
    bool aligned = (((uintptr_t)vis & 15) == 0) && (((uintptr_t)flags & 3) == 0) && (((uintptr_t)out & 3) == 0);
    int64_t n4 = aligned ? n / 4 : 0;
    tc_prof_begin(c, TCP_ELEMENTWISE);
    if (n4 > 0) {
        TC_LAUNCH_NOSYNC(k_flag_nans_zeros_v4, tc_blocks_for(n4, 256), 256, 0, c->stream,
                         (const float4 *)vis, (const uint32_t *)flags, (uint32_t *)out, n4);
        c->launches++;
    }
    int64_t rem = n - n4 * 4;
    if (rem > 0) {
        TC_LAUNCH_NOSYNC(k_flag_nans_zeros, tc_blocks_for(rem, 256), 256, 0, c->stream,
                         (const float2 *)vis, flags, out, n, n4 * 4);
        c->launches++;
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    return TC_OK;
}

// ----------------------------------------------------------------------------
// F2 flag_autos (flagging.py:65-95) and F3 apply_static_mask (98-172) share one
// kernel: per baseline a selector, per channel a mask byte.
//   mode 0: out = flag | (sel[bl] & mask[f])        ("or",  flagging.py:164)
//   mode 1: out = sel[bl] ? mask[f] : flag           ("override", 166)
//   mode 2: out = sel[bl] ? 1 : flag                 (flag_autos, 93)
// Rows are (bl, row, chan) with `rows_per_bl` rows of `nchan` bytes each.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_apply_mask(const u8 *__restrict__ flags, const u8 *__restrict__ bl_sel,
             const u8 *__restrict__ chan_mask, int mode, int64_t rows_per_bl,
             int64_t nchan, int64_t total, u8 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t row = i / nchan;
    int64_t f = i - row * nchan;
    int64_t bl = row / rows_per_bl;
    u8 fl = flags[i];
    u8 sel = bl_sel[bl];
    u8 o;
    if (mode == 0) o = (u8)(fl | ((sel && chan_mask[f]) ? 1 : 0));
    else if (mode == 1) o = sel ? (u8)(chan_mask[f] ? 1 : 0) : fl;
    else o = sel ? (u8)1 : fl;
    out[i] = o;
}

// 16 channels per thread; requires nchan % 16 == 0 and 16-byte aligned pointers
__global__ void __launch_bounds__(256)
k_apply_mask_v16(const uint4 *__restrict__ flags, const u8 *__restrict__ bl_sel,
                 const uint4 *__restrict__ chan_mask01, int mode, int64_t rows_per_bl,
                 int64_t nchan16, int64_t total16, uint4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total16) return;
    int64_t row = i / nchan16;
    int64_t f16 = i - row * nchan16;
    int64_t bl = row / rows_per_bl;
    uint4 fl = flags[i];
    u8 sel = bl_sel[bl];
    if (sel) {
        if (mode == 0) {
            uint4 m = chan_mask01[f16];
            fl.x |= m.x; fl.y |= m.y; fl.z |= m.z; fl.w |= m.w;
        } else if (mode == 1) {
            fl = chan_mask01[f16];
        } else {
            fl.x = fl.y = fl.z = fl.w = 0x01010101u;
        }
    }
    out[i] = fl;
}

// chan_mask must already be normalised to 0/1 bytes (done by the API layer)
static int launch_apply_mask(tc_context *c, const u8 *flags, const u8 *bl_sel_dev,
                             const u8 *chan_mask_dev, int mode, int64_t nbl,
                             int64_t rows_per_bl, int64_t nchan, u8 *out)
{
    int64_t total = nbl * rows_per_bl * nchan;
    if (total == 0) return TC_OK;
    bool vec = (nchan % 16 == 0) && (((uintptr_t)flags & 15) == 0) && (((uintptr_t)out & 15) == 0) &&
               (((uintptr_t)chan_mask_dev & 15) == 0);
    tc_prof_begin(c, TCP_ELEMENTWISE);
    if (vec) {
        int64_t t16 = total / 16;
        TC_LAUNCH_NOSYNC(k_apply_mask_v16, tc_blocks_for(t16, 256), 256, 0, c->stream,
                         (const uint4 *)flags, bl_sel_dev, (const uint4 *)chan_mask_dev, mode,
                         rows_per_bl, nchan / 16, t16, (uint4 *)out);
    } else {
        TC_LAUNCH_NOSYNC(k_apply_mask, tc_blocks_for(total, 256), 256, 0, c->stream, flags,
                         bl_sel_dev, chan_mask_dev, mode, rows_per_bl, nchan, total, out);
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}

// ----------------------------------------------------------------------------
// out = a | b  (strat_executor.py:43,54,59,76); bytes stay 0/1 if inputs are
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_or_v16(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ out, int64_t n16)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    uint4 x = a[i], y = b[i];
    x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w;
    out[i] = x;
}
__global__ void __launch_bounds__(256)
k_or(const u8 *__restrict__ a, const u8 *__restrict__ b, u8 *__restrict__ out, int64_t n, int64_t start)
{
    int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (u8)(a[i] | b[i]);
}
static int launch_or(tc_context *c, const u8 *a, const u8 *b, u8 *out, int64_t n)
{
    bool aligned = ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;
    int64_t n16 = aligned ? n / 16 : 0;
    tc_prof_begin(c, TCP_ELEMENTWISE);
    if (n16 > 0) {
        TC_LAUNCH_NOSYNC(k_or_v16, tc_blocks_for(n16, 256), 256, 0, c->stream, (const uint4 *)a,
                         (const uint4 *)b, (uint4 *)out, n16);
        c->launches++;
    }
    if (n - n16 * 16 > 0) {
        TC_LAUNCH_NOSYNC(k_or, tc_blocks_for(n - n16 * 16, 256), 256, 0, c->stream, a, b, out, n, n16 * 16);
        c->launches++;
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    return TC_OK;
}

// ----------------------------------------------------------------------------
// K2 polarised_intensity / unpolarised_intensity (tricolour/stokes.py:79-209)
// numba evaluates a*(s1*v1 + s2*v2) in complex128 and |.| with glibc's double
// hypot; the result goes back to complex64 with a zero imaginary part.
// ----------------------------------------------------------------------------
#define TC_MAX_STOKES 8
struct StokesTerms {
    int n;
    int c1[TC_MAX_STOKES], c2[TC_MAX_STOKES];
    double ar[TC_MAX_STOKES], ai[TC_MAX_STOKES], s1[TC_MAX_STOKES], s2[TC_MAX_STOKES];
};

// glibc >= 2.35 hypot (sysdeps/ieee754/dbl-64/e_hypot.c, the build without
// __FP_FAST_FMA that x86-64 ships: there is no multiarch variant), operation for
// operation: Borges' "MyHypot3" correction of sqrt(ax^2 + ay^2) with the scaling of
// huge / tiny operands.  Bit-identical to the libm call numba makes for
// np.abs(complex128) (checked against glibc 2.39 on 2e7 operand pairs, including
// random bit patterns), so the Stokes intensities match the reference exactly.
__device__ __forceinline__ double tc_hypot_kernel(double ax, double ay)
{
    double t1, t2;
    double h = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
    if (h <= __dmul_rn(2.0, ay)) {
        const double delta = __dadd_rn(h, -ay);
        t1 = __dmul_rn(ax, __dadd_rn(__dmul_rn(2.0, delta), -ax));
        t2 = __dmul_rn(__dadd_rn(delta, -__dmul_rn(2.0, __dadd_rn(ax, -ay))), delta);
    } else {
        const double delta = __dadd_rn(h, -ax);
        t1 = __dmul_rn(__dmul_rn(2.0, delta), __dadd_rn(ax, -__dmul_rn(2.0, ay)));
        t2 = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(4.0, delta), -ay), ay), __dmul_rn(delta, delta));
    }
    return __dadd_rn(h, -__ddiv_rn(__dadd_rn(t1, t2), __dmul_rn(2.0, h)));
}

__device__ __forceinline__ double tc_hypot(double x, double y)
{
    const double SCALE = 0x1p-600, LARGE_VAL = 0x1p+511, TINY_VAL = 0x1p-459, EPS = 0x1p-54;
    if (isinf(x) || isinf(y)) return INFINITY;
    if (x != x || y != y) return NAN;
    x = fabs(x); y = fabs(y);
    const double ax = x < y ? y : x, ay = x < y ? x : y;
    if (ax > LARGE_VAL) {
        if (ay <= __dmul_rn(ax, EPS)) return __dadd_rn(ax, ay);
        return __ddiv_rn(tc_hypot_kernel(__dmul_rn(ax, SCALE), __dmul_rn(ay, SCALE)), SCALE);
    }
    if (ay < TINY_VAL) {
        if (ax >= __ddiv_rn(ay, EPS)) return __dadd_rn(ax, ay);
        return __dmul_rn(tc_hypot_kernel(__ddiv_rn(ax, SCALE), __ddiv_rn(ay, SCALE)), SCALE);
    }
    if (ax >= __ddiv_rn(ay, EPS)) return __dadd_rn(ax, ay);
    return tc_hypot_kernel(ax, ay);
}

__device__ __forceinline__ double stokes_abs(const float2 *v, const StokesTerms &t, int k)
{
    float2 a = v[t.c1[k]], b = v[t.c2[k]];
    // int * complex64 promotes to complex128 before anything is rounded
    double re = __dadd_rn(__dmul_rn(t.s1[k], (double)a.x), __dmul_rn(t.s2[k], (double)b.x));
    double im = __dadd_rn(__dmul_rn(t.s1[k], (double)a.y), __dmul_rn(t.s2[k], (double)b.y));
    double vr = __dadd_rn(__dmul_rn(t.ar[k], re), -__dmul_rn(t.ai[k], im));
    double vi = __dadd_rn(__dmul_rn(t.ar[k], im), __dmul_rn(t.ai[k], re));
    return tc_hypot(vr, vi);
}

// sqrt(sum |pol term|^2), or (sum |unpol term|) - that, in float64 (stokes.py:132-153, 196-208)
__device__ __forceinline__ double stokes_intensity_exact(const float2 *v, const StokesTerms &pol, const StokesTerms &unpol,
                                                         int with_unpol)
{
    double p = 0.0;
    for (int k = 0; k < pol.n; k++) {
        double a = stokes_abs(v, pol, k);
        p = __dadd_rn(p, __dmul_rn(a, a));
    }
    double r = __dsqrt_rn(p);
    if (with_unpol) {
        double u = 0.0;
        for (int k = 0; k < unpol.n; k++) u = __dadd_rn(u, stokes_abs(v, unpol, k));
        r = __dadd_rn(u, -r);
    }
    return r;
}

// float32 bits of a FINITE value -> the double x * 2^-896 (exact; same trick as the box filter's
// accumulators, k_filter2.cuh): one IMAD.WIDE and a mask on the integer pipes instead of a conversion
// on the 16-lane XU pipe, the 2^896 rides on the coefficient that multiplies it
__device__ __forceinline__ double ew_spread(float x)
{
#ifndef TC_EMU
    int hi, lo;
    asm("{\n\t.reg .s64 w;\n\tmul.wide.s32 w, %2, 536870912;\n\tmov.b64 {%1, %0}, w;\n\t}"
        : "=r"(hi), "=r"(lo) : "r"(__float_as_int(x)));
    return __hiloint2double(hi & (int)0x8fffffff, lo);
#else
    return (double)x * 0x1p-896;
#endif
}

__device__ __forceinline__ void stokes_term_fast(const float2 *v, const StokesTerms &t, int k, double *vr, double *vi)
{
    const float2 a = v[t.c1[k]], b = v[t.c2[k]];
    const double s1 = t.s1[k] * 0x1p896, s2 = t.s2[k] * 0x1p896;
    const double re = s1 * ew_spread(a.x) + s2 * ew_spread(b.x);
    const double im = s1 * ew_spread(a.y) + s2 * ew_spread(b.y);
    *vr = t.ar[k] * re - t.ai[k] * im;
    *vi = t.ar[k] * im + t.ai[k] * re;
}

// The float32 the reference stores.  The exact sequence above costs three double hypots (a square
// root and a division each) and a square root per sample and keeps the kernel on the FP64 pipe at
// 28 % of the HBM roofline.  |a|^2 of a correctly-replayed hypot differs from re^2 + im^2 by a few
// ulps of a double, so the sum of squares -- and with it the float64 result -- is known to within
// 2^-49 relative without any hypot; only when the two ends of that interval round to different
// float32 values (probability ~2^-27 per sample) is the exact sequence replayed.  Same bits out.
__device__ __forceinline__ float stokes_intensity(const float2 *v, const StokesTerms &pol, const StokesTerms &unpol,
                                                  int with_unpol, int ncorr)
{
    // NaN / Inf anywhere: the exact sequence (x * 0 is NaN exactly for those)
    float chk = 0.f;
    for (int cidx = 0; cidx < ncorr && cidx < 8; cidx++) chk = fmaf(v[cidx].x, 0.f, fmaf(v[cidx].y, 0.f, chk));
    if (chk != 0.f) return (float)stokes_intensity_exact(v, pol, unpol, with_unpol);
    double p = 0.0;
    for (int k = 0; k < pol.n; k++) {
        double vr, vi;
        stokes_term_fast(v, pol, k, &vr, &vi);
        p += __fma_rn(vr, vr, vi * vi);
    }
    double r = __dsqrt_rn(p), scale = r;
    if (with_unpol) {
        double u = 0.0;
        for (int k = 0; k < unpol.n; k++) {
            double vr, vi;
            stokes_term_fast(v, unpol, k, &vr, &vi);
            u += __dsqrt_rn(__fma_rn(vr, vr, vi * vi));
        }
        scale = u > r ? u : r;
        r = u - r;
    }
    if (scale == 0.0) return 0.0f;          // every term exactly zero: so is the reference's result
    const double e = scale * 0x1p-46;
    const float lo = (float)(r - e), hi = (float)(r + e);
    // !(lo == hi) also catches NaN (non-finite input) and sends it down the exact path
    if (!(lo == hi) || !(scale > 0x1p-400 && scale < 0x1p400))
        return (float)stokes_intensity_exact(v, pol, unpol, with_unpol);
    return lo;
}

__global__ void __launch_bounds__(256)
k_stokes(const float2 *__restrict__ vis, int64_t n, int ncorr, StokesTerms pol,
         StokesTerms unpol, int with_unpol, float2 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v[8];
    const float2 *src = vis + i * ncorr;
    if (ncorr == 4 && (((uintptr_t)src) & 15) == 0) {
        float4 a = ((const float4 *)src)[0], b = ((const float4 *)src)[1];
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
        v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
    } else {
        for (int c = 0; c < ncorr && c < 8; c++) v[c] = src[c];
    }
    out[i] = make_float2(stokes_intensity(v, pol, unpol, with_unpol, ncorr), 0.0f);
}

// ----------------------------------------------------------------------------
// batched 32x32 tiled transposes: in (nplanes, R, C) -> out (nplanes, C, R)
// ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_transpose(const T *__restrict__ in, T *__restrict__ out, int R, int C)
{
    __shared__ T tile[32][33];
    int64_t plane = (int64_t)blockIdx.z * R * C;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        int r = r0 + k, cc = c0 + tx;
        if (r < R && cc < C) tile[k][tx] = in[plane + (int64_t)r * C + cc];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        int cc = c0 + k, r = r0 + tx;
        if (r < R && cc < C) out[plane + (int64_t)cc * R + r] = tile[tx][k];
    }
}

// byte planes, R and C multiples of 4: 64 x 64 tiles moved as 32-bit words on both sides (the
// 32 x 32 byte tiles of the generic kernel touch one 32-byte sector per warp row: 17 % of the HBM
// roofline on the flag planes)
__global__ void __launch_bounds__(256)
k_transpose_u8x4(const u8 *__restrict__ in, u8 *__restrict__ out, int R, int C)
{
    __shared__ __align__(16) u8 tile[64][68];
    const int64_t plane = (int64_t)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = r0 + ty + 16 * k, cc = c0 + 4 * tx;
        unsigned w = 0u;
        if (r < R && cc < C) w = *reinterpret_cast<const unsigned *>(in + plane + (int64_t)r * C + cc);
        *reinterpret_cast<unsigned *>(&tile[ty + 16 * k][4 * tx]) = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int cc = c0 + ty + 16 * k, r = r0 + 4 * tx;
        if (cc < C && r < R) {
            const int lc = ty + 16 * k;
            const unsigned w = (unsigned)tile[4 * tx][lc] | ((unsigned)tile[4 * tx + 1][lc] << 8) |
                               ((unsigned)tile[4 * tx + 2][lc] << 16) | ((unsigned)tile[4 * tx + 3][lc] << 24);
            *reinterpret_cast<unsigned *>(out + plane + (int64_t)cc * R + r) = w;
        }
    }
}

// 4-byte planes, R and C multiples of 4: 64 x 64 tiles moved as 16-byte vectors on both sides (four times the
// bytes in flight per block of the 32 x 32 kernel, which ran at 2.9 TB/s read + write on the float planes)
__global__ void __launch_bounds__(256)
k_transpose_w32x4(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int R, int C)
{
    __shared__ uint32_t tile[64][65];
    const int64_t plane = (int64_t)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = r0 + ty + 16 * k, cc = c0 + 4 * tx;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (r < R && cc < C) v[k] = *reinterpret_cast<const uint4 *>(in + plane + (int64_t)r * C + cc);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t *row = &tile[ty + 16 * k][4 * tx];
        row[0] = v[k].x; row[1] = v[k].y; row[2] = v[k].z; row[3] = v[k].w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int cc = c0 + ty + 16 * k, r = r0 + 4 * tx;
        if (cc < C && r < R) {
            const int lc = ty + 16 * k;
            const uint4 w = make_uint4(tile[4 * tx][lc], tile[4 * tx + 1][lc], tile[4 * tx + 2][lc], tile[4 * tx + 3][lc]);
            *reinterpret_cast<uint4 *>(out + plane + (int64_t)cc * R + r) = w;
        }
    }
}

template <typename T>
static int launch_transpose(tc_context *c, const T *in, T *out, int64_t nplanes, int R, int C)
{
    if (nplanes == 0 || R == 0 || C == 0) return TC_OK;
    if (sizeof(T) == 4 && (R & 3) == 0 && (C & 3) == 0 && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0) &&
        (R + 63) / 64 <= 65535 && !TC_ENV_FLAG("TC_TRANSPOSE_GENERIC")) {
        for (int64_t p0 = 0; p0 < nplanes; p0 += 65535) {
            int64_t np = nplanes - p0 < 65535 ? nplanes - p0 : 65535;
            dim3 grid((C + 63) / 64, (R + 63) / 64, (unsigned)np);
            tc_prof_begin(c, TCP_TRANSPOSE);
            TC_LAUNCH(k_transpose_w32x4, grid, 256, 0, c->stream, reinterpret_cast<const uint32_t *>(in) + p0 * R * C,
                      reinterpret_cast<uint32_t *>(out) + p0 * R * C, R, C);
            tc_prof_end(c);
            c->launches++;
        }
        TC_KERNEL_CHECK();
        return TC_OK;
    }
    if (sizeof(T) == 1 && (R & 3) == 0 && (C & 3) == 0 && ((((uintptr_t)in | (uintptr_t)out) & 3) == 0) &&
        (R + 63) / 64 <= 65535 && !TC_ENV_FLAG("TC_TRANSPOSE_GENERIC")) {
        for (int64_t p0 = 0; p0 < nplanes; p0 += 65535) {
            int64_t np = nplanes - p0 < 65535 ? nplanes - p0 : 65535;
            dim3 grid((C + 63) / 64, (R + 63) / 64, (unsigned)np);
            tc_prof_begin(c, TCP_TRANSPOSE);
            TC_LAUNCH(k_transpose_u8x4, grid, 256, 0, c->stream, reinterpret_cast<const u8 *>(in) + p0 * R * C,
                      reinterpret_cast<u8 *>(out) + p0 * R * C, R, C);
            tc_prof_end(c);
            c->launches++;
        }
        TC_KERNEL_CHECK();
        return TC_OK;
    }
    TC_REQUIRE((R + 31) / 32 <= 65535, "transpose: more than %d rows per plane are not supported", 65535 * 32);
    // gridDim.z is limited to 65535 planes per launch
    for (int64_t p0 = 0; p0 < nplanes; p0 += 65535) {
        int64_t np = nplanes - p0 < 65535 ? nplanes - p0 : 65535;
        dim3 grid((C + 31) / 32, (R + 31) / 32, (unsigned)np);
        tc_prof_begin(c, TCP_TRANSPOSE);
        TC_LAUNCH(k_transpose<T>, grid, dim3(32, 8, 1), 0, c->stream, in + p0 * R * C,
                  out + p0 * R * C, R, C);
        tc_prof_end(c);
        c->launches++;
    }
    TC_KERNEL_CHECK();
    return TC_OK;
}
