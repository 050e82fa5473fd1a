// k_uvcontsub.cuh -- uvcontsub_flagger (tricolour/flagging.py:989-1073).
//
// Per plane and major cycle the reference (pure numpy) does: mask flagged
// samples, nanmean over time (complex64), FFT along frequency, zero all bins
// >= taylor_degrees, inverse FFT, |vis - smooth|, an unscaled MAD via two
// nanmedians, and flags residual > sigma * mad.  Only the first
// `taylor_degrees` bins survive, so the FFT pair collapses to a K-term direct
// DFT evaluated in float64 (numpy's pocketfft runs in float32; parity on this
// function is therefore "equal up to threshold ties").
#pragma once
#include "tc_common.cuh"
#include "k_select.cuh"

// twiddle table tw[j] = exp(+2 pi i j / F)
__global__ void __launch_bounds__(256)
k_uv_twiddle(double2 *__restrict__ tw, int F)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= F) return;
    double s, co;
    double x = 2.0 * (double)j / (double)F;
#ifdef TC_EMU
    s = sin(M_PI * x); co = cos(M_PI * x);
#else
    sincospi(x, &s, &co);
#endif
    tw[j] = make_double2(co, s);
}

// nanmean over time (flagging.py:1037-1044): sequential float32 complex sum
// per channel, divided the way numpy divides complex64 by an int64 count
// (promote to complex128, multiply by the reciprocal, round back).
// Also counts the unflagged samples of every plane.
__global__ void __launch_bounds__(256)
k_uv_mean(const float2 *__restrict__ vis, const u8 *__restrict__ flags, int T, int F,
          float2 *__restrict__ avg, int *__restrict__ unflagged, int64_t cp0)
{
    __shared__ int s_cnt;
    int64_t cp = cp0 + blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int nun = 0;
    if (f < F) {
        float sr = 0.f, si = 0.f;
        int cnt = 0;
        // eight dumps per trip so that the loads of a trip are all in flight together;
        // the sums stay sequential in t
        const float2 *pv = vis + (cp * T) * (int64_t)F + f;
        const u8 *pf = flags + (cp * T) * (int64_t)F + f;
        for (int t0 = 0; t0 < T; t0 += 8) {
            float2 v[8];
            u8 fl[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const bool in = t0 + k < T;
                fl[k] = in ? pf[(int64_t)k * F] : (u8)1;
                v[k] = in ? pv[(int64_t)k * F] : make_float2(0.f, 0.f);
            }
            pv += 8 * (int64_t)F; pf += 8 * (int64_t)F;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (fl[k]) continue;
                nun++;
                if (v[k].x != v[k].x || v[k].y != v[k].y) continue;
                sr = __fadd_rn(sr, v[k].x);
                si = __fadd_rn(si, v[k].y);
                cnt++;
            }
        }
        float2 o = make_float2(0.f, 0.f);
        if (cnt > 0) {
            double scl = __ddiv_rn(1.0, (double)cnt);
            o.x = (float)__dmul_rn((double)sr, scl);
            o.y = (float)__dmul_rn((double)si, scl);
            if (o.x != o.x || o.y != o.y) o = make_float2(0.f, 0.f);  // avgvis[isnan] = 0
        }
        avg[cp * F + f] = o;
    }
    atomicAdd(&s_cnt, nun);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&unflagged[cp], s_cnt);
}

#define TC_UV_MAXK 64
// (k * f) mod F; 32-bit arithmetic (inline) whenever the product fits, which it
// does for every realistic K (< 64) and F (< 2^26)
__device__ __forceinline__ int uv_tw_index(int k, int f, int F)
{
    if ((int64_t)TC_UV_MAXK * F < ((int64_t)1 << 32)) return (int)(((unsigned)k * (unsigned)f) % (unsigned)F);
    return (int)(((int64_t)k * f) % F);
}

// smooth = ifft(first K bins of fft(avg)); one block per plane
__global__ void __launch_bounds__(256)
k_uv_smooth(const float2 *__restrict__ avg, const double2 *__restrict__ tw, int F, int K,
            float2 *__restrict__ smooth)
{
    __shared__ double2 X[TC_UV_MAXK];
    __shared__ double2 red[8];
    int64_t cp = blockIdx.x;     // gridDim.x goes up to 2^31 - 1
    const float2 *a = avg + cp * F;
    int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = 0; k < K; k++) {
        double re = 0.0, im = 0.0;
        for (int f = tid; f < F; f += blockDim.x) {
            double2 w = tw[uv_tw_index(k, f, F)];
            double ar = (double)a[f].x, ai = (double)a[f].y;
            // a * conj(w)
            re += ar * w.x + ai * w.y;
            im += ai * w.x - ar * w.y;
        }
        for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(TC_FULL_MASK, re, o);
            im += __shfl_xor_sync(TC_FULL_MASK, im, o);
        }
        if (lane == 0) red[wid] = make_double2(re, im);
        __syncthreads();
        if (tid == 0) {
            double sr = 0, si = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) { sr += red[w].x; si += red[w].y; }
            X[k] = make_double2(sr, si);
        }
        __syncthreads();
    }
    double inv = 1.0 / (double)F;
    for (int f = tid; f < F; f += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int k = 0; k < K; k++) {
            double2 w = tw[uv_tw_index(k, f, F)];
            re += X[k].x * w.x - X[k].y * w.y;
            im += X[k].x * w.y + X[k].y * w.x;
        }
        smooth[cp * F + f] = make_float2((float)(re * inv), (float)(im * inv));
    }
}

// absresidual = np.abs(vis - smooth) in float32.  numpy's complex64 absolute
// on FMA hosts is max * sqrtf(fmaf(d, d, 1)), d = min / max.
__device__ __forceinline__ float uv_abs_c64(float vx, float vy, float sx, float sy)
{
    float re = fabsf(__fadd_rn(vx, -sx)), im = fabsf(__fadd_rn(vy, -sy));
    if (re != re || im != im) return (isinf(re) || isinf(im)) ? INFINITY : NAN;
    float mx = re > im ? re : im, mn = re > im ? im : re;
    if (mx == 0.0f) return 0.0f;
    if (isinf(mx)) return INFINITY;
    float d = __fdiv_rn(mn, mx);
    return mx * sqrtf(fmaf(d, d, 1.0f));
}

// grid: (ceil(F / 2 / 256), T, cp); two channels per thread, no index divisions
__global__ void __launch_bounds__(256)
k_uv_absres(const float2 *__restrict__ vis, const float2 *__restrict__ smooth, int T, int F,
            float *__restrict__ out, int t0, int p0)
{
    const int f = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (f >= F) return;
    const int64_t cp = (int64_t)p0 + blockIdx.z;
    const int64_t row = (cp * T + t0 + blockIdx.y) * (int64_t)F;
    const float2 *v = vis + row + f;
    const float2 *s = smooth + cp * F + f;
    if (f + 1 < F && (((uintptr_t)v | (uintptr_t)s) & 15) == 0 && (((uintptr_t)(out + row + f)) & 7) == 0) {
        const float4 vv = *reinterpret_cast<const float4 *>(v), ss = *reinterpret_cast<const float4 *>(s);
        *reinterpret_cast<float2 *>(out + row + f) =
            make_float2(uv_abs_c64(vv.x, vv.y, ss.x, ss.y), uv_abs_c64(vv.z, vv.w, ss.z, ss.w));
    } else {
        out[row + f] = uv_abs_c64(v[0].x, v[0].y, s[0].x, s[0].y);
        if (f + 1 < F) out[row + f + 1] = uv_abs_c64(v[1].x, v[1].y, s[1].x, s[1].y);
    }
}
