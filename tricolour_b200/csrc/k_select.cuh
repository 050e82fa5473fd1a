// k_select.cuh -- exact medians (reference: numba np.median behind
// _time_median flagging.py:226-264, _median_abs 267-279, _median_abs_axis0
// 282-304; numba/np/arraymath.py:1621-1635 for the even-count rule).
//
// np.median is an exact order statistic; with an even count numba returns
// (double)(float)(a + b) / 2 of the two middle values.  Both kernels below do a
// most-significant-bit-first radix select on an order-preserving integer key
// of the float, so the answer does not depend on visit order.
#pragma once
#include "tc_common.cuh"

__device__ __forceinline__ uint32_t f2key(float x)
{
    uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

__device__ __forceinline__ int warp_sum_i(int v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TC_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_max_u(uint32_t v)
{
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t t = __shfl_xor_sync(TC_FULL_MASK, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// numba _median_inner on the two middle order statistics
__device__ __forceinline__ double median_from_pair(float lower, float upper, int n)
{
    if (n & 1) return (double)upper;
    float s = __fadd_rn(lower, upper);
    return (double)s / 2.0;
}

// ----------------------------------------------------------------------------
// Warp-per-line median.  A line is `n` samples at data[base + i*stride]; a
// sample takes part when neither flag array marks it.  Keys are held in
// registers (VPL per lane, n <= 32*VPL).
// ----------------------------------------------------------------------------
enum { LM_TIME_MEDIAN = 0, LM_ST_THRESHOLD = 1 };

struct LineMedianArgs {
    const float *data;
    const u8 *flags;   // may be null
    const u8 *flags2;  // may be null (OR-ed with flags)
    int64_t nlines;
    // line l: outer = l / ninner, inner = l % ninner,
    //   base = outer*outer_stride + inner*inner_stride (+ seg offset below)
    int64_t ninner, outer_stride, inner_stride;
    // optional segmentation of every line into chunks [seg_ends[k], seg_ends[k+1])
    const int64_t *seg_ends;  // device, nseg+1 entries; null -> one segment [0, n)
    int nseg;
    int n;                    // line length when seg_ends == null
    int64_t elem_stride;      // distance between consecutive samples of a line
    int mode;
    int use_abs;
    double thr_scale;   // LM_ST_THRESHOLD: outlier_nsigma * MAD_NORMAL
    float *out;         // [nlines * nseg]
    u8 *out_flags;      // LM_TIME_MEDIAN: 1 where the line had no samples
};

template <int VPL>
__global__ void __launch_bounds__(128)
k_line_median(LineMedianArgs a)
{
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nseg = a.seg_ends ? a.nseg : 1;
    if (warp >= a.nlines * nseg) return;
    int64_t line = warp / nseg;
    int seg = (int)(warp - line * nseg);
    int64_t outer = line / a.ninner, inner = line - outer * a.ninner;
    int64_t s0 = a.seg_ends ? a.seg_ends[seg] : 0;
    int n = a.seg_ends ? (int)(a.seg_ends[seg + 1] - s0) : a.n;
    int64_t base = outer * a.outer_stride + inner * a.inner_stride + s0 * a.elem_stride;

    uint32_t key[VPL];
    uint32_t valid = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < VPL; k++) {
        int i = lane + 32 * k;
        key[k] = 0;
        if (i < n) {
            int64_t idx = base + (int64_t)i * a.elem_stride;
            bool fl = (a.flags && a.flags[idx]) || (a.flags2 && a.flags2[idx]);
            if (!fl) {
                float x = a.data[idx];
                if (a.use_abs) x = fabsf(x);
                key[k] = f2key(x);
                valid |= 1u << k;
                cnt++;
            }
        }
    }
    int total = warp_sum_i(cnt);
    float result;
    if (total == 0) {
        if (a.mode == LM_TIME_MEDIAN) {
            if (lane == 0) { a.out[warp] = 0.0f; a.out_flags[warp] = 1; }
        } else {
            if (lane == 0) a.out[warp] = INFINITY;  // NaN median -> threshold inf (flagging.py:625-626)
        }
        return;
    }
    // select rank `kth` (0-based) = upper median
    int kth = total >> 1;
    int remaining = kth;
    uint32_t prefix = 0;
    uint32_t cand = valid;  // lanes' candidates whose high bits match `prefix`
    for (int bit = 31; bit >= 0; bit--) {
        int c0 = 0;
#pragma unroll
        for (int k = 0; k < VPL; k++)
            c0 += ((cand >> k) & 1u) & (((key[k] >> bit) & 1u) ^ 1u);
        c0 = warp_sum_i(c0);
        uint32_t take1 = remaining >= c0 ? 1u : 0u;
        if (take1) { remaining -= c0; prefix |= 1u << bit; }
        uint32_t nc = 0;
#pragma unroll
        for (int k = 0; k < VPL; k++)
            nc |= ((((key[k] >> bit) & 1u) == take1) ? 1u : 0u) << k;
        cand &= nc;
    }
    // `prefix` is the key of the upper median; `remaining` its rank among equals
    float upper = key2f(prefix);
    float lower = upper;
    if (!(total & 1) && remaining == 0) {
        uint32_t best = 0;
#pragma unroll
        for (int k = 0; k < VPL; k++)
            if (((valid >> k) & 1u) && key[k] < prefix && key[k] > best) best = key[k];
        best = warp_max_u(best);
        lower = key2f(best);
    }
    double med = median_from_pair(lower, upper, total);
    float medf = (float)med;
    if (a.mode == LM_TIME_MEDIAN) {
        result = medf;
        if (lane == 0) { a.out[warp] = result; a.out_flags[warp] = 0; }
    } else {
        // threshold[idx] *= outlier_nsigma * MAD_NORMAL in float32 storage (flagging.py:622-628)
        result = (float)((double)medf * a.thr_scale);
        if (lane == 0) a.out[warp] = result;
    }
}

// generic fallback for lines longer than 32*32 samples: keys are re-read from
// memory on every bit (rare: only freq_chunks = 1 with thousands of channels
// or more than 1024 dumps)
__global__ void __launch_bounds__(128)
k_line_median_long(LineMedianArgs a)
{
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nseg = a.seg_ends ? a.nseg : 1;
    if (warp >= a.nlines * nseg) return;
    int64_t line = warp / nseg;
    int seg = (int)(warp - line * nseg);
    int64_t outer = line / a.ninner, inner = line - outer * a.ninner;
    int64_t s0 = a.seg_ends ? a.seg_ends[seg] : 0;
    int n = a.seg_ends ? (int)(a.seg_ends[seg + 1] - s0) : a.n;
    int64_t base = outer * a.outer_stride + inner * a.inner_stride + s0 * a.elem_stride;

    int cnt = 0;
    for (int i = lane; i < n; i += 32) {
        int64_t idx = base + (int64_t)i * a.elem_stride;
        bool fl = (a.flags && a.flags[idx]) || (a.flags2 && a.flags2[idx]);
        cnt += fl ? 0 : 1;
    }
    int total = warp_sum_i(cnt);
    if (total == 0) {
        if (lane == 0) {
            if (a.mode == LM_TIME_MEDIAN) { a.out[warp] = 0.0f; a.out_flags[warp] = 1; }
            else a.out[warp] = INFINITY;
        }
        return;
    }
    int remaining = total >> 1;
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        // candidates: keys whose bits above `bit` equal prefix
        uint32_t himask = bit == 31 ? 0u : ~((2u << bit) - 1u);
        int c0 = 0;
        for (int i = lane; i < n; i += 32) {
            int64_t idx = base + (int64_t)i * a.elem_stride;
            bool fl = (a.flags && a.flags[idx]) || (a.flags2 && a.flags2[idx]);
            if (fl) continue;
            float x = a.data[idx];
            if (a.use_abs) x = fabsf(x);
            uint32_t k = f2key(x);
            if ((k & himask) == prefix && !((k >> bit) & 1u)) c0++;
        }
        c0 = warp_sum_i(c0);
        if (remaining >= c0) { remaining -= c0; prefix |= 1u << bit; }
    }
    float upper = key2f(prefix), lower = upper;
    if (!(total & 1) && remaining == 0) {
        uint32_t best = 0;
        for (int i = lane; i < n; i += 32) {
            int64_t idx = base + (int64_t)i * a.elem_stride;
            bool fl = (a.flags && a.flags[idx]) || (a.flags2 && a.flags2[idx]);
            if (fl) continue;
            float x = a.data[idx];
            if (a.use_abs) x = fabsf(x);
            uint32_t k = f2key(x);
            if (k < prefix && k > best) best = k;
        }
        best = warp_max_u(best);
        lower = key2f(best);
    }
    float medf = (float)median_from_pair(lower, upper, total);
    if (lane == 0) {
        if (a.mode == LM_TIME_MEDIAN) { a.out[warp] = medf; a.out_flags[warp] = 0; }
        else a.out[warp] = (float)((double)medf * a.thr_scale);
    }
}

static int launch_line_median(tc_context *c, const LineMedianArgs &a, int maxlen)
{
    int nseg = a.seg_ends ? a.nseg : 1;
    int64_t nwarps = a.nlines * nseg;
    if (nwarps == 0) return TC_OK;
    unsigned grid = tc_blocks_for(nwarps * 32, 128);
    tc_prof_begin(c, TCP_LINE_MEDIAN);
    if (maxlen <= 128) TC_LAUNCH(k_line_median<4>, grid, 128, 0, c->stream, a);
    else if (maxlen <= 256) TC_LAUNCH(k_line_median<8>, grid, 128, 0, c->stream, a);
    else if (maxlen <= 512) TC_LAUNCH(k_line_median<16>, grid, 128, 0, c->stream, a);
    else if (maxlen <= 1024) TC_LAUNCH(k_line_median<32>, grid, 128, 0, c->stream, a);
    else TC_LAUNCH(k_line_median_long, grid, 128, 0, c->stream, a);
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}

// ----------------------------------------------------------------------------
// Block-per-range median of |x| over unflagged samples, followed by the
// background rejection update (reference: _get_background2d inner chunk loop,
// flagging.py:556-574).
//
// One block owns one contiguous range resid[lo, hi) of the transposed (F,T)
// residual plane (a frequency chunk of one plane is contiguous there) and
//   1. radix-selects the median of the unflagged |resid| (three digit passes of
//      11/11/10 bits through a shared-memory histogram, plus one max pass when
//      the count is even),
//   2. forms threshold = median * (MAD_NORMAL * reject) in float64,
//   3. sets flags[i] where (double)resid[i] > threshold.
// `resid` already holds |data - background| (written by the filter epilogue).
// ----------------------------------------------------------------------------
enum { CS_REPORT = 0, CS_BACKGROUND = 1, CS_UVCONTSUB = 2 };

struct ChunkSelectArgs {
    const float *resid;
    u8 *flags;
    const int64_t *range_lo;  // device [nranges]
    const int64_t *range_hi;
    int mode;                 // CS_*
    double thr_mult;          // CS_BACKGROUND: MAD_NORMAL * reject_threshold
    int take_abs;             // select on fabsf(x - sub) instead of x
    const double *sub;        // optional [nranges] value subtracted before fabsf
    int skip_nan;             // leave NaN samples out of the selection (np.nanmedian)
    double *medians;          // optional [nranges] (NaN when nothing takes part)
    // CS_UVCONTSUB (flagging.py:1056-1071): flag x > float32(sigma) * mad
    float uv_sigma;
    int uv_replace;           // 1: flags = new, 0: flags |= new
    const int *uv_unflagged;  // [nranges] number of unflagged samples (0 -> plane skipped)
};

#define TC_SEL_BINS 2048

__device__ __forceinline__ float cs_value(const ChunkSelectArgs &a, int64_t i, float sub)
{
    float x = a.resid[i];
    if (a.take_abs) x = fabsf(x - sub);
    return x;
}

__global__ void __launch_bounds__(1024)
k_chunk_select(ChunkSelectArgs a)
{
    __shared__ uint32_t hist[TC_SEL_BINS];
    __shared__ uint32_t s_prefix, s_remaining, s_total, s_best;
    __shared__ double s_thr;
    const int64_t lo = a.range_lo[blockIdx.x], hi = a.range_hi[blockIdx.x];
    const int tid = threadIdx.x, nt = blockDim.x;
    const float sub = a.sub ? (float)a.sub[blockIdx.x] : 0.0f;

    // digit layout over the 32-bit key: [31:21] [20:10] [9:0]
    uint32_t prefix = 0, himask = 0;
    uint32_t remaining = 0, total = 0;
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const uint32_t dmask = pass == 2 ? 1023u : 2047u;
        for (int b = tid; b < TC_SEL_BINS; b += nt) hist[b] = 0;
        __syncthreads();
        for (int64_t i = lo + tid; i < hi; i += nt) {
            if (a.flags[i]) continue;
            float x = cs_value(a, i, sub);
            if (a.skip_nan && x != x) continue;
            uint32_t k = f2key(x);
            if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & dmask], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            if (pass == 0) {
                uint32_t t = 0;
                for (int b = 0; b < TC_SEL_BINS; b++) t += hist[b];
                s_total = t;
                s_remaining = t >> 1;
            }
            uint32_t rem = s_remaining, acc = 0;
            uint32_t digit = 0;
            if (s_total > 0) {
                for (uint32_t b = 0; b <= dmask; b++) {
                    if (rem < acc + hist[b]) { digit = b; break; }
                    acc += hist[b];
                }
                s_remaining = rem - acc;
            }
            s_prefix = prefix | (digit << shift);
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        total = s_total;
        himask |= dmask << shift;
        if (total == 0) break;
    }
    double med;
    if (total == 0) {
        med = NAN;
    } else {
        float upper = key2f(prefix), lower = upper;
        if (!(total & 1u) && remaining == 0) {
            if (tid == 0) s_best = 0;
            __syncthreads();
            uint32_t best = 0;
            for (int64_t i = lo + tid; i < hi; i += nt) {
                if (a.flags[i]) continue;
                float x = cs_value(a, i, sub);
                if (a.skip_nan && x != x) continue;
                uint32_t k = f2key(x);
                if (k < prefix && k > best) best = k;
            }
            best = warp_max_u(best);
            if ((tid & 31) == 0) atomicMax(&s_best, best);
            __syncthreads();
            lower = key2f(s_best);
        }
        med = median_from_pair(lower, upper, (int)total);
    }
    if (a.medians && tid == 0) a.medians[blockIdx.x] = med;
    if (a.mode == CS_REPORT) return;
    if (a.mode == CS_BACKGROUND) {
        // threshold *= MAD_NORMAL * reject (float64); residual > threshold flags
        double thr = med * a.thr_mult;
        if (thr != thr) return;  // NaN threshold never flags
        for (int64_t i = lo + tid; i < hi; i += nt)
            if ((double)a.resid[i] > thr) a.flags[i] = 1;
        return;
    }
    // CS_UVCONTSUB
    if (a.uv_unflagged[blockIdx.x] == 0) return;  // fully flagged plane: untouched
    float thr = a.uv_sigma * (float)med;          // float32 product (NEP 50)
    (void)s_thr;
    for (int64_t i = lo + tid; i < hi; i += nt) {
        bool nf = a.resid[i] > thr;               // false for NaN on either side
        if (a.uv_replace) a.flags[i] = nf ? 1 : 0;
        else if (nf) a.flags[i] = 1;
    }
}

static int launch_chunk_select(tc_context *c, const ChunkSelectArgs &a, int64_t nranges, int64_t max_range)
{
    if (nranges == 0) return TC_OK;
    int bd = 1024;
    if (max_range <= 4096) bd = 128;
    else if (max_range <= 32768) bd = 256;
    else if (max_range <= 131072) bd = 512;
    tc_prof_begin(c, TCP_CHUNK_SELECT);
    TC_LAUNCH(k_chunk_select, (unsigned)nranges, bd, 0, c->stream, a);
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
