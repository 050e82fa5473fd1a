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
#ifndef TC_EMU
    return __reduce_add_sync(TC_FULL_MASK, v);     // one REDUX instruction
#else
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TC_FULL_MASK, v, o);
    return v;
#endif
}
__device__ __forceinline__ uint32_t warp_max_u(uint32_t v)
{
#ifndef TC_EMU
    return __reduce_max_sync(TC_FULL_MASK, v);
#endif
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

// c += key < t as a compare and a predicated add (the compiler's own sequence is
// three instructions on one dependency chain)
__device__ __forceinline__ void lm_count_below(int &c, uint32_t key, uint32_t t)
{
#ifndef TC_EMU
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(key), "r"(t));
#else
    c += key < t ? 1 : 0;
#endif
}

template <int VPL>
__global__ void __launch_bounds__(128)
k_line_median(LineMedianArgs a)
{
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nseg = a.seg_ends ? a.nseg : 1;
    if (warp >= a.nlines * nseg) return;
    int64_t line, outer, inner;
    int seg;
    if (a.nlines * nseg < (int64_t)1 << 31) {
        // 32-bit divisions are inline; 64-bit ones are subroutine calls
        const unsigned w32 = (unsigned)warp, l32 = w32 / (unsigned)nseg, o32 = l32 / (unsigned)a.ninner;
        line = l32; seg = (int)(w32 - l32 * (unsigned)nseg);
        outer = o32; inner = l32 - o32 * (unsigned)a.ninner;
    } else {
        line = warp / nseg;
        seg = (int)(warp - line * nseg);
        outer = line / a.ninner; inner = line - outer * a.ninner;
    }
    int64_t s0 = a.seg_ends ? a.seg_ends[seg] : 0;
    int n = a.seg_ends ? (int)(a.seg_ends[seg + 1] - s0) : a.n;
    int64_t base = outer * a.outer_stride + inner * a.inner_stride + s0 * a.elem_stride;

    uint32_t key[VPL];
    uint32_t valid = 0;
    int cnt = 0;
    // all loads first (independent of each other), then the keys
    float xraw[VPL];
    u8 fraw[VPL];
#pragma unroll
    for (int k = 0; k < VPL; k++) {
        const int i = lane + 32 * k;
        xraw[k] = 0.f;
        fraw[k] = 1;
        if (i < n) {
            const int64_t idx = base + (int64_t)i * a.elem_stride;
            xraw[k] = a.data[idx];
            u8 f = a.flags ? a.flags[idx] : (u8)0;
            if (a.flags2) f |= a.flags2[idx];
            fraw[k] = f;
        }
    }
#pragma unroll
    for (int k = 0; k < VPL; k++) {
        key[k] = 0;
        if (!fraw[k]) {
            float x = xraw[k];
            if (a.use_abs) x = fabsf(x);
            key[k] = f2key(x);
            valid |= 1u << k;
            cnt++;
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
    // select rank `kth` (0-based) = upper median, one bit per round: the answer is
    // the largest value v with |{keys < v}| <= kth.  Flagged slots hold the
    // largest key so that they never count.
#pragma unroll
    for (int k = 0; k < VPL; k++)
        if (!((valid >> k) & 1u)) key[k] = 0xffffffffu;
    const int kth = total >> 1;
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t t = prefix | (1u << bit);
        // four independent partial counts (VPL is a multiple of 4): short dependency chains
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int k = 0; k < VPL; k += 4) {
            lm_count_below(c0, key[k], t);
            lm_count_below(c1, key[k + 1], t);
            lm_count_below(c2, key[k + 2], t);
            lm_count_below(c3, key[k + 3], t);
        }
        const int c = warp_sum_i((c0 + c1) + (c2 + c3));
        if (c <= kth) prefix = t;
    }
    // `prefix` is the key of the upper median; the lower one differs only when the
    // upper median is the first of its equals
    float upper = key2f(prefix);
    float lower = upper;
    if (!(total & 1)) {
        int c = 0;
        uint32_t best = 0;
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            const bool lt = key[k] < prefix;
            c += lt ? 1 : 0;
            if (lt && key[k] > best) best = key[k];
        }
        c = warp_sum_i(c);
        best = warp_max_u(best);
        if (c == kth) lower = key2f(best);
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

// ----------------------------------------------------------------------------
// The same warp-per-line median by interpolation search instead of 32 bit rounds.
//
// With c(t) = |{valid keys < t}| the wanted middle ranks klow <= kth (equal for an
// odd count) stay bracketed by lo < hi with c(lo) <= klow and c(hi) > kth.  A round
// counts the keys below a trial key t placed by linear interpolation of the rank
// between the bracket ends (the key of a float is close to its logarithm, so this
// converges like a quantile estimate; bisection of the key range every other late
// round bounds the worst case) and moves one end.  When at most 32 keys are left
// inside the bracket they are compacted one per lane, sorted by a 15-stage bitonic
// network and the middle ranks are read off with two shuffles.  Measured on
// half-normal / Rayleigh / heavy-tailed lines of 410 ... 1024 samples: 4 - 6 rounds
// on average.  The result is the same exact order statistic as k_line_median's.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_min_u(uint32_t v)
{
#ifndef TC_EMU
    return __reduce_min_sync(TC_FULL_MASK, v);
#endif
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t t = __shfl_xor_sync(TC_FULL_MASK, v, o);
        v = t < v ? t : v;
    }
    return v;
}

template <int VPL>
__global__ void __launch_bounds__(128)
k_line_median2(LineMedianArgs a)
{
    __shared__ uint32_t s_cand[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nseg = a.seg_ends ? a.nseg : 1;
    if (warp >= a.nlines * nseg) return;
    int64_t outer, inner;
    int seg;
    if (a.nlines * nseg < (int64_t)1 << 31) {
        const unsigned w32 = (unsigned)warp, l32 = w32 / (unsigned)nseg, o32 = l32 / (unsigned)a.ninner;
        seg = (int)(w32 - l32 * (unsigned)nseg);
        outer = o32; inner = l32 - o32 * (unsigned)a.ninner;
    } else {
        const int64_t line = warp / nseg;
        seg = (int)(warp - line * nseg);
        outer = line / a.ninner; inner = line - outer * a.ninner;
    }
    const int64_t s0 = a.seg_ends ? a.seg_ends[seg] : 0;
    const int n = a.seg_ends ? (int)(a.seg_ends[seg + 1] - s0) : a.n;
    const int64_t base = outer * a.outer_stride + inner * a.inner_stride + s0 * a.elem_stride;

    uint32_t key[VPL];
    float xraw[VPL];
    u8 fraw[VPL];
    if (a.elem_stride == 1) {
        // contiguous lines (every caller of the library): one pointer per array, the sample index is an
        // immediate offset of the load (the 64-bit index products of the general form were a fifth of the
        // kernel's instructions)
        const float *pd = a.data + base + lane;
        const u8 *pf = a.flags ? a.flags + base + lane : nullptr;
        const u8 *pf2 = a.flags2 ? a.flags2 + base + lane : nullptr;
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            xraw[k] = 0.f;
            fraw[k] = 1;
            if (lane + 32 * k < n) {
                xraw[k] = pd[32 * k];
                u8 f = pf ? pf[32 * k] : (u8)0;
                if (pf2) f |= pf2[32 * k];
                fraw[k] = f;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            const int i = lane + 32 * k;
            xraw[k] = 0.f;
            fraw[k] = 1;
            if (i < n) {
                const int64_t idx = base + (int64_t)i * a.elem_stride;
                xraw[k] = a.data[idx];
                u8 f = a.flags ? a.flags[idx] : (u8)0;
                if (a.flags2) f |= a.flags2[idx];
                fraw[k] = f;
            }
        }
    }
    int cnt = 0;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int k = 0; k < VPL; k++) {
        key[k] = 0xffffffffu;            // flagged slots never count: every trial key is <= 0xffffffff
        if (!fraw[k]) {
            float x = xraw[k];
            if (a.use_abs) x = fabsf(x);
            const uint32_t kk = f2key(x);
            key[k] = kk;
            kmin = kk < kmin ? kk : kmin;
            kmax = kk > kmax ? kk : kmax;
            cnt++;
        }
    }
    const int total = warp_sum_i(cnt);
    if (total == 0) {
        if (lane == 0) {
            if (a.mode == LM_TIME_MEDIAN) { a.out[warp] = 0.0f; a.out_flags[warp] = 1; }
            else a.out[warp] = INFINITY;  // NaN median -> threshold inf (flagging.py:625-626)
        }
        return;
    }
    kmin = warp_min_u(kmin);
    kmax = warp_max_u(kmax);
    const int kth = total >> 1;
    const bool even = !(total & 1);
    const int klow = even ? kth - 1 : kth;
    uint32_t lo = kmin, hi = kmax == 0xffffffffu ? kmax : kmax + 1u;
    int clo = 0, chi = total;
    uint32_t lower = 0, upper = 0;
    bool done = false;
    if (kmax == 0xffffffffu) {
        // a valid key equal to the sentinel (a NaN with every mantissa bit set): leave it to the bit rounds
        lo = 0; hi = 0xffffffffu;
    }
    int round = 0;
    while (chi - clo > 32 && hi - lo > 1u) {
        round++;
        const uint32_t width = hi - lo;
        uint32_t d;
        if (round <= 3 || (round & 1)) {
            // rank interpolation; every lane evaluates the same expression on the same values
            // (an approximate quotient: the trial key only steers the search, every lane computes the same value)
            const float f = __fdividef((float)(kth - clo) + 0.5f, (float)(chi - clo));
            const float wf = (float)width * f;
            d = wf >= 4294967040.0f ? 0xffffff00u : (uint32_t)wf;
        } else {
            d = width >> 1;
        }
        d = d < 1u ? 1u : d;
        d = d > width - 1u ? width - 1u : d;
        const uint32_t t = lo + d;
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int k = 0; k < VPL; k += 4) {
            lm_count_below(c0, key[k], t);
            lm_count_below(c1, key[k + 1], t);
            lm_count_below(c2, key[k + 2], t);
            lm_count_below(c3, key[k + 3], t);
        }
        const int c = warp_sum_i((c0 + c1) + (c2 + c3));
        if (c <= klow) { lo = t; clo = c; }
        else if (c > kth) { hi = t; chi = c; }
        else {
            // even count and t falls between the two middle keys
            uint32_t below = 0u, above = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < VPL; k++) {
                const uint32_t kk = key[k];
                if (kk < t) below = kk > below ? kk : below;
                else above = kk < above ? kk : above;
            }
            lower = warp_max_u(below);
            upper = warp_min_u(above);
            done = true;
            break;
        }
    }
    if (!done) {
        if (hi - lo <= 1u) {
            lower = upper = lo;          // every key of the bracket equals lo
        } else {
            // at most 32 keys inside [lo, hi): one per lane, bitonic sort, read the ranks
            const int m = chi - clo;
            int mine = 0;
#pragma unroll
            for (int k = 0; k < VPL; k++) mine += (key[k] >= lo && key[k] < hi) ? 1 : 0;
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(TC_FULL_MASK, inc, o);
                if (lane >= o) inc += v;
            }
            int pos = inc - mine;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < VPL; k++)
                if (key[k] >= lo && key[k] < hi) { if (pos < 32) s_cand[wib][pos] = key[k]; pos++; }
            __syncwarp();
            uint32_t v = lane < m ? s_cand[wib][lane] : 0xffffffffu;
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const uint32_t o = __shfl_xor_sync(TC_FULL_MASK, v, j);
                    const bool up = (lane & kk) == 0 || kk == 32;
                    const bool first = (lane & j) == 0;
                    const uint32_t mn = v < o ? v : o, mx = v < o ? o : v;
                    v = (first == up) ? mn : mx;
                }
            }
            upper = __shfl_sync(TC_FULL_MASK, v, kth - clo);
            lower = even ? __shfl_sync(TC_FULL_MASK, v, klow - clo) : upper;
        }
    }
    const double med = median_from_pair(key2f(lower), key2f(upper), total);
    const float medf = (float)med;
    if (lane == 0) {
        if (a.mode == LM_TIME_MEDIAN) { a.out[warp] = medf; a.out_flags[warp] = 0; }
        else a.out[warp] = (float)((double)medf * a.thr_scale);   // flagging.py:622-628
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

// Lines longer than 32 * 32 samples (32768-channel mode: frequency chunks of 3277
// channels; more than 1024 dumps): the same interpolation search with the keys re-read
// from memory (L1 / L2 hits after the first sweep) in every round -- about eight sweeps
// instead of the 33 of the bit-per-round form.
template <bool UNIT>          // UNIT: contiguous lines (elem_stride == 1), 32-bit sample offsets from one pointer per array
__global__ void __launch_bounds__(128)
k_line_median2_long(LineMedianArgs a)
{
    __shared__ uint32_t s_cand[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nseg = a.seg_ends ? a.nseg : 1;
    if (warp >= a.nlines * nseg) return;
    const int64_t line = warp / nseg;
    const int seg = (int)(warp - line * nseg);
    const int64_t outer = line / a.ninner, inner = line - outer * a.ninner;
    const int64_t s0 = a.seg_ends ? a.seg_ends[seg] : 0;
    const int n = a.seg_ends ? (int)(a.seg_ends[seg + 1] - s0) : a.n;
    const int64_t base = outer * a.outer_stride + inner * a.inner_stride + s0 * a.elem_stride;

    // key of sample i, or the sentinel when it is flagged.  UNIT: flag and sample are fetched side by side (no
    // load waits on another), so that the unrolled sweeps below keep several samples in flight per lane
    const float *pd = a.data + base;
    const u8 *pf = a.flags ? a.flags + base : nullptr;
    const u8 *pf2 = a.flags2 ? a.flags2 + base : nullptr;
    auto key_at = [&](int i) -> uint32_t {
        if (UNIT) {
            float x = pd[i];
            u8 f = pf ? pf[i] : (u8)0;
            if (pf2) f |= pf2[i];
            if (a.use_abs) x = fabsf(x);
            const uint32_t k = f2key(x);
            return f ? 0xffffffffu : k;
        }
        const int64_t idx = base + (int64_t)i * a.elem_stride;
        u8 f = a.flags ? a.flags[idx] : (u8)0;
        if (a.flags2) f |= a.flags2[idx];
        if (f) return 0xffffffffu;
        float x = a.data[idx];
        if (a.use_abs) x = fabsf(x);
        return f2key(x);
    };
    int cnt = 0;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll 4
    for (int i = lane; i < n; i += 32) {
        const uint32_t kk = key_at(i);
        if (kk != 0xffffffffu) {
            cnt++;
            kmin = kk < kmin ? kk : kmin;
            kmax = kk > kmax ? kk : kmax;
        }
    }
    const int total = warp_sum_i(cnt);
    if (total == 0) {
        if (lane == 0) {
            if (a.mode == LM_TIME_MEDIAN) { a.out[warp] = 0.0f; a.out_flags[warp] = 1; }
            else a.out[warp] = INFINITY;
        }
        return;
    }
    kmin = warp_min_u(kmin);
    kmax = warp_max_u(kmax);
    const int kth = total >> 1;
    const bool even = !(total & 1);
    const int klow = even ? kth - 1 : kth;
    uint32_t lo = kmin, hi = kmax + 1u;
    int clo = 0, chi = total;
    uint32_t lower = 0, upper = 0;
    bool done = false;
    int round = 0;
    while (chi - clo > 32 && hi - lo > 1u) {
        round++;
        const uint32_t width = hi - lo;
        uint32_t d;
        if (round <= 3 || (round & 1)) {
            const float f = __fdividef((float)(kth - clo) + 0.5f, (float)(chi - clo));
            const float wf = (float)width * f;
            d = wf >= 4294967040.0f ? 0xffffff00u : (uint32_t)wf;
        } else {
            d = width >> 1;
        }
        d = d < 1u ? 1u : d;
        d = d > width - 1u ? width - 1u : d;
        const uint32_t t = lo + d;
        int c = 0;
        uint32_t below = 0u, above = 0xffffffffu;
#pragma unroll 4
        for (int i = lane; i < n; i += 32) {
            const uint32_t kk = key_at(i);
            if (kk < t) { c++; below = kk > below ? kk : below; }
            else above = kk < above ? kk : above;
        }
        c = warp_sum_i(c);
        if (c <= klow) { lo = t; clo = c; }
        else if (c > kth) { hi = t; chi = c; }
        else {
            lower = warp_max_u(below);
            upper = warp_min_u(above);
            done = true;
            break;
        }
    }
    if (!done) {
        if (hi - lo <= 1u) {
            lower = upper = lo;
        } else {
            const int m = chi - clo;
            // at most 32 keys in [lo, hi): gather them one sweep, lane-ordered slots
            int mine = 0;
#pragma unroll 4
            for (int i = lane; i < n; i += 32) {
                const uint32_t kk = key_at(i);
                mine += (kk >= lo && kk < hi) ? 1 : 0;
            }
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(TC_FULL_MASK, inc, o);
                if (lane >= o) inc += v;
            }
            int pos = inc - mine;
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const uint32_t kk = key_at(i);
                if (kk >= lo && kk < hi) { if (pos < 32) s_cand[wib][pos] = kk; pos++; }
            }
            __syncwarp();
            uint32_t v = lane < m ? s_cand[wib][lane] : 0xffffffffu;
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const uint32_t o = __shfl_xor_sync(TC_FULL_MASK, v, j);
                    const bool up = (lane & kk) == 0 || kk == 32;
                    const bool first = (lane & j) == 0;
                    const uint32_t mn = v < o ? v : o, mx = v < o ? o : v;
                    v = (first == up) ? mn : mx;
                }
            }
            upper = __shfl_sync(TC_FULL_MASK, v, kth - clo);
            lower = even ? __shfl_sync(TC_FULL_MASK, v, klow - clo) : upper;
        }
    }
    const float medf = (float)median_from_pair(key2f(lower), key2f(upper), total);
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
    if (maxlen <= 1024 && !TC_ENV_FLAG("TC_MEDIAN_BITS")) {
        if (maxlen <= 128) TC_LAUNCH(k_line_median2<4>, grid, 128, 0, c->stream, a);
        else if (maxlen <= 256) TC_LAUNCH(k_line_median2<8>, grid, 128, 0, c->stream, a);
        else if (maxlen <= 512) TC_LAUNCH(k_line_median2<16>, grid, 128, 0, c->stream, a);
        else TC_LAUNCH(k_line_median2<32>, grid, 128, 0, c->stream, a);
    } else if (!TC_ENV_FLAG("TC_MEDIAN_BITS")) {
        if (a.elem_stride == 1) TC_LAUNCH(k_line_median2_long<true>, grid, 128, 0, c->stream, a);
        else TC_LAUNCH(k_line_median2_long<false>, grid, 128, 0, c->stream, a);
    } else
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
    const unsigned *todo;     // optional [nranges]: the sliced radix kernels only touch ranges with todo != 0
    double *medbuf;           // [nranges] where the sliced paths leave the median for k_sel_update
    float brk_k;              // half width of the sample bracket in units of sqrt(sample size)
    int brk_slice;            // samples per collecting block (multiple of 4096)
    int brk_tail_max;         // longest range the collecting sweep's last block redoes itself after a missed bracket
    int update_in_tail;       // the last collecting block of a range applies the range's threshold itself (no k_sel_update launch)
};

#define TC_SEL_BINS 2048

__device__ __forceinline__ void hist_add_agg(uint32_t *hist, uint32_t bin, bool active);

__device__ __forceinline__ float cs_value(const ChunkSelectArgs &a, int64_t i, float sub)
{
    float x = a.resid[i];
    if (a.take_abs) x = fabsf(x - sub);
    return x;
}

// Exact median of the keys of [lo, hi) that take part (unflagged and, with skip_nan, not NaN):
// three 11/11/10-bit radix passes over the range itself, histogram in shared memory, the digit
// holding the wanted rank found by all threads together (thread t owns 2048 / blockDim.x
// neighbouring bins).  Every thread of the block calls it (128 ... 1024 threads) and gets the
// result.  hist: 2048 words, s_wsum: 32 words, s_scal: 4 words of shared memory.
__device__ __forceinline__ double block_range_median(const ChunkSelectArgs &a, int64_t lo, int64_t hi, float sub,
                                                     uint32_t *hist, uint32_t *s_wsum, uint32_t *s_scal)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = TC_SEL_BINS / nt;
    // digit layout over the 32-bit key: [31:21] [20:10] [9:0]
    uint32_t prefix = 0, himask = 0, remaining = 0, total = 0;
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const uint32_t dmask = pass == 2 ? 1023u : 2047u;
        for (int b = tid; b < TC_SEL_BINS; b += nt) hist[b] = 0;
        __syncthreads();
        for (int64_t i0 = lo; i0 < hi; i0 += nt) {
            int64_t i = i0 + tid;
            bool act = i < hi;
            uint32_t bin = 0;
            if (act) {
                act = a.flags[i] == 0;
                if (act) {
                    float x = cs_value(a, i, sub);
                    if (a.skip_nan && x != x) act = false;
                    uint32_t k = f2key(x);
                    if ((k & himask) != prefix) act = false;
                    bin = (k >> shift) & dmask;
                }
            }
            hist_add_agg(hist, bin, act);
        }
        __syncthreads();
        const uint32_t *mine = hist + tid * per;
        uint32_t sum = 0;
        for (int q = 0; q < per; q++) sum += mine[q];
        uint32_t inc = sum;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(TC_FULL_MASK, inc, o);
            if ((tid & 31) >= o) inc += v;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
        __syncthreads();
        uint32_t woff = 0, all = 0;
        for (int w = 0; w < (nt >> 5); w++) {
            const uint32_t ws = s_wsum[w];
            if (w < (tid >> 5)) woff += ws;
            all += ws;
        }
        if (pass == 0) { total = all; remaining = all >> 1; }
        if (total == 0) break;                        // uniform: every thread sees the same total
        const uint32_t excl = woff + inc - sum;
        if (remaining >= excl && remaining < excl + sum) {
            uint32_t acc = excl, digit = tid * per;
            for (int q = 0; q < per; q++) {
                const uint32_t hq = mine[q];
                if (remaining < acc + hq) { digit = tid * per + q; break; }
                acc += hq;
            }
            s_scal[0] = prefix | (digit << shift);
            s_scal[1] = remaining - acc;
        }
        __syncthreads();
        prefix = s_scal[0];
        remaining = s_scal[1];
        himask |= dmask << shift;
        __syncthreads();
    }
    if (total == 0) return NAN;
    float upper = key2f(prefix), lower = upper;
    if (!(total & 1u) && remaining == 0) {
        // even count and the upper middle key is the smallest of its value: the lower one is the
        // largest key below it
        if (tid == 0) s_scal[2] = 0;
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
        if ((tid & 31) == 0) atomicMax(&s_scal[2], best);
        __syncthreads();
        lower = key2f(s_scal[2]);
    }
    return median_from_pair(lower, upper, (int)total);
}

__global__ void __launch_bounds__(1024)
k_chunk_select(ChunkSelectArgs a)
{
    __shared__ uint32_t hist[TC_SEL_BINS];
    __shared__ uint32_t s_wsum[32], s_scal[4];
    if (a.todo && !a.todo[blockIdx.x]) return;  // fallback launch: only ranges whose bracket missed
    const int64_t lo = a.range_lo[blockIdx.x], hi = a.range_hi[blockIdx.x];
    const int tid = threadIdx.x, nt = blockDim.x;
    const float sub = a.sub ? (float)a.sub[blockIdx.x] : 0.0f;
    const double med = block_range_median(a, lo, hi, sub, hist, s_wsum, s_scal);
    if (a.medians && tid == 0) a.medians[blockIdx.x] = med;
    if (a.medbuf && tid == 0) a.medbuf[blockIdx.x] = med;
    if (a.mode == CS_REPORT || a.todo) return;
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
    for (int64_t i = lo + tid; i < hi; i += nt) {
        bool nf = a.resid[i] > thr;               // false for NaN on either side
        if (a.uv_replace) a.flags[i] = nf ? 1 : 0;
        else if (nf) a.flags[i] = 1;
    }
}

// ----------------------------------------------------------------------------
// Multi-block variant for long ranges (whole planes in uvcontsub, very wide
// chunks): the same three digit passes, but every range is cut into slices that
// separate blocks histogram concurrently (shared-memory histogram per block,
// flushed with global atomics), with a tiny "pick" kernel between passes.
// ----------------------------------------------------------------------------
struct SelState {
    uint32_t prefix, remaining, total, best;
    double med;
};

#define TC_SEL_SLICE 32768

// warp-aggregated histogram increment: lanes with the same bin elect a leader
__device__ __forceinline__ void hist_add_agg(uint32_t *hist, uint32_t bin, bool active)
{
#ifdef TC_EMU
    if (active) atomicAdd(&hist[bin], 1u);
#else
    unsigned peers = __match_any_sync(TC_FULL_MASK, active ? bin : 0xffffffffu);
    if (active && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
#endif
}

__global__ void __launch_bounds__(1024)
k_sel_hist(ChunkSelectArgs a, const SelState *__restrict__ st, uint32_t *__restrict__ ghist, int pass)
{
    __shared__ uint32_t hist[TC_SEL_BINS];
    const int range = blockIdx.y;
    if (a.todo && !a.todo[range]) return;
    const int64_t lo = a.range_lo[range] + (int64_t)blockIdx.x * TC_SEL_SLICE;
    int64_t hi = lo + TC_SEL_SLICE;
    if (hi > a.range_hi[range]) hi = a.range_hi[range];
    if (lo >= hi) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (pass > 0 && st[range].total == 0) return;
    const float sub = a.sub ? (float)a.sub[range] : 0.0f;
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const uint32_t dmask = pass == 2 ? 1023u : 2047u;
    const uint32_t himask = pass == 0 ? 0u : (pass == 1 ? (2047u << 21) : ((2047u << 21) | (2047u << 10)));
    const uint32_t prefix = pass == 0 ? 0u : st[range].prefix;
    for (int b = tid; b < TC_SEL_BINS; b += nt) hist[b] = 0;
    __syncthreads();
    for (int64_t i0 = lo; i0 < hi; i0 += nt) {
        int64_t i = i0 + tid;
        bool act = i < hi;
        uint32_t bin = 0;
        if (act) {
            act = a.flags[i] == 0;
            if (act) {
                float x = cs_value(a, i, sub);
                if (a.skip_nan && x != x) act = false;
                uint32_t k = f2key(x);
                if ((k & himask) != prefix) act = false;
                bin = (k >> shift) & dmask;
            }
        }
        hist_add_agg(hist, bin, act);
    }
    __syncthreads();
    uint32_t *gh = ghist + (size_t)range * TC_SEL_BINS;
    for (int b = tid; b < TC_SEL_BINS; b += nt)
        if (hist[b]) atomicAdd(&gh[b], hist[b]);
}

// one block per range: find the digit holding the wanted rank, clear the histogram
__global__ void __launch_bounds__(256)
k_sel_pick(SelState *__restrict__ st, uint32_t *__restrict__ ghist, int pass, const unsigned *__restrict__ todo)
{
    __shared__ uint32_t part[256];
    __shared__ uint32_t s_digit, s_acc;
    const int range = blockIdx.x, tid = threadIdx.x;
    if (todo && !todo[range]) return;
    uint32_t *gh = ghist + (size_t)range * TC_SEL_BINS;
    const int per = TC_SEL_BINS / 256;
    uint32_t loc[TC_SEL_BINS / 256];
    uint32_t sum = 0;
    for (int k = 0; k < per; k++) { loc[k] = gh[tid * per + k]; sum += loc[k]; }
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int t = 0; t < 256; t++) total += part[t];
        if (pass == 0) { st[range].total = total; st[range].remaining = total >> 1; st[range].prefix = 0; st[range].best = 0; }
        uint32_t rem = st[range].remaining, acc = 0;
        int seg = 0;
        if (st[range].total > 0) {
            for (seg = 0; seg < 255; seg++) {
                if (rem < acc + part[seg]) break;
                acc += part[seg];
            }
        }
        s_digit = (uint32_t)seg;
        s_acc = acc;
    }
    __syncthreads();
    if (tid == (int)s_digit && st[range].total > 0) {
        uint32_t rem = st[range].remaining, acc = s_acc;
        uint32_t digit = tid * per;
        for (int k = 0; k < per; k++) {
            if (rem < acc + loc[k]) { digit = tid * per + k; break; }
            acc += loc[k];
        }
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        st[range].remaining = rem - acc;
        st[range].prefix |= digit << shift;
    }
    for (int k = 0; k < per; k++) gh[tid * per + k] = 0;
}

// largest key below the selected one (needed for even counts, see the
// single-block kernel)
__global__ void __launch_bounds__(1024)
k_sel_lower(ChunkSelectArgs a, SelState *__restrict__ st)
{
    const int range = blockIdx.y;
    if (a.todo && !a.todo[range]) return;
    const SelState s = st[range];
    if (s.total == 0 || (s.total & 1u) || s.remaining != 0) return;
    const int64_t lo = a.range_lo[range] + (int64_t)blockIdx.x * TC_SEL_SLICE;
    int64_t hi = lo + TC_SEL_SLICE;
    if (hi > a.range_hi[range]) hi = a.range_hi[range];
    const float sub = a.sub ? (float)a.sub[range] : 0.0f;
    uint32_t best = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (a.flags[i]) continue;
        float x = cs_value(a, i, sub);
        if (a.skip_nan && x != x) continue;
        uint32_t k = f2key(x);
        if (k < s.prefix && k > best) best = k;
    }
    best = warp_max_u(best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&st[range].best, best);
}

__global__ void __launch_bounds__(128)
k_sel_finish(ChunkSelectArgs a, SelState *__restrict__ st, int nranges)
{
    int range = blockIdx.x * blockDim.x + threadIdx.x;
    if (range >= nranges) return;
    if (a.todo && !a.todo[range]) return;
    SelState s = st[range];
    double med = NAN;
    if (s.total > 0) {
        float upper = key2f(s.prefix), lower = upper;
        if (!(s.total & 1u) && s.remaining == 0) lower = key2f(s.best);
        med = median_from_pair(lower, upper, (int)s.total);
    }
    st[range].med = med;
    if (a.medbuf) a.medbuf[range] = med;
    if (a.medians) a.medians[range] = med;
}

// flags of [lo, hi) from the range's median: CS_BACKGROUND flags |= resid > median * thr_mult (float64),
// CS_UVCONTSUB the rule of flagging.py:1056-1071.  Called by every thread of a block.
__device__ __forceinline__ void sel_update_span(const ChunkSelectArgs &a, int range, double med, int64_t lo, int64_t hi,
                                                bool allow_vec)
{
    double thr;
    bool replace = false;
    if (a.mode == CS_BACKGROUND) {
        // threshold *= MAD_NORMAL * reject (float64); residual > threshold flags
        thr = med * a.thr_mult;
        if (thr != thr) return;  // NaN threshold never flags
    } else if (a.mode == CS_UVCONTSUB) {
        if (a.uv_unflagged[range] == 0) return;     // fully flagged plane: untouched
        thr = (double)(a.uv_sigma * (float)med);    // float32 product (NEP 50); NaN never flags
        replace = a.uv_replace != 0;
    } else {
        return;
    }
    const bool vec = allow_vec && ((lo & 3) == 0) && ((((uintptr_t)a.resid) & 15) == 0) && ((((uintptr_t)a.flags) & 3) == 0);
    if (vec) {
        const int64_t n4 = (hi - lo) >> 2;
        // "x > thr" on a float32 sample against the float64 threshold is decided in float32 against the largest
        // float32 <= thr (the same predicate; NaN never flags), two 16-byte loads in flight per thread
        const float thrf = __double2float_rd(thr);
        const float4 *xp = reinterpret_cast<const float4 *>(a.resid + lo);
        uint32_t *fp0 = reinterpret_cast<uint32_t *>(a.flags + lo);
        for (int64_t q0 = threadIdx.x; q0 < n4; q0 += 2 * (int64_t)blockDim.x) {
            float4 x[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int64_t q = q0 + j * (int64_t)blockDim.x;
                x[j] = q < n4 ? xp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int64_t q = q0 + j * (int64_t)blockDim.x;
                if (q >= n4) continue;
                const uint32_t nf = (x[j].x > thrf ? 1u : 0u) | (x[j].y > thrf ? 0x100u : 0u) |
                                    (x[j].z > thrf ? 0x10000u : 0u) | (x[j].w > thrf ? 0x1000000u : 0u);
                if (replace) fp0[q] = nf;
                else if (nf) fp0[q] |= nf;                 // keep bytes 0/1: set byte to 1 where newly flagged
            }
        }
        for (int64_t i = lo + n4 * 4 + threadIdx.x; i < hi; i += blockDim.x) {
            bool nf = (double)a.resid[i] > thr;
            if (replace) a.flags[i] = nf ? 1 : 0;
            else if (nf) a.flags[i] = 1;
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            bool nf = (double)a.resid[i] > thr;
            if (replace) a.flags[i] = nf ? 1 : 0;
            else if (nf) a.flags[i] = 1;
        }
    }
}

__global__ void __launch_bounds__(1024, 2)
k_sel_update(ChunkSelectArgs a)
{
    const int range = blockIdx.y;
    const int64_t lo = a.range_lo[range] + (int64_t)blockIdx.x * TC_SEL_SLICE;
    int64_t hi = lo + TC_SEL_SLICE;
    if (hi > a.range_hi[range]) hi = a.range_hi[range];
    if (lo >= hi) return;
    sel_update_span(a, range, a.medbuf[range], lo, hi, true);
}

static int launch_chunk_select_multi(tc_context *c, const ChunkSelectArgs &a_in, int64_t nranges, int64_t max_range,
                                     bool with_update = true)
{
    tc_mark mark = tc_arena_mark(c);
    SelState *st;
    uint32_t *gh;
    ChunkSelectArgs a = a_in;
    if (!a.medbuf) TC_TRY(tc_alloc(c, (size_t)nranges, &a.medbuf));
    TC_TRY(tc_alloc(c, (size_t)nranges, &st));
    TC_TRY(tc_alloc(c, (size_t)nranges * TC_SEL_BINS, &gh));
    TC_CUDA(cudaMemsetAsync(gh, 0, sizeof(uint32_t) * (size_t)nranges * TC_SEL_BINS, c->stream));
    unsigned slices = (unsigned)((max_range + TC_SEL_SLICE - 1) / TC_SEL_SLICE);
    tc_prof_begin(c, TCP_CHUNK_SELECT);
    for (int64_t r0 = 0; r0 < nranges; r0 += 65535) {
        unsigned nr = (unsigned)(nranges - r0 < 65535 ? nranges - r0 : 65535);
        ChunkSelectArgs b = a;
        b.range_lo += r0; b.range_hi += r0;
        if (b.sub) b.sub += r0;
        if (b.medians) b.medians += r0;
        if (b.uv_unflagged) b.uv_unflagged += r0;
        if (b.todo) b.todo += r0;
        b.medbuf += r0;
        dim3 grid(slices, nr);
        for (int pass = 0; pass < 3; pass++) {
            TC_LAUNCH(k_sel_hist, grid, 1024, 0, c->stream, b, st + r0, gh + r0 * TC_SEL_BINS, pass);
            TC_LAUNCH(k_sel_pick, nr, 256, 0, c->stream, st + r0, gh + r0 * TC_SEL_BINS, pass, b.todo);
        }
        TC_LAUNCH(k_sel_lower, grid, 1024, 0, c->stream, b, st + r0);
        TC_LAUNCH_NOSYNC(k_sel_finish, tc_blocks_for(nr, 128), 128, 0, c->stream, b, st + r0, (int)nr);
        c->launches += 8;
        if (a.mode != CS_REPORT && with_update) {
            TC_LAUNCH_NOSYNC(k_sel_update, grid, 1024, 0, c->stream, b);
            c->launches++;
        }
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    tc_arena_release(c, mark);
    return TC_OK;
}

// ----------------------------------------------------------------------------
// Bracket select: the exact median in two sweeps over the data instead of four.
//   1. k_brk_sample   a stratified sample of 4096 keys per range is gathered in
//                     shared memory; the keys at sample ranks mid -+ delta (two block
//                     radix selects) bracket the true median with overwhelming
//                     probability.  Ranges of at most 4096 samples are settled here.
//   2. k_brk_collect  one sweep: count the keys below the bracket, copy the keys
//                     inside it (a few percent) to a compact buffer.
//   3. k_brk_select   radix select of rank (n/2 - below) inside the compact
//                     buffer; it also yields the lower middle key for even n.
//   4. ranges whose bracket missed (or overflowed the buffer) are redone by the
//      sliced radix select; k_sel_update applies the thresholds.
// The result is the same exact order statistic; only the visiting order differs.
// ----------------------------------------------------------------------------
#define TC_BRK_SAMPLES 4096
#define TC_BRK_TAIL_MAX (1 << 19)   // longest range the collecting sweep redoes itself after a missed bracket
#define TC_BRK_SLICE 32768   // samples per collecting block (its shared stage holds a quarter of that; measured: 16384 +7 %, 65536 equal, 131072 +30 %)

struct BrkState {
    uint32_t lo, hi;        // bracket keys (inclusive)
    uint32_t n_valid, n_below, n_in;
    uint32_t done;          // median already final
    uint32_t pad0, pad1;
};

__device__ __forceinline__ uint32_t brk_hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// key of rank `kth` (0-based) among keys[0, n) in shared memory: three 11/11/10-bit
// radix passes with a shared histogram.  Every thread of the block calls it (256,
// 512 or 1024 threads) and gets the result.
__device__ __forceinline__ uint32_t block_select_smem(const uint32_t *keys, int n, uint32_t kth, uint32_t *hist,
                                                      uint32_t *s_wsum, uint32_t *s_scal)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = TC_SEL_BINS / nt;            // bins per thread: 8, 4 or 2
    uint32_t prefix = 0, himask = 0, remaining = kth;
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const uint32_t dmask = pass == 2 ? 1023u : 2047u;
        for (int q = tid; q < TC_SEL_BINS; q += nt) hist[q] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint32_t k = keys[i];
            if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & dmask], 1u);
        }
        __syncthreads();
        uint32_t loc[8];
        uint32_t sum = 0;
        for (int q = 0; q < per; q++) { loc[q] = hist[tid * per + q]; sum += loc[q]; }
        uint32_t inc = sum;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(TC_FULL_MASK, inc, o);
            if ((tid & 31) >= o) inc += v;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < (tid >> 5); w++) woff += s_wsum[w];
        const uint32_t excl = woff + inc - sum;
        if (remaining >= excl && remaining < excl + sum) {
            uint32_t acc = excl, digit = tid * per;
            for (int q = 0; q < per; q++) {
                if (remaining < acc + loc[q]) { digit = tid * per + q; break; }
                acc += loc[q];
            }
            s_scal[0] = prefix | (digit << shift);
            s_scal[1] = remaining - acc;
        }
        __syncthreads();
        prefix = s_scal[0];
        remaining = s_scal[1];
        himask |= dmask << shift;
        __syncthreads();
    }
    return prefix;
}

// Bracket [ka, kb] of the keys at ranks r_lo <= r_hi among keys[0, n) from TWO histogram passes over the sample
// instead of two exact three-pass selects: the top 22 bits of both keys are found exactly (one 11-bit histogram
// serves both ranks, a second one too whenever they share the top digit -- a +-3 % bracket nearly always does),
// the low 10 bits are rounded outwards.  The bracket only has to contain the median; the collecting sweep
// counts exactly whatever it is given, so rounding its ends outwards by 2^-13 of their value costs nothing.
__device__ __forceinline__ void block_bracket_smem(const uint32_t *keys, int n, uint32_t r_lo, uint32_t r_hi,
                                                   uint32_t *hist, uint32_t *s_wsum, uint32_t *s_scal,
                                                   uint32_t &ka, uint32_t &kb)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = TC_SEL_BINS / nt;
    // digit of `rank` in the current histogram, every thread gets (digit, rank inside the digit's bin)
    auto pick = [&](uint32_t rank, uint32_t &digit, uint32_t &rem) {
        const uint32_t *mine = hist + tid * per;
        uint32_t sum = 0;
        for (int q = 0; q < per; q++) sum += mine[q];
        uint32_t inc = sum;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(TC_FULL_MASK, inc, o);
            if ((tid & 31) >= o) inc += v;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < (tid >> 5); w++) woff += s_wsum[w];
        const uint32_t excl = woff + inc - sum;
        if (rank >= excl && rank < excl + sum) {
            uint32_t acc = excl, d = tid * per;
            for (int q = 0; q < per; q++) {
                if (rank < acc + mine[q]) { d = tid * per + q; break; }
                acc += mine[q];
            }
            s_scal[0] = d;
            s_scal[1] = rank - acc;
        }
        __syncthreads();
        digit = s_scal[0];
        rem = s_scal[1];
        __syncthreads();
    };
    auto fill = [&](uint32_t prefix, bool top) {
        for (int q = tid; q < TC_SEL_BINS; q += nt) hist[q] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint32_t k = keys[i];
            if (top) atomicAdd(&hist[k >> 21], 1u);
            else if ((k >> 21) == prefix) atomicAdd(&hist[(k >> 10) & 2047u], 1u);
        }
        __syncthreads();
    };
    uint32_t d_lo, d_hi, m_lo, m_hi, e_lo, e_hi, t;
    fill(0u, true);
    pick(r_lo, d_lo, m_lo);
    pick(r_hi, d_hi, m_hi);
    fill(d_lo, false);
    pick(m_lo, e_lo, t);
    if (d_hi != d_lo) fill(d_hi, false);               // block-uniform
    pick(m_hi, e_hi, t);
    ka = (d_lo << 21) | (e_lo << 10);
    kb = (d_hi << 21) | (e_hi << 10) | 1023u;
}

__global__ void __launch_bounds__(1024)
k_brk_sample(ChunkSelectArgs a, BrkState *__restrict__ st, unsigned *__restrict__ todo, int update_here)
{
    __shared__ uint32_t keys[TC_BRK_SAMPLES];
    __shared__ uint32_t hist[TC_SEL_BINS];
    __shared__ uint32_t s_wsum[32], s_scal[2];
    __shared__ int s_valid;
    const int range = blockIdx.x, tid = threadIdx.x;
    const int64_t lo = a.range_lo[range], hi = a.range_hi[range];
    const int64_t len = hi - lo;
    const float sub = a.sub ? (float)a.sub[range] : 0.0f;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    const bool exact = len <= TC_BRK_SAMPLES;
    int nv = 0;
    // four samples per thread and trip, flag and value of each loaded side by side (the value of a flagged
    // sample is fetched and dropped): eight independent loads in flight instead of two dependent ones
    for (int j0 = tid; j0 < TC_BRK_SAMPLES; j0 += 4 * (int)blockDim.x) {
        int64_t pos[4];
        u8 fl[4];
        float xv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = j0 + q * (int)blockDim.x;
            pos[q] = -1;
            if (j < TC_BRK_SAMPLES) {
                if (exact) { if (j < len) pos[q] = lo + j; }
                else {
                    // stratified: one sample per stratum of len / 4096 elements
                    int64_t s0 = (int64_t)j * len / TC_BRK_SAMPLES, s1 = (int64_t)(j + 1) * len / TC_BRK_SAMPLES;
                    int64_t w = s1 - s0 > 0 ? s1 - s0 : 1;
                    pos[q] = lo + s0 + (int64_t)(brk_hash((uint32_t)j * 2654435761u ^ (uint32_t)range) % (uint32_t)w);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            fl[q] = 1;
            xv[q] = 0.f;
            if (pos[q] >= 0) { fl[q] = a.flags[pos[q]]; xv[q] = a.resid[pos[q]]; }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = j0 + q * (int)blockDim.x;
            if (j >= TC_BRK_SAMPLES) continue;
            uint32_t k = 0xffffffffu;
            if (!fl[q]) {
                float x = xv[q];
                if (a.take_abs) x = fabsf(x - sub);
                if (!(a.skip_nan && x != x)) { k = f2key(x); nv++; }
            }
            keys[j] = k;
        }
    }
    atomicAdd(&s_valid, nv);
    __syncthreads();
    // two order statistics of the sample are needed (invalid keys = 0xffffffff rank
    // last): the middle pair of an exactly covered range, or the bracket ends
    const int sv = s_valid;
    const int nk = exact ? (int)len : TC_BRK_SAMPLES;
    uint32_t ka = 0, kb = 0;
    bool bracket = false;
    if (exact) {
        if (sv > 0) {
            kb = block_select_smem(keys, nk, (uint32_t)(sv >> 1), hist, s_wsum, s_scal);
            ka = (sv & 1) ? kb : block_select_smem(keys, nk, (uint32_t)((sv >> 1) - 1), hist, s_wsum, s_scal);
        }
    } else if (sv >= 64) {
        const int mid = sv >> 1;
        const int delta = (int)(a.brk_k * sqrtf((float)sv)) + 4;
        bracket = true;
        if (mid - delta >= 0 && mid + delta < sv) {
            block_bracket_smem(keys, nk, (uint32_t)(mid - delta), (uint32_t)(mid + delta), hist, s_wsum, s_scal, ka, kb);
            if (kb > 0xfffffffeu) kb = 0xfffffffeu;
        } else {
            ka = mid - delta >= 0 ? block_select_smem(keys, nk, (uint32_t)(mid - delta), hist, s_wsum, s_scal) : 0u;
            kb = mid + delta < sv ? block_select_smem(keys, nk, (uint32_t)(mid + delta), hist, s_wsum, s_scal) : 0xfffffffeu;
        }
    }
    double med = NAN;
    if (exact && sv > 0) med = median_from_pair(key2f(ka), key2f(kb), sv);
    if (tid == 0) {
        BrkState b;
        b.lo = 0; b.hi = 0xfffffffeu; b.n_valid = 0; b.n_below = 0; b.n_in = 0; b.done = 0; b.pad0 = b.pad1 = 0;
        if (exact) {
            a.medbuf[range] = med;
            if (a.medians) a.medians[range] = med;
            b.done = 1;
            b.n_valid = (uint32_t)sv;
        } else if (bracket) {
            b.lo = ka;
            b.hi = kb;
        }
        st[range] = b;
        todo[range] = 0;
    }
    // a launch whose ranges are all settled here (update_here) applies the thresholds as well:
    // the spectrum stage runs without k_sel_update launches
    if (update_here && exact) sel_update_span(a, range, med, lo, hi, false);
}

// radix select of the wanted rank inside the compact buffer.  Runs as the tail of
// the LAST collecting block of a range (every thread of the block calls it);
// the buffer and the counters were written by other blocks, so they are read
// past L1.
// Returns true when the range's median is settled; `med_out` then holds it in every thread.
__device__ __forceinline__ bool brk_select_tail(const ChunkSelectArgs &a, const BrkState *st, const uint32_t *cbuf,
                                                int64_t cap, unsigned *todo, int range, uint32_t *hist,
                                                uint32_t *s_wsum, uint32_t *s_scal, double &med_out)
{
    uint32_t &s_prefix = s_scal[0], &s_remaining = s_scal[1], &s_best = s_scal[2];
    const int tid = threadIdx.x, nt = blockDim.x;   // 256, 512 or 1024 threads
    BrkState b;
    b.lo = 0; b.hi = 0; b.done = 0; b.pad0 = b.pad1 = 0;
    b.n_valid = __ldcg(&st[range].n_valid);
    b.n_below = __ldcg(&st[range].n_below);
    b.n_in = __ldcg(&st[range].n_in);
    if (b.n_valid == 0) {
        if (tid == 0) { a.medbuf[range] = NAN; if (a.medians) a.medians[range] = NAN; }
        med_out = NAN;
        return true;
    }
    const uint32_t kth = b.n_valid >> 1;
    const bool even = (b.n_valid & 1u) == 0;
    // the wanted rank (and, for even counts, a lower neighbour) must lie inside the buffer
    if ((int64_t)b.n_in > cap || kth < b.n_below || kth >= b.n_below + b.n_in ||
        (even && kth == b.n_below && b.n_below > 0)) {
        // the bracket missed (or the buffer overflowed): ranges of moderate length are redone
        // right here by the plain radix select over the range -- this block is the last one of
        // its range, the others keep sweeping theirs -- longer ones by the fallback launch
        const int64_t rlo = a.range_lo[range], rhi = a.range_hi[range];
        if (rhi - rlo <= a.brk_tail_max) {
            const float sub = a.sub ? (float)a.sub[range] : 0.0f;
            const double med = block_range_median(a, rlo, rhi, sub, hist, s_wsum, s_scal);
            if (tid == 0) { a.medbuf[range] = med; if (a.medians) a.medians[range] = med; }
            med_out = med;
            return true;
        }
        if (tid == 0) todo[range] = 1;
        return false;
    }
    const uint32_t *keys = cbuf + (size_t)range * cap;
    const uint32_t n = b.n_in;
    uint32_t prefix = 0, himask = 0, remaining = kth - b.n_below;
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const uint32_t dmask = pass == 2 ? 1023u : 2047u;
        for (int q = tid; q < TC_SEL_BINS; q += nt) hist[q] = 0;
        __syncthreads();
        // eight loads in flight per thread: the buffer of a long range holds ~10^5 keys and this block is
        // alone with it -- one L2 round trip per key and thread made the tail of the uvcontsub sweeps
        // (2 Mi-sample ranges) as long as the sweep itself (ncu: 0.42 against 0.24 ms)
        for (uint32_t i0 = tid; i0 < n; i0 += 8u * nt) {
            uint32_t kk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t i = i0 + (uint32_t)j * nt;
                kk[j] = i < n ? __ldcg(keys + i) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t i = i0 + (uint32_t)j * nt;
                if (i < n && (kk[j] & himask) == prefix) atomicAdd(&hist[(kk[j] >> shift) & dmask], 1u);
            }
        }
        __syncthreads();
        {
            // parallel search of the digit: thread t owns bins [per t, per (t + 1)), per = 2, 4 or 8
            // (read from the shared histogram both times: no per-thread array, whatever the block size)
            const int per = TC_SEL_BINS / nt;
            const uint32_t *mine = hist + tid * per;
            uint32_t sum = 0;
            for (int q = 0; q < per; q++) sum += mine[q];
            // inclusive scan of `sum` over the block
            uint32_t inc = sum;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(TC_FULL_MASK, inc, o);
                if ((tid & 31) >= o) inc += v;
            }
            if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
            __syncthreads();
            uint32_t woff = 0;
            for (int w = 0; w < (tid >> 5); w++) woff += s_wsum[w];
            const uint32_t excl = woff + inc - sum;
            if (remaining >= excl && remaining < excl + sum) {
                uint32_t acc = excl, digit = tid * per;
                for (int q = 0; q < per; q++) {
                    const uint32_t hq = mine[q];
                    if (remaining < acc + hq) { digit = tid * per + q; break; }
                    acc += hq;
                }
                s_remaining = remaining - acc;
                s_prefix = prefix | (digit << shift);
            }
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        himask |= dmask << shift;
        __syncthreads();
    }
    float upper = key2f(prefix), lower = upper;
    if (even && remaining == 0) {
        if (tid == 0) s_best = 0;
        __syncthreads();
        uint32_t best = 0;
        for (uint32_t i0 = tid; i0 < n; i0 += 8u * nt) {
            uint32_t kk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t i = i0 + (uint32_t)j * nt;
                kk[j] = i < n ? __ldcg(keys + i) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (kk[j] < prefix && kk[j] > best) best = kk[j];
        }
        best = warp_max_u(best);
        if ((tid & 31) == 0 && best) atomicMax(&s_best, best);
        __syncthreads();
        lower = key2f(s_best);
    }
    const double med = median_from_pair(lower, upper, (int)b.n_valid);   // the same value in every thread
    if (tid == 0) {
        a.medbuf[range] = med;
        if (a.medians) a.medians[range] = med;
    }
    med_out = med;
    return true;
}

// sweep: count valid / below-bracket keys, copy in-bracket keys to the compact
// buffer.  In-bracket keys are first gathered in shared memory (one shared
// atomic per warp and iteration), then the block reserves its share of the
// range's buffer with ONE global atomic and copies coalesced.
#define TC_BRK_STAGE 8192
// (two blocks of 1024 threads per SM = 32 registers: the sweep lives on occupancy; the select tail and the
// missed-bracket fallback of the last block may spill)
template <bool TAKE_ABS, bool SKIP_NAN>
__global__ void __launch_bounds__(1024, 2)
k_brk_collect(ChunkSelectArgs a, BrkState *__restrict__ st, uint32_t *__restrict__ cbuf, int64_t cap,
              unsigned *__restrict__ todo)
{
    __shared__ uint32_t stage[TC_BRK_STAGE];
    __shared__ uint32_t s_valid, s_below, s_in, s_base, s_last;
    const int range = blockIdx.y;
    if (st[range].done) return;
    const int64_t lo = a.range_lo[range] + (int64_t)blockIdx.x * a.brk_slice;
    int64_t hi = lo + a.brk_slice;
    if (hi > a.range_hi[range]) hi = a.range_hi[range];
    if (lo >= hi) return;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const uint32_t klo = st[range].lo, khi = st[range].hi;
    const float sub = a.sub ? (float)a.sub[range] : 0.0f;
    uint32_t *out = cbuf + (size_t)range * cap;
    if (tid == 0) { s_valid = 0; s_below = 0; s_in = 0; }
    __syncthreads();
    uint32_t nvalid = 0, nbelow = 0;
    const unsigned lt = (1u << lane) - 1u;
    // four samples of one thread: byte q of `fw` is the flag of sample q (slots that do not exist are passed as
    // flagged), classification, and the warp-level compaction of the in-bracket keys.  Called by every thread of
    // the block the same number of times.
    auto process = [&](uint32_t fw, float x0, float x1, float x2, float x3) {
        const float xv[4] = {x0, x1, x2, x3};
        uint32_t k4[4];
        bool in4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // predicated, no branches: the modes are template parameters
            float x = xv[q];
            if (TAKE_ABS) x = fabsf(x - sub);
            bool valid = !((fw >> (8 * q)) & 0xffu);
            if (SKIP_NAN) valid = valid && !(x != x);
            const uint32_t k = f2key(x);
            const bool below = valid && k < klo;
            const bool inb = valid && !below && k <= khi;
            nvalid += valid ? 1u : 0u;
            nbelow += below ? 1u : 0u;
            in4[q] = inb;
            k4[q] = k;
        }
        // warp-level compaction by ballots: one shared atomic per warp and iteration reserves
        // the warp's slots; a key's slot is the number of in-bracket keys of the earlier
        // sub-samples q plus those of the same q in lower lanes.  The order of the keys in
        // the stage does not matter to the select.
        const unsigned m0 = __ballot_sync(TC_FULL_MASK, in4[0]), m1 = __ballot_sync(TC_FULL_MASK, in4[1]);
        const unsigned m2 = __ballot_sync(TC_FULL_MASK, in4[2]), m3 = __ballot_sync(TC_FULL_MASK, in4[3]);
        if (m0 | m1 | m2 | m3) {
            const uint32_t n0 = __popc(m0), n1 = __popc(m1), n2 = __popc(m2), n3 = __popc(m3);
            const uint32_t total = (n0 + n1) + (n2 + n3);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_in, total);
            base = __shfl_sync(TC_FULL_MASK, base, 0);
            // a stage that overflows is never read (the range goes to the radix fallback)
            if (base + total <= TC_BRK_STAGE) {
                // slots first, then four predicated stores
                const uint32_t p0 = base + __popc(m0 & lt), p1 = base + n0 + __popc(m1 & lt);
                const uint32_t p2 = base + n0 + n1 + __popc(m2 & lt), p3 = base + n0 + n1 + n2 + __popc(m3 & lt);
                if (in4[0]) stage[p0] = k4[0];
                if (in4[1]) stage[p1] = k4[1];
                if (in4[2]) stage[p2] = k4[2];
                if (in4[3]) stage[p3] = k4[3];
            }
        }
    };
    // four samples per thread and iteration: one 4-byte flag word and one 16-byte
    // sample vector (the slice start is 4-aligned whenever the range start is)
    const bool vec = ((lo & 3) == 0) && ((((uintptr_t)a.resid) & 15) == 0) && ((((uintptr_t)a.flags) & 3) == 0);
    if (vec) {
        // iterations in which every thread of the block owns four in-range samples walk two pointers and carry
        // no bounds tests (the mixed loop had both load forms and their predicates in its body: 227 instructions
        // per iteration); what is left of the slice goes through one bounds-checked iteration
        const int64_t len = hi - lo;
        const int nfull = (int)(len / ((int64_t)nt * 4));
        const float4 *px = reinterpret_cast<const float4 *>(a.resid + lo) + tid;
        const uint32_t *pf = reinterpret_cast<const uint32_t *>(a.flags + lo) + tid;
        for (int it = 0; it < nfull; it++) {
            const uint32_t fw = *pf;
            const float4 t4 = *px;
            pf += nt;
            px += nt;
            process(fw, t4.x, t4.y, t4.z, t4.w);
        }
        const int64_t i0 = lo + (int64_t)nfull * nt * 4;
        if (i0 < hi) {                                  // block-uniform
            const int64_t ib = i0 + (int64_t)tid * 4;
            uint32_t fw = 0;
            float xv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int64_t i = ib + q;
                if (i < hi) { fw |= (uint32_t)(a.flags[i] ? 1u : 0u) << (8 * q); xv[q] = a.resid[i]; }
                else fw |= 1u << (8 * q);
            }
            process(fw, xv[0], xv[1], xv[2], xv[3]);
        }
    } else {
        for (int64_t i0 = lo; i0 < hi; i0 += nt) {
            const int64_t i = i0 + tid;
            uint32_t fw = 0x01010100u;                  // one sample per thread, the other three slots do not exist
            float x = 0.f;
            if (i < hi) { fw |= a.flags[i] ? 1u : 0u; x = a.resid[i]; }
            else fw |= 1u;
            process(fw, x, 0.f, 0.f, 0.f);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nvalid += __shfl_xor_sync(TC_FULL_MASK, nvalid, o);
        nbelow += __shfl_xor_sync(TC_FULL_MASK, nbelow, o);
    }
    if (lane == 0) { atomicAdd(&s_valid, nvalid); atomicAdd(&s_below, nbelow); }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&st[range].n_valid, s_valid);
        atomicAdd(&st[range].n_below, s_below);
        // a slice that overflows its stage makes the range's count exceed `cap`,
        // which sends the range to the radix fallback
        uint32_t want = s_in <= TC_BRK_STAGE ? s_in : (uint32_t)(cap + 1);
        s_base = atomicAdd(&st[range].n_in, want);
    }
    __syncthreads();
    const uint32_t cnt = s_in <= TC_BRK_STAGE ? s_in : 0;
    const uint32_t base = s_base;
    for (uint32_t q = tid; q < cnt; q += nt)
        if ((int64_t)(base + q) < cap) out[base + q] = stage[q];
    // the block that finishes a range last selects inside the range's compact buffer.
    // The barrier orders the block's writes before thread 0's device-scope fence
    // (fences are cumulative), the fence orders them before the ticket.
    __syncthreads();
    if (tid == 0) {
        const int64_t len = a.range_hi[range] - a.range_lo[range];
        const unsigned nactive = (unsigned)((len + a.brk_slice - 1) / a.brk_slice);
        __threadfence();
        s_last = atomicAdd(&st[range].pad0, 1u) == nactive - 1 ? 1u : 0u;
        __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    double med;
    const bool settled = brk_select_tail(a, st, cbuf, cap, todo, range, stage, stage + TC_SEL_BINS,
                                         stage + TC_SEL_BINS + 64, med);
    // The block that settles a range applies its threshold as well: every other block of the range has
    // finished (ticket), no other range touches these samples, and the range (a few hundred KB) was read
    // by this launch moments ago, so the update mostly hits L2 instead of being a second sweep over DRAM
    // in a launch of its own.
    if (a.update_in_tail && settled) sel_update_span(a, range, med, a.range_lo[range], a.range_hi[range], true);
}

static int launch_bracket_select(tc_context *c, const ChunkSelectArgs &a_in, int64_t nranges, int64_t max_range)
{
    tc_mark mark = tc_arena_mark(c);
    ChunkSelectArgs a = a_in;
    BrkState *st;
    unsigned *todo;
    TC_TRY(tc_alloc(c, (size_t)nranges, &st));
    TC_TRY(tc_alloc(c, (size_t)nranges, &todo));
    if (!a.medbuf) TC_TRY(tc_alloc(c, (size_t)nranges, &a.medbuf));
    const bool small = max_range <= TC_BRK_SAMPLES;
    static const float env_brk_k = getenv("TC_BRK_K") ? (float)atof(getenv("TC_BRK_K")) : 1.75f;
    a.brk_k = env_brk_k;   // +-3.5 sigma of the sample rank of the median
    a.brk_slice = TC_BRK_SLICE;
    static const int env_tail_max = getenv("TC_BRK_TAIL_MAX") ? atoi(getenv("TC_BRK_TAIL_MAX")) : TC_BRK_TAIL_MAX;
    a.brk_tail_max = env_tail_max;
    // ranges the tail can always settle (no fallback launch) get their thresholds from the tail too;
    // TC_BRK_UPDATE_IN_TAIL=0 keeps the separate k_sel_update launch (A/B knob)
    static const int env_upd_tail = getenv("TC_BRK_UPDATE_IN_TAIL") ? atoi(getenv("TC_BRK_UPDATE_IN_TAIL")) : 1;
    a.update_in_tail = (env_upd_tail && !small && a.mode != CS_REPORT && max_range <= a.brk_tail_max) ? 1 : 0;
    static const int env_brk_slice = getenv("TC_BRK_SLICE") ? atoi(getenv("TC_BRK_SLICE")) : 0;
    if (env_brk_slice >= 4096 && env_brk_slice <= (1 << 20)) a.brk_slice = env_brk_slice & ~4095;
    const int64_t cap = small ? 1 : max_range / 4 + 4096;
    uint32_t *cbuf = nullptr;
    if (!small) TC_TRY(tc_alloc(c, (size_t)nranges * cap, &cbuf));
    unsigned slices = (unsigned)((max_range + TC_SEL_SLICE - 1) / TC_SEL_SLICE);
    tc_prof_begin(c, TCP_CHUNK_SELECT);
    for (int64_t r0 = 0; r0 < nranges; r0 += 65535) {
        unsigned nr = (unsigned)(nranges - r0 < 65535 ? nranges - r0 : 65535);
        ChunkSelectArgs b = a;
        b.range_lo += r0; b.range_hi += r0;
        if (b.sub) b.sub += r0;
        if (b.medians) b.medians += r0;
        if (b.uv_unflagged) b.uv_unflagged += r0;
        b.medbuf += r0;
#ifdef TC_EMU
        const int sample_threads = 256;    // fewer fibers to switch between in the emulated build
#else
        const int sample_threads = 1024;
#endif
        TC_LAUNCH(k_brk_sample, nr, sample_threads, 0, c->stream, b, st + r0, todo + r0,
                  ((small || a.update_in_tail) && a.mode != CS_REPORT) ? 1 : 0);
        c->launches++;
        if (!small) {
            unsigned cslices = (unsigned)((max_range + a.brk_slice - 1) / a.brk_slice);
            static const int env_ct = getenv("TC_BRK_THREADS") ? atoi(getenv("TC_BRK_THREADS")) : 0;
            // measured (chunk_select per step): 32-baseline blocks 1024 threads 68.0 ms, 512 71.7, 256 88.7;
            // 64-baseline blocks 1024 threads 128.9, 512 124.3.  The last block of a range also runs the select
            // tail and the missed-bracket fallback, which wider blocks finish sooner; once the grid is large
            // enough to hide those tails the finer granularity of 512-thread blocks wins.
            const int cthreads = (env_ct == 1024 || env_ct == 512 || env_ct == 256) ? env_ct
                                 : ((int64_t)cslices * nr >= 12288 ? 512 : 1024);
            if (b.take_abs && b.skip_nan)
                TC_LAUNCH((k_brk_collect<true, true>), dim3(cslices, nr), cthreads, 0, c->stream, b, st + r0, cbuf + r0 * cap,
                          cap, todo + r0);
            else if (b.take_abs)
                TC_LAUNCH((k_brk_collect<true, false>), dim3(cslices, nr), cthreads, 0, c->stream, b, st + r0, cbuf + r0 * cap,
                          cap, todo + r0);
            else if (b.skip_nan)
                TC_LAUNCH((k_brk_collect<false, true>), dim3(cslices, nr), cthreads, 0, c->stream, b, st + r0, cbuf + r0 * cap,
                          cap, todo + r0);
            else
                TC_LAUNCH((k_brk_collect<false, false>), dim3(cslices, nr), cthreads, 0, c->stream, b, st + r0, cbuf + r0 * cap,
                          cap, todo + r0);
            c->launches++;
        }
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    if (!small && TC_ENV_FLAG("TC_DEBUG_SELECT")) {
        std::vector<unsigned> h((size_t)nranges);
        std::vector<BrkState> hs((size_t)nranges);
        cudaStreamSynchronize(c->stream);
        cudaMemcpyAsync(h.data(), todo, sizeof(unsigned) * (size_t)nranges, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(hs.data(), st, sizeof(BrkState) * (size_t)nranges, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        int64_t nt = 0, nover = 0, nmiss = 0;
        for (int64_t r = 0; r < nranges; r++)
            if (h[r]) {
                nt++;
                if ((int64_t)hs[r].n_in > cap) nover++; else nmiss++;
            }
        fprintf(stderr, "[tc select] ranges=%lld max_range=%lld mode=%d fallback=%lld (overflow %lld, miss %lld)\n",
                (long long)nranges, (long long)max_range, a.mode, (long long)nt, (long long)nover, (long long)nmiss);
    }
    if (!small && max_range > a.brk_tail_max) {
        // redo the (rare) ranges whose bracket missed with the one-block radix
        // select; blocks of all other ranges exit at once (ranges of at most
        // TC_BRK_TAIL_MAX samples were settled by the collecting sweep's tail)
        ChunkSelectArgs f = a;
        f.todo = todo;
        if (max_range > 64 * TC_SEL_SLICE) {
            // very long ranges: one block per range would crawl, use the sliced radix select
            // (eight launches; below that size the few ranges that need it are cheaper in one launch
            // whose other blocks exit at once -- 40 selects per strategy pass make the difference
            // between ~1100 and ~850 launches per 12-task step)
            TC_TRY(launch_chunk_select_multi(c, f, nranges, max_range, false));
            goto fallback_done;
        }
        tc_prof_begin(c, TCP_CHUNK_SELECT);
        for (int64_t r0 = 0; r0 < nranges; r0 += 65535) {
            unsigned nr = (unsigned)(nranges - r0 < 65535 ? nranges - r0 : 65535);
            ChunkSelectArgs b = f;
            b.range_lo += r0; b.range_hi += r0; b.todo += r0; b.medbuf += r0;
            if (b.sub) b.sub += r0;
            if (b.medians) b.medians += r0;
            TC_LAUNCH(k_chunk_select, nr, 1024, 0, c->stream, b);
            c->launches++;
        }
        tc_prof_end(c);
        TC_KERNEL_CHECK();
    }
fallback_done:
    if (a.mode != CS_REPORT && !small && !a.update_in_tail) {   // small: k_brk_sample applied the thresholds itself
        tc_prof_begin(c, TCP_CHUNK_SELECT);
        for (int64_t r0 = 0; r0 < nranges; r0 += 65535) {
            unsigned nr = (unsigned)(nranges - r0 < 65535 ? nranges - r0 : 65535);
            ChunkSelectArgs b = a;
            b.range_lo += r0; b.range_hi += r0;
            if (b.uv_unflagged) b.uv_unflagged += r0;
            b.medbuf += r0;
            TC_LAUNCH_NOSYNC(k_sel_update, dim3(slices, nr), 1024, 0, c->stream, b);
            c->launches++;
        }
        tc_prof_end(c);
        TC_KERNEL_CHECK();
    }
    tc_arena_release(c, mark);
    return TC_OK;
}

static int launch_chunk_select(tc_context *c, const ChunkSelectArgs &a, int64_t nranges, int64_t max_range)
{
    if (nranges == 0) return TC_OK;
    if (!TC_ENV_FLAG("TC_SELECT_RADIX")) return launch_bracket_select(c, a, nranges, max_range);
    // reference implementation of the select (kept for A/B checks): plain radix select
    if (max_range > 8 * TC_SEL_SLICE || (max_range > TC_SEL_SLICE && nranges < 2 * (int64_t)c->sm_count))
        return launch_chunk_select_multi(c, a, nranges, max_range);
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
