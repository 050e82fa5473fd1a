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
    int flags_transposed;  // FIN_MASKED: `flags` is stored (plane, line, sample) instead of (plane, sample, line)
    int out_transposed;    // outputs are written (plane, line, sample): the layout change is fused into the drain
    int single_axis;    // the only filtered axis of this filter (profile bucket only)
    int role;           // split first-axis kernels: 0 = value and weight blocks alternate, 1 = values only, 2 = weights only
    float *gring;       // global delay-line scratch (when not in shared memory)
    int64_t gring_stride;
};

#define TC_FILT_U 8   // ticks per group: prefetch distance and output staging depth
#define TC_FILT_LPW 4 // lines per warp
#define TC_FILT_SKEW 2 // ticks by which pass p+1 trails pass p (2 takes shuffle+convert off the recurrence)

// Warp-cooperative form.  The 32 lanes of a warp are 4 lines x 2 arrays x 4
// passes: lane = pass*8 + array*4 + line.  Every lane runs the SAME loop body
//     s += entering;  emit = (float)s;  s -= leaving
// on its own accumulator and its own 2r-deep delay line in shared memory
// ([slot][lane], conflict free); pass p+1 picks up what pass p emitted one tick
// earlier with a warp shuffle (lane - 8).  Splitting a line over eight lanes
// keeps the shared-memory footprint per line unchanged (the delay lines) but
// gives the SM eight times as many warps to hide the FP64 / convert latencies.
// Pass-0 lanes fetch the input one group (8 ticks) ahead; pass-3 lanes park
// their outputs in a small staging tile that all 32 lanes drain every 8 ticks
// (division by d^4, background = value / weight, coalesced 16-byte stores).
// shared memory per warp: [staging tile 64][input tiles 2 x 64][ring L x 32] floats
#define TC_FILT_WARP_FIXED 192

template <bool SMEM_RING, int MODE_IN, int MODE_OUT>
__global__ void k_box_filter(FilterArgs a)
{
    TC_DYN_SMEM(float, smem);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;            // warp in block
    const int nwb = blockDim.x >> 5;
    const int pass = lane >> 3;
    const int L = 2 * a.r, r2 = 2 * a.r, r4 = 4 * a.r;
    const int n = a.n;
    const int64_t nj = a.nj;
    const int64_t ngroups = (a.nlines + TC_FILT_LPW - 1) / TC_FILT_LPW;
    const int64_t gwarp = (int64_t)blockIdx.x * nwb + wib;
    const int64_t nwarps = (int64_t)gridDim.x * nwb;
    float *stg = smem + (size_t)wib * (TC_FILT_WARP_FIXED + (SMEM_RING ? (size_t)L * 32 : 0));
    float *tile = stg + 64;                      // 2 x [8 ticks][2 arrays x 4 lines]
    float *ring = SMEM_RING ? (stg + TC_FILT_WARP_FIXED) : (a.gring + (size_t)gwarp * L * 32);
    const int nticks = n + r4 + 3 * TC_FILT_SKEW;
    // window of local ticks in which this pass really receives a sample
    const int add_lo = pass == 3 ? r2 : 0;
    const int add_hi = pass == 0 ? n : (pass == 1 ? n + r2 : 0x7fffffff);
    // load / drain role of this lane: tick kk of a group, line qq of the warp
    const int kk = lane >> 2, qq = lane & 3;

    for (int64_t grp = gwarp; grp < ngroups; grp += nwarps) {
        const int64_t oline = grp * TC_FILT_LPW + qq;
        const bool oline_ok = oline < a.nlines;
        const int64_t oplane = oline_ok ? oline / nj : 0;
        const int64_t obase = oline_ok ? oplane * (int64_t)n * nj + (oline - oplane * nj) : 0;
        // same line in the transposed (plane, line, sample) layout
        const int64_t tbase = oline_ok ? oplane * (int64_t)n * nj + (oline - oplane * nj) * (int64_t)n : 0;
        double s = 0.0;
        float y = 0.f, yold = 0.f;    // this lane's last two outputs (the next pass reads the older one)
        float *rp = ring + lane;
        float *const rend = ring + (size_t)L * 32 + lane;
        float ra = 0.f, rb = 0.f, rc = 0.f;   // raw words of the three in-flight input groups
        unsigned fa = 1u, fb = 1u, fc = 1u;
        float wa = 0.f, wb = 0.f, wc = 0.f;
        int tb = 0;                   // input tile the current group reads

// fetch this lane's sample of the group starting at tick T0 (raw words only:
// nothing may depend on them until TC_FILT_PUBLISH)
#define TC_FILT_FETCH(R, F, W, T0)                                                     \
        {                                                                              \
            const int m_ = (T0) + kk;                                                  \
            R = 0.f; F = 1u; W = 0.f;                                                  \
            if (oline_ok && m_ < n) {                                                  \
                const int64_t idx_ = obase + (int64_t)m_ * nj;                         \
                R = a.data[idx_];                                                      \
                if (MODE_IN == FIN_MASKED) F = a.flags[a.flags_transposed ? tbase + m_ : idx_]; \
                else { F = 0u; W = a.win[idx_]; }                                      \
            }                                                                          \
        }
// turn the raw words into (value, weight) and publish them in input tile B
#define TC_FILT_PUBLISH(R, F, W, B)                                                    \
        {                                                                              \
            float v_, w_;                                                              \
            if (MODE_IN == FIN_MASKED) { v_ = F ? 0.f : R; w_ = F ? 0.f : 1.f; }       \
            else { v_ = R; w_ = W; }                                                   \
            tile[(B) * 64 + kk * 8 + qq] = v_;                                         \
            tile[(B) * 64 + kk * 8 + 4 + qq] = w_;                                     \
        }
// one tick of every lane's pass; FAST drops the warm-up / run-out predicates
#define TC_FILT_TICK(K, T0, B, FAST)                                                   \
        {                                                                              \
            const float prev_ = __shfl_up_sync(TC_FULL_MASK, TC_FILT_SKEW == 2 ? yold : y, 8); \
            float u_ = prev_;                                                          \
            if (pass == 0) u_ = tile[(B) * 64 + (K) * 8 + lane];                       \
            float old_;                                                                \
            if (FAST) {                                                                \
                old_ = *rp;                                                            \
            } else {                                                                   \
                const int m_ = (T0) + (K) - TC_FILT_SKEW * pass;                       \
                u_ = (m_ >= add_lo && m_ < add_hi) ? u_ : 0.f;                         \
                old_ = (m_ >= r2) ? *rp : 0.f;                                         \
            }                                                                          \
            *rp = u_;                                                                  \
            s += (double)u_;                                                           \
            yold = y;                                                                  \
            y = (float)s;                                                              \
            s -= (double)old_;                                                         \
            if (pass == 3) stg[(K) * 8 + (lane - 24)] = y;                             \
            rp += 32;                                                                  \
            if (rp == rend) rp = ring + lane;                                          \
        }
// one group: publish the group fetched two iterations ago, fetch the group
// three ahead, run 8 ticks, drain 32 outputs
#define TC_FILT_ITER(T0, B, RN, FN, WN, RP, FP, WP)                                    \
        {                                                                              \
            TC_FILT_PUBLISH(RP, FP, WP, (B) ^ 1)                                       \
            TC_FILT_FETCH(RN, FN, WN, (T0) + 3 * TC_FILT_U)                            \
            const int jout_ = (T0) + kk - 3 * TC_FILT_SKEW - r4;                       \
            const bool out_ok_ = jout_ >= 0 && jout_ < n && oline_ok;                  \
            const int64_t oidx_ = a.out_transposed ? tbase + (out_ok_ ? jout_ : 0)     \
                                                   : obase + (int64_t)(out_ok_ ? jout_ : 0) * nj; \
            float d2_ = 0.f;                                                           \
            if (MODE_OUT == FOUT_RESID && out_ok_) d2_ = a.data2[oidx_];               \
            __syncwarp();                                                              \
            if ((T0) - 3 * TC_FILT_SKEW >= r2 && (T0) + TC_FILT_U <= n) {              \
                _Pragma("unroll")                                                      \
                for (int k_ = 0; k_ < TC_FILT_U; k_++) TC_FILT_TICK(k_, T0, B, true)   \
            } else {                                                                   \
                _Pragma("unroll")                                                      \
                for (int k_ = 0; k_ < TC_FILT_U; k_++) TC_FILT_TICK(k_, T0, B, false)  \
            }                                                                          \
            __syncwarp();                                                              \
            if (out_ok_) {                                                             \
                const float fv_ = stg[kk * 8 + qq] / a.div, fw_ = stg[kk * 8 + 4 + qq] / a.div; \
                if (MODE_OUT == FOUT_PAIR) {                                           \
                    a.vout[oidx_] = fv_;                                               \
                    a.wout[oidx_] = fw_;                                               \
                } else {                                                               \
                    float bg_ = (fw_ == 0.f) ? NAN : fv_ / fw_;                        \
                    if (MODE_OUT == FOUT_RESID) bg_ = fabsf(d2_ - bg_);                \
                    a.vout[oidx_] = bg_;                                               \
                }                                                                      \
            }                                                                          \
        }

        // prologue: group 0 goes straight into tile 0, groups 1 and 2 stay in flight
        TC_FILT_FETCH(ra, fa, wa, 0)
        TC_FILT_PUBLISH(ra, fa, wa, 0)
        TC_FILT_FETCH(rb, fb, wb, TC_FILT_U)
        TC_FILT_FETCH(rc, fc, wc, 2 * TC_FILT_U)
        tb = 0;
        for (int t0 = 0; t0 < nticks; t0 += 3 * TC_FILT_U) {
            // (publish set, fetch set) rotate b->a, c->b, a->c; the tile alternates
            TC_FILT_ITER(t0, tb, ra, fa, wa, rb, fb, wb)
            tb ^= 1;
            if (t0 + TC_FILT_U >= nticks) break;
            TC_FILT_ITER(t0 + TC_FILT_U, tb, rb, fb, wb, rc, fc, wc)
            tb ^= 1;
            if (t0 + 2 * TC_FILT_U >= nticks) break;
            TC_FILT_ITER(t0 + 2 * TC_FILT_U, tb, rc, fc, wc, ra, fa, wa)
            tb ^= 1;
        }
        __syncwarp();
#undef TC_FILT_FETCH
#undef TC_FILT_PUBLISH
#undef TC_FILT_TICK
#undef TC_FILT_ITER
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
        TC_LAUNCH((k_box_filter<SMEM_RING, MI, MO>), grid, bd, smem, c->stream, a);            \
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
    tc_prof_begin(c, a.single_axis ? TCP_BOX_FILTER_1D : TCP_BOX_FILTER);
    const int64_t ngroups = (a.nlines + TC_FILT_LPW - 1) / TC_FILT_LPW;
    const size_t per_warp = ((size_t)2 * a.r * 32 + TC_FILT_WARP_FIXED) * sizeof(float);
    const size_t smem_cap = (size_t)c->smem_optin - 1024;
    if (per_warp <= smem_cap) {
        // warps per block: whatever packs the most warps into an SM's shared
        // memory (1 KB is reserved per block), at least 2 blocks per SM in flight
        int wpb = 1, best = 0;
        for (int w = 1; w <= 8; w++) {
            size_t need = per_warp * w + 1024;
            if (need > (size_t)c->smem_optin) break;
            int blocks = (int)((size_t)(c->smem_optin + 1024) / need);
            if (blocks > 32) blocks = 32;
            int warps = blocks * w;
            if (warps > 32) warps = 32;      // register file: 64 regs x 32 warps
            if (warps > best || (warps == best && w <= 4)) { best = warps; wpb = w; }
        }
        while (wpb > 1 && (ngroups + wpb - 1) / wpb < 2 * (int64_t)c->sm_count) wpb--;
        size_t smem = per_warp * wpb;
        unsigned grid = (unsigned)((ngroups + wpb - 1) / wpb);
        TC_TRY(launch_box_filter_mode<true>(c, a, grid, wpb * 32, smem));
    } else {
        // delay lines too deep for shared memory: persistent warps, L2-resident ring
        int wpb = 2;
        int64_t blocks = (int64_t)c->sm_count * 4;
        if (blocks > (ngroups + wpb - 1) / wpb) blocks = (ngroups + wpb - 1) / wpb;
        float *g = nullptr;
        TC_TRY(tc_alloc(c, (size_t)blocks * wpb * 2 * a.r * 32, &g));
        a.gring = g;
        a.gring_stride = 0;
        TC_TRY(launch_box_filter_mode<false>(c, a, (unsigned)blocks, wpb * 32, TC_FILT_WARP_FIXED * sizeof(float) * wpb));
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
