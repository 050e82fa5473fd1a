// k_filter5.cuh -- lane-per-chain fused box filter with the passes chained through
// shared memory ("B5"; same reference as k_filter.cuh: _box_gaussian_filter1d
// flagging.py:362-419, masked_gaussian_filter 469-513).
//
// Like k_filter2.cuh's lane-per-chain form a lane is one (stream, pass) chain,
// lane = pass * 8 + stream, and runs the reference's sequence
//     s += entering;  emit (float)s;  s -= leaving
// on its own accumulator, so results are bit-identical to every other form and to
// the oracle.  What changed is everything around those three operations (measured
// on B200: k_box4 issues 26 instructions per chain step, 7 of them are the step):
//
//  * ONE ring per chain holds the samples that enter it.  The producer -- the lane
//    of the pass before, or the input fetch for pass 0 -- writes a group of G = 16
//    samples straight into the consumer's ring one iteration ahead; the consumer
//    reads every slot twice, as the entering sample and 2r ticks later as the
//    leaving one.  No shuffles, no select on the lane's role, no separate input
//    tile, no copy of the entering samples by the consumer: three 16-byte shared
//    memory accesses per 4 ticks and lane.
//  * every lane runs the same straight-line code; the role-specific work (input
//    fetch and masking, division by d^4, value / weight, residual, stores) is
//    spread over all 32 lanes: 16 ticks x 8 streams of input are exactly one
//    16-byte global load per lane and iteration, the outputs a group finishes are
//    one (or half a) 16-byte vector per lane.
//  * float32 -> float64 widening by one IMAD.WIDE (x * 2^29: the two halves of the
//    product are the low and the high word of the double whose exponent field is
//    not rebiased, see k_filter2.cuh) and a sign mask, instead of two shifts and
//    a mask on the ALU pipe.
//  * warm-up / run-out rules (which emits the next pass must not see) are applied
//    by the producer when it stores, in the few iterations where they matter.
//
// Geometry: pass p works on local group j = g - p in iteration g; ring slot
// (i mod Lr) holds the sample with local index i, Lr = roundup(2r + 2G, G), so that
// the group the producer writes (j + 1) never aliases what the consumer still needs
// (entering group j and the leaving samples back to jG - 2r).  One __syncwarp per
// iteration orders all of it.  When 2r mod 4 == 2 the leaving samples straddle
// 16-byte vectors; the last vector is carried in registers (template ODD).
#pragma once
#include "k_filter2.cuh"

#define B5_G 16
#define B5_GQ 4
#define B5_STAGE_ROW 12     // uint4 per staging row: 8 streams + padding (rows 192 bytes apart: conflict-free 8-byte reads)

// float32 bits -> float64 with the value x * 2^-896 (exact), one IMAD.WIDE + one LOP3
__device__ __forceinline__ double b5_spread(unsigned b)
{
#ifndef TC_EMU
    int hi, lo;
    asm("{\n\t.reg .s64 w;\n\tmul.wide.s32 w, %2, 536870912;\n\tmov.b64 {%1, %0}, w;\n\t}" : "=r"(hi), "=r"(lo) : "r"(b));
    return __hiloint2double(hi & (int)0x8fffffff, lo);
#else
    return b2_spread(b);
#endif
}

template <bool INTW> struct B5Acc {
    double s;
    __device__ __forceinline__ void reset() { s = 0.0; }
    __device__ __forceinline__ void add(unsigned u) { s = __fma_rn(b5_spread(u), 0x1p896, s); }
    __device__ __forceinline__ unsigned emit() const { return __float_as_uint(__double2float_rn(s)); }
    __device__ __forceinline__ void sub(unsigned o) { s = __fma_rn(b5_spread(o), -0x1p896, s); }
};
template <> struct B5Acc<true> {
    unsigned s;
    __device__ __forceinline__ void reset() { s = 0u; }
    __device__ __forceinline__ void add(unsigned u) { s += u; }
    __device__ __forceinline__ unsigned emit() const { return s; }
    __device__ __forceinline__ void sub(unsigned o) { s -= o; }
};

// NARR 1: 8 lines of one array per warp (first axis of the 2-D masked filter, the
//         value and the integer weight chains in different warps; pair output)
// NARR 2: 4 lines x (value, weight) per warp (stream = array * 4 + line)
template <int NARR, bool INTW, bool ODD, int MODE_IN, int MODE_OUT>
__device__ __forceinline__ void b5_line_group(const FilterArgs &a, uint4 *wsm, int64_t grp, int lane)
{
    constexpr int G = B5_G, GQ = B5_GQ;
    constexpr int NL = 8 / NARR;
    const int pass = lane >> 3, sidx = lane & 7;
    const int n = a.n, r2 = 2 * a.r, r4 = 4 * a.r;
    const int Lr = (r2 + 3 * G - 1) / G * G, nvec = Lr >> 2;
    const int64_t nj = a.nj;
    uint4 *ring = wsm + lane;                           // vector v of this lane's chain: ring[v * 32]
    uint4 *stage = wsm + (size_t)nvec * 32;             // [2 * GQ rows][B5_STAGE_ROW]: the emits of pass 3, double buffered
    uint4 *ydst = pass < 3 ? ring + 8 : stage + sidx;   // where this lane's emits go: the next pass's ring / the staging rows
    const int ystride = pass < 3 ? 32 : B5_STAGE_ROW;
    // the emits the next pass may see: local indices [ylo, yhi)
    const int ylo = pass == 2 ? r2 : -0x40000000, yhi = pass == 0 ? n + r2 : 0x7fffffff;

    // ---- input role.  Loading side: lane = ls * 4 + lc fetches the 16-byte chunk lc of stream ls, so the four
    // chunks of a stream's group are 64 contiguous bytes fetched by four neighbouring lanes (two cache lines per
    // quarter warp; with lane = chunk * 8 + stream every lane of a quarter warp touched another line: 32 LSU
    // wavefronts per load instead of 8, on a kernel that runs at ~80 % of the LSU data pipe).  Publishing side:
    // lane = fc * 8 + fl stores chunk fc of stream fl into the ring of the pass-0 lane fl (conflict-free: the
    // eight lanes of a quarter warp hit eight different bank groups), after taking it from lane fl * 4 + fc.
    const int ls = lane >> 2, lc = lane & 3;
    const int fl = lane & 7, fc = lane >> 3;
    const int psrc = fl * 4 + fc;
    const int64_t fline = grp * NL + (NARR == 1 ? ls : (ls & 3));
    const bool fok = fline < a.nlines;
    const int64_t fbase = fok ? fline * (int64_t)n : 0;
    const bool lweight = NARR == 2 && ls >= 4;          // the loaded stream is the weight array
    const bool fweight = NARR == 2 && fl >= 4;          // the published stream is the weight array
    const float *fsrc = (MODE_IN == FIN_PAIR && lweight) ? a.win : a.data;
    // two register sets in rotation: a group is fetched two iterations before it is published (one iteration
    // ahead the publish still waited on the load for a tenth of the first-axis kernel's stall samples)
    float4 fqa = make_float4(0.f, 0.f, 0.f, 0.f), fqb = fqa;
    unsigned fga = 0x01010101u, fgb = 0x01010101u;
    // running pointers: group after group of this lane's chunk (recomputing the addresses from the block
    // index every iteration cost ~30 instructions per iteration in the compiled loop)
    const float *fpd = fsrc + fbase + 4 * lc;
    const u8 *fpg = (MODE_IN == FIN_PAIR ? nullptr : a.flags + fbase + 4 * lc);
    int fm = 4 * lc;
    const int fend = fok ? n : 0;
    auto fetch = [&](float4 &fq, unsigned &fg) {
        fq = make_float4(0.f, 0.f, 0.f, 0.f);
        fg = 0x01010101u;
        if (fm < fend) {
            if (MODE_IN == FIN_PAIR) {
                fq = *reinterpret_cast<const float4 *>(fpd);
            } else {
                if (!INTW) fq = *reinterpret_cast<const float4 *>(fpd);
                fg = *reinterpret_cast<const unsigned *>(fpg);
            }
        }
        fm += G;
        fpd += G;
        if (MODE_IN != FIN_PAIR) fpg += G;
    };
    auto publish = [&](int vbase, const float4 &lq, unsigned lg) {
        // loading lane -> publishing lane
        float4 fq = lq;
        unsigned fg = lg;
        if (MODE_IN == FIN_PAIR || !INTW) {
            fq.x = __shfl_sync(TC_FULL_MASK, lq.x, psrc);
            fq.y = __shfl_sync(TC_FULL_MASK, lq.y, psrc);
            fq.z = __shfl_sync(TC_FULL_MASK, lq.z, psrc);
            fq.w = __shfl_sync(TC_FULL_MASK, lq.w, psrc);
        }
        if (MODE_IN != FIN_PAIR) fg = __shfl_sync(TC_FULL_MASK, lg, psrc);
        uint4 o;
        if (MODE_IN == FIN_PAIR) {
            o = make_uint4(__float_as_uint(fq.x), __float_as_uint(fq.y), __float_as_uint(fq.z), __float_as_uint(fq.w));
        } else if (INTW) {
            o = make_uint4((fg & 0xffu) ? 0u : 1u, (fg & 0xff00u) ? 0u : 1u, (fg & 0xff0000u) ? 0u : 1u,
                           (fg & 0xff000000u) ? 0u : 1u);
        } else {
            const unsigned one = 0x3f800000u;
            o.x = (fg & 0xffu) ? 0u : (fweight ? one : __float_as_uint(fq.x));
            o.y = (fg & 0xff00u) ? 0u : (fweight ? one : __float_as_uint(fq.y));
            o.z = (fg & 0xff0000u) ? 0u : (fweight ? one : __float_as_uint(fq.z));
            o.w = (fg & 0xff000000u) ? 0u : (fweight ? one : __float_as_uint(fq.w));
        }
        wsm[(vbase + fc) * 32 + fl] = o;                 // ring of the pass-0 lane of stream fl
    };

    // ---- drain role
    // NARR 1: vector dq (4 ticks) of line dl;  NARR 2: half a vector (2 ticks) of line dl, value and weight.
    // Iteration g drains local group g - 4 of pass 3, whose first tick finishes sample (g - 4) G - 4r: every
    // lane keeps a sample counter and running output pointers that advance by one group per iteration.
    const int dl = NARR == 1 ? (lane & 7) : (lane & 3);
    const int dq = NARR == 1 ? (lane >> 3) : ((lane >> 3) & 3);
    const int dh = NARR == 1 ? 0 : ((lane >> 2) & 1);
    const int64_t dline = grp * NL + dl;
    const bool dok = dline < a.nlines;
    const int64_t dplane = dok ? dline / nj : 0;
    const int64_t dlc = dok ? dline * (int64_t)n : 0;                                     // line-contiguous base
    const int64_t dsm = dok ? dplane * (int64_t)n * nj + (dline - dplane * nj) : 0;       // sample-major base
    const int64_t omul = a.out_transposed ? 1 : nj;                                       // output stride of a sample
    int js = -4 * G + 4 * dq + 2 * dh - r4;
    const int64_t ooff = (a.out_transposed ? dlc : dsm) + (int64_t)js * omul;
    float *pv = (NARR == 1 && INTW ? a.wout : a.vout) + ooff;
    float *pw = (NARR == 2 && MODE_OUT == FOUT_PAIR) ? a.wout + ooff : nullptr;
    const float *pd2 = MODE_OUT == FOUT_RESID ? a.data2 + dlc + js : nullptr;             // unfiltered samples, line-contiguous
    const int64_t ostep = (int64_t)G * omul;
    const unsigned jsmax = dok ? (unsigned)n : 0u;
    B2Div dv;
    dv.init(a.div);
    // The drain is split so that its arithmetic is straight-line code the compiler can interleave with the
    // chain steps: drain_math runs the call-free division sequences on every lane and records whether an
    // operand was outside the range where they are exact; drain_store redoes those lanes with the plain
    // divisions (one warp vote; never taken on amplitude data) and stores.
    // exact ranges: B2Div::fast for |x| in [2^-87, 2^123) or 0; b2_div_fast for quotients by d^4 in [2^-60, 2^60)
    // (the value may be 0) -- which also pins the undivided sums inside the first range.
    float o[4];
    unsigned yraw[4];
    bool dbad = false;
    // the unfiltered samples the residual needs are fetched at the top of the iteration that stores them:
    // loaded where they are used they cost every warp one exposed DRAM round trip per iteration (ncu source
    // view: a sixth of the stall samples of a lone warp sat on that one FADD)
    float2 d2v = make_float2(0.f, 0.f);
    auto d2_prefetch = [&]() {
        if (MODE_OUT == FOUT_RESID && (unsigned)js < jsmax) d2v = *reinterpret_cast<const float2 *>(pd2);
    };
    auto in_q_range = [](float q) { return __float_as_uint(q) - ((127u - 60u) << 23) < (120u << 23); };
    auto drain_math = [&](int buf) {
        const uint4 *row = stage + (buf * GQ + dq) * B5_STAGE_ROW;
        if (NARR == 1) {
            const uint4 v = row[dl];
            yraw[0] = v.x; yraw[1] = v.y; yraw[2] = v.z; yraw[3] = v.w;
            dbad = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (INTW) {
                    o[k] = dv.of_count(yraw[k]);
                } else {
                    o[k] = dv.fast(__uint_as_float(yraw[k]));
                    dbad = dbad || !(yraw[k] == 0u || yraw[k] - (40u << 23) < (210u << 23));
                }
            }
        } else {
            const uint2 v = reinterpret_cast<const uint2 *>(row + dl)[dh];
            const uint2 w = reinterpret_cast<const uint2 *>(row + 4 + dl)[dh];
            yraw[0] = v.x; yraw[1] = v.y; yraw[2] = w.x; yraw[3] = w.y;
            dbad = false;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float fv = dv.fast(__uint_as_float(yraw[k]));
                const float fw = dv.fast(__uint_as_float(yraw[2 + k]));
                dbad = dbad || !((fv == 0.f && yraw[k] == 0u) || in_q_range(fv)) || !((fw == 0.f && yraw[2 + k] == 0u) || in_q_range(fw));
                if (MODE_OUT == FOUT_PAIR) {
                    o[k] = fv;
                    o[2 + k] = fw;
                } else {
                    o[k] = (fw == 0.f) ? NAN : b2_div_fast(fv, fw);
                }
            }
        }
    };
    auto drain_store = [&]() {
        const bool valid = (unsigned)js < jsmax;
        if (!INTW && __ballot_sync(TC_FULL_MASK, dbad && valid) != 0u) {
            if (dbad) {
                if (NARR == 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++) o[k] = dv(__uint_as_float(yraw[k]));
                } else {
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const float fv = dv(__uint_as_float(yraw[k]));
                        const float fw = dv(__uint_as_float(yraw[2 + k]));
                        if (MODE_OUT == FOUT_PAIR) {
                            o[k] = fv;
                            o[2 + k] = fw;
                        } else {
                            o[k] = (fw == 0.f) ? NAN : fv / fw;
                        }
                    }
                }
            }
        }
        if (valid) {
            if (NARR == 1) {
#pragma unroll
                for (int k = 0; k < 4; k++) pv[(int64_t)k * omul] = o[k];
            } else if (MODE_OUT == FOUT_PAIR) {
                pv[0] = o[0]; pv[omul] = o[1];
                pw[0] = o[2]; pw[omul] = o[3];
            } else {
                if (MODE_OUT == FOUT_RESID) {
                    o[0] = fabsf(d2v.x - o[0]);
                    o[1] = fabsf(d2v.y - o[1]);
                }
                pv[0] = o[0];
                pv[omul] = o[1];
            }
        }
        js += G;
        pv += ostep;
        if (NARR == 2 && MODE_OUT == FOUT_PAIR) pw += ostep;
        if (MODE_OUT == FOUT_RESID) pd2 += G;
    };

    // ---- chain state
    B5Acc<INTW> acc;
    acc.reset();
    uint4 car = make_uint4(0u, 0u, 0u, 0u);
    const int ngr = nvec / GQ;
    int ev = ((ngr - pass % ngr) % ngr) * GQ;            // vector base of the entering group (local group -pass)
    int lv;                                              // first vector the next leaving group loads
    {
        int le = (-pass * G - r2) % Lr;
        if (le < 0) le += Lr;
        lv = ((le + (ODD ? 2 : 0)) >> 2) % nvec;
    }
    for (int v = 0; v < nvec; v++) ring[v * 32] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    fetch(fqa, fga);
    publish(0, fqa, fga);
    fetch(fqa, fga);
    fetch(fqb, fgb);
    __syncwarp();

    const int niter = (n + r4 + G - 1) / G + 3;          // pass 3 finishes local group niter - 4 in the last iteration
    int pubv = GQ % nvec;
    // the leaving samples of the next iteration are loaded one iteration ahead whenever they are all older
    // than the group the producer is writing (2r >= G); the entering ones only exist after the __syncwarp
    const bool ahead = r2 >= G;
    auto load_leaving = [&](uint4 *dst) {
        int rq = lv;
#pragma unroll
        for (int q = 0; q < GQ; q++) {
            dst[q] = ring[rq * 32];
            rq++; if (rq == nvec) rq = 0;
        }
        lv = rq;
    };
    // one iteration; `cur` holds the leaving samples of this iteration when they were loaded ahead and
    // `nx` receives those of the next one.  The loop below is unrolled by two with the roles of the two
    // buffers swapped, so that no registers are copied between iterations (32 moves per iteration in the
    // rolled form: a tenth of the loop, ncu source view).
    auto body = [&](int g, uint4 *cur, uint4 *nx, float4 &fq, unsigned &fg) {
        d2_prefetch();
        publish(pubv, fq, fg);                           // group g + 1 of the input
        pubv += GQ; if (pubv == nvec) pubv = 0;
        fetch(fq, fg);                                   // group g + 3

        uint4 e[GQ];
#pragma unroll
        for (int q = 0; q < GQ; q++) e[q] = ring[(ev + q) * 32];
        if (ahead) load_leaving(nx);
        else load_leaving(cur);
        drain_math((g - 1) & 1);                         // what pass 3 staged in the previous iteration
        unsigned in[G], old[G], y[G];
#pragma unroll
        for (int q = 0; q < GQ; q++) {
            in[4 * q] = e[q].x; in[4 * q + 1] = e[q].y; in[4 * q + 2] = e[q].z; in[4 * q + 3] = e[q].w;
        }
        if (ODD) {
            old[0] = car.z; old[1] = car.w;
#pragma unroll
            for (int q = 0; q < GQ; q++) {
                old[4 * q + 2] = cur[q].x; old[4 * q + 3] = cur[q].y;
                if (q + 1 < GQ) { old[4 * q + 4] = cur[q].z; old[4 * q + 5] = cur[q].w; }
            }
            car.z = cur[GQ - 1].z; car.w = cur[GQ - 1].w;
        } else {
#pragma unroll
            for (int q = 0; q < GQ; q++) {
                old[4 * q] = cur[q].x; old[4 * q + 1] = cur[q].y; old[4 * q + 2] = cur[q].z; old[4 * q + 3] = cur[q].w;
            }
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            acc.add(in[k]);
            y[k] = acc.emit();
            acc.sub(old[k]);
        }
        const int i0 = (g - pass) * G;
        if (i0 < ylo || i0 + G > yhi) {
#pragma unroll
            for (int k = 0; k < G; k++)
                if (i0 + k < ylo || i0 + k >= yhi) y[k] = 0u;
        }
        const int yv = pass < 3 ? ev : (g & 1) * GQ;
#pragma unroll
        for (int q = 0; q < GQ; q++)
            ydst[(yv + q) * ystride] = make_uint4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        ev += GQ; if (ev == nvec) ev = 0;
        drain_store();
        __syncwarp();
    };
    uint4 bufa[GQ], bufb[GQ];
    if (ahead) load_leaving(bufa);
    int g = 0;
    for (; g + 1 < niter; g += 2) {
        body(g, bufa, bufb, fqa, fga);
        body(g + 1, bufb, bufa, fqb, fgb);
    }
    if (g < niter) body(g, bufa, bufb, fqa, fga);
    d2_prefetch();
    drain_math((niter - 1) & 1);
    drain_store();
}

// first axis of the 2-D masked filter: even blocks filter the values (float64 chains)
// into vout, odd blocks the weights (uint32 chains) into wout (a.role as in k_box8)
template <bool ODD>
__global__ void k_box5a(FilterArgs a)
{
    TC_DYN_SMEM(uint4, smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int Lr = (2 * a.r + 3 * B5_G - 1) / B5_G * B5_G;
    uint4 *wsm = smem + (size_t)wib * ((size_t)Lr * 8 + 2 * B5_GQ * B5_STAGE_ROW);
    const int64_t ngroups = (a.nlines + 7) / 8;
    const int64_t grp = (int64_t)(a.role ? blockIdx.x : blockIdx.x >> 1) * nwb + wib;
    if (grp >= ngroups) return;
    const bool weights = a.role ? a.role == 2 : (blockIdx.x & 1) != 0;
    if (weights) b5_line_group<1, true, ODD, FIN_MASKED, FOUT_PAIR>(a, wsm, grp, lane);
    else b5_line_group<1, false, ODD, FIN_MASKED, FOUT_PAIR>(a, wsm, grp, lane);
}

// value and weight arrays of 4 lines in one warp, every in / out mode
template <bool ODD, int MODE_IN, int MODE_OUT>
__global__ void k_box5b(FilterArgs a)
{
    TC_DYN_SMEM(uint4, smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int Lr = (2 * a.r + 3 * B5_G - 1) / B5_G * B5_G;
    uint4 *wsm = smem + (size_t)wib * ((size_t)Lr * 8 + 2 * B5_GQ * B5_STAGE_ROW);
    const int64_t ngroups = (a.nlines + 3) / 4;
    const int64_t grp = (int64_t)blockIdx.x * nwb + wib;
    if (grp >= ngroups) return;
    b5_line_group<2, false, ODD, MODE_IN, MODE_OUT>(a, wsm, grp, lane);
}

// ---------------------------------------------------------------- launching ----
static size_t b5_per_warp(int r)
{
    const int Lr = (2 * r + 3 * B5_G - 1) / B5_G * B5_G;
    return ((size_t)Lr * 8 + 2 * B5_GQ * B5_STAGE_ROW) * sizeof(uint4);
}

// same contract as launch_box_filter2: every input array line-contiguous, lines a
// multiple of 4 samples long and 16-byte aligned
static bool b5_supported(tc_context *c, const FilterArgs &a)
{
    if (a.r < 1 || (a.n & 3) || TC_ENV_FLAG("TC_FILTER_NO_B5") || TC_ENV_FLAG("TC_FILTER_OLD")) return false;
    if ((((uintptr_t)a.data | (uintptr_t)a.win | (uintptr_t)a.flags) & 15) != 0) return false;
    return b5_per_warp(a.r) + 1024 <= (size_t)c->smem_optin;
}

static int launch_box_filter5(tc_context *c, FilterArgs a)
{
    if (a.nlines == 0 || a.n == 0) return TC_OK;
    TC_REQUIRE(b5_supported(c, a), "internal: B5 filter launched on an unsupported shape");
    a.div = tc_f32_pow4(2 * (int64_t)a.r + 1);
    if (TC_ENV_FLAG("TC_FILTER_TRACE"))
        fprintf(stderr, "b5 filter: n=%d nj=%d r=%d in=%d out=%d tr=%d\n", a.n, a.nj, a.r, a.mode_in, a.mode_out,
                a.out_transposed);
    const bool odd = (a.r & 1) != 0;
    const size_t per_warp = b5_per_warp(a.r);
    const bool split = a.mode_in == FIN_MASKED && a.mode_out == FOUT_PAIR && a.r <= B2_INTW_MAX_R &&
                       !TC_ENV_FLAG("TC_FILTER_NO_INTW");
    tc_prof_begin(c, split ? TCP_BOX_FILTER8 : (a.single_axis ? TCP_BOX_FILTER_1D : TCP_BOX_FILTER));
    if (split) {
        const int64_t ngroups = (a.nlines + 7) / 8;
        const int wpb = b2_warps_per_block(c, per_warp, (a.role ? 1 : 2) * ngroups, 24);
        const unsigned grid = (unsigned)((a.role ? 1 : 2) * ((ngroups + wpb - 1) / wpb));
        if (odd) TC_TRY(b2_launch(c, k_box5a<true>, a, grid, wpb, per_warp * wpb));
        else TC_TRY(b2_launch(c, k_box5a<false>, a, grid, wpb, per_warp * wpb));
    } else {
        const int64_t ngroups = (a.nlines + 3) / 4;
        const int wpb = b2_warps_per_block(c, per_warp, ngroups, 24);
        const unsigned grid = (unsigned)((ngroups + wpb - 1) / wpb);
        const size_t smem = per_warp * wpb;
#define B5_CASE(MI, MO)                                                                  \
        if (a.mode_in == MI && a.mode_out == MO) {                                       \
            if (odd) TC_TRY(b2_launch(c, k_box5b<true, MI, MO>, a, grid, wpb, smem));     \
            else TC_TRY(b2_launch(c, k_box5b<false, MI, MO>, a, grid, wpb, smem));        \
        }
        B5_CASE(FIN_MASKED, FOUT_PAIR)
        B5_CASE(FIN_MASKED, FOUT_BG)
        B5_CASE(FIN_MASKED, FOUT_RESID)
        B5_CASE(FIN_PAIR, FOUT_PAIR)
        B5_CASE(FIN_PAIR, FOUT_BG)
        B5_CASE(FIN_PAIR, FOUT_RESID)
#undef B5_CASE
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
