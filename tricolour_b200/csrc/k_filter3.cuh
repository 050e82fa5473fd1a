// k_filter3.cuh -- thread-per-line fused box filter with the passes skewed by one
// tick IN REGISTERS ("TPL"; same reference as k_filter.cuh: _box_gaussian_filter1d
// flagging.py:362-419, masked_gaussian_filter 469-513).
//
// One thread runs all four box passes of a line (one array, or the value and the
// weight array together).  Pass p+1 consumes what pass p emitted in the PREVIOUS
// tick, so within a tick the 4 (or 8) running sums of a thread are independent
// dependency chains: the float64 add -> round -> widen latency of one pass no longer
// sits in front of the next pass, and a warp keeps the issue slots busy on its own.
// That is what lets the kernel run at the low occupancy the delay lines force (one
// warp per scheduler) -- the unskewed thread-per-line form of k_filter2.cuh needed
// three to four times as many resident warps -- and so covers radii up to ~54.
// No shuffles, no per-lane roles, fully coalesced sample-major global accesses.
//
// Per accumulator the order of floating point operations is the reference's
//     s += entering;  emit (float)s;  s -= leaving
// with the integer-pipe widening of k_filter2.cuh (B2Acc), so results are
// bit-identical to the other forms and to the oracle.
//
// Tick bookkeeping (t = 0 .. n + 4r + 2, groups of 4 ticks; P1..P4 are the padded
// arrays of the reference after each pass):
//   pass 0 enters x[t] (zero for t >= n)              emits P1[2r + t]
//   pass 1 enters pass 0's previous emit, while t - 1 < n + 2r   emits P2[t - 1]
//   pass 2 enters pass 1's previous emit                          emits P3[t - 2 - 2r]
//   pass 3 enters pass 2's previous emit, once t - 3 >= 2r        emits P4[t - 3 - 4r]
// A sample leaves a running sum 2r ticks after it entered: delay lines of
// Lp = roundup(2r, 4) slots per chain in shared memory ([chain][vector][lane], 16-byte
// accesses, conflict free), written and read once per group exactly like the rings of
// k_filter2.cuh's thread-per-line kernel (template ODD when Lp - 2r == 2).
//
//   k_box_tpl_a   first axis of the 2-D masked filter: masked input (samples
//                 sample-major, flags line-contiguous), value blocks (float64 chains)
//                 and weight blocks (uint32 chains) alternate; output pair sample-major
//                 (or line-contiguous when out_transposed).
//   k_box_tpl_b   second axis: the (value, weight) pair of the first axis, stored
//                 line-contiguous for this axis, is fetched by the warp in 16-sample
//                 tiles (16-byte loads, 64 contiguous bytes per line) and turned through
//                 a shared tile into per-thread samples; 8 float64 chains per thread;
//                 the drain divides by d^4, forms value / weight (NaN where the weight
//                 is zero), optionally |data - background|, and writes sample-major.
#pragma once
#include "k_filter2.cuh"

#define TPL_TILE_ROW 5      // uint4 per tile row: 16 samples + 4 words of padding (conflict-free 16-byte reads)

template <int NARR, bool INTW, bool ODD, int MODE_IN, int MODE_OUT>
__device__ __forceinline__ void tpl_lines(const FilterArgs &a, uint4 *wring, uint4 *wtile, int64_t line0, int lane)
{
    constexpr int NCH = 4 * NARR;
    const int n = a.n, r2 = 2 * a.r, r4 = 4 * a.r;
    const int Lp = (r2 + 3) & ~3, nvec = Lp >> 2;
    const int64_t nj = a.nj;
    const int nticks = n + r4 + 3;
    int64_t line = line0 + lane;
    const bool lok = line < a.nlines;
    if (!lok) line = a.nlines - 1;        // a redundant copy of the last line, never stored
    const int64_t plane = line / nj;
    const int64_t sm_base = plane * (int64_t)n * nj + (line - plane * nj);   // sample-major: + i * nj
    const int64_t lc_base = line * (int64_t)n;                               // line-contiguous: + i
    uint4 *ring = wring + lane;           // vector v of chain c: ring[(v * NCH + c) * 32] (a group's accesses: constant offsets)
    B2Div dv;
    dv.init(a.div);

    B2Acc<INTW> acc[NCH];
    unsigned yprev[NCH], old[NCH][4];
    uint4 car[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        acc[c].reset();
        yprev[c] = 0u;
        car[c] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < 4; k++) old[c][k] = 0u;
    }
    for (int v = 0; v < NCH * nvec; v++) ring[v * 32] = make_uint4(0u, 0u, 0u, 0u);
    int wv = 0, rv = (ODD ? 2 : 1) % nvec;

    // ---- input pipelines
    // masked input: samples three groups ahead, flags one 16-tick block ahead
    float xa[4], xb[4], xc[4];
    uint4 fcur = make_uint4(0u, 0u, 0u, 0u), fnxt = fcur;
    const float *xp = a.data + sm_base;
    const u8 *fp = MODE_IN == FIN_MASKED ? a.flags + lc_base : nullptr;
    auto load_x = [&](float *x, int t0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int m = t0 + k;
            m = m < n ? m : n - 1;                     // past the end: any valid sample, masked below
            x[k] = INTW ? 0.f : xp[(int64_t)m * nj];
        }
    };
    auto load_f = [&](int t0) {
        uint4 f = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
        if (t0 < n) f = *reinterpret_cast<const uint4 *>(fp + t0);
        return f;
    };
    // pair input: the warp's next 16-sample tile of both arrays, in registers until it is parked
    const int fl = lane >> 2, fc = lane & 3;
    uint4 pv[4], pw[4];
    int64_t tl[4];
    auto fetch_tile = [&](int s0) {
        const int m = s0 + 4 * fc;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            pv[i] = make_uint4(0u, 0u, 0u, 0u);
            pw[i] = pv[i];
            if (m < n) {
                pv[i] = *reinterpret_cast<const uint4 *>(a.data + tl[i] + m);
                pw[i] = *reinterpret_cast<const uint4 *>(a.win + tl[i] + m);
            }
        }
    };
    if (MODE_IN == FIN_MASKED) {
        load_x(xa, 0);
        load_x(xb, 4);
        load_x(xc, 8);
        fcur = load_f(0);
        fnxt = load_f(16);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int64_t l = line0 + fl + 8 * i;
            l = l < a.nlines ? l : a.nlines - 1;
            tl[i] = l * (int64_t)n;
        }
        fetch_tile(0);
    }

    // ---- output addressing: sample j of this line at obase + j * omul
    const int64_t obase = a.out_transposed ? lc_base : sm_base;
    const int64_t omul = a.out_transposed ? 1 : nj;
    int64_t ooff = obase + (int64_t)(-r4 - 3) * omul;      // sample of tick t0 + k: ooff + k * omul
    float d2[4] = {0.f, 0.f, 0.f, 0.f};

    const int ngroups = (nticks + 3) >> 2;
    for (int g = 0; g < ngroups; g++) {
        const int t0 = g * 4;
        // ---- this group's samples; refill the pipelines
        unsigned u0[NARR][4];
        if (MODE_IN == FIN_MASKED) {
            const unsigned fw = fcur.x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool flg = ((fw >> (8 * k)) & 0xffu) != 0u || t0 + k >= n;
                u0[0][k] = INTW ? (flg ? 0u : 1u) : (flg ? 0u : __float_as_uint(xa[k]));
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { xa[k] = xb[k]; xb[k] = xc[k]; }
            load_x(xc, t0 + 12);
            if ((g & 3) == 3) { fcur = fnxt; fnxt = load_f(t0 + 20); }
            else { fcur.x = fcur.y; fcur.y = fcur.z; fcur.z = fcur.w; }
        } else {
            if ((g & 3) == 0) {
                // park the fetched tile (rows of 16 samples + padding), start fetching the next one
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    wtile[(fl + 8 * i) * TPL_TILE_ROW + fc] = pv[i];
                    wtile[(32 + fl + 8 * i) * TPL_TILE_ROW + fc] = pw[i];
                }
                __syncwarp();
                fetch_tile(t0 + 16);
            }
            const uint4 xv = wtile[lane * TPL_TILE_ROW + (g & 3)];
            const uint4 xw = wtile[(32 + lane) * TPL_TILE_ROW + (g & 3)];
            u0[0][0] = xv.x; u0[0][1] = xv.y; u0[0][2] = xv.z; u0[0][3] = xv.w;
            u0[NARR - 1][0] = xw.x; u0[NARR - 1][1] = xw.y; u0[NARR - 1][2] = xw.z; u0[NARR - 1][3] = xw.w;
        }
        // residual mode: the unfiltered samples of the outputs this group produces
        if (MODE_OUT == FOUT_RESID) {
            const int j0 = t0 - r4 - 3;
            if (j0 >= 0 && j0 + 3 < n) {
#pragma unroll
                for (int k = 0; k < 4; k++) d2[k] = a.data2[ooff + (int64_t)k * omul];
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((unsigned)(j0 + k) < (unsigned)n) d2[k] = a.data2[ooff + (int64_t)k * omul];
            }
        }

        // ---- four ticks, every chain one step per tick; passes in descending order so that
        // a pass reads its predecessor's emit of the previous tick before it is replaced
        unsigned un[NCH][4], y3[NARR][4];
        const bool fast = t0 >= r2 + 3 && t0 + 2 < n + r2;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int ar = 0; ar < NARR; ar++) {
#pragma unroll
                for (int p = 3; p >= 0; p--) {
                    const int c = ar * 4 + p;
                    unsigned u = p == 0 ? u0[ar][k] : yprev[c - 1];
                    if (!fast) {
                        if (p == 1 && t0 + k - 1 >= n + r2) u = 0u;
                        if (p == 3 && t0 + k - 3 < r2) u = 0u;
                    }
                    un[c][k] = u;
                    acc[c].add(u);
                    yprev[c] = acc[c].emit();
                    acc[c].sub(old[c][k]);
                }
                y3[ar][k] = yprev[ar * 4 + 3];
            }
        }
        // ---- delay lines: park this group's samples, pick up the ones that leave next group
        {
            uint4 *rw = ring + (size_t)wv * (NCH * 32);
            const uint4 *rr = ring + (size_t)rv * (NCH * 32);
#pragma unroll
            for (int c = 0; c < NCH; c++) rw[c * 32] = make_uint4(un[c][0], un[c][1], un[c][2], un[c][3]);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint4 nw = rr[c * 32];
                if (ODD) {
                    old[c][0] = car[c].z; old[c][1] = car[c].w; old[c][2] = nw.x; old[c][3] = nw.y;
                    car[c] = nw;
                } else {
                    old[c][0] = nw.x; old[c][1] = nw.y; old[c][2] = nw.z; old[c][3] = nw.w;
                }
            }
        }
        wv++; if (wv == nvec) wv = 0;
        rv++; if (rv == nvec) rv = 0;

        // ---- outputs: tick t0 + k finishes sample j = t0 + k - 4r - 3
        const int j0 = t0 - r4 - 3;
        if (lok && j0 + 3 >= 0 && j0 < n) {
            if (MODE_OUT == FOUT_PAIR) {
                float *out = INTW ? a.wout : a.vout;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if ((unsigned)(j0 + k) >= (unsigned)n) continue;
                    out[ooff + (int64_t)k * omul] = INTW ? dv.of_count(y3[0][k]) : dv(__uint_as_float(y3[0][k]));
                }
            } else {
                // value / d^4, weight / d^4 and their quotient, each correctly rounded.  All three
                // take the call-free correction sequences when the four sums of the group lie in
                // [2^-40, 2^50) or are zero -- no intermediate can then leave the normal range --
                // and the whole group falls back to the plain divisions otherwise (one uniform
                // test per group instead of three data-dependent branches per sample).
                float res[4];
                bool safe = true;
#pragma unroll
                for (int k = 0; k < 4; k++) safe = safe && dv.safe(y3[0][k]) && dv.safe(y3[NARR - 1][k]);
                if (safe) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float fv = dv.fast(__uint_as_float(y3[0][k]));
                        const float fw = dv.fast(__uint_as_float(y3[NARR - 1][k]));
                        res[k] = (fw == 0.f) ? NAN : b2_div_fast(fv, fw);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float fv = dv(__uint_as_float(y3[0][k]));
                        const float fw = dv(__uint_as_float(y3[NARR - 1][k]));
                        res[k] = (fw == 0.f) ? NAN : fv / fw;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if ((unsigned)(j0 + k) >= (unsigned)n) continue;
                    float bg = res[k];
                    if (MODE_OUT == FOUT_RESID) bg = fabsf(d2[k] - bg);
                    a.vout[ooff + (int64_t)k * omul] = bg;
                }
            }
        }
        ooff += 4 * omul;
    }
}

// first axis: even blocks run the value chains, odd blocks the weight chains (a.role as in k_box8)
template <bool ODD>
__global__ void k_box_tpl_a(FilterArgs a)
{
    TC_DYN_SMEM(uint4, smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int Lp = (2 * a.r + 3) & ~3;
    uint4 *wring = smem + (size_t)wib * Lp * 32;            // 4 chains x Lp / 4 vectors x 32 lanes
    const int64_t line0 = ((int64_t)(a.role ? blockIdx.x : blockIdx.x >> 1) * nwb + wib) * 32;
    if (line0 >= a.nlines) return;
    const bool weights = a.role ? a.role == 2 : (blockIdx.x & 1) != 0;
    if (weights) tpl_lines<1, true, ODD, FIN_MASKED, FOUT_PAIR>(a, wring, nullptr, line0, lane);
    else tpl_lines<1, false, ODD, FIN_MASKED, FOUT_PAIR>(a, wring, nullptr, line0, lane);
}

// second axis: value and weight chains of a line in one thread
template <bool ODD, int MODE_OUT>
__global__ void k_box_tpl_b(FilterArgs a)
{
    TC_DYN_SMEM(uint4, smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int Lp = (2 * a.r + 3) & ~3;
    const size_t per_warp = (size_t)2 * Lp * 32 + 64 * TPL_TILE_ROW;   // uint4: 8 chains x Lp / 4 x 32 + the two tiles
    uint4 *wring = smem + (size_t)wib * per_warp;
    uint4 *wtile = wring + (size_t)2 * Lp * 32;
    const int64_t line0 = ((int64_t)blockIdx.x * nwb + wib) * 32;
    if (line0 >= a.nlines) return;
    tpl_lines<2, false, ODD, FIN_PAIR, MODE_OUT>(a, wring, wtile, line0, lane);
}

// ---------------------------------------------------------------- launching ----
// smallest number of resident warps per SM the kernels are launched with (one per scheduler)
#define TPL_MIN_WARPS_SM 4

static size_t tpl_a_per_warp(int r) { return (size_t)((2 * r + 3) & ~3) * 32 * sizeof(uint4); }
static size_t tpl_b_per_warp(int r)
{
    return ((size_t)2 * ((2 * r + 3) & ~3) * 32 + 64 * TPL_TILE_ROW) * sizeof(uint4);
}

static int tpl_env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static bool tpl_a_supported(tc_context *c, const FilterArgs &a)
{
    static const int max_r = tpl_env_int("TC_TPL_A_MAXR", 54);
    // measured slower than the k_filter2.cuh forms at every radius of default.yaml
    // (profiles/r02_filter_ab.txt): an experiment knob, off by default
    if (!TC_ENV_FLAG("TC_TPL_A") || TC_ENV_FLAG("TC_FILTER_OLD")) return false;
    if (a.r < 2 || a.r > max_r || (a.n & 15)) return false;
    if (a.mode_in != FIN_MASKED || a.mode_out != FOUT_PAIR || !a.flags_transposed) return false;
    return (tpl_a_per_warp(a.r) + 1024) * TPL_MIN_WARPS_SM <= (size_t)c->smem_optin + 1024;
}

static bool tpl_b_supported(tc_context *c, const FilterArgs &a)
{
    static const int max_r = tpl_env_int("TC_TPL_B_MAXR", 34);
    static const int min_warps = tpl_env_int("TC_TPL_B_MINW", 3);
    if (!TC_ENV_FLAG("TC_TPL_B") || TC_ENV_FLAG("TC_FILTER_OLD")) return false;
    if (a.r < 2 || a.r > max_r || (a.n & 3)) return false;
    if (a.mode_in != FIN_PAIR || (a.mode_out != FOUT_BG && a.mode_out != FOUT_RESID)) return false;
    if ((((uintptr_t)a.data | (uintptr_t)a.win) & 15) != 0) return false;
    return (tpl_b_per_warp(a.r) + 1024) * (size_t)min_warps <= (size_t)c->smem_optin + 1024;
}

// data: sample-major; flags: line-contiguous; outputs per out_transposed
static int launch_box_tpl_a(tc_context *c, FilterArgs a)
{
    if (a.nlines == 0 || a.n == 0) return TC_OK;
    a.div = tc_f32_pow4(2 * (int64_t)a.r + 1);
    const bool odd = (a.r & 1) != 0;
    const size_t per_warp = tpl_a_per_warp(a.r);
    const int64_t nwarps = (a.nlines + 31) / 32;
    const int wpb = b2_warps_per_block(c, per_warp, (a.role ? 1 : 2) * nwarps, 32);
    const unsigned grid = (unsigned)((a.role ? 1 : 2) * ((nwarps + wpb - 1) / wpb));
    if (TC_ENV_FLAG("TC_FILTER_TRACE"))
        fprintf(stderr, "tpl a filter: n=%d nj=%d r=%d tr=%d wpb=%d\n", a.n, a.nj, a.r, a.out_transposed, wpb);
    tc_prof_begin(c, TCP_BOX_FILTER8);
    if (odd) TC_TRY(b2_launch(c, k_box_tpl_a<true>, a, grid, wpb, per_warp * wpb));
    else TC_TRY(b2_launch(c, k_box_tpl_a<false>, a, grid, wpb, per_warp * wpb));
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}

// data / win: line-contiguous pair; output sample-major (or line-contiguous when out_transposed)
static int launch_box_tpl_b(tc_context *c, FilterArgs a)
{
    if (a.nlines == 0 || a.n == 0) return TC_OK;
    a.div = tc_f32_pow4(2 * (int64_t)a.r + 1);
    const bool odd = (a.r & 1) != 0;
    const size_t per_warp = tpl_b_per_warp(a.r);
    const int64_t nwarps = (a.nlines + 31) / 32;
    const int wpb = b2_warps_per_block(c, per_warp, nwarps, 32);
    const unsigned grid = (unsigned)((nwarps + wpb - 1) / wpb);
    if (TC_ENV_FLAG("TC_FILTER_TRACE"))
        fprintf(stderr, "tpl b filter: n=%d nj=%d r=%d out=%d tr=%d wpb=%d\n", a.n, a.nj, a.r, a.mode_out,
                a.out_transposed, wpb);
    tc_prof_begin(c, a.single_axis ? TCP_BOX_FILTER_1D : TCP_BOX_FILTER);
    if (a.mode_out == FOUT_RESID) {
        if (odd) TC_TRY(b2_launch(c, k_box_tpl_b<true, FOUT_RESID>, a, grid, wpb, per_warp * wpb));
        else TC_TRY(b2_launch(c, k_box_tpl_b<false, FOUT_RESID>, a, grid, wpb, per_warp * wpb));
    } else {
        if (odd) TC_TRY(b2_launch(c, k_box_tpl_b<true, FOUT_BG>, a, grid, wpb, per_warp * wpb));
        else TC_TRY(b2_launch(c, k_box_tpl_b<false, FOUT_BG>, a, grid, wpb, per_warp * wpb));
    }
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
