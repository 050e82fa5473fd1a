// k_pack.cuh -- MS row order <-> (bl, corr, time, chan) windows and the flag
// counts behind the window statistics.
//   P1 _numba_pack_data              tricolour/packing.py:243-278
//   P2 _unpack_data / _numpy_unpack_transpose   packing.py:369-415
//   W1 _window_stats (counting part) tricolour/window_statistics.py:12-66
// A row (r, :, :) is nchan*ncorr contiguous samples with corr fastest; a window
// row (bl, c, t, :) is nchan contiguous samples.  One thread moves the ncorr
// samples of one (row, chan): the read is one contiguous ncorr-vector, the
// writes are ncorr streams that are each contiguous across the warp.
#pragma once
#include "tc_common.cuh"
#include "k_elementwise.cuh"

// window defaults (packing.py:96-98, 116-117): vis = NaN + NaNj, flag = 1
__global__ void __launch_bounds__(256)
k_fill_windows(float2 *__restrict__ vis_win, u8 *__restrict__ flag_win, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (vis_win) vis_win[i] = make_float2(NAN, NAN);
    if (flag_win) flag_win[i] = 1;
}

// the same defaults for the (baseline, time) slots no row writes: `slots` holds bl * ntime + t,
// a slot is ncorr runs of nchan samples.  Everything else is overwritten by the pack itself, so
// a window that every row covers is written once instead of twice.
__global__ void __launch_bounds__(256)
k_fill_slots(const int32_t *__restrict__ slots, int64_t nslots, int ncorr, int ntime, int nchan,
             float2 *__restrict__ vis_win, u8 *__restrict__ flag_win)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)ncorr * nchan;
    if (g >= nslots * per) return;
    const int64_t s = g / per;
    const int64_t rem = g - s * per;
    const int cc = (int)(rem / nchan), f = (int)(rem - (int64_t)cc * nchan);
    const int slot = slots[s];
    const int bl = slot / ntime, t = slot - bl * ntime;
    const int64_t dst = (((int64_t)bl * ncorr + cc) * ntime + t) * nchan + f;
    if (vis_win) vis_win[dst] = make_float2(NAN, NAN);
    if (flag_win) flag_win[dst] = 1;
}

__global__ void __launch_bounds__(256)
k_pack(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
       const float2 *__restrict__ vis, const u8 *__restrict__ flags, int nchan, int ncorr,
       int ntime, float2 *__restrict__ vis_win, u8 *__restrict__ flag_win)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan) return;
    int64_t r = g / nchan;
    int f = (int)(g - r * nchan);
    int bl = row_bl[r];
    if (bl < 0) return;
    int t = row_t[r];
    int64_t src = g * ncorr;
    for (int c = 0; c < ncorr; c++) {
        int64_t dst = (((int64_t)bl * ncorr + c) * ntime + t) * nchan + f;
        if (vis_win) vis_win[dst] = vis[src + c];
        if (flag_win) flag_win[dst] = flags[src + c];
    }
}

// ncorr == 4 fast path: 4 channels per thread so that flag bytes move as
// 32-bit words and visibilities as 16-byte vectors
__global__ void __launch_bounds__(256)
k_pack_c4(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
          const float4 *__restrict__ vis, const uint4 *__restrict__ flags, int nchan4, int ntime,
          float2 *__restrict__ vis_win, uint32_t *__restrict__ flag_win)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan4) return;
    int64_t r = g / nchan4;
    int f4 = (int)(g - r * nchan4);
    int bl = row_bl[r];
    if (bl < 0) return;
    int t = row_t[r];
    int nchan = nchan4 * 4;
    if (vis_win) {
        // 4 channels x 4 corr complex64 = 8 float4
        float2 v[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float4 a = vis[g * 8 + 2 * k], b = vis[g * 8 + 2 * k + 1];
            v[k][0] = make_float2(a.x, a.y); v[k][1] = make_float2(a.z, a.w);
            v[k][2] = make_float2(b.x, b.y); v[k][3] = make_float2(b.z, b.w);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int64_t dst = (((int64_t)bl * 4 + c) * ntime + t) * nchan + f4 * 4;
            float4 *o = (float4 *)(vis_win + dst);
            o[0] = make_float4(v[0][c].x, v[0][c].y, v[1][c].x, v[1][c].y);
            o[1] = make_float4(v[2][c].x, v[2][c].y, v[3][c].x, v[3][c].y);
        }
    }
    if (flag_win) {
        uint4 w = flags[g];  // bytes: chan k (word), corr c (byte within word)
        uint32_t in[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t o = ((in[0] >> (8 * c)) & 0xffu) | (((in[1] >> (8 * c)) & 0xffu) << 8) |
                         (((in[2] >> (8 * c)) & 0xffu) << 16) | (((in[3] >> (8 * c)) & 0xffu) << 24);
            int64_t dst = (((int64_t)bl * 4 + c) * ntime + t) * nchan4 + f4;
            flag_win[dst] = o;
        }
    }
}

// gather back; ELEM is float2 (vis) or u8 (flags).  Rows without a window slot
// stay zero (packing.py:396).  General form (any ncorr): one thread per (row, chan).
template <typename ELEM>
__global__ void __launch_bounds__(256)
k_unpack(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
         const ELEM *__restrict__ win, int nchan, int ncorr, int ntime, ELEM *__restrict__ out)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan) return;
    int64_t r = g / nchan;
    int f = (int)(g - r * nchan);
    int bl = row_bl[r];
    int t = row_t[r];
    ELEM zero;
    memset(&zero, 0, sizeof(ELEM));
    for (int c = 0; c < ncorr; c++) {
        ELEM v = zero;
        if (bl >= 0) v = win[(((int64_t)bl * ncorr + c) * ntime + t) * nchan + f];
        out[g * ncorr + c] = v;
    }
}

// ncorr == 4 flags: a thread gathers 4 channels of the 4 correlation rows as four
// 32-bit words (each window row read is contiguous across the warp), transposes the
// 4 x 4 bytes in registers and writes the 16 row-order bytes with one store.
//   MODE 0: plain gather                        (packing.py:369-415)
//   MODE 1: any(corr) broadcast over the 4 correlations (app.py:479-480)
template <int MODE>
__global__ void __launch_bounds__(256)
k_unpack_flags_c4(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
                  const uint32_t *__restrict__ win, int nchan4, int ntime, uint4 *__restrict__ out)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan4) return;
    const int64_t r = g / nchan4;
    const int f4 = (int)(g - r * nchan4);
    const int bl = row_bl[r];
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (bl >= 0) {
        const int t = row_t[r];
        uint32_t w[4];
#pragma unroll
        for (int c = 0; c < 4; c++) w[c] = win[(((int64_t)bl * 4 + c) * ntime + t) * nchan4 + f4];
        if (MODE == 1) {
            uint32_t any = w[0] | w[1] | w[2] | w[3];
            // bytes -> 0/1, then every channel's byte replicated over the 4 correlations
            any = (any | (any >> 4)) & 0x0f0f0f0fu;
            any = (any | (any >> 2)) & 0x03030303u;
            any = (any | (any >> 1)) & 0x01010101u;
            o.x = (any & 0xffu) * 0x01010101u;
            o.y = ((any >> 8) & 0xffu) * 0x01010101u;
            o.z = ((any >> 16) & 0xffu) * 0x01010101u;
            o.w = (any >> 24) * 0x01010101u;
        } else {
            // out word k (channel k) = bytes k of w[0..3]
            o.x = (w[0] & 0xffu) | ((w[1] & 0xffu) << 8) | ((w[2] & 0xffu) << 16) | ((w[3] & 0xffu) << 24);
            o.y = ((w[0] >> 8) & 0xffu) | (((w[1] >> 8) & 0xffu) << 8) | (((w[2] >> 8) & 0xffu) << 16) |
                  (((w[3] >> 8) & 0xffu) << 24);
            o.z = ((w[0] >> 16) & 0xffu) | (((w[1] >> 16) & 0xffu) << 8) | (((w[2] >> 16) & 0xffu) << 16) |
                  (((w[3] >> 16) & 0xffu) << 24);
            o.w = (w[0] >> 24) | ((w[1] >> 24) << 8) | ((w[2] >> 24) << 16) | ((w[3] >> 24) << 24);
        }
    }
    out[g] = o;
}

// one-correlation window (polarised flagging, app.py:415-432) back to rows of NOUT = 4
// correlations: 16 channels per thread (one 16-byte window read), every flag byte
// normalised to 0/1 and replicated over the correlations (four 16-byte stores)
__global__ void __launch_bounds__(256)
k_unpack_flags_c1_to4(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
                      const uint4 *__restrict__ win, int nchan16, int ntime, uint4 *__restrict__ out)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan16) return;
    const int64_t r = g / nchan16;
    const int f16 = (int)(g - r * nchan16);
    const int bl = row_bl[r];
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (bl >= 0) w = win[((int64_t)bl * ntime + row_t[r]) * nchan16 + f16];
    const uint32_t in[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t any = in[k];
        any = (any | (any >> 4)) & 0x0f0f0f0fu;
        any = (any | (any >> 2)) & 0x03030303u;
        any = (any | (any >> 1)) & 0x01010101u;
        out[g * 4 + k] = make_uint4((any & 0xffu) * 0x01010101u, ((any >> 8) & 0xffu) * 0x01010101u,
                                    ((any >> 16) & 0xffu) * 0x01010101u, (any >> 24) * 0x01010101u);
    }
}

// ncorr == 4 visibilities: a thread gathers one channel of the 4 correlation rows
// (8-byte reads, contiguous across the warp per row) and writes its 32 row-order bytes
__global__ void __launch_bounds__(256)
k_unpack_vis_c4(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
                const float2 *__restrict__ win, int nchan, int ntime, float4 *__restrict__ out)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan) return;
    const int64_t r = g / nchan;
    const int f = (int)(g - r * nchan);
    const int bl = row_bl[r];
    float2 v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = make_float2(0.f, 0.f);
    if (bl >= 0) {
        const int t = row_t[r];
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = win[(((int64_t)bl * 4 + c) * ntime + t) * nchan + f];
    }
    out[g * 2] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    out[g * 2 + 1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}

// flags only, any ncorr_win -> ncorr_out: out[r, f, :] = any over the window's
// correlations (app.py:479-480 fused into the gather); general fallback
__global__ void __launch_bounds__(256)
k_unpack_any_corr(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t,
                  int64_t nrow, const u8 *__restrict__ win, int nchan, int ncorr, int ntime,
                  int ncorr_out, u8 *__restrict__ out)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan) return;
    int64_t r = g / nchan;
    int f = (int)(g - r * nchan);
    int bl = row_bl[r];
    int t = row_t[r];
    u8 any = 0;
    if (bl >= 0)
        for (int c = 0; c < ncorr; c++) any |= win[(((int64_t)bl * ncorr + c) * ntime + t) * nchan + f];
    any = any ? 1 : 0;
    for (int c = 0; c < ncorr_out; c++) out[g * ncorr_out + c] = any;
}

// ----------------------------------------------------------------------------
// K2 + N2 + P1 fused (polarised / total-power flagging, app.py:415-432 followed by
// pack_data): per (row, chan) the Stokes intensity of the correlations
// (stokes.py:157-209, or 79-154 with `with_unpol`) and the OR of their flags go
// straight into the one-correlation windows (nbl, 1, T, F).  The rows are read once
// (32 + 4 bytes per sample for 4 correlations), 9 bytes are written.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_stokes_pack(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t, int64_t nrow,
              const float2 *__restrict__ vis, const u8 *__restrict__ flags, int nchan, int ncorr, int ntime,
              StokesTerms pol, StokesTerms unpol, int with_unpol, float2 *__restrict__ vis_win,
              u8 *__restrict__ flag_win)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nrow * nchan) return;
    const int64_t r = g / nchan;
    const int f = (int)(g - r * nchan);
    const int bl = row_bl[r];
    if (bl < 0) return;
    const int t = row_t[r];
    float2 v[8];
    const float2 *src = vis + g * ncorr;
    u8 any = 0;
    if (ncorr == 4 && (((uintptr_t)src) & 15) == 0 && (((uintptr_t)(flags + g * 4)) & 3) == 0) {
        const float4 a = ((const float4 *)src)[0], b = ((const float4 *)src)[1];
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
        v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
        any = *reinterpret_cast<const uint32_t *>(flags + g * 4) ? 1 : 0;
    } else {
        for (int c = 0; c < ncorr && c < 8; c++) { v[c] = src[c]; any |= flags[g * ncorr + c]; }
        any = any ? 1 : 0;
    }
    const int64_t dst = ((int64_t)bl * ntime + t) * nchan + f;
    vis_win[dst] = make_float2(stokes_intensity(v, pol, unpol, with_unpol, ncorr), 0.0f);
    flag_win[dst] = any;
}

// W1: sums of the flag bytes per baseline and per channel (window_statistics.py:
// 26-64 sums the flag array itself, so a uint8 window contributes its byte values).
// General form: grid = (segments, nbl); a block walks `rows_per_seg` window rows of one
// baseline, every thread owning a strided set of channels, then folds with warp shuffles.
__global__ void __launch_bounds__(256)
k_window_counts(const u8 *__restrict__ flags, int64_t rows_per_bl, int rows_per_seg, int F,
                unsigned long long *__restrict__ bl_counts, unsigned long long *__restrict__ chan_counts)
{
    __shared__ unsigned long long s_tot;
    int64_t bl = blockIdx.y;
    int64_t row0 = (int64_t)blockIdx.x * rows_per_seg;
    int64_t row1 = row0 + rows_per_seg < rows_per_bl ? row0 + rows_per_seg : rows_per_bl;
    if (threadIdx.x == 0) s_tot = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        unsigned int col = 0;
        for (int64_t row = row0; row < row1; row++) col += flags[(bl * rows_per_bl + row) * F + f];
        if (col) atomicAdd(&chan_counts[f], (unsigned long long)col);
        mine += col;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(TC_FULL_MASK, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_tot, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_tot) atomicAdd(&bl_counts[bl], s_tot);
}

// Vector form (F % 16 == 0): a thread owns 16 adjacent channels, a block of 256 threads a
// tile of 4096 channels, and walks up to TC_WC_ROWS window rows of one baseline with one
// 16-byte load per row (eight rows in flight): every row is read fully coalesced.  The 16
// column sums of a thread live in eight registers as packed 16-bit fields (128 rows x 255
// cannot overflow them).  A block adds its non-zero column sums to the channel counts with
// 64-bit reductions (consecutive threads, consecutive channels) and its total to the
// baseline with one more: a single launch, no partial sums in memory.
#define TC_WC_ROWS 128
__global__ void __launch_bounds__(256)
k_window_counts_v16(const uint4 *__restrict__ flags, int rows_per_bl, int F16,
                    unsigned long long *__restrict__ chan_counts, unsigned long long *__restrict__ bl_counts)
{
    __shared__ unsigned s_tot[8];
    const int seg = blockIdx.x, tile = blockIdx.y;
    const int64_t bl = blockIdx.z;
    const int f16 = tile * 256 + (int)threadIdx.x;
    const int row0 = seg * TC_WC_ROWS;
    const int row1 = row0 + TC_WC_ROWS < rows_per_bl ? row0 + TC_WC_ROWS : rows_per_bl;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0u;
    if (f16 < F16) {
        const uint4 *p = flags + (bl * rows_per_bl + row0) * (int64_t)F16 + f16;
        int row = row0;
        for (; row + 8 <= row1; row += 8) {
            uint4 w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) w[q] = p[(int64_t)q * F16];
            p += 8 * (int64_t)F16;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                acc[0] += w[q].x & 0x00ff00ffu; acc[1] += (w[q].x >> 8) & 0x00ff00ffu;
                acc[2] += w[q].y & 0x00ff00ffu; acc[3] += (w[q].y >> 8) & 0x00ff00ffu;
                acc[4] += w[q].z & 0x00ff00ffu; acc[5] += (w[q].z >> 8) & 0x00ff00ffu;
                acc[6] += w[q].w & 0x00ff00ffu; acc[7] += (w[q].w >> 8) & 0x00ff00ffu;
            }
        }
        for (; row < row1; row++) {
            const uint4 w = *p;
            p += F16;
            acc[0] += w.x & 0x00ff00ffu; acc[1] += (w.x >> 8) & 0x00ff00ffu;
            acc[2] += w.y & 0x00ff00ffu; acc[3] += (w.y >> 8) & 0x00ff00ffu;
            acc[4] += w.z & 0x00ff00ffu; acc[5] += (w.z >> 8) & 0x00ff00ffu;
            acc[6] += w.w & 0x00ff00ffu; acc[7] += (w.w >> 8) & 0x00ff00ffu;
        }
    }
    // column k of the thread: word k / 4, byte k % 4 -> acc[2 * (k / 4) + (k & 1)], field (k % 4) / 2
    unsigned mine = 0;
    if (f16 < F16) {
        unsigned long long *out = chan_counts + (int64_t)f16 * 16;
        uint32_t col[16];
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            col[4 * wd + 0] = acc[2 * wd] & 0xffffu;
            col[4 * wd + 1] = acc[2 * wd + 1] & 0xffffu;
            col[4 * wd + 2] = acc[2 * wd] >> 16;
            col[4 * wd + 3] = acc[2 * wd + 1] >> 16;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            mine += col[k];
            if (col[k]) atomicAdd(out + k, (unsigned long long)col[k]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(TC_FULL_MASK, mine, o);
    if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; k++) t += s_tot[k];
        if (t) atomicAdd(&bl_counts[bl], t);
    }
}

