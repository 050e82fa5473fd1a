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

// window defaults (packing.py:96-98, 116-117): vis = NaN + NaNj, flag = 1
__global__ void __launch_bounds__(256)
k_fill_windows(float2 *__restrict__ vis_win, u8 *__restrict__ flag_win, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (vis_win) vis_win[i] = make_float2(NAN, NAN);
    if (flag_win) flag_win[i] = 1;
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
// stay zero (packing.py:396).
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

// flags only: out[r, f, :] = any over corr (app.py:479-480 fused into the gather)
__global__ void __launch_bounds__(256)
k_unpack_any_corr(const int32_t *__restrict__ row_bl, const int32_t *__restrict__ row_t,
                  int64_t nrow, const u8 *__restrict__ win, int nchan, int ncorr, int ntime,
                  u8 *__restrict__ out)
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
    for (int c = 0; c < ncorr; c++) out[g * ncorr + c] = any;
}

// W1: sums of the flag bytes per baseline and per channel.  grid = (segments,
// nbl); a block walks `rows_per_seg` window rows of one baseline, every thread
// owning a strided set of channels, then folds with warp shuffles.
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
