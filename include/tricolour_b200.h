/*
 * tricolour_b200.h -- C ABI of the B200-native tricolour flagging hot path.
 *
 * Every entry point replaces one function of the reference (ratt-ru/tricolour,
 * paths below are relative to the reference checkout) at the numpy boundary
 * that tricolour/dask_wrappers.py:9-18 binds.  Plain pointers and sizes only.
 *
 * Conventions
 *   - windows are C-contiguous (bl, corr, time, chan) exactly as
 *     tricolour/packing.py:15; a "plane" is one (bl, corr) slice of shape (T, F)
 *     and ncp = nbl * ncorr.
 *   - flags are one byte per sample, 0 = clear, non-zero = flagged.
 *   - `space` says where the array arguments live: TC_HOST (the library stages
 *     them through its own device workspace and returns after the result is
 *     back in host memory) or TC_DEVICE (device pointers, work is enqueued on
 *     the context's stream and the call returns without synchronising).
 *     Small parameter tables (windows, masks, selectors) are always host memory.
 *   - every function returns TC_OK or an error code; tc_last_error() gives the
 *     thread-local message.  Nothing aborts the process.
 *   - a tc_context owns a stream and a workspace arena and must not be used by
 *     two threads at once; create one per calling thread (the reference is
 *     called from a dask ThreadPool, tricolour/apps/tricolour/app.py:266-271).
 */
#ifndef TRICOLOUR_B200_H
#define TRICOLOUR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_OK 0
#define TC_ERR_VALUE 1 /* bad argument -> Python ValueError */
#define TC_ERR_CUDA 2  /* CUDA runtime failure -> RuntimeError */
#define TC_ERR_NOGPU 3 /* no usable device -> RuntimeError */

#define TC_HOST 0
#define TC_DEVICE 1

#define TC_VIS_COMPLEX64 0
#define TC_VIS_FLOAT32 1

typedef struct tc_context tc_context;

/* ---- runtime ---------------------------------------------------------- */
const char *tc_last_error(void);
int tc_device_count(void);
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream; pass
 * cudaStreamLegacy, (void *)1, for the default stream), or NULL to let the
 * context create its own non-blocking stream. */
int tc_context_create(int device, void *stream, tc_context **out);
void tc_context_destroy(tc_context *ctx);
int tc_synchronize(tc_context *ctx);
/* number of kernels this context has launched so far */
unsigned long long tc_launch_count(tc_context *ctx);
/* Optional timing of the kernel families with CUDA events on the context's
 * stream (used by bench.py for the roofline of the dominant kernel).
 * tc_profile_read synchronises on the recorded events. */
int tc_profile_enable(tc_context *ctx, int on);
int tc_profile_reset(tc_context *ctx);
int tc_profile_count(void);
const char *tc_profile_name(int id);
int tc_profile_read(tc_context *ctx, int id, double *total_ms, long long *launches);
/* workspace high-water mark in bytes */
size_t tc_workspace_peak(tc_context *ctx);
/* device memory the context's arena holds right now, and the context's share of the
 * DEVICE-WIDE workspace budget (environment TC_WORKSPACE_MB, default 49152): the budget is
 * divided between the contexts that hold an arena on the device, because the reference
 * calls these functions from ThreadPool(nworkers) threads (apps/tricolour/app.py:266-271)
 * and every thread owns a context.  tc_sum_threshold sizes its plane batches to the share;
 * an arena that outgrew the share is released at the start of the next call. */
size_t tc_workspace_held(tc_context *ctx);
size_t tc_workspace_share(tc_context *ctx);
/* waits for the context's stream and hands the arena back to the driver */
int tc_context_trim(tc_context *ctx);
int tc_alloc_pinned(size_t nbytes, void **out);
int tc_free_pinned(void *ptr);
/* asynchronous copy on the context's stream: kind 0 = host -> device, 1 = device -> host.
 * With page-locked host memory the call returns at once; it is what the pipelined
 * executor (tricolour_b200/strategy.py) stages blocks with -- the reference moves its
 * blocks through dask (apps/tricolour/app.py:451-467). */
int tc_memcpy_async(tc_context *ctx, void *dst, const void *src, size_t nbytes, int kind);
/* 1 when the library was built with the CPU SIMT emulator (tests only) */
int tc_is_emulated(void);

/* ---- F1: flag_nans_and_zeros, tricolour/flagging.py:29-62 --------------- */
int tc_flag_nans_zeros(tc_context *ctx, const void *vis_c64, const uint8_t *flags,
                       uint8_t *out, int64_t n, int space);

/* ---- F2: flag_autos, tricolour/flagging.py:65-95 ----------------------- */
/* auto_sel[bl] != 0 where ubl[bl,1] == ubl[bl,2]; plane_elems = ncorr*T*F */
int tc_flag_autos(tc_context *ctx, const uint8_t *flags, const uint8_t *auto_sel,
                  int64_t nbl, int64_t plane_elems, uint8_t *out, int space);

/* ---- F3: apply_static_mask, tricolour/flagging.py:98-172 ---------------- */
/* bl_sel[bl]: baseline inside uvrange (lines 141-150); chan_mask[f]: channel
 * hit by the (combined) static mask (lines 157-160).  mode 0 = "or",
 * 1 = "override".  rows_per_bl = ncorr*T. */
int tc_apply_channel_mask(tc_context *ctx, const uint8_t *flags, const uint8_t *bl_sel,
                          const uint8_t *chan_mask, int mode, int64_t nbl,
                          int64_t rows_per_bl, int64_t nchan, uint8_t *out, int space);

/* ---- S13: sum_threshold_flagger, tricolour/flagging.py:1076-1196 -------- */
/* Parameters already conditioned the way lines 1160-1179 condition them. */
typedef struct tc_st_params {
    double outlier_nsigma;
    int32_t nwin_time;
    int32_t nwin_freq;
    const int64_t *windows_time; /* [nwin_time] */
    const double *tf_time;       /* pow(rho, log2(w)), flagging.py:641 */
    const float *scale_time;     /* float32(1.0 / w), flagging.py:664 */
    const int64_t *windows_freq; /* [nwin_freq] */
    const double *tf_freq;
    const float *scale_freq;
    double background_reject;
    int32_t background_iterations;
    int32_t nchunk_ends;
    const int64_t *radii_spec;      /* [(iterations+1)*2] box radii, sigma=(0, ef*spike_f) */
    const int64_t *radii_2d;        /* [(iterations+1)*2] box radii, sigma=ef*(spike_t, spike_f) */
    const int64_t *freq_chunk_ends; /* [nchunk_ends], flagging.py:1172-1173 */
    int64_t time_extend;
    int64_t freq_extend;
    int64_t average_freq;
    double flag_all_time_frac;
    double flag_all_freq_frac;
    int32_t num_major_iterations;
    int32_t reserved;
} tc_st_params;

/* vis: (ncp, T, F) complex64 or float32 (vis_kind); flags/out: (ncp, T, F).
 * Returns the LAST major iteration's flags only (flagging.py:1196). */
int tc_sum_threshold(tc_context *ctx, const tc_st_params *p, const void *vis,
                     int vis_kind, const uint8_t *flags, int64_t ncp, int64_t T,
                     int64_t F, uint8_t *out, int space);

/* ---- U1: uvcontsub_flagger, tricolour/flagging.py:989-1073 -------------- */
int tc_uvcontsub(tc_context *ctx, const void *vis_c64, const uint8_t *flags,
                 int64_t ncp, int64_t T, int64_t F, int major_cycles,
                 int or_original_from_cycle, int taylor_degrees, double sigma,
                 uint8_t *out, int space);

/* ---- K2: polarised / unpolarised intensity, tricolour/stokes.py:79-209 -- */
/* vis: (nrowchan, ncorr) complex64 -> out (nrowchan) complex64 (imag = 0).
 * term k: value = (a_re + i a_im) * (s1 * vis[c1] + s2 * vis[c2]) with
 * idx[2k..] = (c1, c2) and coef[4k..] = (a_re, a_im, s1, s2). */
int tc_polarised_intensity(tc_context *ctx, const void *vis_c64, int64_t nrowchan,
                           int ncorr, const int32_t *pol_idx, const double *pol_coef,
                           int npol, void *out_c64, int space);
int tc_unpolarised_intensity(tc_context *ctx, const void *vis_c64, int64_t nrowchan,
                             int ncorr, const int32_t *unpol_idx,
                             const double *unpol_coef, int nunpol,
                             const int32_t *pol_idx, const double *pol_coef, int npol,
                             void *out_c64, int space);

/* ---- P1/P2: pack / unpack, tricolour/packing.py:243-278, 369-415 -------- */
/* row_bl[r]: window baseline slot of MS row r (or -1 to skip the row);
 * row_t[r] = time_inv[r].  Rows that lose a (bl, t) collision to a later row
 * must be marked -1 by the caller (the reference's loop order makes the last
 * row win).  fill != 0 first writes the window defaults of packing.py:96-98,
 * 116-117 (vis NaN+NaNj, flag 1).  data: (nrow, nchan, ncorr). */
int tc_pack(tc_context *ctx, const int32_t *row_bl, const int32_t *row_t, int64_t nrow,
            const void *vis_c64, const uint8_t *flags, int64_t nchan, int64_t ncorr,
            int64_t ntime, int64_t nbl, void *vis_win, uint8_t *flag_win, int fill,
            int space);
/* elem_size 1 (flags) or 8 (complex64); rows with row_bl < 0 are zero filled */
int tc_unpack(tc_context *ctx, const int32_t *row_bl, const int32_t *row_t, int64_t nrow,
              const void *window, int elem_size, int64_t nchan, int64_t ncorr,
              int64_t ntime, int64_t nbl, void *out, int space);
/* corr-equalising unpack: out[r,f,:] = any_c window[bl,c,t,f]
 * (tricolour/apps/tricolour/app.py:479-480 fused into the gather) */
int tc_unpack_flags_any_corr(tc_context *ctx, const int32_t *row_bl, const int32_t *row_t,
                             int64_t nrow, const uint8_t *window, int64_t nchan,
                             int64_t ncorr, int64_t ntime, int64_t nbl, uint8_t *out,
                             int space);
/* the same gather with the any(corr) result broadcast over ncorr_out correlations: a
 * one-correlation window (polarised / total-power flagging, app.py:415-432) goes back to
 * rows of the measurement set's correlation count (app.py:479-480 broadcast_to). */
int tc_unpack_flags_broadcast(tc_context *ctx, const int32_t *row_bl, const int32_t *row_t,
                              int64_t nrow, const uint8_t *window, int64_t nchan,
                              int64_t ncorr_win, int64_t ncorr_out, int64_t ntime,
                              int64_t nbl, uint8_t *out, int space);
/* polarised_intensity (stokes.py:157-209; unpolarised 79-154 when nunpol > 0),
 * flags.any(axis=2) (app.py:420, 432) and _numba_pack_data (packing.py:243-278) in one
 * pass over the rows: vis/flags (row, chan, ncorr) -> windows (nbl, 1, ntime, nchan).
 * Terms as for tc_polarised_intensity; fill as for tc_pack. */
int tc_stokes_pack(tc_context *ctx, const int32_t *row_bl, const int32_t *row_t, int64_t nrow,
                   const void *vis_c64, const uint8_t *flags, int64_t nchan, int64_t ncorr,
                   int64_t ntime, int64_t nbl, const int32_t *unpol_idx,
                   const double *unpol_coef, int nunpol, const int32_t *pol_idx,
                   const double *pol_coef, int npol, void *vis_win, uint8_t *flag_win,
                   int fill, int space);

/* ---- W1: window statistics, tricolour/window_statistics.py:12-66 -------- */
/* bl_counts[nbl], chan_counts[nchan]: sums of the flag bytes, always returned
 * to HOST memory (they are a few KB). */
int tc_window_counts(tc_context *ctx, const uint8_t *flags, int64_t nbl, int64_t ncorr,
                     int64_t T, int64_t F, uint64_t *bl_counts, uint64_t *chan_counts,
                     int space);

/* ---- device-resident helpers for the strategy executor --- */
/* out = a | b  (strat_executor.py:43, 54, 59, 76) */
int tc_flags_or(tc_context *ctx, const uint8_t *a, const uint8_t *b, uint8_t *out,
                int64_t n, int space);

/* ---- stage-level entry points ------------------------------------------ */
/* These expose the private numba kernels of the reference one by one so that
 * the parity tests can mirror tricolour/tests/test_flagging.py.  All arrays
 * are (ncp, T, F) row-major unless stated. */
/* _average_freq, flagging.py:819-875 -> data (ncp,T,Fa) f32, flags (ncp,T,Fa) */
int tc_stage_average_freq(tc_context *ctx, const void *vis, int vis_kind,
                          const uint8_t *flags, int64_t ncp, int64_t T, int64_t F,
                          int64_t factor, float *out_data, uint8_t *out_flags, int space);
/* _time_median, flagging.py:226-264 -> (ncp, F) */
int tc_stage_time_median(tc_context *ctx, const float *data, const uint8_t *flags,
                         int64_t ncp, int64_t T, int64_t F, float *out_data,
                         uint8_t *out_flags, int space);
/* _median_abs per frequency chunk, flagging.py:267-279 -> (ncp, nchunks) f64 */
int tc_stage_chunk_median_abs(tc_context *ctx, const float *data, const uint8_t *flags,
                              int64_t ncp, int64_t T, int64_t F,
                              const int64_t *chunk_ends, int nchunk_ends, double *out,
                              int space);
/* masked_gaussian_filter, flagging.py:469-513 (radii from line 451) */
int tc_stage_masked_filter(tc_context *ctx, const float *data, const uint8_t *flags,
                           int64_t ncp, int64_t T, int64_t F, int64_t r0, int64_t r1,
                           float *out, int space);
/* _linearly_interpolate_nans, flagging.py:347-359 */
int tc_stage_interp_nans(tc_context *ctx, const float *data, int64_t ncp, int64_t T,
                         int64_t F, float *out, int space);
/* _get_background2d, flagging.py:516-579; radii [(iterations+1)*2] */
int tc_stage_background2d(tc_context *ctx, const float *data, const uint8_t *flags,
                          int64_t ncp, int64_t T, int64_t F, int iterations,
                          const int64_t *radii, double reject_threshold,
                          const int64_t *chunk_ends, int nchunk_ends, float *out,
                          int space);
/* _sum_threshold, flagging.py:684-742; chunk_ends NULL -> [0, n] */
int tc_stage_sum_threshold(tc_context *ctx, const float *data, const uint8_t *flags,
                           int64_t ncp, int64_t T, int64_t F, int axis,
                           const int64_t *windows, const double *tf, const float *scale,
                           int nwin, double outlier_nsigma, const int64_t *chunk_ends,
                           int nchunk_ends, uint8_t *out, int space);
/* _combine_flags + _unaverage_freq, flagging.py:784-816, 878-918.
 * spec (ncp, Fa); time/freq (ncp, T, Fa); out (ncp, T, F). */
int tc_stage_combine_unaverage(tc_context *ctx, const uint8_t *spec, const uint8_t *time_f,
                               const uint8_t *freq_f, int64_t ncp, int64_t T, int64_t Fa,
                               int64_t F, int64_t time_extend, int64_t freq_extend,
                               int64_t average_freq, double flag_all_time_frac,
                               double flag_all_freq_frac, uint8_t *out, int space);

#ifdef __cplusplus
}
#endif
#endif /* TRICOLOUR_B200_H */
