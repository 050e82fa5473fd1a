// tricolour_b200.cu -- the C ABI declared in include/tricolour_b200.h.
// Single translation unit: kernels live in the k_*.cuh headers, the stage
// sequencing in st_driver.cuh.
#include "st_driver.cuh"
#include "k_uvcontsub.cuh"
#include "k_pack.cuh"

#include <stdlib.h>

extern "C" {

const char *tc_last_error(void) { return g_tc_err; }

int tc_is_emulated(void)
{
#ifdef TC_EMU
    return 1;
#else
    return 0;
#endif
}

int tc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int tc_context_create(int device, void *stream, tc_context **out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return tc_fail(TC_ERR_NOGPU, "no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return tc_fail(TC_ERR_VALUE, "device %d out of range [0, %d)", device, n);
    TC_CUDA(cudaSetDevice(device));
    tc_context *c = new tc_context();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; return tc_fail(TC_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e2)); }
        c->own_stream = true;
    }
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0)
        c->sm_count = v;
#ifndef TC_EMU
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess && v > 0)
        c->smem_optin = v;
#endif
    *out = c;
    return TC_OK;
}

void tc_context_destroy(tc_context *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    tc_arena_free_all(c);
#ifndef TC_EMU
    for (int k = 0; k < tc_context::NSLAB; k++) {
        if (c->slab[k]) cudaFreeHost(c->slab[k]);
        if (c->slab_ev[k]) cudaEventDestroy(c->slab_ev[k]);
    }
#endif
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int tc_synchronize(tc_context *c)
{
    TC_CUDA(cudaStreamSynchronize(c->stream));
    return TC_OK;
}

unsigned long long tc_launch_count(tc_context *c) { return c->launches; }

int tc_profile_enable(tc_context *c, int on)
{
    tc_prof_collect(c);
    c->prof = on != 0;
    return TC_OK;
}
int tc_profile_reset(tc_context *c)
{
    tc_prof_collect(c);
    for (int i = 0; i < TCP_NIDS; i++) { c->prof_ms[i] = 0; c->prof_cnt[i] = 0; }
    return TC_OK;
}
int tc_profile_count(void) { return TCP_NIDS; }
const char *tc_profile_name(int id) { return id >= 0 && id < TCP_NIDS ? tc_prof_names[id] : ""; }
int tc_profile_read(tc_context *c, int id, double *ms, long long *count)
{
    TC_REQUIRE(id >= 0 && id < TCP_NIDS, "bad profile id");
    tc_prof_collect(c);
    *ms = c->prof_ms[id];
    *count = c->prof_cnt[id];
    return TC_OK;
}
size_t tc_workspace_peak(tc_context *c) { return c->peak; }
size_t tc_workspace_held(tc_context *c) { return c->held; }
size_t tc_workspace_share(tc_context *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    return tc_ws_share(c);
}
int tc_context_trim(tc_context *c)
{
    if (!c) return tc_fail(TC_ERR_VALUE, "null context");
    TC_CUDA(cudaSetDevice(c->device));
    TC_CUDA(cudaStreamSynchronize(c->stream));
    tc_arena_free_all(c);
    c->peak = 0;
    return TC_OK;
}

int tc_alloc_pinned(size_t nbytes, void **out)
{
    TC_CUDA(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return TC_OK;
}
int tc_free_pinned(void *p)
{
    TC_CUDA(cudaFreeHost(p));
    return TC_OK;
}
int tc_memcpy_async(tc_context *c, void *dst, const void *src, size_t nbytes, int kind)
{
    if (!c) return tc_fail(TC_ERR_VALUE, "null context");
    TC_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (host to device) or 1 (device to host)");
    TC_CUDA(cudaSetDevice(c->device));
    if (nbytes)
        TC_CUDA(cudaMemcpyAsync(dst, src, nbytes, kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                c->stream));
    return TC_OK;
}

static int tc_begin(tc_context *c)
{
    if (!c) return tc_fail(TC_ERR_VALUE, "null context");
    TC_CUDA(cudaSetDevice(c->device));
    tc_slab_rotate(c);
    return tc_arena_reset(c);
}

// ------------------------------------------------------------------ F1 ------
int tc_flag_nans_zeros(tc_context *c, const void *vis, const uint8_t *flags, uint8_t *out,
                       int64_t n, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(n >= 0, "negative size");
    const float2 *dvis; const u8 *dfl; u8 *dout;
    TC_TRY(tc_stage_in(c, (const float2 *)vis, (size_t)n, space, &dvis));
    TC_TRY(tc_stage_in(c, flags, (size_t)n, space, &dfl));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &dout));
    if (n) TC_TRY(launch_flag_nans_zeros(c, dvis, dfl, dout, n));
    return tc_stage_out_end(c, out, dout, (size_t)n, space);
}

// ------------------------------------------------------------- F2 / F3 ------
static int apply_mask_common(tc_context *c, const u8 *flags, const u8 *bl_sel, const u8 *chan_mask,
                             int mode, int64_t nbl, int64_t rows_per_bl, int64_t nchan, u8 *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(nbl >= 0 && rows_per_bl >= 0 && nchan >= 0, "negative size");
    int64_t total = nbl * rows_per_bl * nchan;
    const u8 *dfl; u8 *dout;
    TC_TRY(tc_stage_in(c, flags, (size_t)total, space, &dfl));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)total, space, &dout));
    // small tables: normalise to 0/1 and upload
    std::vector<u8> sel((size_t)(nbl > 0 ? nbl : 1)), cm((size_t)(nchan > 0 ? nchan : 1), 0);
    for (int64_t i = 0; i < nbl; i++) sel[i] = bl_sel[i] ? 1 : 0;
    if (chan_mask) for (int64_t i = 0; i < nchan; i++) cm[i] = chan_mask[i] ? 1 : 0;
    u8 *dsel, *dcm;
    TC_TRY(tc_alloc(c, sel.size(), &dsel));
    TC_TRY(tc_alloc(c, cm.size(), &dcm));
    TC_TRY(tc_upload_small(c, sel.data(), sel.size(), dsel));
    TC_TRY(tc_upload_small(c, cm.data(), cm.size(), dcm));
    TC_TRY(launch_apply_mask(c, dfl, dsel, dcm, mode, nbl, rows_per_bl, nchan, dout));
    return tc_stage_out_end(c, out, dout, (size_t)total, space);
}

int tc_flag_autos(tc_context *c, const uint8_t *flags, const uint8_t *auto_sel, int64_t nbl,
                  int64_t plane_elems, uint8_t *out, int space)
{
    // one "row" of plane_elems "channels" per baseline
    return apply_mask_common(c, flags, auto_sel, nullptr, 2, nbl, 1, plane_elems, out, space);
}

int tc_apply_channel_mask(tc_context *c, const uint8_t *flags, const uint8_t *bl_sel,
                          const uint8_t *chan_mask, int mode, int64_t nbl, int64_t rows_per_bl,
                          int64_t nchan, uint8_t *out, int space)
{
    if (mode != 0 && mode != 1)
        return tc_fail(TC_ERR_VALUE, "Invalid accumulation_mode. Should be 'or' or 'override'");
    return apply_mask_common(c, flags, bl_sel, chan_mask, mode, nbl, rows_per_bl, nchan, out, space);
}

int tc_flags_or(tc_context *c, const uint8_t *a, const uint8_t *b, uint8_t *out, int64_t n, int space)
{
    TC_TRY(tc_begin(c));
    const u8 *da, *db; u8 *dout;
    TC_TRY(tc_stage_in(c, a, (size_t)n, space, &da));
    TC_TRY(tc_stage_in(c, b, (size_t)n, space, &db));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &dout));
    if (n) TC_TRY(launch_or(c, da, db, dout, n));
    return tc_stage_out_end(c, out, dout, (size_t)n, space);
}

// ----------------------------------------------------------------- S13 ------
static size_t st_workspace_per_plane(int64_t T, int64_t F, int64_t Fa, int nchunks, int64_t maxw)
{
    size_t N = (size_t)T * Fa;
    size_t pad = (size_t)(nchunks > 0 ? nchunks : 1) * 2 * (size_t)(maxw > 0 ? maxw : 1) * (size_t)(T > F ? T : F);
    // data(8) + flags(2) + bg work(18) + bg/dres(8) + ST flags(3) + scan scratch(9 + pad) + out staging
    return N * 48 + pad * 9 + (size_t)T * F * 2 + (size_t)(T + F) * 8 + (1 << 16);
}


static int st_validate(const tc_st_params *p)
{
    TC_REQUIRE(p != nullptr, "null parameters");
    TC_REQUIRE(p->average_freq >= 1, "average_freq must be >= 1");
    TC_REQUIRE(p->background_iterations >= 0, "background_iterations must be >= 0");
    TC_REQUIRE(p->nchunk_ends >= 1, "freq_chunk_ends must have at least one entry");
    TC_REQUIRE(p->time_extend >= 0 && p->freq_extend >= 0, "extend sizes must be >= 0");
    return TC_OK;
}

int tc_sum_threshold(tc_context *c, const tc_st_params *p, const void *vis, int vis_kind,
                     const uint8_t *flags, int64_t ncp, int64_t T, int64_t F, uint8_t *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_TRY(st_validate(p));
    TC_REQUIRE(vis_kind == TC_VIS_COMPLEX64 || vis_kind == TC_VIS_FLOAT32, "unsupported visibility type");
    TC_REQUIRE(ncp >= 0 && T >= 0 && F >= 0, "negative shape");
    TC_REQUIRE(T < (1 << 30) && F < (1 << 30), "plane too large");
    int64_t total = ncp * T * F;
    size_t esz = vis_kind == TC_VIS_COMPLEX64 ? 8 : 4;
    u8 *dout;
    TC_TRY(tc_stage_out_begin(c, out, (size_t)total, space, &dout));
    if (total == 0) return tc_stage_out_end(c, out, dout, 0, space);
    if (p->num_major_iterations <= 0) {
        TC_CUDA(cudaMemsetAsync(dout, 0, (size_t)total, c->stream));
        return tc_stage_out_end(c, out, dout, (size_t)total, space);
    }
    int64_t Fa = (F + p->average_freq - 1) / p->average_freq;
    int64_t maxw = 1;
    for (int k = 0; k < p->nwin_time; k++) if (p->windows_time[k] > maxw) maxw = p->windows_time[k];
    for (int k = 0; k < p->nwin_freq; k++) if (p->windows_freq[k] > maxw) maxw = p->windows_freq[k];
    size_t per_plane = st_workspace_per_plane(T, F, Fa, p->nchunk_ends - 1, maxw);
    size_t io_per_plane = (size_t)T * F * (2 + (space == TC_HOST ? esz + 1 : 0));
    int64_t batch = (int64_t)(tc_ws_share(c) / (per_plane + io_per_plane));
    if (batch < 1) batch = 1;
    if (batch > ncp) batch = ncp;

    for (int64_t p0 = 0; p0 < ncp; p0 += batch) {
        int64_t np = ncp - p0 < batch ? ncp - p0 : batch;
        int64_t n = np * T * F;
        tc_mark mark = tc_arena_mark(c);
        const char *dvis;
        const u8 *dfl;
        TC_TRY(tc_stage_in(c, (const char *)vis + (size_t)p0 * T * F * esz, (size_t)n * esz, space, &dvis));
        TC_TRY(tc_stage_in(c, flags + p0 * T * F, (size_t)n, space, &dfl));
        u8 *iter_flags;
        TC_TRY(tc_alloc(c, (size_t)n, &iter_flags));
        TC_TRY(launch_norm_flags(c, dfl, iter_flags, n));
        c->launches++;
        TC_KERNEL_CHECK();
        PassTables tab;
        TC_TRY(dev_pass_tables(c, p, np, (int)T, (int)F, &tab));
        for (int it = 0; it < p->num_major_iterations; it++)
            TC_TRY(dev_get_flags_pass(c, p, dvis, vis_kind, iter_flags, np, (int)T, (int)F,
                                      dout + p0 * T * F, iter_flags, &tab));
        tc_arena_release(c, mark);
    }
    return tc_stage_out_end(c, out, dout, (size_t)total, space);
}

// ------------------------------------------------------------------ U1 ------
int tc_uvcontsub(tc_context *c, const void *vis, const uint8_t *flags, int64_t ncp, int64_t T,
                 int64_t F, int major_cycles, int or_original_from_cycle, int taylor_degrees,
                 double sigma, uint8_t *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(ncp >= 0 && T >= 0 && F >= 0, "negative shape");
    TC_REQUIRE(taylor_degrees >= 0, "taylor_degrees must be >= 0");
    TC_REQUIRE(T < (1 << 30) && F < (1 << 30), "plane too large");
    int64_t total = ncp * T * F;
    const float2 *dvis; const u8 *dfl; u8 *dout;
    TC_TRY(tc_stage_in(c, (const float2 *)vis, (size_t)total, space, &dvis));
    TC_TRY(tc_stage_in(c, flags, (size_t)total, space, &dfl));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)total, space, &dout));
    if (total == 0) return tc_stage_out_end(c, out, dout, 0, space);
    // result_flags = flags.copy() (boolean semantics)
    TC_TRY(launch_norm_flags(c, dfl, dout, total));
    c->launches++;
    int K = taylor_degrees < (int)F ? taylor_degrees : (int)F;
    TC_REQUIRE(K <= TC_UV_MAXK, "taylor_degrees above %d is not supported", TC_UV_MAXK);
    double2 *tw; float2 *avg, *smooth; float *absres; int *unfl; double *med1;
    int64_t *lo, *hi;
    TC_TRY(tc_alloc(c, (size_t)F, &tw));
    TC_TRY(tc_alloc(c, (size_t)ncp * F, &avg));
    TC_TRY(tc_alloc(c, (size_t)ncp * F, &smooth));
    TC_TRY(tc_alloc(c, (size_t)total, &absres));
    TC_TRY(tc_alloc(c, (size_t)ncp, &unfl));
    TC_TRY(tc_alloc(c, (size_t)ncp, &med1));
    std::vector<int64_t> hlo((size_t)ncp), hhi((size_t)ncp);
    for (int64_t p = 0; p < ncp; p++) { hlo[p] = p * T * F; hhi[p] = (p + 1) * T * F; }
    TC_TRY(dev_upload_i64(c, hlo.data(), hlo.size(), &lo));
    TC_TRY(dev_upload_i64(c, hhi.data(), hhi.size(), &hi));
    TC_LAUNCH_NOSYNC(k_uv_twiddle, tc_blocks_for(F, 256), 256, 0, c->stream, tw, (int)F);
    c->launches++;
    for (int mi = 0; mi < major_cycles; mi++) {
        tc_prof_begin(c, TCP_UVCONTSUB);
        TC_CUDA(cudaMemsetAsync(unfl, 0, sizeof(int) * (size_t)ncp, c->stream));
        for (int64_t p0 = 0; p0 < ncp; p0 += 65535)       // gridDim.y is limited to 65535
            TC_LAUNCH(k_uv_mean, dim3(tc_blocks_for(F, 256), (unsigned)(ncp - p0 < 65535 ? ncp - p0 : 65535)), 256, 0,
                      c->stream, dvis, dout, (int)T, (int)F, avg, unfl, p0);
        TC_LAUNCH(k_uv_smooth, (unsigned)ncp, 256, 0, c->stream, avg, tw, (int)F, K, smooth);
        // grid (channel pairs, dumps, planes): at most 65535 in y and z per launch
        for (int64_t p0 = 0; p0 < ncp; p0 += 65535)
            for (int64_t t0 = 0; t0 < T; t0 += 65535) {
                const unsigned np_ = (unsigned)(ncp - p0 < 65535 ? ncp - p0 : 65535);
                const unsigned nt_ = (unsigned)(T - t0 < 65535 ? T - t0 : 65535);
                TC_LAUNCH_NOSYNC(k_uv_absres, dim3(tc_blocks_for((F + 1) / 2, 256), nt_, np_), 256, 0, c->stream,
                                 dvis, smooth, (int)T, (int)F, absres, (int)t0, (int)p0);
            }
        tc_prof_end(c);
        c->launches += 3;
        ChunkSelectArgs s;
        memset(&s, 0, sizeof(s));
        s.resid = absres; s.flags = dout; s.range_lo = lo; s.range_hi = hi;
        s.mode = CS_REPORT; s.take_abs = 1; s.sub = nullptr; s.skip_nan = 1; s.medians = med1;
        TC_TRY(launch_chunk_select(c, s, ncp, T * F));
        s.mode = CS_UVCONTSUB; s.sub = med1; s.medians = nullptr;
        s.uv_sigma = (float)sigma; s.uv_replace = mi < or_original_from_cycle ? 1 : 0; s.uv_unflagged = unfl;
        TC_TRY(launch_chunk_select(c, s, ncp, T * F));
    }
    TC_KERNEL_CHECK();
    return tc_stage_out_end(c, out, dout, (size_t)total, space);
}

// ------------------------------------------------------------------ K2 ------
static int fill_terms(StokesTerms *t, const int32_t *idx, const double *coef, int n, int ncorr)
{
    TC_REQUIRE(n >= 0 && n <= TC_MAX_STOKES, "at most %d stokes terms are supported", TC_MAX_STOKES);
    t->n = n;
    for (int k = 0; k < n; k++) {
        TC_REQUIRE(idx[2 * k] >= 0 && idx[2 * k] < ncorr && idx[2 * k + 1] >= 0 && idx[2 * k + 1] < ncorr,
                   "correlation index out of range");
        t->c1[k] = idx[2 * k]; t->c2[k] = idx[2 * k + 1];
        t->ar[k] = coef[4 * k]; t->ai[k] = coef[4 * k + 1]; t->s1[k] = coef[4 * k + 2]; t->s2[k] = coef[4 * k + 3];
    }
    return TC_OK;
}

static int stokes_common(tc_context *c, const void *vis, int64_t n, int ncorr, const StokesTerms &pol,
                         const StokesTerms &unpol, int with_unpol, void *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(ncorr >= 1 && ncorr <= 8, "between 1 and 8 correlations are supported");
    const float2 *dvis; float2 *dout;
    TC_TRY(tc_stage_in(c, (const float2 *)vis, (size_t)n * ncorr, space, &dvis));
    TC_TRY(tc_stage_out_begin(c, (float2 *)out, (size_t)n, space, &dout));
    if (n) {
        tc_prof_begin(c, TCP_ELEMENTWISE);
        TC_LAUNCH_NOSYNC(k_stokes, tc_blocks_for(n, 256), 256, 0, c->stream, dvis, n, ncorr, pol, unpol,
                         with_unpol, dout);
        tc_prof_end(c);
        c->launches++;
        TC_KERNEL_CHECK();
    }
    return tc_stage_out_end(c, (float2 *)out, dout, (size_t)n, space);
}

int tc_polarised_intensity(tc_context *c, const void *vis, int64_t nrowchan, int ncorr,
                           const int32_t *pol_idx, const double *pol_coef, int npol, void *out, int space)
{
    StokesTerms pol, unpol;
    memset(&unpol, 0, sizeof(unpol));
    TC_TRY(fill_terms(&pol, pol_idx, pol_coef, npol, ncorr));
    return stokes_common(c, vis, nrowchan, ncorr, pol, unpol, 0, out, space);
}

int tc_unpolarised_intensity(tc_context *c, const void *vis, int64_t nrowchan, int ncorr,
                             const int32_t *unpol_idx, const double *unpol_coef, int nunpol,
                             const int32_t *pol_idx, const double *pol_coef, int npol, void *out, int space)
{
    StokesTerms pol, unpol;
    TC_TRY(fill_terms(&pol, pol_idx, pol_coef, npol, ncorr));
    TC_TRY(fill_terms(&unpol, unpol_idx, unpol_coef, nunpol, ncorr));
    return stokes_common(c, vis, nrowchan, ncorr, pol, unpol, 1, out, space);
}

// -------------------------------------------------------------- P1 / P2 -----
static int upload_i32(tc_context *c, const int32_t *h, size_t n, int32_t **d)
{
    TC_TRY(tc_alloc(c, n ? n : 1, d));
    if (n) TC_TRY(tc_upload_small(c, h, n * sizeof(int32_t), *d));
    return TC_OK;
}

// window defaults before a pack: only the (baseline, time) slots that no row writes
static int pack_fill(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow, int64_t nbl,
                     int64_t ntime, int64_t ncorr_win, int64_t nchan, float2 *dvw, u8 *dfw)
{
    const int64_t nslot = nbl * ntime, nwin = nslot * ncorr_win * nchan;
    if (nwin == 0) return TC_OK;
    if (nslot >= ((int64_t)1 << 31) || TC_ENV_FLAG("TC_PACK_FULL_FILL")) {
        TC_LAUNCH_NOSYNC(k_fill_windows, tc_blocks_for(nwin, 256), 256, 0, c->stream, dvw, dfw, nwin);
        c->launches++;
        return TC_OK;
    }
    std::vector<u8> covered((size_t)nslot, 0);
    for (int64_t r = 0; r < nrow; r++)
        if (row_bl[r] >= 0) covered[(size_t)row_bl[r] * ntime + row_t[r]] = 1;
    std::vector<int32_t> open_slots;
    for (int64_t s = 0; s < nslot; s++)
        if (!covered[(size_t)s]) open_slots.push_back((int32_t)s);
    if (open_slots.empty()) return TC_OK;
    int32_t *dslots;
    TC_TRY(tc_alloc(c, open_slots.size(), &dslots));
    TC_TRY(tc_upload_small(c, open_slots.data(), open_slots.size() * sizeof(int32_t), dslots));
    const int64_t work = (int64_t)open_slots.size() * ncorr_win * nchan;
    TC_LAUNCH_NOSYNC(k_fill_slots, tc_blocks_for(work, 256), 256, 0, c->stream, dslots, (int64_t)open_slots.size(),
                     (int)ncorr_win, (int)ntime, (int)nchan, dvw, dfw);
    c->launches++;
    return TC_OK;
}

int tc_pack(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow, const void *vis,
            const uint8_t *flags, int64_t nchan, int64_t ncorr, int64_t ntime, int64_t nbl,
            void *vis_win, uint8_t *flag_win, int fill, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(nrow >= 0 && nchan >= 0 && ncorr >= 0 && ntime >= 0 && nbl >= 0, "negative shape");
    for (int64_t r = 0; r < nrow; r++) {
        TC_REQUIRE(row_bl[r] < nbl, "row %lld: baseline slot %d out of range", (long long)r, row_bl[r]);
        TC_REQUIRE(row_bl[r] < 0 || (row_t[r] >= 0 && row_t[r] < ntime), "row %lld: time index %d out of range",
                   (long long)r, row_t[r]);
    }
    int64_t nin = nrow * nchan * ncorr, nwin = nbl * ncorr * ntime * nchan;
    int32_t *dbl, *dt;
    TC_TRY(upload_i32(c, row_bl, (size_t)nrow, &dbl));
    TC_TRY(upload_i32(c, row_t, (size_t)nrow, &dt));
    const float2 *dvis = nullptr; const u8 *dfl = nullptr; float2 *dvw = nullptr; u8 *dfw = nullptr;
    if (vis) {
        TC_TRY(tc_stage_in(c, (const float2 *)vis, (size_t)nin, space, &dvis));
        TC_TRY(tc_stage_out_begin(c, (float2 *)vis_win, (size_t)nwin, space, &dvw));
        if (space == TC_HOST && !fill)
            TC_CUDA(cudaMemcpyAsync(dvw, vis_win, (size_t)nwin * 8, cudaMemcpyHostToDevice, c->stream));
    }
    if (flags) {
        TC_TRY(tc_stage_in(c, flags, (size_t)nin, space, &dfl));
        TC_TRY(tc_stage_out_begin(c, flag_win, (size_t)nwin, space, &dfw));
        if (space == TC_HOST && !fill)
            TC_CUDA(cudaMemcpyAsync(dfw, flag_win, (size_t)nwin, cudaMemcpyHostToDevice, c->stream));
    }
    tc_prof_begin(c, TCP_PACK);
    if (fill && nwin) TC_TRY(pack_fill(c, row_bl, row_t, nrow, nbl, ntime, ncorr, nchan, dvw, dfw));
    if (nin) {
        bool c4 = ncorr == 4 && nchan % 4 == 0 && (((uintptr_t)dvis | (uintptr_t)dfl | (uintptr_t)dvw | (uintptr_t)dfw) & 15) == 0;
        if (c4) {
            int64_t work = nrow * (nchan / 4);
            TC_LAUNCH_NOSYNC(k_pack_c4, tc_blocks_for(work, 256), 256, 0, c->stream, dbl, dt, nrow,
                             (const float4 *)dvis, (const uint4 *)dfl, (int)(nchan / 4), (int)ntime, dvw,
                             (uint32_t *)dfw);
        } else {
            TC_LAUNCH_NOSYNC(k_pack, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow, dvis, dfl,
                             (int)nchan, (int)ncorr, (int)ntime, dvw, dfw);
        }
        c->launches++;
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    if (vis && space == TC_HOST)
        TC_CUDA(cudaMemcpyAsync(vis_win, dvw, (size_t)nwin * 8, cudaMemcpyDeviceToHost, c->stream));
    if (flags && space == TC_HOST)
        TC_CUDA(cudaMemcpyAsync(flag_win, dfw, (size_t)nwin, cudaMemcpyDeviceToHost, c->stream));
    if (space == TC_HOST) TC_CUDA(cudaStreamSynchronize(c->stream));
    return TC_OK;
}

int tc_unpack(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow, const void *window,
              int elem_size, int64_t nchan, int64_t ncorr, int64_t ntime, int64_t nbl, void *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(elem_size == 1 || elem_size == 8, "elem_size must be 1 (flags) or 8 (complex64)");
    for (int64_t r = 0; r < nrow; r++) {
        TC_REQUIRE(row_bl[r] < nbl, "row %lld: baseline slot %d out of range", (long long)r, row_bl[r]);
        TC_REQUIRE(row_bl[r] < 0 || (row_t[r] >= 0 && row_t[r] < ntime), "row %lld: time index out of range", (long long)r);
    }
    int64_t nout = nrow * nchan * ncorr, nwin = nbl * ncorr * ntime * nchan;
    int32_t *dbl, *dt;
    TC_TRY(upload_i32(c, row_bl, (size_t)nrow, &dbl));
    TC_TRY(upload_i32(c, row_t, (size_t)nrow, &dt));
    const char *dwin; char *dout;
    TC_TRY(tc_stage_in(c, (const char *)window, (size_t)nwin * elem_size, space, &dwin));
    TC_TRY(tc_stage_out_begin(c, (char *)out, (size_t)nout * elem_size, space, &dout));
    if (nout) {
        tc_prof_begin(c, TCP_PACK);
        const bool al16 = ((((uintptr_t)dwin) | ((uintptr_t)dout)) & 15) == 0;
        if (elem_size == 8 && ncorr == 4 && al16)
            TC_LAUNCH_NOSYNC(k_unpack_vis_c4, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow,
                             (const float2 *)dwin, (int)nchan, (int)ntime, (float4 *)dout);
        else if (elem_size == 1 && ncorr == 4 && (nchan & 3) == 0 && al16)
            TC_LAUNCH_NOSYNC(k_unpack_flags_c4<0>, tc_blocks_for(nrow * (nchan / 4), 256), 256, 0, c->stream, dbl, dt,
                             nrow, (const uint32_t *)dwin, (int)(nchan / 4), (int)ntime, (uint4 *)dout);
        else if (elem_size == 8)
            TC_LAUNCH_NOSYNC(k_unpack<float2>, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow,
                             (const float2 *)dwin, (int)nchan, (int)ncorr, (int)ntime, (float2 *)dout);
        else
            TC_LAUNCH_NOSYNC(k_unpack<u8>, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow,
                             (const u8 *)dwin, (int)nchan, (int)ncorr, (int)ntime, (u8 *)dout);
        tc_prof_end(c);
        c->launches++;
        TC_KERNEL_CHECK();
    }
    return tc_stage_out_end(c, (char *)out, dout, (size_t)nout * elem_size, space);
}

int tc_unpack_flags_broadcast(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow,
                              const uint8_t *window, int64_t nchan, int64_t ncorr_win, int64_t ncorr_out,
                              int64_t ntime, int64_t nbl, uint8_t *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(ncorr_win >= 1 && ncorr_out >= 1, "correlation counts must be >= 1");
    for (int64_t r = 0; r < nrow; r++) {
        TC_REQUIRE(row_bl[r] < nbl, "row %lld: baseline slot %d out of range", (long long)r, row_bl[r]);
        TC_REQUIRE(row_bl[r] < 0 || (row_t[r] >= 0 && row_t[r] < ntime), "row %lld: time index out of range", (long long)r);
    }
    int64_t nout = nrow * nchan * ncorr_out, nwin = nbl * ncorr_win * ntime * nchan;
    int32_t *dbl, *dt;
    TC_TRY(upload_i32(c, row_bl, (size_t)nrow, &dbl));
    TC_TRY(upload_i32(c, row_t, (size_t)nrow, &dt));
    const u8 *dwin; u8 *dout;
    TC_TRY(tc_stage_in(c, window, (size_t)nwin, space, &dwin));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)nout, space, &dout));
    if (nout) {
        tc_prof_begin(c, TCP_PACK);
        const bool al16 = ((((uintptr_t)dwin) | ((uintptr_t)dout)) & 15) == 0;
        if (ncorr_win == 4 && ncorr_out == 4 && (nchan & 3) == 0 && al16)
            TC_LAUNCH_NOSYNC(k_unpack_flags_c4<1>, tc_blocks_for(nrow * (nchan / 4), 256), 256, 0, c->stream, dbl, dt,
                             nrow, (const uint32_t *)dwin, (int)(nchan / 4), (int)ntime, (uint4 *)dout);
        else if (ncorr_win == 1 && ncorr_out == 4 && (nchan & 15) == 0 && al16)
            TC_LAUNCH_NOSYNC(k_unpack_flags_c1_to4, tc_blocks_for(nrow * (nchan / 16), 256), 256, 0, c->stream, dbl, dt,
                             nrow, (const uint4 *)dwin, (int)(nchan / 16), (int)ntime, (uint4 *)dout);
        else
            TC_LAUNCH_NOSYNC(k_unpack_any_corr, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow,
                             dwin, (int)nchan, (int)ncorr_win, (int)ntime, (int)ncorr_out, dout);
        tc_prof_end(c);
        c->launches++;
        TC_KERNEL_CHECK();
    }
    return tc_stage_out_end(c, out, dout, (size_t)nout, space);
}

int tc_unpack_flags_any_corr(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow,
                             const uint8_t *window, int64_t nchan, int64_t ncorr, int64_t ntime, int64_t nbl,
                             uint8_t *out, int space)
{
    return tc_unpack_flags_broadcast(c, row_bl, row_t, nrow, window, nchan, ncorr, ncorr, ntime, nbl, out, space);
}

// ---------------------------------------------------------- K2 + N2 + P1 ----
int tc_stokes_pack(tc_context *c, const int32_t *row_bl, const int32_t *row_t, int64_t nrow, const void *vis,
                   const uint8_t *flags, int64_t nchan, int64_t ncorr, int64_t ntime, int64_t nbl,
                   const int32_t *unpol_idx, const double *unpol_coef, int nunpol, const int32_t *pol_idx,
                   const double *pol_coef, int npol, void *vis_win, uint8_t *flag_win, int fill, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(nrow >= 0 && nchan >= 0 && ntime >= 0 && nbl >= 0, "negative shape");
    TC_REQUIRE(ncorr >= 1 && ncorr <= 8, "between 1 and 8 correlations are supported");
    TC_REQUIRE(vis && flags && vis_win && flag_win, "null array");
    StokesTerms pol, unpol;
    memset(&unpol, 0, sizeof(unpol));
    TC_TRY(fill_terms(&pol, pol_idx, pol_coef, npol, (int)ncorr));
    if (nunpol > 0) TC_TRY(fill_terms(&unpol, unpol_idx, unpol_coef, nunpol, (int)ncorr));
    for (int64_t r = 0; r < nrow; r++) {
        TC_REQUIRE(row_bl[r] < nbl, "row %lld: baseline slot %d out of range", (long long)r, row_bl[r]);
        TC_REQUIRE(row_bl[r] < 0 || (row_t[r] >= 0 && row_t[r] < ntime), "row %lld: time index %d out of range",
                   (long long)r, row_t[r]);
    }
    int64_t nin = nrow * nchan * ncorr, nwin = nbl * ntime * nchan;
    int32_t *dbl, *dt;
    TC_TRY(upload_i32(c, row_bl, (size_t)nrow, &dbl));
    TC_TRY(upload_i32(c, row_t, (size_t)nrow, &dt));
    const float2 *dvis; const u8 *dfl; float2 *dvw; u8 *dfw;
    TC_TRY(tc_stage_in(c, (const float2 *)vis, (size_t)nin, space, &dvis));
    TC_TRY(tc_stage_in(c, flags, (size_t)nin, space, &dfl));
    TC_TRY(tc_stage_out_begin(c, (float2 *)vis_win, (size_t)nwin, space, &dvw));
    TC_TRY(tc_stage_out_begin(c, flag_win, (size_t)nwin, space, &dfw));
    if (space == TC_HOST && !fill) {
        TC_CUDA(cudaMemcpyAsync(dvw, vis_win, (size_t)nwin * 8, cudaMemcpyHostToDevice, c->stream));
        TC_CUDA(cudaMemcpyAsync(dfw, flag_win, (size_t)nwin, cudaMemcpyHostToDevice, c->stream));
    }
    tc_prof_begin(c, TCP_PACK);
    if (fill && nwin) TC_TRY(pack_fill(c, row_bl, row_t, nrow, nbl, ntime, 1, nchan, dvw, dfw));
    if (nin) {
        TC_LAUNCH_NOSYNC(k_stokes_pack, tc_blocks_for(nrow * nchan, 256), 256, 0, c->stream, dbl, dt, nrow, dvis, dfl,
                         (int)nchan, (int)ncorr, (int)ntime, pol, unpol, nunpol > 0 ? 1 : 0, dvw, dfw);
        c->launches++;
    }
    tc_prof_end(c);
    TC_KERNEL_CHECK();
    TC_TRY(tc_stage_out_end(c, (float2 *)vis_win, dvw, (size_t)nwin, space));
    return tc_stage_out_end(c, flag_win, dfw, (size_t)nwin, space);
}

// ------------------------------------------------------------------ W1 ------
int tc_window_counts(tc_context *c, const uint8_t *flags, int64_t nbl, int64_t ncorr, int64_t T, int64_t F,
                     uint64_t *bl_counts, uint64_t *chan_counts, int space)
{
    TC_TRY(tc_begin(c));
    int64_t total = nbl * ncorr * T * F;
    const u8 *dfl;
    TC_TRY(tc_stage_in(c, flags, (size_t)total, space, &dfl));
    unsigned long long *dbl, *dch;
    TC_TRY(tc_alloc(c, (size_t)(nbl > 0 ? nbl : 1), &dbl));
    TC_TRY(tc_alloc(c, (size_t)(F > 0 ? F : 1), &dch));
    TC_CUDA(cudaMemsetAsync(dbl, 0, sizeof(unsigned long long) * (size_t)(nbl > 0 ? nbl : 1), c->stream));
    TC_CUDA(cudaMemsetAsync(dch, 0, sizeof(unsigned long long) * (size_t)(F > 0 ? F : 1), c->stream));
    if (total) {
        int64_t rows_per_bl = ncorr * T;
        tc_prof_begin(c, TCP_STATS);
        if ((F & 15) == 0 && (((uintptr_t)dfl) & 15) == 0 && rows_per_bl < ((int64_t)1 << 30)) {
            // 16-byte row reads, column sums added to the channel counts by 64-bit reductions: one launch
            const int F16 = (int)(F / 16);
            const unsigned segs = tc_blocks_for(rows_per_bl, TC_WC_ROWS), tiles = tc_blocks_for(F16, 256);
            TC_REQUIRE(tiles <= 65535, "too many channels");
            for (int64_t b0 = 0; b0 < nbl; b0 += 65535) {
                const int64_t nb = nbl - b0 < 65535 ? nbl - b0 : 65535;
                TC_LAUNCH(k_window_counts_v16, dim3(segs, tiles, (unsigned)nb), 256, 0, c->stream,
                          (const uint4 *)(dfl + b0 * rows_per_bl * F), (int)rows_per_bl, F16, dch, dbl + b0);
                c->launches++;
            }
        } else {
            int rows_per_seg = 64;
            // keep the per-thread column sum below 2^32 (flag bytes are <= 255)
            unsigned segs = tc_blocks_for(rows_per_bl, rows_per_seg);
            for (int64_t b0 = 0; b0 < nbl; b0 += 65535) {
                int64_t nb = nbl - b0 < 65535 ? nbl - b0 : 65535;
                TC_LAUNCH(k_window_counts, dim3(segs, (unsigned)nb), 256, 0, c->stream, dfl + b0 * rows_per_bl * F,
                          rows_per_bl, rows_per_seg, (int)F, dbl + b0, dch);
                c->launches++;
            }
        }
        tc_prof_end(c);
        TC_KERNEL_CHECK();
    }
    if (nbl) TC_CUDA(cudaMemcpyAsync(bl_counts, dbl, sizeof(uint64_t) * (size_t)nbl, cudaMemcpyDeviceToHost, c->stream));
    if (F) TC_CUDA(cudaMemcpyAsync(chan_counts, dch, sizeof(uint64_t) * (size_t)F, cudaMemcpyDeviceToHost, c->stream));
    TC_CUDA(cudaStreamSynchronize(c->stream));
    return TC_OK;
}

// =========================================================================
// stage-level entry points (parity tests mirror tricolour/tests/test_flagging.py)
// =========================================================================
int tc_stage_average_freq(tc_context *c, const void *vis, int vis_kind, const uint8_t *flags, int64_t ncp,
                          int64_t T, int64_t F, int64_t factor, float *out_data, uint8_t *out_flags, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(factor >= 1, "factor must be >= 1");
    int64_t Fa = (F + factor - 1) / factor;
    int64_t nin = ncp * T * F, nout = ncp * T * Fa;
    size_t esz = vis_kind == TC_VIS_COMPLEX64 ? 8 : 4;
    const char *dvis; const u8 *dfl; float *dd; u8 *df;
    TC_TRY(tc_stage_in(c, (const char *)vis, (size_t)nin * esz, space, &dvis));
    TC_TRY(tc_stage_in(c, flags, (size_t)nin, space, &dfl));
    TC_TRY(tc_stage_out_begin(c, out_data, (size_t)nout, space, &dd));
    TC_TRY(tc_stage_out_begin(c, out_flags, (size_t)nout, space, &df));
    if (nout) {
        TC_LAUNCH_NOSYNC(k_prep, tc_blocks_for(nout, 256), 256, 0, c->stream, (const void *)dvis, vis_kind, dfl, nout,
                         (int)F, (int)Fa, (int)factor, dd, df);
        c->launches++;
        TC_KERNEL_CHECK();
    }
    TC_TRY(tc_stage_out_end(c, out_data, dd, (size_t)nout, space));
    return tc_stage_out_end(c, out_flags, df, (size_t)nout, space);
}

// helper: stage (ncp,T,F) data+flags and build their FT transposes
struct StageIn {
    const float *d_TF; const u8 *f_TF; float *d_FT; u8 *f_FT;
};
static int stage_planes(tc_context *c, const float *data, const u8 *flags, int64_t ncp, int64_t T, int64_t F,
                        int space, StageIn *s)
{
    int64_t n = ncp * T * F;
    TC_TRY(tc_stage_in(c, data, (size_t)n, space, &s->d_TF));
    s->f_TF = nullptr; s->f_FT = nullptr;
    TC_TRY(tc_alloc(c, (size_t)n, &s->d_FT));
    TC_TRY(launch_transpose<float>(c, s->d_TF, s->d_FT, ncp, (int)T, (int)F));
    if (flags) {
        const u8 *raw;
        u8 *norm;
        TC_TRY(tc_stage_in(c, flags, (size_t)n, space, &raw));
        TC_TRY(tc_alloc(c, (size_t)n, &norm));
        if (n) {
            TC_TRY(launch_norm_flags(c, raw, norm, n));
            c->launches++;
        }
        s->f_TF = norm;
        TC_TRY(tc_alloc(c, (size_t)n, &s->f_FT));
        TC_TRY(launch_transpose<u8>(c, s->f_TF, s->f_FT, ncp, (int)T, (int)F));
    }
    return TC_OK;
}

int tc_stage_time_median(tc_context *c, const float *data, const uint8_t *flags, int64_t ncp, int64_t T,
                         int64_t F, float *out_data, uint8_t *out_flags, int space)
{
    TC_TRY(tc_begin(c));
    StageIn s;
    TC_TRY(stage_planes(c, data, flags, ncp, T, F, space, &s));
    float *dd; u8 *df;
    TC_TRY(tc_stage_out_begin(c, out_data, (size_t)(ncp * F), space, &dd));
    TC_TRY(tc_stage_out_begin(c, out_flags, (size_t)(ncp * F), space, &df));
    LineMedianArgs m;
    memset(&m, 0, sizeof(m));
    m.data = s.d_FT; m.flags = s.f_FT; m.nlines = ncp * F; m.ninner = F; m.outer_stride = T * F;
    m.inner_stride = T; m.elem_stride = 1; m.n = (int)T; m.mode = LM_TIME_MEDIAN; m.out = dd; m.out_flags = df;
    TC_TRY(launch_line_median(c, m, (int)T));
    TC_TRY(tc_stage_out_end(c, out_data, dd, (size_t)(ncp * F), space));
    return tc_stage_out_end(c, out_flags, df, (size_t)(ncp * F), space);
}

int tc_stage_chunk_median_abs(tc_context *c, const float *data, const uint8_t *flags, int64_t ncp, int64_t T,
                              int64_t F, const int64_t *chunk_ends, int nce, double *out, int space)
{
    TC_TRY(tc_begin(c));
    StageIn s;
    TC_TRY(stage_planes(c, data, flags, ncp, T, F, space, &s));
    int nch = nce - 1;
    int64_t *lo, *hi, mr;
    TC_TRY(dev_make_ranges(c, ncp, T, F, chunk_ends, nce, &lo, &hi, &mr));
    double *dm;
    TC_TRY(tc_stage_out_begin(c, out, (size_t)(ncp * nch), space, &dm));
    ChunkSelectArgs a;
    memset(&a, 0, sizeof(a));
    a.resid = s.d_FT; a.flags = s.f_FT; a.range_lo = lo; a.range_hi = hi; a.mode = CS_REPORT; a.take_abs = 1;
    a.medians = dm;
    TC_TRY(launch_chunk_select(c, a, ncp * nch, mr));
    return tc_stage_out_end(c, out, dm, (size_t)(ncp * nch), space);
}

int tc_stage_masked_filter(tc_context *c, const float *data, const uint8_t *flags, int64_t ncp, int64_t T,
                           int64_t F, int64_t r0, int64_t r1, float *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(r0 >= 0 && r1 >= 0, "radii must be >= 0");
    int64_t n = ncp * T * F;
    StageIn s;
    TC_TRY(stage_planes(c, data, flags, ncp, T, F, space, &s));
    BgWork w;
    TC_TRY(dev_bg_work_alloc(c, n, true, &w));
    TC_TRY(tc_copy_d2d(c, w.fl_FT, s.f_FT, (int64_t)n));
    float *o_FT, *o_TF;
    TC_TRY(tc_alloc(c, (size_t)n, &o_FT));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &o_TF));
    if (n) {
        TC_TRY(dev_masked_filter(c, ncp, (int)T, (int)F, s.d_TF, s.d_FT, w, r0, r1, 0, o_FT));
        TC_TRY(launch_transpose<float>(c, o_FT, o_TF, ncp, (int)F, (int)T));
    }
    return tc_stage_out_end(c, out, o_TF, (size_t)n, space);
}

int tc_stage_interp_nans(tc_context *c, const float *data, int64_t ncp, int64_t T, int64_t F, float *out, int space)
{
    TC_TRY(tc_begin(c));
    int64_t n = ncp * T * F;
    const float *din; float *dout;
    TC_TRY(tc_stage_in(c, data, (size_t)n, space, &din));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &dout));
    if (n) {
        int *rv;
        TC_TRY(tc_alloc(c, (size_t)n, &rv));
        TC_LAUNCH(k_interp_nans_rows, tc_blocks_for(ncp * T * 32, 128), 128, 0, c->stream, din,
                  (const float *)nullptr, dout, rv, ncp * T, (int)F);
        c->launches++;
        TC_KERNEL_CHECK();
    }
    return tc_stage_out_end(c, out, dout, (size_t)n, space);
}

int tc_stage_background2d(tc_context *c, const float *data, const uint8_t *flags, int64_t ncp, int64_t T,
                          int64_t F, int iterations, const int64_t *radii, double reject_threshold,
                          const int64_t *chunk_ends, int nce, float *out, int space)
{
    TC_TRY(tc_begin(c));
    TC_REQUIRE(iterations >= 0 && nce >= 1, "bad background parameters");
    int64_t n = ncp * T * F;
    StageIn s;
    TC_TRY(stage_planes(c, data, flags, ncp, T, F, space, &s));
    int64_t *lo, *hi, mr;
    TC_TRY(dev_make_ranges(c, ncp, T, F, chunk_ends, nce, &lo, &hi, &mr));
    float *o_FT, *o_TF;
    TC_TRY(tc_alloc(c, (size_t)n, &o_FT));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &o_TF));
    if (n)
        TC_TRY(dev_background2d(c, ncp, (int)T, (int)F, s.d_TF, s.d_FT, s.f_TF, s.f_FT, iterations, radii,
                                reject_threshold, lo, hi, nce - 1, mr, o_FT, o_TF, nullptr));
    return tc_stage_out_end(c, out, o_TF, (size_t)n, space);
}

int tc_stage_sum_threshold(tc_context *c, const float *data, const uint8_t *flags, int64_t ncp, int64_t T,
                           int64_t F, int axis, const int64_t *windows, const double *tf, const float *scale,
                           int nwin, double outlier_nsigma, const int64_t *chunk_ends, int nce, uint8_t *out,
                           int space)
{
    TC_TRY(tc_begin(c));
    if (axis != 0 && axis != 1) return tc_fail(TC_ERR_VALUE, "axis must be 0 or 1");
    int64_t n = ncp * T * F;
    StageIn s;
    TC_TRY(stage_planes(c, data, flags, ncp, T, F, space, &s));
    int64_t dflt[2] = {0, axis == 0 ? T : F};
    if (!chunk_ends) { chunk_ends = dflt; nce = 2; }
    u8 *o_nat, *o_TF;
    TC_TRY(tc_alloc(c, (size_t)n, &o_nat));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)n, space, &o_TF));
    if (n) {
        TC_TRY(dev_sum_threshold(c, ncp, (int)T, (int)F, axis, s.d_TF, s.d_FT, s.f_TF, s.f_FT, nullptr, windows, tf,
                                 scale, nwin, outlier_nsigma, chunk_ends, nce, axis == 0 ? o_TF : o_nat));
        if (axis == 1) TC_TRY(launch_transpose<u8>(c, o_nat, o_TF, ncp, (int)F, (int)T));
    }
    return tc_stage_out_end(c, out, o_TF, (size_t)n, space);
}

int tc_stage_combine_unaverage(tc_context *c, const uint8_t *spec, const uint8_t *time_f, const uint8_t *freq_f,
                               int64_t ncp, int64_t T, int64_t Fa, int64_t F, int64_t time_extend,
                               int64_t freq_extend, int64_t average_freq, double flag_all_time_frac,
                               double flag_all_freq_frac, uint8_t *out, int space)
{
    TC_TRY(tc_begin(c));
    int64_t N = ncp * T * Fa, NF = ncp * T * F;
    const u8 *ds, *dt, *df;
    TC_TRY(tc_stage_in(c, spec, (size_t)(ncp * Fa), space, &ds));
    TC_TRY(tc_stage_in(c, time_f, (size_t)N, space, &dt));
    TC_TRY(tc_stage_in(c, freq_f, (size_t)N, space, &df));
    u8 *c1, *dout;
    TC_TRY(tc_alloc(c, (size_t)N, &c1));
    TC_TRY(tc_stage_out_begin(c, out, (size_t)NF, space, &dout));
    if (NF)
        TC_TRY(dev_combine_flags(c, ncp, (int)T, (int)Fa, (int)F, (int)average_freq, (int)time_extend,
                                 (int)freq_extend, flag_all_time_frac, flag_all_freq_frac, ds, dt, df, c1,
                                 (const void *)nullptr, 0, dout, (u8 *)nullptr));
    return tc_stage_out_end(c, out, dout, (size_t)NF, space);
}

}  // extern "C"
