// k_filter5t.cuh -- the B5 filter of k_filter5.cuh for the second filtered axis with
// its global traffic on the TMA unit ("B5T"; same reference: _box_gaussian_filter1d
// flagging.py:362-419, masked_gaussian_filter 469-513).
//
// ncu on k_box5b (profiles/r02_box5_*): the kernel is bound by the LSU data pipe
// (85-95 % of its wavefront peak), and 45 % of those wavefronts are not the delay
// rings but the warp's global accesses -- 16-byte chunks of eight lines coming in
// and 16-byte runs of four adjacent lines going out cost about one wavefront per
// 32-byte sector.  Here none of that passes through the LSU:
//
//  * input: one cp.async.bulk.tensor (TMA) load per warp and iteration brings the next
//    16 samples of the value and the weight array of the warp's 4 lines -- a
//    (16 samples, 4 lines, 2 arrays) box of a 3-D tensor map -- straight into the
//    rings of the eight pass-0 chains, two groups ahead, completing on an mbarrier;
//    out-of-range samples and lines arrive as zeros, which is what the filter wants;
//  * the unfiltered samples of the residual come the same way (a (16, 4) box);
//  * output: the lanes park the 16 x 4 finished samples in a small tile and one TMA
//    store writes it ((4 lines, 16 samples, 1 plane) box, or (16, 4) when the output
//    is line-contiguous); samples beyond n are clipped by the tensor map, so the drain
//    needs no predicates (the one tile that starts before sample 0 is stored directly:
//    bulk tensor stores, unlike loads, fault on negative coordinates --
//    tools/probes/tma_probe.cu).
//
// Ring layout (all 32 chains, so that every lane runs the same code): group block of
// 2048 bytes = [lane][4 x 16-byte chunks], chunk index XORed with bits 1-2 of the lane
// -- the hardware's 64-byte swizzle pattern (address bits 4-5 ^= bits 7-8), so the box
// the TMA unit writes for lanes 0-7 and the conflict-free 16-byte reads of all lanes
// agree.  The arithmetic (accumulators, rounding, order of operations) is B5's.
#pragma once
#include "k_filter5.cuh"

#ifndef TC_EMU
#include <cuda.h>

#define B5T_GROUP_BYTES 2048
#define B5T_YSTAGE_BYTES (2 * B5_GQ * B5_STAGE_ROW * 16)     // 1536
#define B5T_D2_BYTES (2 * 256)
#define B5T_OUT_BYTES 1024                                   // 3 x 256 output tiles, 2 mbarriers at +768
#define B5T_FIXED_BYTES (B5T_YSTAGE_BYTES + B5T_D2_BYTES + B5T_OUT_BYTES)

__device__ __forceinline__ unsigned b5t_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void b5t_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void b5t_mbar_expect(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void b5t_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void b5t_load3(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void b5t_load2(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void b5t_store3(const CUtensorMap *map, int c0, int c1, int c2, unsigned src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void b5t_store2(const CUtensorMap *map, int c0, int c1, unsigned src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void b5t_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct B5TMaps {
    CUtensorMap in;      // (sample, line, array): the first axis's (value, weight) pair, line-contiguous, 64-byte swizzle
    CUtensorMap d2;      // (sample, line): the unfiltered samples, line-contiguous
    CUtensorMap out;     // (line in plane, sample, plane), or (sample, line) when the output is line-contiguous
};

template <bool ODD, int MODE_OUT>
__global__ void __launch_bounds__(128)
k_box5t(FilterArgs a, const __grid_constant__ B5TMaps maps)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int G = B5_G, GQ = B5_GQ;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int n = a.n, r2 = 2 * a.r, r4 = 4 * a.r;
    const int ngr = (r2 + 4 * G - 1) / G;                   // ring groups: 2r + 3G samples (the loads run two groups ahead)
    const int64_t grp = (int64_t)blockIdx.x * nwb + wib;
    if (grp * 4 >= a.nlines) return;
    unsigned char *wbase = smem_raw + (size_t)wib * ((size_t)ngr * B5T_GROUP_BYTES + B5T_FIXED_BYTES);
    const unsigned ring_s = b5t_smem_u32(wbase);
    unsigned char *ystage_p = wbase + (size_t)ngr * B5T_GROUP_BYTES;
    const unsigned ystage_s = ring_s + ngr * B5T_GROUP_BYTES;
    const unsigned d2_s = ystage_s + B5T_YSTAGE_BYTES;
    const unsigned out_s = d2_s + B5T_D2_BYTES;
    const unsigned bar_s = out_s + 768;
    const float *d2_p = reinterpret_cast<const float *>(ystage_p + B5T_YSTAGE_BYTES);
    float *out_p = reinterpret_cast<float *>(ystage_p + B5T_YSTAGE_BYTES + B5T_D2_BYTES);

    const int pass = lane >> 3, sidx = lane & 7;
    const int ylo = pass == 2 ? r2 : -0x40000000, yhi = pass == 0 ? n + r2 : 0x7fffffff;
    // byte offsets of this lane's four chunks inside a group block, and of the chunks its emits go to
    const unsigned sw = (lane >> 1) & 3;
    unsigned eo[GQ], yo[GQ];
#pragma unroll
    for (int q = 0; q < GQ; q++) {
        eo[q] = lane * 64 + ((q ^ sw) << 4);
        const unsigned t = lane + 8;
        yo[q] = pass < 3 ? t * 64 + ((q ^ ((t >> 1) & 3)) << 4) : (unsigned)((q * B5_STAGE_ROW + sidx) * 16);
    }
    // leaving samples of local group j start at element jG - 2r: chunk phase c of the first vector that is loaded
    int le = (-pass * G - r2) % (ngr * G);
    if (le < 0) le += ngr * G;
    le += ODD ? 2 : 0;
    if (le >= ngr * G) le -= ngr * G;
    const int lc = (le >> 2) & 3;
    unsigned lo[GQ];
    bool lhi[GQ];
#pragma unroll
    for (int k = 0; k < GQ; k++) {
        lo[k] = lane * 64 + ((((lc + k) & 3) ^ sw) << 4);
        lhi[k] = lc + k >= 4;
    }
    int lg = le >> 4;                                        // group the first leaving vector lives in
    int eg = (ngr - pass % ngr) % ngr;                       // group of the entering samples (local group -pass)

    // drain role: 2 samples (4 dq + 2 dh + {0, 1}) of line dl, value and weight
    const int dl = lane & 3, dq = lane >> 3, dh = (lane >> 2) & 1;
    const int ds = 4 * dq + 2 * dh;
    const int64_t line0 = grp * 4;
    const int plane = (int)(line0 / a.nj), lip0 = (int)(line0 - (int64_t)plane * a.nj);
    B2Div dv;
    dv.init(a.div);

    // ---- prologue: barriers, zeroed rings, the first two groups on their way
    for (int v = lane; v < ngr * (B5T_GROUP_BYTES / 16); v += 32) reinterpret_cast<uint4 *>(wbase)[v] = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0) {
        b5t_mbar_init(bar_s, 1);
        b5t_mbar_init(bar_s + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    b5t_fence_async();
    __syncwarp();
    const int niter = (n + r4 + G - 1) / G + 3;
    const unsigned tx = 512u + (MODE_OUT == FOUT_RESID ? 256u : 0u);
    const unsigned ring_end = ring_s + (unsigned)ngr * B5T_GROUP_BYTES;
    // everything the loop indexes by g is kept as a running value (no divisions, no multiplies by g)
    unsigned ld_s = ring_s;                                  // ring group the next load fills
    int ld_c = 0;                                            // its first sample
    unsigned ld_par = 0;                                     // (group index) & 1: barrier / residual buffer it uses
    auto issue_loads = [&]() {
        // the next input group, and the unfiltered samples of the iteration that consumes it
        const unsigned bar = bar_s + 8 * ld_par;
        b5t_mbar_expect(bar, tx);
        b5t_load3(ld_s, &maps.in, ld_c, (int)line0, 0, bar);
        if (MODE_OUT == FOUT_RESID) b5t_load2(d2_s + 256 * ld_par, &maps.d2, ld_c - 4 * G - r4, (int)line0, bar);
        ld_s += B5T_GROUP_BYTES; if (ld_s == ring_end) ld_s = ring_s;
        ld_c += G;
        ld_par ^= 1u;
    };
    if (lane == 0) {
        issue_loads();
        issue_loads();
    }

    B5Acc<false> acc;
    acc.reset();
    uint4 car = make_uint4(0u, 0u, 0u, 0u);
    unsigned e_s = ring_s + (unsigned)eg * B5T_GROUP_BYTES;  // entering group of this lane's chain
    unsigned l_s = ring_s + (unsigned)lg * B5T_GROUP_BYTES;  // group of its first leaving vector
    int jb = -4 * G - r4;                                    // first sample the current iteration finishes
    int i0 = -pass * G;                                      // first local index of this lane's group
    unsigned par = 0, ph = 0;                                // g & 1, (g >> 1) & 1
    unsigned t_cur = 0, t_prev = 512;                        // byte offsets of the output tiles g % 3, (g - 1) % 3
    for (int g = 0;; g++) {
        const bool live = jb + G > 0 && jb < n;
        // the tile iteration g - 1 parked goes out; the one of g - 2 has been read by then
        if (lane == 0) {
            const int pb = jb - G;
            if (pb >= 0 && pb < n) {
                if (a.out_transposed) b5t_store2(&maps.out, pb, (int)line0, out_s + t_prev);
                else b5t_store3(&maps.out, lip0, pb, plane, out_s + t_prev);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
        }
        b5t_mbar_wait(bar_s + 8 * par, ph);
        if (live) {
            const uint4 *row = reinterpret_cast<const uint4 *>(ystage_p) + ((par ^ 1u) * GQ + dq) * B5_STAGE_ROW;
            const uint2 v = reinterpret_cast<const uint2 *>(row + dl)[dh];
            const uint2 w = reinterpret_cast<const uint2 *>(row + 4 + dl)[dh];
            const unsigned yv[2] = {v.x, v.y}, yw[2] = {w.x, w.y};
            float fv[2], fw[2], res[2];
            // one path for every sample: a group-level fast / plain split diverges wherever flagged
            // regions make some weights tiny, and then costs both
#pragma unroll
            for (int k = 0; k < 2; k++) {
                fv[k] = dv(__uint_as_float(yv[k]));
                fw[k] = dv(__uint_as_float(yw[k]));
                res[k] = (fw[k] == 0.f) ? NAN : fv[k] / fw[k];
            }
            if (MODE_OUT == FOUT_RESID) {
                const float2 d2 = *reinterpret_cast<const float2 *>(d2_p + 64 * par + dl * 16 + ds);
                res[0] = fabsf(d2.x - res[0]);
                res[1] = fabsf(d2.y - res[1]);
            }
            if (jb < 0) {
                // the one tile that starts before sample 0: bulk tensor stores take no negative
                // coordinates (measured: illegal instruction), so its valid samples are stored directly
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int js = jb + ds + k;
                    if (js < 0 || js >= n) continue;
                    if (a.out_transposed) a.vout[(line0 + dl) * (int64_t)n + js] = res[k];
                    else a.vout[((int64_t)plane * n + js) * a.nj + lip0 + dl] = res[k];
                }
            } else {
                float *tile = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(out_p) + t_cur);
                if (a.out_transposed) {
                    *reinterpret_cast<float2 *>(tile + dl * 16 + ds) = make_float2(res[0], res[1]);
                } else {
                    tile[ds * 4 + dl] = res[0];
                    tile[ds * 4 + 4 + dl] = res[1];
                }
                b5t_fence_async();
            }
        }
        if (g == niter) break;

        // ---- the chain step of B5 on the swizzled rings
        unsigned l_s1 = l_s + B5T_GROUP_BYTES;
        if (l_s1 == ring_end) l_s1 = ring_s;
        uint4 e[GQ], nw[GQ];
#pragma unroll
        for (int q = 0; q < GQ; q++)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(e[q].x), "=r"(e[q].y), "=r"(e[q].z), "=r"(e[q].w) : "r"(e_s + eo[q]));
#pragma unroll
        for (int k = 0; k < GQ; k++)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(nw[k].x), "=r"(nw[k].y), "=r"(nw[k].z), "=r"(nw[k].w) : "r"((lhi[k] ? l_s1 : l_s) + lo[k]));
        l_s = l_s1;
        unsigned in[G], old[G], y[G];
#pragma unroll
        for (int q = 0; q < GQ; q++) {
            in[4 * q] = e[q].x; in[4 * q + 1] = e[q].y; in[4 * q + 2] = e[q].z; in[4 * q + 3] = e[q].w;
        }
        if (ODD) {
            old[0] = car.z; old[1] = car.w;
#pragma unroll
            for (int q = 0; q < GQ; q++) {
                old[4 * q + 2] = nw[q].x; old[4 * q + 3] = nw[q].y;
                if (q + 1 < GQ) { old[4 * q + 4] = nw[q].z; old[4 * q + 5] = nw[q].w; }
            }
            car = nw[GQ - 1];
        } else {
#pragma unroll
            for (int q = 0; q < GQ; q++) {
                old[4 * q] = nw[q].x; old[4 * q + 1] = nw[q].y; old[4 * q + 2] = nw[q].z; old[4 * q + 3] = nw[q].w;
            }
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            acc.add(in[k]);
            y[k] = acc.emit();
            acc.sub(old[k]);
        }
        if (i0 < ylo || i0 + G > yhi) {
#pragma unroll
            for (int k = 0; k < G; k++)
                if (i0 + k < ylo || i0 + k >= yhi) y[k] = 0u;
        }
        const unsigned ybase = pass < 3 ? e_s : ystage_s + par * (GQ * B5_STAGE_ROW * 16);
#pragma unroll
        for (int q = 0; q < GQ; q++)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                         ::"r"(ybase + yo[q]), "r"(y[4 * q]), "r"(y[4 * q + 1]), "r"(y[4 * q + 2]), "r"(y[4 * q + 3]) : "memory");
        e_s += B5T_GROUP_BYTES; if (e_s == ring_end) e_s = ring_s;
        i0 += G;
        jb += G;
        ph ^= par;
        par ^= 1u;
        t_prev = t_cur;
        t_cur = t_cur == 512 ? 0 : t_cur + 256;
        // input two groups ahead (and nothing that would still be in flight when the warp exits)
        if (lane == 0 && g + 2 <= niter) issue_loads();
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {
        const int pb = (niter - 4) * G - r4;
        if (pb >= 0 && pb < n) {
            const unsigned src = out_s + 256 * (niter % 3);   // == out_s + t_cur
            if (a.out_transposed) b5t_store2(&maps.out, pb, (int)line0, src);
            else b5t_store3(&maps.out, lip0, pb, plane, src);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
}

// ---------------------------------------------------------------- launching ----
typedef CUresult (*b5t_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static b5t_encode_fn b5t_encoder()
{
    static b5t_encode_fn fn = []() -> b5t_encode_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (b5t_encode_fn)p;
    }();
    return fn;
}

static size_t b5t_per_warp(int r)
{
    const int ngr = (2 * r + 4 * B5_G - 1) / B5_G;
    return (size_t)ngr * B5T_GROUP_BYTES + B5T_FIXED_BYTES;
}

// second axis of the 2-D masked filter (pair input, background / residual output) on lines that
// come in whole groups of 4 per plane; the pair must live in one allocation (win after data)
static bool b5t_supported(tc_context *c, const FilterArgs &a)
{
    // measured on B200 (profiles/r02_filter_probe.txt): ahead of the plain-load form while the rings are short enough for
    // ~18 resident warps per SM (r = 8: 0.92 against 0.96 ms), behind it from r = 17 (11 warps: 1.21 against 1.08 ms)
    // Since the plain-load form fetches its residual samples an iteration ahead and runs its loop unrolled by
    // two (k_filter5.cuh), it is ahead at every radius (second axis 122.0 -> 117.8 ms per 32-baseline step with
    // this form switched off): the TMA form is opt-in, TC_FILTER_TMA=1.
    static const int max_r = tpl_env_int("TC_B5T_MAXR", 12);
    if (!TC_ENV_FLAG("TC_FILTER_TMA") || TC_ENV_FLAG("TC_FILTER_NO_B5") || TC_ENV_FLAG("TC_FILTER_OLD")) return false;
    if (a.mode_in != FIN_PAIR || (a.mode_out != FOUT_BG && a.mode_out != FOUT_RESID)) return false;
    if (a.r < 1 || a.r > max_r || (a.n & 3) || (a.nj & 3) || a.nlines % a.nj) return false;
    if (a.win <= a.data || ((uintptr_t)a.win - (uintptr_t)a.data) % 16 ||
        (uintptr_t)a.win - (uintptr_t)a.data >= ((uintptr_t)1 << 40)) return false;
    if ((((uintptr_t)a.data | (uintptr_t)a.vout | (uintptr_t)a.data2) & 15) != 0) return false;
    if (a.nlines >= ((int64_t)1 << 31) || (int64_t)a.n + 4 * a.r + 8 * B5_G >= ((int64_t)1 << 30)) return false;
    if (b5t_per_warp(a.r) + 2048 > (size_t)c->smem_optin) return false;
    return b5t_encoder() != nullptr;
}

static int b5t_encode(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                      const cuuint32_t *box, CUtensorMapSwizzle swz)
{
    const cuuint32_t ones[3] = {1, 1, 1};
    CUresult rc = b5t_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, (cuuint32_t)rank, const_cast<void *>(base), dims,
                                strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return tc_fail(TC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return TC_OK;
}

static int launch_box_filter5t(tc_context *c, FilterArgs a)
{
    if (a.nlines == 0 || a.n == 0) return TC_OK;
    TC_REQUIRE(b5t_supported(c, a), "internal: B5T filter launched on an unsupported shape");
    a.div = tc_f32_pow4(2 * (int64_t)a.r + 1);
    if (TC_ENV_FLAG("TC_FILTER_TRACE"))
        fprintf(stderr, "b5t filter: n=%d nj=%d r=%d out=%d tr=%d\n", a.n, a.nj, a.r, a.mode_out, a.out_transposed);
    B5TMaps maps;
    memset(&maps, 0, sizeof(maps));
    const cuuint64_t n = (cuuint64_t)a.n, nl = (cuuint64_t)a.nlines, nj = (cuuint64_t)a.nj, np = nl / nj;
    {
        const cuuint64_t dims[3] = {n, nl, 2};
        const cuuint64_t strides[2] = {n * 4, (cuuint64_t)((uintptr_t)a.win - (uintptr_t)a.data)};
        const cuuint32_t box[3] = {B5_G, 4, 2};
        TC_TRY(b5t_encode(&maps.in, a.data, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
    }
    if (a.mode_out == FOUT_RESID) {
        const cuuint64_t dims[2] = {n, nl};
        const cuuint64_t strides[1] = {n * 4};
        const cuuint32_t box[2] = {B5_G, 4};
        TC_TRY(b5t_encode(&maps.d2, a.data2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
    if (a.out_transposed) {
        const cuuint64_t dims[2] = {n, nl};
        const cuuint64_t strides[1] = {n * 4};
        const cuuint32_t box[2] = {B5_G, 4};
        TC_TRY(b5t_encode(&maps.out, a.vout, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    } else {
        const cuuint64_t dims[3] = {nj, n, np};
        const cuuint64_t strides[2] = {nj * 4, n * nj * 4};
        const cuuint32_t box[3] = {4, B5_G, 1};
        TC_TRY(b5t_encode(&maps.out, a.vout, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
    const bool odd = (a.r & 1) != 0;
    const size_t per_warp = b5t_per_warp(a.r);
    const int64_t nwarps = (a.nlines + 3) / 4;
    int wpb = b2_warps_per_block(c, per_warp, nwarps, 24);
    if (wpb > 4) wpb = 4;
    const unsigned grid = (unsigned)((nwarps + wpb - 1) / wpb);
    const size_t smem = per_warp * wpb;
    tc_prof_begin(c, TCP_BOX_FILTER);
#define B5T_CASE(OD, MO)                                                                                     \
    if (odd == OD && a.mode_out == MO) {                                                                     \
        if (smem > 48 * 1024)                                                                                \
            TC_CUDA(cudaFuncSetAttribute(k_box5t<OD, MO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_box5t<OD, MO><<<grid, wpb * 32, smem, c->stream>>>(a, maps);                                        \
    }
    B5T_CASE(true, FOUT_BG)
    B5T_CASE(false, FOUT_BG)
    B5T_CASE(true, FOUT_RESID)
    B5T_CASE(false, FOUT_RESID)
#undef B5T_CASE
    tc_prof_end(c);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}
#else
static bool b5t_supported(tc_context *, const FilterArgs &) { return false; }
static int launch_box_filter5t(tc_context *, FilterArgs) { return TC_ERR_VALUE; }
#endif
