// tc_common.cuh -- context, workspace arena and error plumbing.
#pragma once
#include "tc_rt.h"
#include "../../include/tricolour_b200.h"

#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <string>
#include <atomic>

typedef uint8_t u8;

// ---------------------------------------------------------------- errors ----
static thread_local char g_tc_err[512] = "";

static int tc_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_tc_err, sizeof(g_tc_err), fmt, ap);
    va_end(ap);
    return code;
}

#define TC_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess)                                                     \
            return tc_fail(TC_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                           cudaGetErrorString(e_), __FILE__, __LINE__);            \
    } while (0)

#define TC_TRY(call)                 \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != TC_OK) return rc_; \
    } while (0)

#define TC_KERNEL_CHECK() TC_CUDA(cudaGetLastError())

#define TC_REQUIRE(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) return tc_fail(TC_ERR_VALUE, __VA_ARGS__); \
    } while (0)

// ------------------------------------------------------ environment knobs ----
// read once per process (getenv walks the whole environment; the launch helpers sit on
// the host path of every kernel launch)
#define TC_ENV_FLAG(name)                                           \
    ([]() -> bool {                                                 \
        static const bool v = getenv(name) != nullptr;              \
        return v;                                                   \
    }())

// ------------------------------------------------------- device ledger ------
// The workspace budget is per DEVICE, not per context: dask's ThreadPool(nworkers)
// gives every worker thread its own context (stream + arena), and each of them sizing
// its plane batches to the whole budget would exhaust HBM.  The ledger counts the
// contexts that hold an arena and the bytes they hold; tc_ws_share() divides
// the budget between them.
#define TC_MAX_DEVICES 64
struct tc_device_ledger {
    std::atomic<long long> arena_bytes{0};
    std::atomic<int> arenas{0};
};
static tc_device_ledger g_tc_ledger[TC_MAX_DEVICES];

// --------------------------------------------------------------- context ----
struct tc_block { char *ptr; size_t size; };

struct tc_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int smem_optin = 227 * 1024;
    std::vector<tc_block> blocks;  // arena blocks; blocks[0] is the main one
    size_t held = 0;               // bytes of all arena blocks (mirrored in the device ledger)
    size_t cur_block = 0, cur_off = 0;
    size_t used_total = 0;         // bytes handed out since the last reset
    size_t peak = 0;
    unsigned long long launches = 0;  // kernels launched through this context
    // optional per-kernel-family timing with CUDA events (tc_profile_*)
    bool prof = false;
    struct ProfRec { int id; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    double prof_ms[32] = {0};
    long long prof_cnt[32] = {0};
    // page-locked staging slabs for the small host tables a call uploads (chunk
    // boundaries, range tables, masks): copies from pageable memory would make
    // the host wait for the stream and serialise callers that use several streams
    enum { NSLAB = 4, SLAB_BYTES = 512 * 1024 };
    char *slab[NSLAB] = {nullptr, nullptr, nullptr, nullptr};
    char *slab_dev[NSLAB] = {nullptr, nullptr, nullptr, nullptr};   // the same slabs as the device sees them
    cudaEvent_t slab_ev[NSLAB] = {nullptr, nullptr, nullptr, nullptr};
    bool slab_busy[NSLAB] = {false, false, false, false};
    int slab_cur = 0;
    size_t slab_off = 0;
};

enum { TCP_BOX_FILTER = 0, TCP_CHUNK_SELECT, TCP_LINE_MEDIAN, TCP_ST_SCAN, TCP_TRANSPOSE, TCP_PREP,
       TCP_COMBINE, TCP_INTERP, TCP_ELEMENTWISE, TCP_UVCONTSUB, TCP_PACK, TCP_STATS, TCP_BOX_FILTER8, TCP_BOX_FILTER_1D,
       TCP_NIDS };
static const char *const tc_prof_names[TCP_NIDS] = {
    "box_filter", "chunk_select", "line_median", "st_scan", "transpose", "prep", "combine", "interp_nans",
    "elementwise", "uvcontsub", "pack_unpack", "window_counts", "box_filter_axis0", "box_filter_single_axis"};

static inline void tc_prof_begin(tc_context *c, int id)
{
#ifndef TC_EMU
    if (!c->prof) return;
    tc_context::ProfRec r;
    r.id = id;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, c->stream);
    c->prof_recs.push_back(r);
#else
    (void)c; (void)id;
#endif
}
static inline void tc_prof_end(tc_context *c)
{
#ifndef TC_EMU
    if (!c->prof || c->prof_recs.empty()) return;
    cudaEventRecord(c->prof_recs.back().b, c->stream);
#else
    (void)c;
#endif
}
static inline void tc_prof_collect(tc_context *c)
{
#ifndef TC_EMU
    for (auto &r : c->prof_recs) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.id] += ms; c->prof_cnt[r.id]++; }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->prof_recs.clear();
#else
    (void)c;
#endif
}

static inline size_t tc_align(size_t n, size_t a = 256) { return (n + a - 1) / a * a; }

// start of an API call: retire the staging slab of the previous call and move on
static inline void tc_slab_rotate(tc_context *c)
{
#ifndef TC_EMU
    if (c->slab_off && c->slab[c->slab_cur]) {
        if (!c->slab_ev[c->slab_cur]) cudaEventCreateWithFlags(&c->slab_ev[c->slab_cur], cudaEventDisableTiming);
        cudaEventRecord(c->slab_ev[c->slab_cur], c->stream);
        c->slab_busy[c->slab_cur] = true;
        c->slab_cur = (c->slab_cur + 1) % tc_context::NSLAB;
    }
    if (c->slab_busy[c->slab_cur]) {
        cudaEventSynchronize(c->slab_ev[c->slab_cur]);
        c->slab_busy[c->slab_cur] = false;
    }
#endif
    c->slab_off = 0;
}

#ifndef TC_EMU
// copies a staged table out of the page-locked slab (read through its device
// mapping) into device memory: 16 bytes per thread, then the byte tail
__global__ void __launch_bounds__(256)
k_upload_small(const unsigned char *src, unsigned char *dst, unsigned bytes, int vec)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const unsigned n16 = bytes >> 4;
        if (i < n16) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
        const unsigned t = (n16 << 4) + i;
        if (i < 16 && t < bytes) dst[t] = src[t];
    } else if (i < bytes) dst[i] = src[i];
}
#endif

// host -> device copy of a small table, asynchronous with respect to the host.
// The table is staged in a page-locked slab and fetched by a kernel of the
// context's stream instead of the copy engine: a cudaMemcpyAsync would queue
// behind whatever bulk upload another stream has in flight (the next block of a
// pipelined executor: 4.8 GB, 90 ms) and stall the flagging kernels behind it.
static int tc_upload_small(tc_context *c, const void *h, size_t bytes, void *d)
{
    if (!bytes) return TC_OK;
#ifndef TC_EMU
    size_t need = tc_align(bytes, 16);
    if (c->slab_off + need <= (size_t)tc_context::SLAB_BYTES) {
        if (!c->slab[c->slab_cur]) {
            void *p = nullptr, *dp = nullptr;
            if (cudaHostAlloc(&p, tc_context::SLAB_BYTES, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
                p = nullptr;
            if (p && (cudaHostGetDevicePointer(&dp, p, 0) != cudaSuccess || !dp)) { cudaFreeHost(p); p = nullptr; }
            c->slab[c->slab_cur] = (char *)p;
            c->slab_dev[c->slab_cur] = (char *)dp;
            cudaGetLastError();
        }
        if (c->slab[c->slab_cur]) {
            char *stage = c->slab[c->slab_cur] + c->slab_off;
            const unsigned char *stage_dev = (const unsigned char *)c->slab_dev[c->slab_cur] + c->slab_off;
            memcpy(stage, h, bytes);
            c->slab_off += need;
            const int vec = ((uintptr_t)d & 15) == 0;
            const unsigned nthreads = vec ? (unsigned)(bytes >> 4) + 16u : (unsigned)bytes;
            k_upload_small<<<(nthreads + 255) / 256, 256, 0, c->stream>>>(stage_dev, (unsigned char *)d, (unsigned)bytes, vec);
            TC_KERNEL_CHECK();
            c->launches++;
            return TC_OK;
        }
    }
#endif
    TC_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return TC_OK;
}

// device -> device copy on the SMs (the copy engines stay free for the bulk
// transfers of other streams); both pointers 16-byte aligned
__global__ void __launch_bounds__(256)
k_copy_bytes(const unsigned char *src, unsigned char *dst, int64_t bytes)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n16 = bytes >> 4;
    if (i < n16) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    const int64_t t = (n16 << 4) + i;
    if (i < 16 && t < bytes) dst[t] = src[t];
}

static int tc_copy_d2d(tc_context *c, void *dst, const void *src, int64_t bytes)
{
    if (bytes <= 0) return TC_OK;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) != 0) {
        TC_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, c->stream));
        return TC_OK;
    }
    const int64_t nthreads = (bytes >> 4) + 16;
    TC_LAUNCH_NOSYNC(k_copy_bytes, (unsigned)((nthreads + 255) / 256), 256, 0, c->stream, (const unsigned char *)src,
                     (unsigned char *)dst, bytes);
    c->launches++;
    TC_KERNEL_CHECK();
    return TC_OK;
}

static inline tc_device_ledger &tc_ledger(const tc_context *c)
{
    return g_tc_ledger[(c->device >= 0 && c->device < TC_MAX_DEVICES) ? c->device : 0];
}

static void tc_arena_note_block(tc_context *c, char *p, size_t size)
{
    if (c->blocks.empty()) tc_ledger(c).arenas.fetch_add(1);
    c->blocks.push_back({p, size});
    c->held += size;
    tc_ledger(c).arena_bytes.fetch_add((long long)size);
}

// hands the whole arena back to the driver (the stream must be idle)
static void tc_arena_free_all(tc_context *c)
{
    if (c->blocks.empty()) return;
    for (auto &b : c->blocks) cudaFree(b.ptr);
    c->blocks.clear();
    tc_ledger(c).arena_bytes.fetch_sub((long long)c->held);
    tc_ledger(c).arenas.fetch_sub(1);
    c->held = 0;
    c->cur_block = 0; c->cur_off = 0; c->used_total = 0;
}

// device-wide workspace limit: TC_WORKSPACE_MB (default 48 GiB); read once per API call
static size_t tc_workspace_limit()
{
    const char *e = getenv("TC_WORKSPACE_MB");
    size_t mb = e ? (size_t)atoll(e) : 49152;
    if (mb < 64) mb = 64;
    return mb << 20;
}

// this context's share of the device-wide workspace: the limit divided by the number
// of contexts that hold (or are about to hold) an arena on the device, and never more
// than what the device can still give
static size_t tc_ws_share(tc_context *c)
{
    int n = tc_ledger(c).arenas.load();
    if (c->blocks.empty()) n++;
    if (n < 1) n = 1;
    size_t share = tc_workspace_limit() / (size_t)n;
#ifndef TC_EMU
    size_t fre = 0, tot = 0;
    if (cudaMemGetInfo(&fre, &tot) == cudaSuccess) {
        const size_t reserve = (size_t)256 << 20;
        size_t avail = c->held + (fre > reserve ? fre - reserve : 0) / 10 * 9;
        if (avail < share) share = avail;
    } else {
        cudaGetLastError();
    }
#endif
    if (share < ((size_t)64 << 20)) share = (size_t)64 << 20;
    return share;
}

// start of an API call: if the previous call overflowed into extra blocks, fold
// everything into one block of the peak size; an arena that has outgrown the context's
// share of the device (more worker threads have started since) is handed back.
static int tc_arena_reset(tc_context *c)
{
    bool oversized = false;
    if (c->held > ((size_t)256 << 20) && tc_ledger(c).arenas.load() > 1) {
        // only other arenas on the device can shrink this context's share
        const size_t share = tc_workspace_limit() / (size_t)tc_ledger(c).arenas.load();
        oversized = c->held > share + (share >> 1);
    }
    if (c->blocks.size() > 1 || oversized) {
        TC_CUDA(cudaStreamSynchronize(c->stream));
        tc_arena_free_all(c);
        if (!oversized) {
            size_t want = tc_align(c->peak + (c->peak >> 3), 1 << 20);
            char *p = nullptr;
            TC_CUDA(cudaMalloc((void **)&p, want));
            tc_arena_note_block(c, p, want);
        } else {
            c->peak = 0;
        }
    }
    c->cur_block = 0;
    c->cur_off = 0;
    c->used_total = 0;
    return TC_OK;
}

static int tc_arena_alloc(tc_context *c, size_t bytes, void **out)
{
    bytes = tc_align(bytes ? bytes : 1);
    while (true) {
        if (c->cur_block < c->blocks.size()) {
            tc_block &b = c->blocks[c->cur_block];
            if (c->cur_off + bytes <= b.size) {
                *out = b.ptr + c->cur_off;
                c->cur_off += bytes;
                c->used_total += bytes;
                if (c->used_total > c->peak) c->peak = c->used_total;
                return TC_OK;
            }
            c->cur_block++;
            c->cur_off = 0;
            continue;
        }
        size_t want = tc_align(bytes > (size_t)(64 << 20) ? bytes : (size_t)(64 << 20), 1 << 20);
        char *p = nullptr;
        cudaError_t e = cudaMalloc((void **)&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return tc_fail(TC_ERR_CUDA, "cudaMalloc(%zu) failed: %s (arena holds %zu bytes; %d arenas hold %lld bytes on "
                           "device %d; lower TC_WORKSPACE_MB or close idle contexts)", want, cudaGetErrorString(e),
                           c->held, tc_ledger(c).arenas.load(), tc_ledger(c).arena_bytes.load(), c->device);
        }
        tc_arena_note_block(c, p, want);
    }
}

// stack discipline inside one API call: everything allocated after a mark is
// handed back by the matching release (stream order makes the reuse safe)
struct tc_mark { size_t block, off, used; };
static inline tc_mark tc_arena_mark(tc_context *c) { return tc_mark{c->cur_block, c->cur_off, c->used_total}; }
static inline void tc_arena_release(tc_context *c, tc_mark m)
{
    c->cur_block = m.block; c->cur_off = m.off; c->used_total = m.used;
}

template <typename T> static int tc_alloc(tc_context *c, size_t count, T **out)
{
    void *p = nullptr;
    TC_TRY(tc_arena_alloc(c, count * sizeof(T), &p));
    *out = (T *)p;
    return TC_OK;
}

// Stage a host or device array on the device.  space == TC_HOST copies through
// the context stream; TC_DEVICE uses the pointer as is.
template <typename T>
static int tc_stage_in(tc_context *c, const T *src, size_t count, int space, const T **dev)
{
    if (space == TC_DEVICE) { *dev = src; return TC_OK; }
    T *d = nullptr;
    TC_TRY(tc_alloc(c, count, &d));
    if (count) TC_CUDA(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dev = d;
    return TC_OK;
}

template <typename T>
static int tc_stage_out_begin(tc_context *c, T *dst, size_t count, int space, T **dev)
{
    if (space == TC_DEVICE) { *dev = dst; return TC_OK; }
    return tc_alloc(c, count, dev);
}

template <typename T>
static int tc_stage_out_end(tc_context *c, T *dst, const T *dev, size_t count, int space)
{
    if (space == TC_DEVICE) return TC_OK;
    if (count) TC_CUDA(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    TC_CUDA(cudaStreamSynchronize(c->stream));
    return TC_OK;
}

static inline unsigned tc_blocks_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }
