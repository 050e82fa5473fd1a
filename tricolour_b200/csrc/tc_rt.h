// tc_rt.h -- thin runtime layer for the tricolour_b200 CUDA library.
//
// Product build (nvcc, sm_100a): plain CUDA runtime calls and <<<>>> launches.
//
// TC_EMU build (g++, tests only): the same kernel sources are compiled for the
// host and every launch is executed by a cooperative SIMT emulator (one
// ucontext fiber per CUDA thread, __syncthreads/warp shuffles honoured).  It
// exists so that the kernels' index arithmetic can be checked against the
// oracle inside the GPU-less build container.  The Python package never loads
// the emulated library: only tests/ do, explicitly.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <math.h>

#ifndef TC_EMU
// ------------------------------------------------------------------ CUDA ----
#include <cuda_runtime.h>

#define TC_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type *name = reinterpret_cast<type *>(name##_raw_)

#define TC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// kernels that never synchronise inside a block (hint for the emulator only)
#define TC_LAUNCH_NOSYNC(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)

#define TC_FULL_MASK 0xffffffffu

#else
// -------------------------------------------------------------- emulation ---
#include <string.h>
#include <stdlib.h>
#include <functional>
#include <algorithm>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
extern uint3_emu threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
extern unsigned char *tc_emu_dyn_smem;

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };

static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)
#define TC_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(tc_emu_dyn_smem)
#define TC_FULL_MASK 0xffffffffu

// fiber scheduler entry points (cuda_emu.cpp)
void tc_emu_run_grid(dim3 grid, dim3 block, size_t smem, bool nosync,
                     const std::function<void()> &body);
void __syncthreads();
void __syncwarp(unsigned mask = 0xffffffffu);
uint64_t tc_emu_shfl_u64(uint64_t v, int srclane, int mode, int width);
unsigned __ballot_sync(unsigned mask, int pred);

template <typename T> static inline T tc_emu_shfl(T v, int lane, int mode, int width)
{
    static_assert(sizeof(T) <= 8, "shfl type too wide");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = tc_emu_shfl_u64(raw, lane, mode, width);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { return tc_emu_shfl(v, src, 0, width); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { return tc_emu_shfl(v, m, 1, width); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) { return tc_emu_shfl(v, (int)d, 2, width); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) { return tc_emu_shfl(v, (int)d, 3, width); }

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline double __dsqrt_rn(double x) { return sqrt(x); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float __double2float_rd(double a)
{
    float f = (float)a;
    if ((double)f > a) f = nextafterf(f, -INFINITY);
    return f;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline void __threadfence() {}
static inline float fminf_emu(float a, float b) { return a < b ? a : b; }

#define TC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    tc_emu_run_grid(dim3(grid), dim3(block), (smem), false, [&]() { kernel(__VA_ARGS__); })
#define TC_LAUNCH_NOSYNC(kernel, grid, block, smem, stream, ...) \
    tc_emu_run_grid(dim3(grid), dim3(block), (smem), true, [&]() { kernel(__VA_ARGS__); })

#endif  // TC_EMU
