// ubench.cu -- issue-rate microbenchmarks for the ops the box filter lives on
// (B200, sm_100a): DADD, F2F f64->f32, F2F f32->f64, LDS/STS, Veltkamp rounding.
// Prints operations per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

__global__ void k_dadd(double *out, double x)
{
    double a[NACC];
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) a[i] = __dadd_rn(a[i], x);
    double s = 0;
    for (int i = 0; i < NACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_f2f_round(double *out, double x)
{
    // d -> f -> d round trip, NACC independent chains (2 conversions per step)
    double a[NACC];
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 1e-3 + i + x;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            float f = __double2float_rn(a[i]);
            a[i] = (double)f * 1.0000001;   // DMUL keeps the chain from folding
        }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_d2f_only(float *out, double x)
{
    double a[NACC];
    float acc = 0.f;
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 1e-3 + i + x;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            a[i] = __dadd_rn(a[i], x);
            acc += __double2float_rn(a[i]);
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_f2d_only(double *out, float x)
{
    float a[NACC];
    double acc[NACC];
    for (int i = 0; i < NACC; i++) { a[i] = threadIdx.x * 1e-3f + i + x; acc[i] = 0; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            a[i] = a[i] * 1.0000001f + x;
            acc[i] = __dadd_rn(acc[i], (double)a[i]);
        }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_veltkamp(double *out, double x)
{
    // round-to-24-bit via Veltkamp split: 1 DMUL + 2 DADD per rounding
    double a[NACC];
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 1e-3 + i + x;
    const double C = 536870913.0;  // 2^29 + 1
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            double p = __dmul_rn(a[i], C);
            double q = __dadd_rn(a[i], -p);
            a[i] = __dadd_rn(q, p) + x;
        }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds(float *out)
{
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8 * blockDim.x; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float acc = 0.f;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float v = sm[i * blockDim.x + idx];
            sm[i * blockDim.x + idx] = v + 1.0f;
            acc += v;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> static float timeit(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *d; float *f;
    cudaMalloc(&d, sizeof(double) * sms * 8 * 1024);
    cudaMalloc(&f, sizeof(float) * sms * 8 * 1024);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, nominal %d MHz\n", p.name, sms, clk_khz / 1000);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int bd = warps * 32, grid = sms;
        double ops = (double)grid * bd * ITERS * NACC;
        float t;
        t = timeit([&] { k_dadd<<<grid, bd>>>(d, 1e-9); });
        printf("warps/SM %2d  DADD            %.1f ops/clk/SM (at %d MHz)  %.3f ms\n", warps, ops / (t * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, t);
        t = timeit([&] { k_f2f_round<<<grid, bd>>>(d, 1e-9); });
        printf("warps/SM %2d  d2f+f2d+DMUL    %.1f triples/clk/SM  %.3f ms\n", warps, ops / (t * 1e-3) / sms / (clk_khz * 1e3), t);
        t = timeit([&] { k_d2f_only<<<grid, bd>>>(f, 1e-9); });
        printf("warps/SM %2d  DADD+d2f+FADD   %.1f /clk/SM  %.3f ms\n", warps, ops / (t * 1e-3) / sms / (clk_khz * 1e3), t);
        t = timeit([&] { k_f2d_only<<<grid, bd>>>(d, 1e-9f); });
        printf("warps/SM %2d  FFMA+f2d+DADD   %.1f /clk/SM  %.3f ms\n", warps, ops / (t * 1e-3) / sms / (clk_khz * 1e3), t);
        t = timeit([&] { k_veltkamp<<<grid, bd>>>(d, 1e-9); });
        printf("warps/SM %2d  veltkamp(3op)+DADD %.1f /clk/SM  %.3f ms\n", warps, ops / (t * 1e-3) / sms / (clk_khz * 1e3), t);
        t = timeit([&] { k_lds<<<grid, bd, 8 * bd * 4>>>(f); });
        printf("warps/SM %2d  LDS+STS+2FADD   %.1f pairs/clk/SM  %.3f ms\n", warps, (double)grid * bd * ITERS * 8 / (t * 1e-3) / sms / (clk_khz * 1e3), t);
    }
    return 0;
}
