// Stand-alone probe of the TMA pieces k_filter5t.cuh relies on (bulk tensor loads with the
// 64-byte swizzle into a per-warp ring, mbarrier completion, bulk tensor stores with clipping).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/tma_probe tools/probes/tma_probe.cu
// run:   tools/probes/tma_probe [variant]   (variants exercise one mechanism each; prints PASS/FAIL per check)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Maps { CUtensorMap in, d2, out; };

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ Maps maps, int variant, int n, int nj, unsigned *dump)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    const int lane = threadIdx.x & 31;
    const unsigned base = s32(sm);
    const unsigned bar = base + 4096;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = lane; i < 1024; i += 32) reinterpret_cast<unsigned *>(sm)[i] = 0xdeadbeefu;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (variant >= 1 && lane == 0) {
        unsigned tx = variant == 1 ? 256u : 512u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
        if (variant == 1)          // 2-D box (16 samples, 4 lines), no swizzle, coords (-8, 0): first 8 samples out of range
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(base), "l"(&maps.d2), "r"(-8), "r"(0), "r"(bar) : "memory");
        else                       // 3-D box (16, 4, 2), 64-byte swizzle
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(base), "l"(&maps.in), "r"(16), "r"(4), "r"(0), "r"(bar) : "memory");
    }
    if (variant >= 1) {
        asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(0) : "memory");
    }
    __syncwarp();
    for (int i = lane; i < 128; i += 32) dump[i] = reinterpret_cast<unsigned *>(sm)[i];
    if (variant >= 3) {
        // park a (16 samples x 4 lines) tile and store it at sample -4 (clipped) of plane 1, lines 4..7
        float *tile = reinterpret_cast<float *>(sm + 2048);
        for (int i = lane; i < 64; i += 32) tile[i] = 1000.f + i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            if (variant == 5)        // 2-D store, in range
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&maps.d2), "r"(16), "r"(4), "r"(base + 2048) : "memory");
            else
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&maps.out), "r"(4), "r"(variant == 4 ? 8 : (variant == 6 ? n - 4 : -4)), "r"(1), "r"(base + 2048) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
}

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    int variant = argc > 1 ? atoi(argv[1]) : 3;
    const int n = 64, nj = 8, np = 2, nl = nj * np;
    enc_fn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q));
    if (!enc || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    std::vector<float> hv(2 * nl * n);
    for (int a = 0; a < 2; a++) for (int l = 0; l < nl; l++) for (int i = 0; i < n; i++) hv[(a * nl + l) * n + i] = a * 10000.f + l * 100.f + i;
    float *dv, *dout; unsigned *ddump;
    CK(cudaMalloc(&dv, hv.size() * 4)); CK(cudaMalloc(&dout, nl * n * 4)); CK(cudaMalloc(&ddump, 128 * 4));
    CK(cudaMemcpy(dv, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, nl * n * 4));
    Maps maps; memset(&maps, 0, sizeof(maps));
    const cuuint32_t ones[3] = {1, 1, 1};
    {
        cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)nl, 2}, strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)nl * n * 4};
        cuuint32_t box[3] = {16, 4, 2};
        CUresult rc = enc(&maps.in, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dv, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode in: %d\n", (int)rc);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)nl}, strides[1] = {(cuuint64_t)n * 4};
        cuuint32_t box[2] = {16, 4};
        CUresult rc = enc(&maps.d2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dv, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode d2: %d\n", (int)rc);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)nj, (cuuint64_t)n, (cuuint64_t)np}, strides[2] = {(cuuint64_t)nj * 4, (cuuint64_t)n * nj * 4};
        cuuint32_t box[3] = {4, 16, 1};
        CUresult rc = enc(&maps.out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dout, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode out: %d\n", (int)rc);
    }
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
    k_probe<<<1, 32, 8192>>>(maps, variant, n, nj, ddump);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: sync -> %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<unsigned> hd(128);
    CK(cudaMemcpy(hd.data(), ddump, 128 * 4, cudaMemcpyDeviceToHost));
    if (variant == 1) {
        // row l at 64 B * l: samples -8 .. 7 of line l
        int bad = 0;
        for (int l = 0; l < 4; l++) for (int i = 0; i < 16; i++) {
            float got; memcpy(&got, &hd[l * 16 + i], 4);
            float want = i < 8 ? 0.f : l * 100.f + (i - 8);
            if (got != want) bad++;
        }
        printf("2-D load with clipping: %s\n", bad ? "FAIL" : "PASS");
    } else if (variant >= 2) {
        // row = array * 4 + line at 64 B * row, 16-byte chunk c at (c ^ ((row >> 1) & 3))
        int bad = 0;
        for (int row = 0; row < 8; row++) for (int c = 0; c < 4; c++) for (int e2 = 0; e2 < 4; e2++) {
            float got; memcpy(&got, &hd[row * 16 + ((c ^ ((row >> 1) & 3)) * 4) + e2], 4);
            int a = row >> 2, l = 4 + (row & 3), i = 16 + c * 4 + e2;
            float want = a * 10000.f + l * 100.f + i;
            if (got != want) bad++;
        }
        printf("3-D load, 64-byte swizzle as assumed: %s\n", bad ? "FAIL" : "PASS");
        if (bad) for (int row = 0; row < 8; row++) { for (int w = 0; w < 16; w++) { float g; memcpy(&g, &hd[row * 16 + w], 4); printf("%7.0f", g); } printf("\n"); }
    }
    if (variant >= 3) {
        std::vector<float> ho(nl * n);
        CK(cudaMemcpy(ho.data(), dout, nl * n * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int p = 0; p < np; p++) for (int s = 0; s < n; s++) for (int l = 0; l < nj; l++) {
            float got = ho[(p * n + s) * nj + l], want = 0.f;
            int ts = s + (variant == 4 ? -8 : (variant == 6 ? 4 - n : 4)), tl = l - 4;       // tile sample / line
            if (variant == 5) continue;
            if (p == 1 && ts >= (variant == 3 ? 4 : 0) && ts < 16 && tl >= 0 && tl < 4) want = 1000.f + ts * 4 + tl;
            if (got != want) bad++;
        }
        printf("3-D store with clipping: %s\n", bad ? "FAIL" : "PASS");
    }
    return 0;
}
