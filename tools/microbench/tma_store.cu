// Micro-benchmark: throughput of TMA box stores (cp.async.bulk.tensor.3d global<-shared) of a constant pattern, as the
// fused clear of k_raster issues them, for several box shapes and numbers of issuing threads, beside plain 16-byte
// vector stores.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tma_store.cu -o tma_store
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct __align__(64) Map { CUtensorMap m; };

__device__ __forceinline__ void tma_store_box(const CUtensorMap *m, const void *smem, int cx, int cy, int cv)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 :: "l"((unsigned long long)m), "r"(cx), "r"(cy), "r"(cv), "r"((unsigned)__cvta_generic_to_shared(smem)) : "memory");
}

// One CTA per (tile column block, row block): the tensor is [V][H][WF] floats; each CTA owns a region of RX x RY floats
// and covers it with boxes of BX x BY, issued by `issuers` threads (lane 0 of the first warps).
__global__ void __launch_bounds__(256) k_tma(const __grid_constant__ Map M, int WF, int H, int RX, int RY, int BX, int BY, int issuers)
{
    extern __shared__ __align__(128) float pat[];
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) pat[i] = 1e6f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int regionsX = WF / RX;
    const int rx = blockIdx.x % regionsX, ry = blockIdx.x / regionsX, view = blockIdx.y;
    const int nbx = RX / BX, nby = RY / BY;
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < issuers) {
        for (int b = threadIdx.x >> 5; b < nbx * nby; b += issuers)
            tma_store_box(&M.m, pat, rx * RX + (b % nbx) * BX, ry * RY + (b / nbx) * BY, view);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) k_plain(float *out, int WF, int H, int RX, int RY)
{
    const int regionsX = WF / RX;
    const int rx = blockIdx.x % regionsX, ry = blockIdx.x / regionsX, view = blockIdx.y;
    const int q4 = RX / 4;
    for (int i = threadIdx.x; i < RY * q4; i += blockDim.x) {
        const int r = i / q4, q = i % q4;
        reinterpret_cast<float4 *>(out + ((long long)view * H + ry * RY + r) * WF + rx * RX)[q] = make_float4(1e6f, 1e6f, 1e6f, 1e6f);
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    const int V = 32, H = 1024;
    float *buf;
    CK(cudaMalloc(&buf, (size_t)V * H * 3072 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Cfg { int WF, RX, RY, BX, BY, issuers; };
    const Cfg cfgs[] = {
        {1024, 32, 32, 32, 8, 1},  {1024, 32, 32, 32, 8, 4},  {1024, 32, 32, 32, 32, 1},
        {3072, 96, 32, 96, 8, 1},  {3072, 96, 32, 96, 8, 4},  {3072, 96, 32, 96, 32, 1},
        {3072, 192, 32, 192, 32, 1}, {3072, 256, 64, 256, 64, 1}, {3072, 256, 64, 256, 8, 8}, {3072, 96, 32, 32, 8, 8},
    };
    for (const Cfg &c : cfgs) {
        Map M;
        const cuuint64_t dims[3] = {(cuuint64_t)c.WF, (cuuint64_t)H, (cuuint64_t)V};
        const cuuint64_t strides[2] = {(cuuint64_t)c.WF * 4, (cuuint64_t)c.WF * 4 * H};
        const cuuint32_t box[3] = {(cuuint32_t)c.BX, (cuuint32_t)c.BY, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&M.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d for box %dx%d\n", (int)r, c.BX, c.BY); continue; }
        const dim3 grid((c.WF / c.RX) * (H / c.RY), V);
        const size_t smem = (size_t)c.BX * c.BY * 4;
        CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        for (int it = 0; it < 3; ++it) k_tma<<<grid, 256, smem>>>(M, c.WF, H, c.RX, c.RY, c.BX, c.BY, c.issuers);
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 10; ++it) k_tma<<<grid, 256, smem>>>(M, c.WF, H, c.RX, c.RY, c.BX, c.BY, c.issuers);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)V * H * c.WF * 4 * 10;
        printf("tma   WF=%4d region %3dx%2d box %3dx%2d (%5zu B) issuers %d: %7.1f GB/s\n", c.WF, c.RX, c.RY, c.BX, c.BY, smem, c.issuers, bytes / ms / 1e6);
        for (int it = 0; it < 3; ++it) k_plain<<<grid, 256>>>(buf, c.WF, H, c.RX, c.RY);
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 10; ++it) k_plain<<<grid, 256>>>(buf, c.WF, H, c.RX, c.RY);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("plain WF=%4d region %3dx%2d                              : %7.1f GB/s\n", c.WF, c.RX, c.RY, bytes / ms / 1e6);
    }
    CK(cudaGetLastError());
    return 0;
}
