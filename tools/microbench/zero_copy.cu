// Micro-benchmark: how fast can kernels write into mapped pinned HOST memory (the sparse read-back of k_readback), beside
// the copy engine?  Variants: 16-byte vector stores (512 bytes per warp instruction, what k_readback does), 1-D bulk stores
// from shared memory (cp.async.bulk.global.shared::cta) of 128 / 384 / 1024 / 4096-byte pieces, and cudaMemcpyAsync D2H.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a zero_copy.cu -o zero_copy
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// every warp copies 512-byte pieces: src (device) -> dst (host mapped), piece index strided over all warps
__global__ void __launch_bounds__(256) k_vec(const float4 *src, float4 *dst, long long pieces)
{
    const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (long long)gridDim.x * 8;
    const int lane = threadIdx.x & 31;
    for (long long p = warp; p < pieces; p += nwarps) dst[p * 32 + lane] = src[p * 32 + lane];
}

// every warp stages `bytes` from device memory into its shared-memory slot and one lane bulk-stores it to the host
__global__ void __launch_bounds__(256) k_bulk(const float4 *src, float4 *dst, long long pieces, int bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *slot = reinterpret_cast<float4 *>(smem + (size_t)wid * bytes);
    const long long warp = (long long)blockIdx.x * 8 + wid, nwarps = (long long)gridDim.x * 8;
    const int v = bytes / 16;
    for (long long p = warp; p < pieces; p += nwarps) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int i = lane; i < v; i += 32) slot[i] = src[p * v + i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + p * v),
                         "r"((unsigned)__cvta_generic_to_shared(slot)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const size_t total = 256ull << 20;
    float4 *src, *hst, *hdev;
    CK(cudaMalloc(&src, total));
    CK(cudaMemset(src, 1, total));
    CK(cudaHostAlloc(&hst, total, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&hdev, hst, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(hst, src, total, cudaMemcpyDeviceToHost));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) printf("copy engine D2H              %7.2f GB/s\n", total / ms / 1e6);
    }
    for (int grid : {37, 74, 148, 296, 592}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_vec<<<grid, 256>>>(src, hdev, (long long)(total / 512));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("16-byte stores   grid %4d    %7.2f GB/s\n", grid, total / ms / 1e6);
        }
    }
    for (int bytes : {128, 384, 1024, 4096}) {
        for (int grid : {37, 148, 592}) {
            CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                k_bulk<<<grid, 256, 8 * bytes>>>(src, hdev, (long long)(total / bytes), bytes);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep) printf("bulk %4d B      grid %4d    %7.2f GB/s\n", bytes, grid, total / ms / 1e6);
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
