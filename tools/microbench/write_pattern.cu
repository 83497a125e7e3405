// Micro-benchmark: how fast can 28 B/pixel of frame buffers (z [H,W], colour [H,W,3], normals [H,W,3], V views) be
// written on B200 for different tile shapes?  Guides the tile geometry of k_raster.  Build: nvcc -O3 -arch=sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) fill_tiles(float *z, float *c, float *n, int W, int H)
{
    const int tilesX = W / TW;
    const int tx = blockIdx.x % tilesX, ty = blockIdx.x / tilesX, view = blockIdx.y;
    const long long slab = (long long)view * W * H;
    const int x0 = tx * TW, y0 = ty * TH;
    constexpr int ZQ = TW / 4, CQ = TW * 3 / 4;
    for (int i = threadIdx.x; i < TH * ZQ; i += NT) {
        const int r = i / ZQ, q = i % ZQ;
        reinterpret_cast<float4 *>(z + slab + (long long)(y0 + r) * W + x0)[q] = make_float4(1e6f, 1e6f, 1e6f, 1e6f);
    }
    for (int i = threadIdx.x; i < TH * CQ; i += NT) {
        const int r = i / CQ, q = i % CQ;
        const long long o = (slab + (long long)(y0 + r) * W + x0) * 3;
        reinterpret_cast<float4 *>(c + o)[q] = make_float4(0, 0, 0, 0);
        reinterpret_cast<float4 *>(n + o)[q] = make_float4(0, 0, 0, 0);
    }
}

// the store mapping k_raster uses: thread -> (row = tid/8, 16-byte column tid%8 (+8,+16)); optional persistent walk,
// optional static shared memory footprint (limits resident CTAs like the real kernel)
template <int SMEM, bool PERSIST, int MINB>
__global__ void __launch_bounds__(256, MINB) fill_like_raster(float *z, float *c, float *n, int W, int H, int V, const unsigned *cnt)
{
    __shared__ float pad[SMEM / 4 + 1];
    if (SMEM && threadIdx.x == 9999) pad[0] = 1.f;
    const int tilesX = W / 32, nTiles = tilesX * (H / 32);
    const long long nAll = (long long)nTiles * V;
    const long long step = PERSIST ? gridDim.x : nAll;
    unsigned n_next = cnt ? cnt[blockIdx.x] : 0u;
    for (long long t = blockIdx.x; t < nAll; t += step) {
        const unsigned nn = n_next;
        if (cnt && t + step < nAll) n_next = cnt[t + step];
        if (nn) continue;
        const int view = (int)(t / nTiles), tile = (int)(t % nTiles);
        const int ty = tile / tilesX, tx = tile % tilesX;
        const long long slab = (long long)view * W * H;
        const int q = threadIdx.x & 7;
        for (int r = threadIdx.x >> 3; r < 32; r += 32) {
            const long long rowpix = slab + (long long)(ty * 32 + r) * W + tx * 32;
            reinterpret_cast<float4 *>(z + rowpix)[q] = make_float4(1e6f, 1e6f, 1e6f, 1e6f);
            float4 *o = reinterpret_cast<float4 *>(c + rowpix * 3);
            o[q] = make_float4(0, 0, 0, 0); o[q + 8] = make_float4(0, 0, 0, 0); o[q + 16] = make_float4(0, 0, 0, 0);
            o = reinterpret_cast<float4 *>(n + rowpix * 3);
            o[q] = make_float4(0, 0, 0, 0); o[q + 8] = make_float4(0, 0, 0, 0); o[q + 16] = make_float4(0, 0, 0, 0);
        }
    }
}

// persistent walk in contiguous groups of G tiles per CTA visit: one coalesced load brings G counts, the group's
// counts for the NEXT visit are prefetched a whole visit (G tiles of stores) ahead
template <int SMEM, int MINB, int G>
__global__ void __launch_bounds__(256, MINB) fill_grouped(float *z, float *c, float *n, int W, int H, int V, const unsigned *cnt)
{
    __shared__ float pad[SMEM / 4 + 1];
    if (SMEM && threadIdx.x == 9999) pad[0] = 1.f;
    const int tilesX = W / 32, nTiles = tilesX * (H / 32);
    const long long nAll = (long long)nTiles * V;
    const int lane = threadIdx.x & 31;
    long long base = (long long)blockIdx.x * G;
    unsigned mine_next = (lane < G && base + lane < nAll) ? cnt[base + lane] : 1u;
    for (; base < nAll; base += (long long)gridDim.x * G) {
        const unsigned mine = mine_next;
        const long long nb = base + (long long)gridDim.x * G;
        mine_next = (lane < G && nb + lane < nAll) ? cnt[nb + lane] : 1u;
        for (int k = 0; k < G; ++k) {
            const unsigned nn = __shfl_sync(0xffffffffu, mine, k);
            if (nn) continue;
            const long long t = base + k;
            const int view = (int)(t / nTiles), tile = (int)(t % nTiles);
            const int ty = tile / tilesX, tx = tile % tilesX;
            const long long slab = (long long)view * W * H;
            const int q = threadIdx.x & 7;
            const int r = threadIdx.x >> 3;
            const long long rowpix = slab + (long long)(ty * 32 + r) * W + tx * 32;
            reinterpret_cast<float4 *>(z + rowpix)[q] = make_float4(1e6f, 1e6f, 1e6f, 1e6f);
            float4 *o = reinterpret_cast<float4 *>(c + rowpix * 3);
            o[q] = make_float4(0, 0, 0, 0); o[q + 8] = make_float4(0, 0, 0, 0); o[q + 16] = make_float4(0, 0, 0, 0);
            o = reinterpret_cast<float4 *>(n + rowpix * 3);
            o[q] = make_float4(0, 0, 0, 0); o[q + 8] = make_float4(0, 0, 0, 0); o[q + 16] = make_float4(0, 0, 0, 0);
        }
    }
}
template <int SMEM, int MINB, int G>
void run_grouped(const char *name, float *z, float *c, float *n, int W, int H, int V, const unsigned *cnt)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = 148 * MINB;
    for (int i = 0; i < 3; ++i) fill_grouped<SMEM, MINB, G><<<grid, 256>>>(z, c, n, W, H, V, cnt);
    CK(cudaEventRecord(a));
    for (int i = 0; i < 10; ++i) fill_grouped<SMEM, MINB, G><<<grid, 256>>>(z, c, n, W, H, V, cnt);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-44s %7.2f us/view  %7.0f GB/s\n", name, ms / 10 / V * 1000, 28.0 * W * H * V / (ms / 10 / 1000) / 1e9);
}

template <int SMEM, bool PERSIST, int MINB>
void run_like(const char *name, float *z, float *c, float *n, int W, int H, int V, const unsigned *cnt)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = PERSIST ? 148 * MINB : (W / 32) * (H / 32) * V;
    for (int i = 0; i < 3; ++i) fill_like_raster<SMEM, PERSIST, MINB><<<grid, 256>>>(z, c, n, W, H, V, cnt);
    CK(cudaEventRecord(a));
    for (int i = 0; i < 10; ++i) fill_like_raster<SMEM, PERSIST, MINB><<<grid, 256>>>(z, c, n, W, H, V, cnt);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-44s %7.2f us/view  %7.0f GB/s\n", name, ms / 10 / V * 1000, 28.0 * W * H * V / (ms / 10 / 1000) / 1e9);
}

__global__ void fill_linear(float4 *p, long long n4)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_float4(0, 0, 0, 0);
}

template <int TW, int TH, int NT>
void run(const char *name, float *z, float *c, float *n, int W, int H, int V)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    dim3 grid((W / TW) * (H / TH), V);
    for (int i = 0; i < 3; ++i) fill_tiles<TW, TH, NT><<<grid, NT>>>(z, c, n, W, H);
    CK(cudaEventRecord(a));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) fill_tiles<TW, TH, NT><<<grid, NT>>>(z, c, n, W, H);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = 28.0 * W * H * V;
    printf("%-28s %7.2f us/view  %7.0f GB/s\n", name, ms / reps / V * 1000, bytes / (ms / reps / 1000) / 1e9);
}

int main()
{
    const int W = 1024, H = 1024, V = 32;
    float *z, *c, *n;
    CK(cudaMalloc(&z, (size_t)W * H * V * 4)); CK(cudaMalloc(&c, (size_t)W * H * V * 12)); CK(cudaMalloc(&n, (size_t)W * H * V * 12));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float ms;
    for (int k = 0; k < 2; ++k) {
        CK(cudaEventRecord(a));
        for (int i = 0; i < 10; ++i) { CK(cudaMemsetAsync(z, 0, (size_t)W * H * V * 4)); CK(cudaMemsetAsync(c, 0, (size_t)W * H * V * 12)); CK(cudaMemsetAsync(n, 0, (size_t)W * H * V * 12)); }
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("%-28s %7.2f us/view  %7.0f GB/s\n", "cudaMemsetAsync x3", ms / 10 / V * 1000, 28.0 * W * H * V / (ms / 10 / 1000) / 1e9);
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        for (int k = 0; k < 2; ++k) {
            CK(cudaEventRecord(a));
            for (int i = 0; i < 10; ++i) { fill_linear<<<blocks, 256>>>((float4 *)z, (long long)W * H * V / 4); fill_linear<<<blocks, 256>>>((float4 *)c, (long long)W * H * V * 3 / 4); fill_linear<<<blocks, 256>>>((float4 *)n, (long long)W * H * V * 3 / 4); }
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("fill_linear grid=%-5d        %7.2f us/view  %7.0f GB/s\n", blocks, ms / 10 / V * 1000, 28.0 * W * H * V / (ms / 10 / 1000) / 1e9);
    }
    run<32, 32, 256>("tile 32x32 nt256", z, c, n, W, H, V);
    run<32, 32, 128>("tile 32x32 nt128", z, c, n, W, H, V);
    run<32, 16, 256>("tile 32x16 nt256", z, c, n, W, H, V);
    run<32, 8, 256>("tile 32x8 nt256", z, c, n, W, H, V);
    run<64, 16, 256>("tile 64x16 nt256", z, c, n, W, H, V);
    run<64, 32, 256>("tile 64x32 nt256", z, c, n, W, H, V);
    run<128, 8, 256>("tile 128x8 nt256", z, c, n, W, H, V);
    run<128, 16, 256>("tile 128x16 nt256", z, c, n, W, H, V);
    run<256, 4, 256>("tile 256x4 nt256", z, c, n, W, H, V);
    run<256, 8, 256>("tile 256x8 nt256", z, c, n, W, H, V);
    run<1024, 1, 256>("tile 1024x1 nt256", z, c, n, W, H, V);
    run<1024, 4, 256>("tile 1024x4 nt256", z, c, n, W, H, V);
    run<64, 64, 256>("tile 64x64 nt256", z, c, n, W, H, V);
    unsigned *cnt; CK(cudaMalloc(&cnt, 1024 * 32 * 4)); CK(cudaMemset(cnt, 0, 1024 * 32 * 4));
    run_like<0, false, 8>("raster map, one CTA per tile", z, c, n, W, H, V, nullptr);
    run_like<0, false, 8>("raster map, one CTA per tile, cnt load", z, c, n, W, H, V, cnt);
    run_like<34000, false, 5>("raster map, CTA/tile, 34KB smem", z, c, n, W, H, V, cnt);
    run_like<0, true, 8>("raster map, persistent 8/SM", z, c, n, W, H, V, cnt);
    run_like<34000, true, 5>("raster map, persistent 5/SM 34KB", z, c, n, W, H, V, cnt);
    run_like<34000, true, 4>("raster map, persistent 4/SM 34KB", z, c, n, W, H, V, cnt);
    run_like<34000, true, 6>("raster map, persistent 6/SM 34KB", z, c, n, W, H, V, cnt);
    run_like<0, true, 8>("persistent 8/SM, no cnt load", z, c, n, W, H, V, nullptr);
    run_grouped<34000, 5, 2>("persistent 5/SM grouped x2", z, c, n, W, H, V, cnt);
    run_grouped<34000, 5, 4>("persistent 5/SM grouped x4", z, c, n, W, H, V, cnt);
    run_grouped<34000, 5, 8>("persistent 5/SM grouped x8", z, c, n, W, H, V, cnt);
    run_grouped<34000, 5, 32>("persistent 5/SM grouped x32", z, c, n, W, H, V, cnt);
    run_grouped<0, 8, 8>("persistent 8/SM grouped x8", z, c, n, W, H, V, cnt);
    return 0;
}
