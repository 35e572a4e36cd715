// Micro-benchmark: how long does a block wait for its rows? One thread issues the k_rows_t load set (per row: 8 KB +
// 8 KB + 4 KB bulk copies, two rows = 40 KB) and the block waits on the mbarrier; clock64 around it, per block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bulk_latency scripts/ubench/bulk_latency.cu
//   ./bulk_latency [blocks_per_sm] [mode]      mode 0 = bulk copies (1 thread), 1 = the same split over 2 mbarriers,
//                                              2 = cp.async 16 B per thread, 3 = bulk copies after an L2 warm-up pass
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph)
{
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

constexpr int N = 1024, ROW = N * 8, OM = N * 4, SLOT = 2 * ROW + OM;

__global__ void __launch_bounds__(96) k(const char* h0, const char* om, long long* t_wait, long long* t_half, int mode, float* sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * SLOT);
    const int tid = threadIdx.x;
    const uint32_t r0 = blockIdx.x % (N / 2), tile = blockIdx.x / (N / 2);
    const uint32_t rows[2] = {r0 ? r0 : 0u, r0 ? N - r0 : N / 2};
    const char* H = h0 + size_t(tile) * N * ROW;
    const char* W = om + size_t(tile) * N * OM;
    long long t0 = clock64(), t1 = 0, t2 = 0;
    if (mode != 2) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (mode == 1) {
                mbar_expect(bar, SLOT);
                mbar_expect(bar + 1, SLOT);
            } else
                mbar_expect(bar, 2 * SLOT);
            for (int s = 0; s < 2; ++s) {
                uint64_t* b = mode == 1 ? bar + s : bar;
                bulk(sm + s * SLOT, H + size_t(rows[s]) * ROW, ROW, b);
                bulk(sm + s * SLOT + ROW, H + size_t(N - 1 - rows[s]) * ROW, ROW, b);
                bulk(sm + s * SLOT + 2 * ROW, W + size_t(rows[s]) * OM, OM, b);
            }
        }
        __syncthreads();
        mbar_wait(bar, 0);
        t1 = clock64();
        if (mode == 1) mbar_wait(bar + 1, 0);
        t2 = clock64();
    } else {
        for (int s = 0; s < 2; ++s) {
            for (int o = tid * 16; o < ROW; o += 96 * 16) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + s * SLOT + o)), "l"(H + size_t(rows[s]) * ROW + o));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + s * SLOT + ROW + o)), "l"(H + size_t(N - 1 - rows[s]) * ROW + o));
            }
            for (int o = tid * 16; o < OM; o += 96 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + s * SLOT + 2 * ROW + o)), "l"(W + size_t(rows[s]) * OM + o));
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        t1 = clock64();
        __syncthreads();
        t2 = clock64();
    }
    // consume, and hold the SM for a while like phases A and B would (so that blocks overlap as in the real kernel)
    float acc = 0.f;
    const float* f = reinterpret_cast<const float*>(sm);
    for (int rep = 0; rep < 40; ++rep)
        for (int i = tid; i < 2 * SLOT / 4; i += 96) acc = fmaf(acc, 1.0001f, f[i]);
    if (acc == 123.456f) *sink = acc;
    if (tid == 0) {
        t_half[blockIdx.x] = t1 - t0;
        t_wait[blockIdx.x] = t2 - t0;
    }
}

int main(int argc, char** argv)
{
    const int tiles = 8, blocks = tiles * N / 2;
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    char *h0, *om;
    long long *tw, *th;
    float* sink;
    cudaMalloc(&h0, size_t(tiles) * N * ROW);
    cudaMalloc(&om, size_t(tiles) * N * OM);
    cudaMemset(h0, 0, size_t(tiles) * N * ROW);
    cudaMemset(om, 0, size_t(tiles) * N * OM);
    cudaMalloc(&tw, blocks * 8);
    cudaMalloc(&th, blocks * 8);
    cudaMalloc(&sink, 4);
    char* flush;
    cudaMalloc(&flush, 512 << 20);
    const int smem = 2 * SLOT + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int rep = 0; rep < 3; ++rep) {
        if (mode != 3) cudaMemset(flush, rep, 512 << 20);               // evict h0 / omega from L2 (mode 3: leave what fits)
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<blocks, 96, smem>>>(h0, om, tw, th, mode == 3 ? 0 : mode, sink);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> w(blocks), h(blocks);
        cudaMemcpy(w.data(), tw, blocks * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(h.data(), th, blocks * 8, cudaMemcpyDeviceToHost);
        std::vector<long long> first(w.begin(), w.begin() + 740), rest(w.begin() + 740, w.end()), hrest(h.begin() + 740, h.end());
        std::sort(first.begin(), first.end());
        std::sort(rest.begin(), rest.end());
        std::sort(hrest.begin(), hrest.end());
        auto us = [&](long long c) { return c * 1e3 / clk; };
        printf("mode %d rep %d kernel %.1f us | wait cycles->us (max clock %d kHz): first wave median %.2f  later blocks p10 %.2f median %.2f p90 %.2f | first-half median %.2f\n",
               mode, rep, ms * 1e3, clk, us(first[370]), us(rest[rest.size() / 10]), us(rest[rest.size() / 2]), us(rest[rest.size() * 9 / 10]), us(hrest[hrest.size() / 2]));
    }
    return 0;
}
