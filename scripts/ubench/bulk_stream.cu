// micro-benchmark: how fast can ONE SM stream HBM into shared memory with cp.async.bulk, as a function of the copy
// size, the number of stages in flight and the number of issuing warps?  (design input for the STREAM pipeline)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_stream bulk_stream.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// each CTA streams [blockIdx.x * per_cta, +per_cta) bytes; `split` copies per tile (tile bytes are cut into `split` equal copies)
__global__ void k_stream(const unsigned char* src, size_t per_cta, uint32_t tile, uint32_t nst, uint32_t npw, uint32_t split) {
    extern __shared__ __align__(128) unsigned char sm[];
    const uint32_t bars = smem_u32(sm + (size_t)tile * nst);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
    const uint32_t ntiles = (uint32_t)(per_cta / tile);
    if (warp < npw && lane == 0) {
        for (uint32_t t = warp; t < ntiles; t += npw) {
            const uint32_t st = t % nst, round = t / nst;
            if (round > 0) mbar_wait(bars + 8 * st, (round - 1) & 1);  // the previous copy into this stage has landed
            mbar_expect_tx(bars + 8 * st, tile);
            const uint32_t part = tile / split;
            for (uint32_t c = 0; c < split; ++c)
                bulk_g2s(smem_u32(sm) + st * tile + c * part, base + (size_t)t * tile + c * part, part, bars + 8 * st);
        }
        // drain: wait for the last round of every stage this warp touched
        for (uint32_t t = ntiles > nst ? ntiles - nst : 0; t < ntiles; ++t)
            if (t % npw == warp) mbar_wait(bars + 8 * (t % nst), (t / nst) & 1);
    }
}

int main() {
    const size_t total = (size_t)2 << 30;
    unsigned char* d;
    cudaMalloc(&d, total);
    cudaMemset(d, 1, total);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    printf("tile_B stages producers split  ring_KB   ms   GB/s  GB/s_per_SM\n");
    const uint32_t tiles[] = {2048, 4096, 6144, 8192, 12288, 16384, 24576, 32768};
    for (uint32_t tile : tiles)
        for (uint32_t ring_kb : {48u, 96u, 192u})
            for (uint32_t npw : {1u, 2u, 4u})
                for (uint32_t split : {1u, 2u}) {
                    const uint32_t nst = ring_kb * 1024 / tile;
                    if (nst < 2 || nst < npw) continue;
                    const size_t per_cta = total / sms / tile * tile;
                    const size_t smem = (size_t)tile * nst + 8 * nst;
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(a);
                        k_stream<<<sms, 32 * npw, smem>>>(d, per_cta, tile, nst, npw, split);
                        cudaEventRecord(b);
                        cudaEventSynchronize(b);
                    }
                    float ms = 0;
                    cudaEventElapsedTime(&ms, a, b);
                    cudaError_t e = cudaGetLastError();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    const double gbs = (double)per_cta * sms / ms / 1e6;
                    printf("%6u %6u %9u %5u %8u %7.3f %7.0f %7.1f\n", tile, nst, npw, split, ring_kb, ms, gbs, gbs / sms);
                }
    return 0;
}
