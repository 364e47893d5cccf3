// micro-benchmark: issue rate of the legacy tensor path (mma.sync.m16n8k8 tf32, SASS HMMA) on sm_100a, per SM, as a
// function of resident warps — design input for a 3xTF32 Gram in the ALS kernel (1 MMA = 16*8*8 = 1024 MAC)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void k_mma(float* out, int iters) {
    float acc[NACC][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) mma_tf32(acc[n], a, b);
    }
    float s = 0.f;
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += acc[n][i];
    if (s == 123.456f) out[0] = s;
}

__global__ void k_ffma(float* out, int iters) {
    float acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    const float x = 1.0001f, y = 0.5f + threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

int main() {
    float* d;
    cudaMalloc(&d, 4);
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    printf("sms %d clock %.0f MHz\n", sms, khz / 1e3);
    for (int warps : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(a); k_mma<8><<<sms, warps * 32>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b); }
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        const double mma_per_sm = (double)warps * 8 * iters;
        const double cyc = ms * 1e-3 * khz * 1e3;
        printf("mma.sync tf32: %2d warps/SM: %.3f ms  %.2f cycles per MMA per SM (%.0f MAC/clk/SM, %.1f TFLOP/s TF32 dense)\n", warps, ms,
               cyc / mma_per_sm, 1024.0 * mma_per_sm / cyc, 2.0 * 1024.0 * mma_per_sm * sms / (ms * 1e-3) / 1e12);
    }
    for (int warps : {8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(a); k_ffma<<<sms, warps * 32>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b); }
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        const double fma_per_sm = (double)warps * 32 * 16 * iters;
        const double cyc = ms * 1e-3 * khz * 1e3;
        printf("ffma: %2d warps/SM: %.3f ms  %.0f FMA/clk/SM, %.1f TFLOP/s FP32\n", warps, ms, fma_per_sm / cyc, 2.0 * fma_per_sm * sms / (ms * 1e-3) / 1e12);
    }
    return 0;
}
