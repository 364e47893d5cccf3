// rmse.cu — test-set RMSE, fully on device.
// Reference: calrmse / dot (src/tools.cpp:184-198, 235-248) on the CPU path; GPU_rmse
// (cuda_src/CUDA_AUX.cu:3-27) plus a D2H copy of one float per test rating and a serial host sum
// (cuda_src/CCD_CUDA.cu:385-401) on the GPU path.  Here: one thread per test rating, prediction =
// sum over ranks of the FP32 product W*H promoted into a double accumulator (the CPU path's
// arithmetic), squared error in double, warp + CTA reduction, one partial per CTA; the last CTA to
// finish adds the partials in CTA order (fixed tree: the sum does not depend on scheduling, so two
// runs over the same factors give the same bits), and a single scalar comes back to the host.
#include "rmse.cuh"

namespace mf {
namespace {

// CTA-wide sum of `local`, one partial per CTA, the last CTA to finish adds the partials in CTA order into acc[0]
// (acc[1 .. gridDim.x] = partials, ticket word behind them: rmse_scratch_doubles)
__device__ __forceinline__ void finish_sum(double local, double* __restrict__ acc) {
    __shared__ double part[8];
    __shared__ bool s_last;
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            acc[1 + blockIdx.x] = v;
            __threadfence();
            unsigned* ticket = reinterpret_cast<unsigned*>(acc + 1 + gridDim.x);
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        double v = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) v += __ldcg(acc + 1 + b);
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            acc[0] = v;
            *reinterpret_cast<unsigned*>(acc + 1 + gridDim.x) = 0u;
        }
    }
}

// Per-rank incremental test RMSE — calrmse_r1 (src/tools.cpp:260-270; its caller is the commented-out verbose block of
// src/CCD.cpp:141-148): the test residual loses the rank's new product and gets its old one back, in FP32 like the
// reference (res -= u*v - u_old*v_old), the squares are summed in double.
__global__ void __launch_bounds__(256) k_rmse_r1(int64_t nt, const uint32_t* __restrict__ trow, const uint32_t* __restrict__ tcol,
                                                  float* __restrict__ tres, const float* __restrict__ u, const float* __restrict__ v,
                                                  const float* __restrict__ u_old, const float* __restrict__ v_old,
                                                  double* __restrict__ acc) {
    double local = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nt; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = trow[e], c = tcol[e];
        const float x = __fsub_rn(tres[e], __fsub_rn(__fmul_rn(u[r], v[c]), __fmul_rn(u_old[r], v_old[c])));
        tres[e] = x;
        local += (double)__fmul_rn(x, x);
    }
    finish_sum(local, acc);
}

// Function decrease of one solve sweep (the -e stop rule of CCDR1, restated in oracle/mf_oracle.c orc_ccdpp_ex): sum over
// the segments of (lambda*deg + h) * (old - new)^2, h re-added from the segment's partial-sum slots in slot order.
__global__ void __launch_bounds__(256) k_fundec(int64_t nseg, const uint32_t* __restrict__ seg_ptr, const uint32_t* __restrict__ slot_ptr,
                                                 const float2* __restrict__ partials, float lambda, const float* __restrict__ v_old,
                                                 const float* __restrict__ v_new, double* __restrict__ acc) {
    double local = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t deg = seg_ptr[s + 1] - seg_ptr[s];
        if (deg == 0u) continue;
        float h = 0.0f;
        for (uint32_t q = slot_ptr[s]; q < slot_ptr[s + 1]; ++q) h += __ldcg(partials + q).y;
        const double d = (double)v_old[s] - (double)v_new[s];
        local += (double)(lambda * deg + h) * d * d;
    }
    finish_sum(local, acc);
}

__global__ void __launch_bounds__(256) k_rmse(int64_t nt, const uint32_t* __restrict__ trow, const uint32_t* __restrict__ tcol,
                                               const float* __restrict__ tval, const float* __restrict__ W,
                                               const float* __restrict__ H, int k, int64_t w_rank_stride,
                                               int64_t w_row_stride, int64_t h_rank_stride, int64_t h_row_stride,
                                               double* __restrict__ acc) {
    // acc[0] = result, acc[1 .. gridDim.x] = per-CTA partials, ticket counter behind them (rmse_scratch_doubles)
    double local = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nt; e += (int64_t)gridDim.x * blockDim.x) {
        const float* w = W + (int64_t)trow[e] * w_row_stride;
        const float* h = H + (int64_t)tcol[e] * h_row_stride;
        double pred = 0.0;
        for (int t = 0; t < k; ++t) pred += (double)__fmul_rn(w[t * w_rank_stride], h[t * h_rank_stride]);
        double err = -(double)tval[e];
        err += pred;
        local += err * err;
    }
    finish_sum(local, acc);
}

// predictions for arbitrary pairs: the same arithmetic, one thread per pair, the double goes out unrounded
__global__ void __launch_bounds__(256) k_predict(int64_t n, const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                                  const float* __restrict__ W, const float* __restrict__ H, int k,
                                                  int64_t w_rank_stride, int64_t w_row_stride, int64_t h_rank_stride,
                                                  int64_t h_row_stride, double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float* w = W + (int64_t)row[e] * w_row_stride;
        const float* h = H + (int64_t)col[e] * h_row_stride;
        double pred = 0.0;
        for (int t = 0; t < k; ++t) pred += (double)__fmul_rn(w[t * w_rank_stride], h[t * h_rank_stride]);
        out[e] = pred;
    }
}

}  // namespace

size_t rmse_scratch_doubles(int sm_count) { return 1 + (size_t)sm_count * 8 + 1; }

int rmse_accumulate(int64_t nt, const uint32_t* trow, const uint32_t* tcol, const float* tval, const float* W,
                    const float* H, int k, int64_t w_rank_stride, int64_t w_row_stride, int64_t h_rank_stride,
                    int64_t h_row_stride, double* d_acc, int sm_count, cudaStream_t st) {
    if (nt <= 0) {
        MF_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double), st));
        return MF_OK;
    }
    int64_t blocks = (nt + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    MF_CUDA(cudaMemsetAsync(d_acc + 1 + blocks, 0, sizeof(double), st));  // ticket
    k_rmse<<<(unsigned)blocks, 256, 0, st>>>(nt, trow, tcol, tval, W, H, k, w_rank_stride, w_row_stride, h_rank_stride,
                                            h_row_stride, d_acc);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int rmse_r1_accumulate(int64_t nt, const uint32_t* trow, const uint32_t* tcol, float* tres, const float* u, const float* v,
                       const float* u_old, const float* v_old, double* d_acc, int sm_count, cudaStream_t st) {
    if (nt <= 0) {
        MF_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double), st));
        return MF_OK;
    }
    int64_t blocks = (nt + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    MF_CUDA(cudaMemsetAsync(d_acc + 1 + blocks, 0, sizeof(double), st));  // ticket
    k_rmse_r1<<<(unsigned)blocks, 256, 0, st>>>(nt, trow, tcol, tres, u, v, u_old, v_old, d_acc);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int fundec_accumulate(int64_t nseg, const uint32_t* seg_ptr, const uint32_t* slot_ptr, const float2* partials, float lambda,
                      const float* v_old, const float* v_new, double* d_acc, int sm_count, cudaStream_t st) {
    if (nseg <= 0) {
        MF_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double), st));
        return MF_OK;
    }
    int64_t blocks = (nseg + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    MF_CUDA(cudaMemsetAsync(d_acc + 1 + blocks, 0, sizeof(double), st));  // ticket
    k_fundec<<<(unsigned)blocks, 256, 0, st>>>(nseg, seg_ptr, slot_ptr, partials, lambda, v_old, v_new, d_acc);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int predict_pairs(int64_t n, const uint32_t* row, const uint32_t* col, const float* W, const float* H, int k,
                  int64_t w_rank_stride, int64_t w_row_stride, int64_t h_rank_stride, int64_t h_row_stride, double* out,
                  int sm_count, cudaStream_t st) {
    if (n <= 0) return MF_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    k_predict<<<(unsigned)blocks, 256, 0, st>>>(n, row, col, W, H, k, w_rank_stride, w_row_stride, h_rank_stride, h_row_stride, out);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

}  // namespace mf
