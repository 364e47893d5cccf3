#pragma once
#include "common.cuh"
namespace mf {
// *d_acc <- sum over test ratings of (w_i . h_j - r)^2.  Strides describe the factor layout:
// CCD++ (W[t*ld + i]): rank_stride = ld, row_stride = 1;  ALS (W[i*k + t]): rank_stride = 1, row_stride = k.
// d_acc needs rmse_scratch_doubles(sm_count) doubles: [0] result, per-CTA partials, ticket word.
size_t rmse_scratch_doubles(int sm_count);
int rmse_accumulate(int64_t nt, const uint32_t* trow, const uint32_t* tcol, const float* tval, const float* W,
                    const float* H, int k, int64_t w_rank_stride, int64_t w_row_stride, int64_t h_rank_stride,
                    int64_t h_row_stride, double* d_acc, int sm_count, cudaStream_t st);
// per-rank incremental RMSE (calrmse_r1, src/tools.cpp:260-270): tres[e] -= u[r]*v[c] - u_old[r]*v_old[c]; *d_acc <- sum tres^2
int rmse_r1_accumulate(int64_t nt, const uint32_t* trow, const uint32_t* tcol, float* tres, const float* u, const float* v,
                       const float* u_old, const float* v_old, double* d_acc, int sm_count, cudaStream_t st);
// function decrease of one solve sweep over this shard's segments: *d_acc <- sum (lambda*deg + h) * (old - new)^2
// (v_old / v_new already offset to the shard's first segment; h re-added from the partial-sum slots)
int fundec_accumulate(int64_t nseg, const uint32_t* seg_ptr, const uint32_t* slot_ptr, const float2* partials, float lambda,
                      const float* v_old, const float* v_new, double* d_acc, int sm_count, cudaStream_t st);
// out[e] <- w_row[e] . h_col[e] (FP32 products, FP64 sum in rank order — src/extras.cpp:165-168); device pointers
int predict_pairs(int64_t n, const uint32_t* row, const uint32_t* col, const float* W, const float* H, int k,
                  int64_t w_rank_stride, int64_t w_row_stride, int64_t h_rank_stride, int64_t h_row_stride, double* out,
                  int sm_count, cudaStream_t st);
}  // namespace mf
