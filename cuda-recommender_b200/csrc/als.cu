// als.cu — ALS half-step: per-segment Gram assembly + Cholesky solve (sm_100a).
//
// Reference: ALS_OMP row/column bodies (src/ALS.cpp:98-158, :161-219) with Mt_byM_multiply (:66-79) and
// the explicit Cholesky inverse (:6-64); GPU kernels being replaced: updateW_overH_kernel /
// updateH_overW_kernel (cuda_src/ALS_CUDA.cu:81-181), which give every row to ONE thread and
// malloc() k*k floats per thread on the device heap.
//
// Here one CTA owns a segment (a user row or an item column) at a time:
//   * segments are visited longest-first (degree-binned order) through an atomic queue;
//   * the segment's factor rows Y[idx] are staged through shared memory in batches of 32 rows with asynchronous
//     copies (cp.async, double-buffered: batch b+1 is in flight while batch b is accumulated; the row ids of
//     batch b+2 travel in registers);
//   * A = Y_O^T Y_O is accumulated in registers as 4x4 tiles of the lower triangle (each thread owns
//     up to MAXT tiles for the whole segment), b = Y_O^T r alongside;
//   * A + lambda*I (lambda NOT scaled by |O|, src/ALS.cpp:120-122) is factored in shared memory
//     (in-place lower Cholesky) and x is obtained by two triangular solves — no explicit inverse;
//   * an empty segment writes zeros (src/ALS.cpp:151-157).
#include "session.cuh"

namespace mf {
namespace {

constexpr int kBatch = 32;  // factor rows staged per step

// segment order: longest-first by degree bin (bit length of the degree), via per-bin cursors
__global__ void k_order_by_bin(int64_t nseg, const uint32_t* __restrict__ ptr, unsigned* __restrict__ cursor /*[33]*/,
                               uint32_t* __restrict__ order) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t d = ptr[s + 1] - ptr[s];
    const int b = 32 - __clz(d);
    order[atomicAdd(&cursor[32 - b], 1u)] = (uint32_t)s;
}
__global__ void k_bin_count(int64_t nseg, const uint32_t* __restrict__ ptr, unsigned* __restrict__ count /*[33]*/) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t d = ptr[s + 1] - ptr[s];
    atomicAdd(&count[32 - (32 - __clz(d))], 1u);
}
__global__ void k_bin_scan(unsigned* __restrict__ count, unsigned* __restrict__ cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned acc = 0;
        for (int i = 0; i < 33; ++i) { cursor[i] = acc; acc += count[i]; }
    }
}

// 4/8/16-byte asynchronous global->shared copies (LDGSTS): the factor-row gather of the next batch runs while the
// current batch is being accumulated
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

// VW = floats per copy (4 when k % 4 == 0, 2 when k is even, else 1): every factor row starts VW-aligned
template <int TPS, int MAXT, int VW>
__global__ void __launch_bounds__(TPS) k_als_half(int64_t nseg, const uint32_t* __restrict__ order, unsigned* __restrict__ queue,
                                                  const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx,
                                                  const float* __restrict__ val, const float* __restrict__ Y,
                                                  float* __restrict__ X, int k, int kp, float lambda) {
    extern __shared__ __align__(16) float sm[];
    float* A = sm;                           // [kp*kp]  lower triangle used
    float* bvec = A + kp * kp;               // [kp]
    float* Ys = bvec + kp;                   // [2][kBatch*kp]  double-buffered staged factor rows
    float* rs = Ys + 2 * kBatch * kp;        // [2][kBatch]     ratings of the staged rows
    uint32_t* sidx = reinterpret_cast<uint32_t*>(rs + 2 * kBatch);  // [2][kBatch] row ids of the batch to be fetched next
    __shared__ unsigned s_next;

    const int tid = threadIdx.x;
    const int nb = kp >> 2;                   // 4x4 tile grid
    const int ntiles = nb * (nb + 1) / 2;
    int ti[MAXT], tj[MAXT];
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
        int q = tid + m * TPS;
        if (q < ntiles) {
            int I = (int)((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
            while ((I + 1) * (I + 2) / 2 <= q) ++I;
            while (I * (I + 1) / 2 > q) --I;
            ti[m] = I;
            tj[m] = q - I * (I + 1) / 2;
        } else {
            ti[m] = -1;
            tj[m] = 0;
        }
    }
    // the pad columns k..kp-1 of both staging buffers stay zero for the whole kernel (the copies never touch them)
    for (int e = tid; e < 2 * kBatch * kp; e += TPS) Ys[e] = 0.0f;
    const int vec_per_row = k / VW;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(queue, 1u);
        __syncthreads();
        const unsigned qi = s_next;
        if (qi >= nseg) break;
        const int64_t s = order[qi];
        const uint32_t lo = ptr[s], hi = ptr[s + 1];
        float* x = X + s * k;
        if (hi == lo) {
            for (int c = tid; c < k; c += TPS) x[c] = 0.0f;
            continue;
        }
        float acc[MAXT][16];
#pragma unroll
        for (int m = 0; m < MAXT; ++m)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[m][e] = 0.0f;
        float bacc = 0.0f;

        const int nbatch = (int)((hi - lo + kBatch - 1) / kBatch);
        // issue the asynchronous gather of batch `b` into buffer b&1, row ids taken from sidx[b&1]
        auto issue_rows = [&](int b) {
            const int nrow = (int)min((uint32_t)kBatch, hi - (lo + (uint32_t)b * kBatch));
            float* dst = Ys + (b & 1) * kBatch * kp;
            const uint32_t* ids = sidx + (b & 1) * kBatch;
            for (int e = tid; e < nrow * vec_per_row; e += TPS) {
                const int r = e / vec_per_row, c = (e - r * vec_per_row) * VW;
                cp_async<VW * 4>(dst + r * kp + c, Y + (size_t)ids[r] * k + c);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // prologue: ids + ratings of batches 0 and 1, rows of batch 0
        if (tid < kBatch) {
            const uint32_t e0 = lo + tid, e1 = lo + kBatch + tid;
            if (e0 < hi) { sidx[tid] = __ldg(idx + e0); rs[tid] = __ldg(val + e0); }
            if (e1 < hi) { sidx[kBatch + tid] = __ldg(idx + e1); rs[kBatch + tid] = __ldg(val + e1); }
        }
        __syncthreads();
        issue_rows(0);

        for (int b = 0; b < nbatch; ++b) {
            const int nrow = (int)min((uint32_t)kBatch, hi - (lo + (uint32_t)b * kBatch));
            if (b + 1 < nbatch) issue_rows(b + 1); else asm volatile("cp.async.commit_group;" ::: "memory");
            // ids + ratings of batch b+2 travel in registers while batch b is accumulated
            uint32_t nid = 0; float nr = 0.0f;
            const uint32_t e2 = lo + (uint32_t)(b + 2) * kBatch + tid;
            const bool has2 = tid < kBatch && b + 2 < nbatch && e2 < hi;
            if (has2) { nid = __ldg(idx + e2); nr = __ldg(val + e2); }
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // batch b has landed (batch b+1 may be in flight)
            __syncthreads();
            const float* Yb = Ys + (b & 1) * kBatch * kp;
            const float* rb = rs + (b & 1) * kBatch;
#pragma unroll
            for (int m = 0; m < MAXT; ++m) {
                if (ti[m] < 0) continue;
                const float* yi = Yb + 4 * ti[m];
                const float* yj = Yb + 4 * tj[m];
                for (int r = 0; r < nrow; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(yi + r * kp);
                    const float4 c = *reinterpret_cast<const float4*>(yj + r * kp);
                    acc[m][0] = fmaf(a.x, c.x, acc[m][0]);   acc[m][1] = fmaf(a.x, c.y, acc[m][1]);
                    acc[m][2] = fmaf(a.x, c.z, acc[m][2]);   acc[m][3] = fmaf(a.x, c.w, acc[m][3]);
                    acc[m][4] = fmaf(a.y, c.x, acc[m][4]);   acc[m][5] = fmaf(a.y, c.y, acc[m][5]);
                    acc[m][6] = fmaf(a.y, c.z, acc[m][6]);   acc[m][7] = fmaf(a.y, c.w, acc[m][7]);
                    acc[m][8] = fmaf(a.z, c.x, acc[m][8]);   acc[m][9] = fmaf(a.z, c.y, acc[m][9]);
                    acc[m][10] = fmaf(a.z, c.z, acc[m][10]); acc[m][11] = fmaf(a.z, c.w, acc[m][11]);
                    acc[m][12] = fmaf(a.w, c.x, acc[m][12]); acc[m][13] = fmaf(a.w, c.y, acc[m][13]);
                    acc[m][14] = fmaf(a.w, c.z, acc[m][14]); acc[m][15] = fmaf(a.w, c.w, acc[m][15]);
                }
            }
            if (tid < k)
                for (int r = 0; r < nrow; ++r) bacc = fmaf(rb[r], Yb[r * kp + tid], bacc);
            __syncthreads();  // everyone is done with buffer b&1 and its ratings / ids
            if (has2) { sidx[(b & 1) * kBatch + tid] = nid; rs[(b & 1) * kBatch + tid] = nr; }
            __syncthreads();  // ids of batch b+2 are in place before the next iteration issues its rows
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // tiles -> A (lower triangle incl. the whole diagonal tiles), + lambda on the diagonal
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            if (ti[m] < 0) continue;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int i = 4 * ti[m] + (e >> 2), j = 4 * tj[m] + (e & 3);
                float v = acc[m][e];
                if (i == j) v += (i < k) ? lambda : 1.0f;  // padded diagonal -> 1 keeps the factorisation finite
                A[i * kp + j] = v;
            }
        }
        if (tid < kp) bvec[tid] = tid < k ? bacc : 0.0f;
        __syncthreads();

        // in-place lower Cholesky, right-looking
        for (int j = 0; j < k; ++j) {
            if (tid == 0) A[j * kp + j] = sqrtf(A[j * kp + j]);
            __syncthreads();
            const float inv = 1.0f / A[j * kp + j];
            for (int i = j + 1 + tid; i < k; i += TPS) A[i * kp + j] *= inv;
            __syncthreads();
            const int n = k - 1 - j;
            for (int e = tid; e < n * n; e += TPS) {
                const int a = e / n, b = e - a * n;
                if (b <= a) {
                    const int i = j + 1 + a, l = j + 1 + b;
                    A[i * kp + l] = fmaf(-A[i * kp + j], A[l * kp + j], A[i * kp + l]);
                }
            }
            __syncthreads();
        }
        // L y = b, then L^T x = y, by warp 0
        if (tid < 32) {
            for (int i = 0; i < k; ++i) {
                float p = 0.0f;
                for (int q = tid; q < i; q += 32) p = fmaf(A[i * kp + q], bvec[q], p);
                p = warp_sum(p);
                if (tid == 0) bvec[i] = (bvec[i] - p) / A[i * kp + i];
                __syncwarp();
            }
            for (int i = k - 1; i >= 0; --i) {
                float p = 0.0f;
                for (int q = i + 1 + tid; q < k; q += 32) p = fmaf(A[q * kp + i], bvec[q], p);
                p = warp_sum(p);
                if (tid == 0) bvec[i] = (bvec[i] - p) / A[i * kp + i];
                __syncwarp();
            }
            for (int c = tid; c < k; c += 32) x[c] = bvec[c];
        }
    }
}

template <int TPS, int MAXT, int VW>
int launch_als_vw(int64_t nseg, const uint32_t* order, unsigned* queue, const Side& s, const float* Y, float* X, int k,
                  int kp, float lambda, int sm_count, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)kp * kp + kp + 2 * (size_t)kBatch * kp + 2 * kBatch) + sizeof(uint32_t) * 2 * kBatch;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        MF_CUDA(cudaFuncSetAttribute(k_als_half<TPS, MAXT, VW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    int per_sm = 1;
    MF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_als_half<TPS, MAXT, VW>, TPS, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count * per_sm;
    if (grid > nseg) grid = nseg;
    k_als_half<TPS, MAXT, VW><<<(unsigned)grid, TPS, smem, st>>>(nseg, order, queue, s.ptr, s.idx, s.val, Y, X, k, kp, lambda);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int TPS, int MAXT>
int launch_als(int64_t nseg, const uint32_t* order, unsigned* queue, const Side& s, const float* Y, float* X, int k,
               int kp, float lambda, int sm_count, cudaStream_t st) {
    if (k % 4 == 0) return launch_als_vw<TPS, MAXT, 4>(nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    if (k % 2 == 0) return launch_als_vw<TPS, MAXT, 2>(nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    return launch_als_vw<TPS, MAXT, 1>(nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
}

}  // namespace

int als_half_step(const Side& s, const float* Y, float* X, int k, float lambda, int sm_count, cudaStream_t st) {
    if (s.nseg <= 0) return MF_OK;
    const int kp = (k + 3) / 4 * 4;
    const int nb = kp / 4, ntiles = nb * (nb + 1) / 2;
    // scratch: bin counters, cursors, queue head, longest-first order (rebuilt per call: a few microseconds)
    unsigned* scratch = nullptr;
    uint32_t* order = nullptr;
    MF_TRY(dev_alloc(&scratch, 33 + 33 + 1));
    MF_TRY(dev_alloc(&order, (size_t)s.nseg));
    MF_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned) * 67, st));
    const unsigned g = (unsigned)((s.nseg + 255) / 256);
    k_bin_count<<<g, 256, 0, st>>>(s.nseg, s.ptr, scratch);
    k_bin_scan<<<1, 32, 0, st>>>(scratch, scratch + 33);
    k_order_by_bin<<<g, 256, 0, st>>>(s.nseg, s.ptr, scratch + 33, order);
    MF_CUDA(cudaGetLastError());
    unsigned* queue = scratch + 66;
    int rc;
    if (ntiles <= 32)       rc = launch_als<32, 1>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else if (ntiles <= 64)  rc = launch_als<64, 1>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else if (ntiles <= 128) rc = launch_als<128, 1>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else if (ntiles <= 256) rc = launch_als<256, 1>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else if (ntiles <= 512) rc = launch_als<256, 2>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else if (ntiles <= 768) rc = launch_als<256, 3>(s.nseg, order, queue, s, Y, X, k, kp, lambda, sm_count, st);
    else { set_error("ALS: k=%d not supported yet (k <= 152)", k); rc = MF_ERR_UNSUPPORTED; }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    cudaFree(order);
    if (rc == MF_OK && e != cudaSuccess) { set_error("ALS half-step failed: %s", cudaGetErrorString(e)); rc = MF_ERR_CUDA; }
    return rc;
}

}  // namespace mf
