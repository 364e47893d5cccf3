// ingest.cu — integer tier: COO -> CSR (sorted by row, col) + CSC (sorted by col, row) on the GPU.
//
// The reference has no builder: it loads ready-made CSR *and* CSC files (src/tools.cpp:36-59,
// src/pmf_util.h:108-136) and "transposes" by swapping pointers (pmf_util.h:66-81).  This is the
// ingest the new build adds; its oracle is the stable counting sort in
// oracle/mf_oracle.c (orc_coo_to_csr_csc) and the result must match it bit for bit.
//
// Because (row, col) pairs are unique, the sorted order is unique, so no stable sort is needed:
//   1. histogram of the major key (atomics) -> exclusive scan -> ptr
//   2. scatter every entry into its segment's range in arrival order (atomic cursor)
//   3. per segment, one CTA marks the minor keys in a shared-memory bitmap, prefix-sums the word
//      popcounts, and moves every entry to ptr[seg] + rank(minor key); a minor dimension too wide for one
//      bitmap (> ~928 K keys in 227 KB) is handled in several passes over windows of the key range, each
//      window's entries placed behind those of the windows before it
// Step 3 makes the output independent of the arrival order of step 2.  A duplicate (row, col) pair shows up as a
// segment whose distinct keys are fewer than its entries: the build fails with MF_ERR_ARG instead of leaving holes.
#include <algorithm>

#include "common.cuh"
#include "layout.cuh"

namespace mf {
namespace {

__global__ void k_hist(int64_t nnz, const uint32_t* __restrict__ key, uint32_t* __restrict__ count) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&count[key[e]], 1u);
}

__global__ void k_scatter(int64_t nnz, const uint32_t* __restrict__ key, const uint32_t* __restrict__ other,
                          const float* __restrict__ val, const uint32_t* __restrict__ ptr, uint32_t* __restrict__ cursor,
                          uint32_t* __restrict__ tmp_idx, float* __restrict__ tmp_val) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = key[e];
        const uint32_t d = ptr[s] + atomicAdd(&cursor[s], 1u);
        tmp_idx[d] = other[e];
        tmp_val[d] = val[e];
    }
}

// one CTA per segment (grid-stride); bitmap of `words` 32-bit words in dynamic shared memory
__global__ void __launch_bounds__(256) k_rank_place(int64_t nseg, uint32_t words, uint32_t nminor, const uint32_t* __restrict__ ptr,
                                                    const uint32_t* __restrict__ tmp_idx, const float* __restrict__ tmp_val,
                                                    uint32_t* __restrict__ out_idx, float* __restrict__ out_val, int* __restrict__ dup) {
    extern __shared__ uint32_t sm[];
    uint32_t* bits = sm;            // [words]
    uint32_t* pre = sm + words;     // [words] exclusive prefix of popcounts
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int64_t s = blockIdx.x; s < nseg; s += gridDim.x) {
        const uint32_t lo = ptr[s], hi = ptr[s + 1];
        if (hi == lo) continue;
        if (hi - lo == 1) {
            if (tid == 0) { out_idx[lo] = tmp_idx[lo]; out_val[lo] = tmp_val[lo]; }
            continue;
        }
        uint32_t placed = 0;  // entries of the windows already done (CTA-uniform)
        for (uint64_t wb = 0; wb < nminor; wb += (uint64_t)words * 32u) {  // one window of the key range per pass
        const uint32_t wbeg = (uint32_t)wb;  // (keys are compared as m - wbeg, unsigned: below the window wraps to a huge value)
        __syncthreads();
        for (uint32_t w = tid; w < words; w += 256) bits[w] = 0u;
        if (tid == 0) carry_s = 0u;
        __syncthreads();
        for (uint32_t e = lo + tid; e < hi; e += 256) {
            const uint32_t m = tmp_idx[e] - wbeg;
            if (m < words * 32u) atomicOr(&bits[m >> 5], 1u << (m & 31));
        }
        __syncthreads();
        // exclusive scan of popcounts, 256 words per round
        for (uint32_t base = 0; base < words; base += 256) {
            const uint32_t w = base + tid;
            const uint32_t c = w < words ? __popc(bits[w]) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) warp_tot[wid] = inc;
            __syncthreads();
            uint32_t wbase = 0;
            for (int i = 0; i < wid; ++i) wbase += warp_tot[i];
            const uint32_t carry = carry_s;
            if (w < words) pre[w] = carry + wbase + inc - c;
            __syncthreads();
            if (tid == 255) carry_s = carry + wbase + inc;
            __syncthreads();
        }
        for (uint32_t e = lo + tid; e < hi; e += 256) {
            const uint32_t m = tmp_idx[e] - wbeg;
            if (m < words * 32u) {
                const uint32_t rank = placed + pre[m >> 5] + __popc(bits[m >> 5] & ((1u << (m & 31)) - 1u));
                out_idx[lo + rank] = m + wbeg;
                out_val[lo + rank] = tmp_val[e];
            }
        }
        __syncthreads();
        placed += carry_s;  // distinct keys of this window
        }
        if (tid == 0 && placed != hi - lo) *dup = 1;  // fewer distinct keys than entries: a (row, col) pair occurs twice
    }
}

int build_one(int64_t nmajor, int64_t nminor, int64_t nnz, const uint32_t* d_key, const uint32_t* d_other,
              const float* d_val, uint32_t* d_ptr, uint32_t* d_idx, float* d_outval, uint32_t* d_tmp_idx, float* d_tmp_val,
              uint32_t* d_count, uint32_t* d_scan_tmp, int sm_count) {
    MF_CUDA(cudaMemset(d_count, 0, sizeof(uint32_t) * (size_t)nmajor));
    if (nnz > 0) k_hist<<<sm_count * 8, 256>>>(nnz, d_key, d_count);
    MF_CUDA(cudaGetLastError());
    MF_TRY(exclusive_scan_u32(d_count, d_ptr, (size_t)nmajor, d_scan_tmp, 0));
    MF_CUDA(cudaMemset(d_count, 0, sizeof(uint32_t) * (size_t)nmajor));
    if (nnz > 0) {
        k_scatter<<<sm_count * 8, 256>>>(nnz, d_key, d_other, d_val, d_ptr, d_count, d_tmp_idx, d_tmp_val);
        const uint32_t max_words = (uint32_t)((227 * 1024 - 256) / (2 * sizeof(uint32_t)));  // bitmap + prefix in shared memory
        const uint32_t words = std::min<uint32_t>((uint32_t)((nminor + 31) / 32), max_words);  // one window of the key range
        const size_t smem = sizeof(uint32_t) * 2 * (size_t)words;
        static size_t attr[64] = {0};  // per device
        int dev = 0;
        MF_CUDA(cudaGetDevice(&dev));
        if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > attr[dev])) {
            MF_CUDA(cudaFuncSetAttribute(k_rank_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr[dev] = smem;
        }
        int* d_dup = nullptr;
        MF_TRY(dev_alloc(&d_dup, 1));
        MF_CUDA(cudaMemset(d_dup, 0, sizeof(int)));
        int64_t grid = nmajor < (int64_t)sm_count * 8 ? nmajor : (int64_t)sm_count * 8;
        k_rank_place<<<(unsigned)grid, 256, smem>>>(nmajor, words, (uint32_t)nminor, d_ptr, d_tmp_idx, d_tmp_val, d_idx, d_outval, d_dup);
        int h_dup = 0;
        const cudaError_t e = cudaMemcpy(&h_dup, d_dup, sizeof(int), cudaMemcpyDeviceToHost);
        dev_free(d_dup);
        if (e != cudaSuccess) { set_error("mf_build_csr_csc: %s", cudaGetErrorString(e)); return MF_ERR_CUDA; }
        if (h_dup) { set_error("mf_build_csr_csc: a (row, col) pair occurs more than once"); return MF_ERR_ARG; }
    }
    MF_CUDA(cudaGetLastError());
    MF_CUDA(cudaDeviceSynchronize());
    return MF_OK;
}

}  // namespace
}  // namespace mf

extern "C" int mf_build_csr_csc(int64_t rows, int64_t cols, int64_t nnz, const uint32_t* coo_row, const uint32_t* coo_col,
                                const float* coo_val, uint32_t* csr_row_ptr, uint32_t* csr_col_idx, float* csr_val,
                                uint32_t* csc_col_ptr, uint32_t* csc_row_idx, float* csc_val, int device) {
    using namespace mf;
    MF_REQUIRE(rows > 0 && cols > 0 && nnz >= 0 && rows < ((int64_t)1 << 32) && cols < ((int64_t)1 << 32) && nnz < ((int64_t)1 << 32),
               "mf_build_csr_csc: bad shape");
    MF_REQUIRE(csr_row_ptr && csc_col_ptr && (nnz == 0 || (coo_row && coo_col && coo_val && csr_col_idx && csr_val && csc_row_idx && csc_val)),
               "mf_build_csr_csc: NULL argument");
    MF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MF_CUDA(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    const int64_t nmax = rows > cols ? rows : cols;
    uint32_t *d_r = nullptr, *d_c = nullptr, *d_ptr = nullptr, *d_idx = nullptr, *d_tmp_idx = nullptr, *d_count = nullptr, *d_scan = nullptr;
    float *d_v = nullptr, *d_out = nullptr, *d_tmp_val = nullptr;
    int rc = MF_OK;
    auto A = [&](auto** p, size_t n) { if (rc == MF_OK) rc = dev_alloc(p, n); };
    A(&d_r, (size_t)nnz); A(&d_c, (size_t)nnz); A(&d_v, (size_t)nnz);
    A(&d_ptr, (size_t)nmax + 1); A(&d_idx, (size_t)nnz); A(&d_out, (size_t)nnz);
    A(&d_tmp_idx, (size_t)nnz); A(&d_tmp_val, (size_t)nnz); A(&d_count, (size_t)nmax); A(&d_scan, scan_tmp_elems((size_t)nmax));
    auto copy = [&](void* dst, const void* src, size_t bytes) {
        if (rc == MF_OK && bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyDefault) != cudaSuccess) {
            set_error("mf_build_csr_csc: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MF_ERR_CUDA;
        }
    };
    copy(d_r, coo_row, sizeof(uint32_t) * (size_t)nnz);
    copy(d_c, coo_col, sizeof(uint32_t) * (size_t)nnz);
    copy(d_v, coo_val, sizeof(float) * (size_t)nnz);
    if (rc == MF_OK) {  // the counting kernels index with these: they must be indices
        bool ok_r = true, ok_c = true;
        rc = check_below(d_r, nnz, (uint64_t)rows, &ok_r, nullptr);
        if (rc == MF_OK) rc = check_below(d_c, nnz, (uint64_t)cols, &ok_c, nullptr);
        if (rc == MF_OK && (!ok_r || !ok_c)) { set_error("mf_build_csr_csc: a %s index is outside the %lld x %lld matrix", ok_r ? "column" : "row", (long long)rows, (long long)cols); rc = MF_ERR_ARG; }
    }
    if (rc == MF_OK) rc = build_one(rows, cols, nnz, d_r, d_c, d_v, d_ptr, d_idx, d_out, d_tmp_idx, d_tmp_val, d_count, d_scan, sms);
    copy(csr_row_ptr, d_ptr, sizeof(uint32_t) * ((size_t)rows + 1));
    copy(csr_col_idx, d_idx, sizeof(uint32_t) * (size_t)nnz);
    copy(csr_val, d_out, sizeof(float) * (size_t)nnz);
    if (rc == MF_OK) rc = build_one(cols, rows, nnz, d_c, d_r, d_v, d_ptr, d_idx, d_out, d_tmp_idx, d_tmp_val, d_count, d_scan, sms);
    copy(csc_col_ptr, d_ptr, sizeof(uint32_t) * ((size_t)cols + 1));
    copy(csc_row_idx, d_idx, sizeof(uint32_t) * (size_t)nnz);
    copy(csc_val, d_out, sizeof(float) * (size_t)nnz);
    void* ptrs[] = {d_r, d_c, d_v, d_ptr, d_idx, d_out, d_tmp_idx, d_tmp_val, d_count, d_scan};
    for (void* p : ptrs) if (p) dev_free(p);
    return rc;
}
