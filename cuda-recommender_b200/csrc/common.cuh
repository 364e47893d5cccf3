// common.cuh — error plumbing and small device helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mf_abi.h"

namespace mf {

// MF_TRACE=1 in the environment: host-side phase timings on stderr (time since the previous mark)
void trace_mark(const char* what);

// thread-local last-error text behind mf_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

// The reference's convention is print-and-continue (cuda_src/CUDA_AUX.h:11-18); here every CUDA
// call is checked and the C-ABI returns MF_ERR_CUDA with the message.
#define MF_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::mf::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,       \
                            __LINE__, cudaGetErrorString(_e));                                 \
            return MF_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define MF_TRY(expr)                   \
    do {                               \
        int _rc = (expr);              \
        if (_rc != MF_OK) return _rc;  \
    } while (0)

#define MF_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            ::mf::set_error(__VA_ARGS__);     \
            return MF_ERR_ARG;                \
        }                                     \
    } while (0)

// Device arena: a session allocates its long-lived arrays (ratings, panel layout, test set) by bumping a pointer
// inside a few large cudaMalloc chunks instead of ~40 cudaMalloc / cudaFree calls — the frees alone cost 31 ms per
// mf_ccdpp_train call on the Netflix shape (cudaFree synchronises and unmaps).  While an arena is bound to the calling
// thread, dev_alloc() takes from it; dev_free() is a no-op for arena memory (it goes away with the arena) and
// cudaFree for everything else.  Arrays that are exported through CUDA IPC are allocated unbound.
struct DeviceArena;
DeviceArena* arena_create(size_t first_chunk_bytes, cudaStream_t st);  // st: chunks from the stream-ordered pool
void arena_destroy(DeviceArena* a);
void arena_bind(DeviceArena* a);  // nullptr: unbind
int dev_alloc_bytes(void** p, size_t bytes);
void dev_free(void* p);
struct ArenaScope {  // binds for a scope
    explicit ArenaScope(DeviceArena* a) { arena_bind(a); }
    ~ArenaScope() { arena_bind(nullptr); }
};

template <typename T>
static inline int dev_alloc(T** p, size_t n) {
    *p = nullptr;
    return dev_alloc_bytes((void**)p, (n > 0 ? n : 1) * sizeof(T));
}

// stream-ordered scratch (cudaMallocAsync pool): for temporaries that live on one stream only
template <typename T>
static inline int tmp_alloc(T** p, size_t n, cudaStream_t st) {
    *p = nullptr;
    MF_CUDA(cudaMallocAsync((void**)p, (n > 0 ? n : 1) * sizeof(T), st));
    return MF_OK;
}
static inline void tmp_free(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

// upload.cu: host -> device copy ordered on `st`; a pageable host source is staged through page-locked bounce buffers by
// several host threads (returns once the source has been read), anything else is one cudaMemcpyAsync
int upload_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st);
bool host_pageable(const void* p);
void upload_release_cached();

static inline uint32_t ceil_div_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// exclusive scan of n uint32 values into out[0..n] (out[n] = total); in and out may alias only if
// out == in is NOT used.  tmp: device scratch of at least scan_tmp_elems(n) uint32.
size_t scan_tmp_elems(size_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace mf
