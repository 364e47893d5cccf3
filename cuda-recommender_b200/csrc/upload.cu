// upload.cu — host -> device copies of the rating arrays for the one-shot trainers.
//
// The reference's loader hands kernel_wrapper_ccdpp_NV ordinary `new[]` arrays (src/pmf_util.h:53-64) and copies them
// with six blocking cudaMemcpy calls (cuda_src/CCD_CUDA.cu:292-316).  From pageable memory such a copy is staged by
// the driver through one internal bounce buffer by one thread: a fraction of what the link carries.  Here a pageable
// source is cut into chunks that several host threads copy into their own page-locked bounce buffers and send with
// cudaMemcpyAsync on their own streams — the host memcpy of one chunk overlaps the DMA of the others; page-locked or
// device sources go straight to cudaMemcpyAsync on the caller's stream.  The bounce buffers are kept for the next call
// of the process (a few MB per thread) and go away with mf_release_cached_memory().
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mf {
namespace {

constexpr size_t kChunk = (size_t)4 << 20;  // bytes per bounce buffer
constexpr int kBuffersPerLane = 2;
constexpr int kMaxLanes = 8;

struct Lane {
    int device = -1;
    cudaStream_t st = nullptr;
    void* buf[kBuffersPerLane] = {nullptr, nullptr};
    cudaEvent_t done[kBuffersPerLane] = {nullptr, nullptr};
    cudaError_t err = cudaSuccess;
};
std::vector<Lane> g_lanes;

int lanes_for(int device) {
    unsigned hw = std::thread::hardware_concurrency();
    int want = (int)std::min<unsigned>(kMaxLanes, std::max<unsigned>(2u, hw / 2u));
    if (const char* e = getenv("MF_UPLOAD_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxLanes) want = v;
    }
    if (!g_lanes.empty() && (g_lanes[0].device != device || (int)g_lanes.size() != want)) upload_release_cached();
    if (g_lanes.empty()) {
        g_lanes.resize((size_t)want);
        for (Lane& l : g_lanes) {
            l.device = device;
            MF_CUDA(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
            for (int b = 0; b < kBuffersPerLane; ++b) {
                MF_CUDA(cudaHostAlloc(&l.buf[b], kChunk, cudaHostAllocDefault));
                MF_CUDA(cudaEventCreateWithFlags(&l.done[b], cudaEventDisableTiming));
            }
        }
    }
    return MF_OK;
}

}  // namespace

void upload_release_cached() {
    for (Lane& l : g_lanes) {
        if (l.st) cudaStreamSynchronize(l.st);
        for (int b = 0; b < kBuffersPerLane; ++b) {
            if (l.buf[b]) cudaFreeHost(l.buf[b]);
            if (l.done[b]) cudaEventDestroy(l.done[b]);
        }
        if (l.st) cudaStreamDestroy(l.st);
    }
    g_lanes.clear();
}

bool host_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// dst (device) <- src (host pageable / host page-locked / device), ordered on `st`: work enqueued on `st` after this call
// sees the data.  For a pageable source the call returns once the source has been read completely.
int upload_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return MF_OK;
    if (bytes < 4 * kChunk || !host_pageable(src)) {
        MF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
        return MF_OK;
    }
    int device = 0;
    MF_CUDA(cudaGetDevice(&device));
    MF_TRY(lanes_for(device));
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    const int nl = (int)g_lanes.size();
    // the destination may still be in use by earlier work on `st` (it is not: fresh allocations), but keep the order
    cudaEvent_t ready = nullptr;
    MF_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    MF_CUDA(cudaEventRecord(ready, st));
    std::vector<std::thread> workers;
    for (int li = 0; li < nl; ++li) {
        workers.emplace_back([=]() {
            Lane& l = g_lanes[(size_t)li];
            l.err = cudaSetDevice(device);
            if (l.err == cudaSuccess) l.err = cudaStreamWaitEvent(l.st, ready, 0);
            int b = 0;
            for (size_t c = (size_t)li; c < nchunks && l.err == cudaSuccess; c += (size_t)nl, b ^= 1) {
                const size_t off = c * kChunk, n = std::min(kChunk, bytes - off);
                l.err = cudaEventSynchronize(l.done[b]);  // the buffer's previous DMA has finished (no-op the first time)
                if (l.err != cudaSuccess) break;
                memcpy(l.buf[b], (const char*)src + off, n);
                l.err = cudaMemcpyAsync((char*)dst + off, l.buf[b], n, cudaMemcpyHostToDevice, l.st);
                if (l.err == cudaSuccess) l.err = cudaEventRecord(l.done[b], l.st);
            }
        });
    }
    for (std::thread& t : workers) t.join();
    cudaEventDestroy(ready);
    for (Lane& l : g_lanes) {
        if (l.err != cudaSuccess) {
            set_error("staged upload failed: %s", cudaGetErrorString(l.err));
            return MF_ERR_CUDA;
        }
        // `st` continues after every lane's last copy
        for (int b = 0; b < kBuffersPerLane; ++b) MF_CUDA(cudaStreamWaitEvent(st, l.done[b], 0));
    }
    return MF_OK;
}

}  // namespace mf
