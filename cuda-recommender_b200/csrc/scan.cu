// scan.cu — device-wide exclusive prefix sum over uint32 (integer tier plumbing for the
// layout builders: piece offsets, work-item offsets, slot offsets, cost prefix).
// Three launches: per-tile reduce, single-CTA scan of the tile sums, per-tile scan + offset.
#include <stdarg.h>
#include <stdlib.h>

#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mf {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

void trace_mark(const char* what) {
    // (multi-process launches: only the process whose RANK is MF_TRACE's value — "1" also means rank 0 — reports)
    static const bool on = getenv("MF_TRACE") != nullptr &&
                           (getenv("RANK") == nullptr || atoi(getenv("RANK")) == (atoi(getenv("MF_TRACE")) == 1 ? 0 : atoi(getenv("MF_TRACE"))));
    static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mf trace] %-34s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}


// ---- device arena (common.cuh) -------------------------------------------------------------------------------
struct DeviceArena {
    struct Chunk { char* base; size_t size, used; bool pooled; };
    std::vector<Chunk> chunks;
    size_t next_chunk = (size_t)256 << 20;
    cudaStream_t st = nullptr;  // chunks come from the stream-ordered pool on this stream (nullptr: plain cudaMalloc)
};
namespace {
std::mutex g_arena_mu;
std::vector<DeviceArena*> g_arenas;  // live arenas (for dev_free's membership test)
thread_local DeviceArena* t_arena = nullptr;
constexpr size_t kArenaAlign = 256;

// Chunks are taken from the device's stream-ordered pool (cudaMallocAsync; the sessions keep its release threshold at
// "never"): freeing one is a pointer hand-back instead of a cudaFree, which was measured to stall for 0.5-0.9 s now and
// then on a 3.7 GB block, and the next session of the process gets the same block back without a cudaMalloc (~22 ms).
// mf_release_cached_memory() trims the pool.
bool arena_add_chunk(DeviceArena* a, size_t bytes) {
    void* p = nullptr;
    bool pooled = false;
    if (a->st != nullptr && cudaMallocAsync(&p, bytes, a->st) == cudaSuccess && cudaStreamSynchronize(a->st) == cudaSuccess) {
        pooled = true;
    } else {
        cudaGetLastError();
        p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    std::lock_guard<std::mutex> lk(g_arena_mu);
    a->chunks.push_back({(char*)p, bytes, 0, pooled});
    return true;
}
}  // namespace

DeviceArena* arena_create(size_t first_chunk_bytes, cudaStream_t st) {
    DeviceArena* a = new DeviceArena();
    a->st = st;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        g_arenas.push_back(a);
    }
    if (first_chunk_bytes > 0) arena_add_chunk(a, (first_chunk_bytes + kArenaAlign - 1) / kArenaAlign * kArenaAlign);  // best effort
    return a;
}

void arena_destroy(DeviceArena* a) {
    if (!a) return;
    if (t_arena == a) t_arena = nullptr;
    std::vector<DeviceArena::Chunk> chunks;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        chunks.swap(a->chunks);
        for (size_t i = 0; i < g_arenas.size(); ++i)
            if (g_arenas[i] == a) { g_arenas.erase(g_arenas.begin() + i); break; }
    }
    for (auto& c : chunks) {
        if (c.pooled) cudaFreeAsync(c.base, a->st);  // stream-ordered: after everything the session enqueued
        else cudaFree(c.base);
    }
    delete a;
}

void arena_bind(DeviceArena* a) { t_arena = a; }

int dev_alloc_bytes(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 1;
    DeviceArena* a = t_arena;
    if (a) {
        const size_t need = (bytes + kArenaAlign - 1) / kArenaAlign * kArenaAlign;
        for (int attempt = 0; attempt < 2; ++attempt) {
            {
                std::lock_guard<std::mutex> lk(g_arena_mu);
                for (auto& c : a->chunks)
                    if (c.size - c.used >= need) {
                        *p = c.base + c.used;
                        c.used += need;
                        return MF_OK;
                    }
            }
            if (attempt == 0 && !arena_add_chunk(a, need > a->next_chunk ? need : a->next_chunk)) break;
        }
        // the arena could not grow: fall through to a plain allocation
    }
    MF_CUDA(cudaMalloc(p, bytes));
    return MF_OK;
}

void dev_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        for (DeviceArena* a : g_arenas)
            for (auto& c : a->chunks)
                if ((char*)p >= c.base && (char*)p < c.base + c.size) return;  // released with the arena
    }
    cudaFree(p);
}

namespace {
constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 4;
constexpr int kScanTile = kScanThreads * kScanPerThread;

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        warp_tot[lane] = winc - w;  // exclusive prefix of warp totals
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    uint32_t res = warp_tot[wid] + inc - v;
    __syncthreads();  // warp_tot / *total may be reused by the caller's next round
    return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t* __restrict__ in, size_t n,
                                                             uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t tot;
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPerThread;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i)
        if (base + i < n) s += in[base + i];
    (void)block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

// single CTA: tile_sum[0..nt) -> exclusive prefix in place; grand total -> *total_out
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(uint32_t* __restrict__ tile_sum, size_t nt,
                                                            uint32_t* __restrict__ total_out) {
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    for (size_t base = 0; base < nt; base += kScanThreads) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nt ? tile_sum[i] : 0u;
        uint32_t ex = block_exclusive_scan(v, &tot);
        if (i < nt) tile_sum[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* __restrict__ in, size_t n,
                                                            const uint32_t* __restrict__ tile_off,
                                                            uint32_t* __restrict__ out) {
    __shared__ uint32_t tot;
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        v[i] = base + i < n ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t ex = block_exclusive_scan(s, &tot) + tile_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}
}  // namespace

size_t scan_tmp_elems(size_t n) { return (n + kScanTile - 1) / kScanTile + 1; }

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t st) {
    if (n == 0) {
        MF_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), st));
        return MF_OK;
    }
    size_t nt = (n + kScanTile - 1) / kScanTile;
    k_scan_reduce<<<(unsigned)nt, kScanThreads, 0, st>>>(in, n, tmp);
    k_scan_tiles<<<1, kScanThreads, 0, st>>>(tmp, nt, out + n);
    k_scan_apply<<<(unsigned)nt, kScanThreads, 0, st>>>(in, n, tmp, out);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

}  // namespace mf
