// ccd_kernels.cu — CCD++ sweeps over one compressed-sparse copy of the ratings (sm_100a).
//
// One kernel template covers the three things the reference does to a copy, alone or fused:
//   SUB    residual -= g_old[idx] * s_old[seg]      UpdateRating(add=false)  src/CCD.cpp:18-43, :133-134
//   ADD    residual += g_add[idx] * s_add[seg]      UpdateRating(add=true)   src/CCD.cpp:100-103
//   SOLVE  (g, h) = sum g_new[idx]*residual, sum g_new[idx]^2  -> out = g / (lambda*deg + h)
//                                                   RankOneUpdate            src/CCD.cpp:6-16, :110-121
// (GPU counterparts being replaced: cuda_src/CCD_CUDA.cu:3-104.)
// Residual products and sums are rounded separately (__fmul_rn / __fadd_rn), exactly like the CPU
// path, so both copies stay bit-identical to the oracle's residual; g and h use FMA and a fixed
// reduction tree (lane-serial over 8 entries, xor-butterfly over lanes, slots in order).
//
// PANEL kernel (layout.cuh): persistent CTAs, one per SM.  A CTA owns an equal-cost contiguous range
// of work items; for every panel its range touches it stages that panel of the gathered factor
// vector(s) in shared memory, then its warps pull batches of four items from a shared-memory counter
// (the next batch's descriptors are fetched one batch ahead).  Each 8-lane group of the warp streams
// one item in steps of 32 entries — a lane owns 4 consecutive entries of a step (8-byte index vector,
// 16-byte value vector), a group's load covers one contiguous 64/128-byte span — with four steps in
// flight per group (register ring).  Inside a panel the items are length-ranked and dealt into lanes
// (degree-binned order, prep.cu), so the four items of a batch have nearly equal length (the groups
// stay in step) and every CTA range holds the same mix of lengths (the CTAs finish together).
// Values are written back with 16-byte stores; the factor gathers never leave shared memory.
// Reduction tree of an item: lane-serial over its steps, xor-butterfly over the 8 lanes.
#include "ccd_kernels.cuh"

namespace mf {
namespace {

constexpr unsigned kFull = 0xffffffffu;

// geometry of the register-ring sweep (A/B knobs: scripts/build_variant.sh): threads per CTA (one CTA per SM) and steps
// in flight per 8-lane group
#ifndef MF_SWEEP_THREADS
#define MF_SWEEP_THREADS 1024
#endif
#ifndef MF_SWEEP_RING
#define MF_SWEEP_RING 4
#endif
constexpr int kSweepThreads = MF_SWEEP_THREADS;
constexpr uint32_t kRing = MF_SWEEP_RING;

// One lane's share of a 32-entry step of its 8-lane group: 4 consecutive entries (8 bytes of indices,
// 16 bytes of values), so that a group's load covers one contiguous 64- or 128-byte span.
struct Step {
    uint2 i;   // 4 x uint16 panel-local indices
    float4 v;  // 4 values
};
__device__ __forceinline__ Step load_step(const uint16_t* __restrict__ idx16, const float* __restrict__ val, uint32_t pos) {
    Step e;
#if defined(MF_EXP_NOLOAD)
    // timing experiment only (results are meaningless): no global loads, pseudo-random panel indices
    const uint32_t r0 = (pos * 2654435761u) >> 17, r1 = pos * 40503u + 12345u;
    e.i = make_uint2((r0 & 0x7ffcu) | ((r1 & 0x7ffcu) << 16), ((r0 >> 3) & 0x7ffcu) | (((r1 * 7u) & 0x7ffcu) << 16));
    e.v = make_float4(1.f, 2.f, 3.f, 4.f);
    return e;
#endif
#if defined(MF_STREAM_NO_ALLOCATE)
    // streaming loads that do not allocate in L1 (the shared-memory panels leave little L1 behind)
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(e.i.x), "=r"(e.i.y) : "l"(idx16 + pos));
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(e.v.x), "=f"(e.v.y), "=f"(e.v.z), "=f"(e.v.w) : "l"(val + pos));
#else
    e.i = __ldcs(reinterpret_cast<const uint2*>(idx16 + pos));
    e.v = __ldcs(reinterpret_cast<const float4*>(val + pos));
#endif
    return e;
}

// idx16 holds the panel-local index already multiplied by 4 (the byte offset into the shared-memory
// panel, layout.cuh), so each gather address costs one instruction: mask or shift, then LDS.
__device__ __forceinline__ float panel_at(const float* __restrict__ sm, uint32_t byte_off) {
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(sm) + byte_off);
}

template <int MODE>
__device__ __forceinline__ void calc4(Step& e, const float* __restrict__ sm_new, const float* __restrict__ sm_add,
                                      const float* __restrict__ sm_old, float s_add, float s_old, float& g, float& h) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve;
    float v[4] = {e.v.x, e.v.y, e.v.z, e.v.w};
#if defined(MF_EXP_NOGATHER)
    // timing experiment only: every lane reads the same shared-memory word (broadcast, no bank conflicts)
    const uint32_t ix[4] = {e.i.x & 4u, (e.i.x >> 16) & 4u, e.i.y & 4u, (e.i.y >> 16) & 4u};
#else
    const uint32_t ix[4] = {e.i.x & 0xffffu, e.i.x >> 16, e.i.y & 0xffffu, e.i.y >> 16};
#endif
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float x = v[j];
        if (SUB) x = __fsub_rn(x, __fmul_rn(panel_at(sm_old, ix[j]), s_old));
        if (ADD) x = __fadd_rn(x, __fmul_rn(panel_at(sm_add, ix[j]), s_add));
        if (SOLVE) {
            const float un = panel_at(sm_new, ix[j]);
            g = fmaf(un, x, g);
            h = fmaf(un, un, h);
        }
        v[j] = x;
    }
    if (SUB || ADD) e.v = make_float4(v[0], v[1], v[2], v[3]);
}

// The panel holding item `ib` = the number of panels that end at or before it (panel_item_ptr is non-decreasing).
// Counted by the whole CTA at once: a serial walk is one dependent L2 round trip per panel, i.e. the CTAs of the
// last of 30 panels started ~12 us late.
__device__ __forceinline__ int first_panel(const uint32_t* __restrict__ panel_item_ptr, int npanels, uint32_t ib) {
    int p = 0;
    for (int base = 0; base < npanels; base += (int)blockDim.x) {
        const int j = base + (int)threadIdx.x;
        p += __syncthreads_count(j < npanels && panel_item_ptr[j + 1] <= ib);
    }
    return p;
}

// One panel of a factor vector -> shared memory, zeros beyond `cnt` (the panel's last valid entry) up to `stride`.
// 16-byte loads, four in flight per thread: a 64 KB panel costs a 1024-thread CTA one L2 round trip instead of 16
// dependent ones (a scalar loop here was ~6 us per panel and vector: the fixed cost that kept the sweeps from scaling
// across GPUs).  g + base is 32-byte aligned (panels are multiples of 8 entries, factor rows 128-byte aligned) and the
// factor rows are padded to 32 entries, so the vector load that straddles `cnt` stays inside the allocation.
__device__ __forceinline__ void stage_panel(float* __restrict__ sm, const float* __restrict__ g, int64_t base, uint32_t cnt,
                                            uint32_t stride, const uint32_t tid, const uint32_t nthr) {
    const float4* __restrict__ src = reinterpret_cast<const float4*>(g + base);
    float4* __restrict__ dst = reinterpret_cast<float4*>(sm);
    const uint32_t n4 = stride >> 2;
    for (uint32_t i = tid; i < n4; i += 4u * nthr) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t j = i + (uint32_t)u * nthr;
            v[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (j < n4 && 4u * j < cnt) v[u] = __ldcg(src + j);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t j = i + (uint32_t)u * nthr;
            if (j < n4) {
                if (4u * j + 3u >= cnt) {  // the vector that straddles the end of the panel
                    if (4u * j + 1u >= cnt) v[u].y = 0.0f;
                    if (4u * j + 2u >= cnt) v[u].z = 0.0f;
                    v[u].w = 0.0f;
                    if (4u * j >= cnt) v[u].x = 0.0f;
                }
                dst[j] = v[u];
            }
        }
    }
}

// The sweep body reads its arguments through accessors, so that the per-launch kernels (arguments = kernel parameters)
// and the persistent kernel (per-side constants = __constant__ memory, per-phase vectors = a shared-memory record) share it.
struct LaunchView {
    const PanelSweepArgs& a;
    __device__ __forceinline__ const uint16_t* idx16() const { return a.idx16; }
    __device__ __forceinline__ float* val() const { return a.val; }
    __device__ __forceinline__ const WorkItem* items() const { return a.items; }
    __device__ __forceinline__ const uint32_t* panel_item_ptr() const { return a.panel_item_ptr; }
    __device__ __forceinline__ int npanels() const { return a.npanels; }
    __device__ __forceinline__ uint32_t panel_rows() const { return a.panel_rows; }
    __device__ __forceinline__ int64_t gdim() const { return a.gdim; }
    __device__ __forceinline__ int64_t seg_offset() const { return a.seg_offset; }
    __device__ __forceinline__ const float* g_new() const { return a.g_new; }
    __device__ __forceinline__ const float* g_add() const { return a.g_add; }
    __device__ __forceinline__ const float* g_old() const { return a.g_old; }
    __device__ __forceinline__ const float* s_add() const { return a.s_add; }
    __device__ __forceinline__ const float* s_old() const { return a.s_old; }
    __device__ __forceinline__ float2* partials() const { return a.partials; }
    __device__ __forceinline__ uint32_t pf_dist() const { return a.pf_dist; }
    __device__ __forceinline__ uint32_t npad() const { return a.npad; }
};

// Streams the (up to four) items of a batch through the 8-lane groups of a warp: a lane owns 4 consecutive
// entries of its group's 32-entry step.  kRing steps are in flight per group (register ring e[0..kRing)): the load
// of step s+kRing is issued right after step s is consumed.  While every group still has a full ring round ahead
// (o + 32*kRing <= minlen) the steps are consumed without per-lane checks; the ragged end runs predicated.  A lane
// consumes its entries in storage order whatever kRing is, so the reduction tree of an item does not depend on it.
template <int MODE, class Args>
__device__ __forceinline__ void stream_batch(const Args& a, uint32_t pos, uint32_t len, uint32_t minlen,
                                             uint32_t maxlen, uint32_t lane_off, const float* __restrict__ sm_new,
                                             const float* __restrict__ sm_add, const float* __restrict__ sm_old, float s_add,
                                             float s_old, float& g, float& h) {
    constexpr bool WRITE = (MODE & kSub) || (MODE & kAdd);
    constexpr uint32_t R = kRing;
    Step e[R];
#pragma unroll
    for (uint32_t r = 0; r < R; ++r) {
        // defined on every path: a register that is only conditionally written is live back to the function entry — and,
        // inside the persistent kernel's phase loop, around the whole loop, for every inlined mode at once
        e[r].i = make_uint2(0u, 0u);
        e[r].v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (32u * r + lane_off < len) e[r] = load_step(a.idx16(), a.val(), pos + 32u * r);
    }
    uint32_t o = 0;
#pragma unroll 1
    for (; o + 32u * R <= minlen; o += 32u * R) {
#pragma unroll
        for (uint32_t r = 0; r < R; ++r) {
            calc4<MODE>(e[r], sm_new, sm_add, sm_old, s_add, s_old, g, h);
            if (WRITE) __stcs(reinterpret_cast<float4*>(a.val() + pos + o + 32u * r), e[r].v);
            if (o + 32u * (R + r) + lane_off < len) e[r] = load_step(a.idx16(), a.val(), pos + o + 32u * (R + r));
        }
    }
#pragma unroll 1
    for (; o < maxlen; o += 32u * R) {
#pragma unroll
        for (uint32_t r = 0; r < R; ++r) {
            if (o + 32u * r + lane_off < len) {
                calc4<MODE>(e[r], sm_new, sm_add, sm_old, s_add, s_old, g, h);
                if (WRITE) __stcs(reinterpret_cast<float4*>(a.val() + pos + o + 32u * r), e[r].v);
            }
            if (o + 32u * (R + r) + lane_off < len) e[r] = load_step(a.idx16(), a.val(), pos + o + 32u * (R + r));
        }
    }
}

// Adds a segment's slots, applies the regulariser, divides: out = g / (lambda*deg + h), with lambda*deg a
// float*unsigned product as at src/CCD.cpp:112,120; empty segment -> 0 (src/CCD.cpp:8).
// LANES = 1: one thread per segment (slots added in order); LANES = 32: one warp per segment, lane l adds
// slots l, l+32, ... in order and the lanes are combined by an xor-butterfly — both are fixed trees.
// Multi-GPU epilogue (fused solve -> exchange over NVLink, CUDA IPC mappings of the
// peers' buffers, dist.cu).  Low-latency protocol, as NCCL's LL: every freshly solved coordinate is stored into
// each peer's receive buffer as ONE 8-byte word {value bits, epoch} — 8-byte stores arrive whole, so the epoch
// half is the "data valid" flag and no fence or separate signal is needed.  The receiving side (k_ll_unpack)
// polls each word of the blocks it does not own until the epoch matches and writes the value into its factor
// vector.  (Measured alternatives, 2 GPUs, per exchange: NCCL grouped broadcast 21 us; 4-byte remote stores +
// __threadfence_system + flag 19 us; flag + peer pull 47 us — a system-scope fence alone costs >10 us here.)
// One receive buffer per factor matrix suffices: a rank can only produce the next generation of a vector after
// it has unpacked everybody's blocks of the other vector, and everybody sends those only after their last sweep
// that read the old generation (sweeps alternate u / v).  flag-based signalling remains for the rare barrier.
struct PushArgs {
    unsigned long long* const* peer_ll;  // [nranks] LL receive buffer (for this factor matrix) of every rank; nullptr = no push
    unsigned* const* peer_flags;         // [nranks] flag words of every rank (barrier only)
    unsigned* ticket;                    // local CTA counter (barrier only)
    int64_t vec_off;                     // index of out[0] inside the factor vector (this shard's first segment)
    int rank, nranks;
    unsigned epoch;
    int barrier;                         // 1: no values, publish the epoch in the peers' flag words
};

// grid-stride over segments, four segments per thread and round with their metadata loads issued together (the
// finalize is a chain of dependent L2 round trips: pointers -> slots -> result); `partials` is read through L2
// (written by other CTAs of the same launch when the finalize runs inside the sweep kernel)
// segment metadata of a thread's first finalize round (layout data: it can be fetched before the grid barrier, which
// takes one dependent L2 round trip out of every solving sweep)
struct FinalizeMeta {
    uint32_t deg[4], lo[4], hi[4];
};
template <int LANES>
__device__ __forceinline__ void finalize_load_meta(int64_t nseg, const uint32_t* __restrict__ slot_ptr, const uint32_t* __restrict__ seg_ptr,
                                                   int64_t t0, FinalizeMeta& m) {
    const int64_t total = nseg * LANES, stride = (int64_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t tid = t0 + u * stride;
        m.deg[u] = 0u; m.lo[u] = 0u; m.hi[u] = 0u;
        if (tid < total) {
            const int64_t s = tid / LANES;
            m.deg[u] = seg_ptr[s + 1] - seg_ptr[s];
            m.lo[u] = slot_ptr[s];
            m.hi[u] = slot_ptr[s + 1];
        }
    }
}

template <int LANES>
__device__ __forceinline__ void finalize_segments(int64_t nseg, const uint32_t* __restrict__ slot_ptr, const float2* partials,
                                                  const uint32_t* __restrict__ seg_ptr, float lambda, int nmf,
                                                  float* __restrict__ out, unsigned long long* const* peer_ll, int64_t vec_off,
                                                  int rank, int nranks, unsigned epoch, const FinalizeMeta* first = nullptr) {
    constexpr int U = 4;
    const int64_t total = nseg * LANES, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tfirst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t0 = tfirst; t0 < total; t0 += U * stride) {
        FinalizeMeta m;
        if (first != nullptr && t0 == tfirst) m = *first;
        else finalize_load_meta<LANES>(nseg, slot_ptr, seg_ptr, t0, m);
        uint32_t (&deg)[U] = m.deg, (&lo)[U] = m.lo, (&hi)[U] = m.hi;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tid = t0 + u * stride;
            if (tid >= total) break;  // warp-uniform for LANES = 32 (a warp shares its segment)
            const int64_t s = tid / LANES;
            const int l = (int)(tid % LANES);
            float g = 0.0f, h = 0.0f;
            if (deg[u] != 0u) {
                for (uint32_t q = lo[u] + l; q < hi[u]; q += LANES) {
                    const float2 pr = __ldcg(partials + q);
                    g += pr.x;
                    h += pr.y;
                }
            }
            if (LANES > 1) {
#pragma unroll
                for (int o = 1; o < LANES; o <<= 1) {
                    g += __shfl_xor_sync(0xffffffffu, g, o);
                    h += __shfl_xor_sync(0xffffffffu, h, o);
                }
            }
            if (l == 0) {
                float r = 0.0f;
                if (deg[u] != 0u) {
                    r = g / (lambda * deg[u] + h);
                    if (nmf) r = fmaxf(r, 0.0f);
                }
                out[s] = r;
                if (peer_ll != nullptr) {
                    const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(r);
                    for (int p = 0; p < nranks; ++p)
                        if (p != rank) {
                            unsigned long long* dst = peer_ll[p] + vec_off + s;
                            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
                        }
                }
            }
        }
    }
}

// Bounded waits.  A wait that does not complete within kWaitLimitNs (a protocol bug, a lost peer, a device shared with
// another grid-barrier kernel) records a code in the session's status word and makes the kernel return; the host then
// reports MF_ERR_STATE.  Without a status word (legacy per-launch kernels) the wait traps.
constexpr unsigned long long kWaitLimitNs = 4000000000ull;  // 4 s
enum : unsigned { kStatusBarrierTimeout = 1u, kStatusExchangeTimeout = 2u };
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned status_peek(const unsigned* status) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(status) : "memory");
    return v;
}

// receiving side of the LL exchange: grid-stride over the factor entries owned by peers; polls the entry's receive
// word until its epoch half matches, then stores the value half into the factor vector.  Returns false on timeout.
__device__ __forceinline__ bool ll_unpack_entries(const unsigned long long* ll, float* __restrict__ vec, int64_t dim, int64_t own_lo,
                                                  int64_t own_hi, unsigned epoch, unsigned* status = nullptr) {
    const int64_t nother = dim - (own_hi - own_lo);
    bool ok = true;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nother && ok; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = j < own_lo ? j : j + (own_hi - own_lo);
        unsigned long long w;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(ll + i) : "memory");
            if ((unsigned)(w >> 32) == epoch) break;
            if ((++spins & 0x3fffu) == 0u) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kWaitLimitNs || (status != nullptr && status_peek(status) != 0u)) {
                    if (status == nullptr) __trap();
                    atomicCAS(status, 0u, kStatusExchangeTimeout);
                    ok = false;
                    break;
                }
            }
        }
        if (ok) vec[i] = __uint_as_float((unsigned)(w & 0xffffffffull));
    }
    return ok;
}

// Grid-wide barrier on a monotonic counter (every CTA of the launch is resident: one CTA per SM, checked with the occupancy
// API at session creation, or enforced by a cooperative launch).  Bounded: a wait that does not complete (the device is
// shared with another grid-barrier kernel, a lost peer) records a code in the session's status word and the kernel
// returns; the host then reports MF_ERR_STATE — no trap, the CUDA context stays usable.
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned target, unsigned* status, int* s_abort) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v, spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int)(v - target) >= 0) break;
            if ((++spins & 0x3fffu) == 0u) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kWaitLimitNs || status_peek(status) != 0u) {
                    atomicCAS(status, 0u, kStatusBarrierTimeout);
                    *s_abort = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    return *s_abort == 0;
}

// L2 prefetch of a stretch of the rating stream (TMA engine, no shared memory involved): `bytes` a multiple of 16
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One batch of (up to) four items through the 8-lane groups of a warp: group `grp` streams the item whose descriptor is d.
template <int MODE, class Args>
__device__ __forceinline__ void group_batch(const Args& a, const uint4 d, const int sl, const float* __restrict__ sm_new,
                                            const float* __restrict__ sm_add, const float* __restrict__ sm_old) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve;
    const uint32_t len = d.y;
    const uint32_t lane_off = 4u * (uint32_t)sl;     // this lane's 4 entries inside a 32-entry step
    // The storage is in work-list order, so what this CTA will read pf_dist entries from now is simply this
    // item's stretch shifted by pf_dist: one lane per item asks the L2 for it (the per-lane loads of the
    // register ring then hit L2 instead of paying the DRAM latency under load).
    if (a.pf_dist() != 0u && sl == 0 && len != 0u) {
        const uint32_t q = d.x + a.pf_dist();
        if (q < a.npad()) {
            const uint32_t n = q + len <= a.npad() ? len : a.npad() - q;  // multiples of 8 entries
            l2_prefetch(a.idx16() + q, n * 2u);
            l2_prefetch(a.val() + q, n * 4u);
        }
    }
    float s_add = 0.0f, s_old = 0.0f;
    if (len != 0u) {
        // L2 loads: the finalize of this very launch rewrites the vector s_add points into (after the grid
        // barrier) and the persistent kernel reads vectors other CTAs wrote earlier in the launch, so the
        // non-coherent path and L1 are off limits
        if (ADD) s_add = __ldcg(a.s_add() + a.seg_offset() + d.z);
        if (SUB) s_old = __ldcg(a.s_old() + a.seg_offset() + d.z);
    }
    const uint32_t maxlen = __reduce_max_sync(kFull, len);
    const uint32_t minlen = __reduce_min_sync(kFull, len);
    float g = 0.0f, h = 0.0f;
    stream_batch<MODE>(a, d.x + lane_off, len, minlen, maxlen, lane_off, sm_new, sm_add, sm_old, s_add, s_old, g, h);
    if (SOLVE) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            g += __shfl_xor_sync(kFull, g, o);
            h += __shfl_xor_sync(kFull, h, o);
        }
        if (sl == 0 && len != 0u) a.partials()[d.w] = make_float2(g, h);
    }
}

// One item of at most 32 entries, streamed by ONE lane, with the arithmetic of an 8-lane group: the group's lane q owns
// entries 4q .. 4q+3 and starts its (g, h) at zero; the butterfly then adds the eight lane sums as
// ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)), lanes without entries contributing +0.  Replayed here in that order.
template <int MODE, class Args>
__device__ __forceinline__ void lane_item(const Args& a, const uint4 d, const float* __restrict__ sm_new,
                                          const float* __restrict__ sm_add, const float* __restrict__ sm_old) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve;
    constexpr bool WRITE = SUB || ADD;
    const uint32_t len = d.y;
    float s_add = 0.0f, s_old = 0.0f;
    if (len != 0u) {
        if (ADD) s_add = __ldcg(a.s_add() + a.seg_offset() + d.z);
        if (SUB) s_old = __ldcg(a.s_old() + a.seg_offset() + d.z);
    }
    float pg[8], ph[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { pg[q] = 0.0f; ph[q] = 0.0f; }
#pragma unroll
    for (int q0 = 0; q0 < 8; q0 += 2) {  // two 4-entry groups (24 bytes each) in flight per lane
        Step e[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            e[u].i = make_uint2(0u, 0u);
            e[u].v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (4u * (uint32_t)(q0 + u) < len) e[u] = load_step(a.idx16(), a.val(), d.x + 4u * (uint32_t)(q0 + u));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (4u * (uint32_t)(q0 + u) < len) {
                calc4<MODE>(e[u], sm_new, sm_add, sm_old, s_add, s_old, pg[q0 + u], ph[q0 + u]);
                if (WRITE) __stcs(reinterpret_cast<float4*>(a.val() + d.x + 4u * (uint32_t)(q0 + u)), e[u].v);
            }
        }
    }
    if (SOLVE && len != 0u) {
        const float g = ((pg[0] + pg[1]) + (pg[2] + pg[3])) + ((pg[4] + pg[5]) + (pg[6] + pg[7]));
        const float h = ((ph[0] + ph[1]) + (ph[2] + ph[3])) + ((ph[4] + ph[5]) + (ph[6] + ph[7]));
        a.partials()[d.w] = make_float2(g, h);
    }
}

// One CTA's share of a sweep: the work items [ib, ie) of the list, starting in panel p (the panel that holds item ib).
// smem: the staged panel vectors; s_ctr: a shared-memory counter the warps pull batches from.
template <int MODE, bool SHORT = false, class Args>
__device__ __forceinline__ void sweep_cta_range(const Args& a, float* smem, unsigned* s_ctr, uint32_t ib, uint32_t ie, int p, const unsigned tid,
                                                const unsigned nthr, uint32_t pend_first = 0xffffffffu) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    const uint32_t PR = a.panel_rows();
    const uint32_t stride = PR + 8;  // 8 zeroed floats behind each panel: the padding slot idx16 == PR
    // shared-memory vectors, in this order: [new] [add (only when separate)] [old]
    constexpr bool NEEDNEW = SOLVE || (ADD && !ADDSEP);
    float* sm_new = smem;
    float* sm_add = smem;
    float* sm_old = smem;
    {
        int n = 0;
        if (NEEDNEW) { sm_new = smem + n * stride; ++n; }
        if (ADD) { if (ADDSEP) { sm_add = smem + n * stride; ++n; } else sm_add = sm_new; }
        if (SUB) { sm_old = smem + n * stride; ++n; }
    }
    const float* g_add = ADDSEP ? a.g_add() : a.g_new();

    const int lane = tid & 31;
    const int grp = lane >> 3, sl = lane & 7;
    const uint4* __restrict__ items = reinterpret_cast<const uint4*>(a.items());

    bool first = pend_first != 0xffffffffu;  // the end of the first panel may have been fetched by the caller
    while (ib < ie && p < a.npanels()) {
        const uint32_t pend = first ? pend_first : a.panel_item_ptr()[p + 1];
        first = false;
        const uint32_t pe = ie < pend ? ie : pend;
        if (pe > ib) {
            __syncthreads();  // every warp is done with the previous panel and counter
            const int64_t base = (int64_t)p * PR;
            const uint32_t cnt = (uint32_t)((a.gdim() - base) < (int64_t)PR ? (a.gdim() - base) : (int64_t)PR);
            if (NEEDNEW) stage_panel(sm_new, a.g_new(), base, cnt, stride, tid, nthr);
            if (ADD && ADDSEP) stage_panel(sm_add, g_add, base, cnt, stride, tid, nthr);
            if (SUB) stage_panel(sm_old, a.g_old(), base, cnt, stride, tid, nthr);
            if (tid == 0) *s_ctr = ib;
            __syncthreads();

            if (!SHORT) {
                // Batches of four items (one per 8-lane group); neighbours in the list have (nearly) the same length.  The
                // next batch's descriptors are fetched while this batch is processed.
                uint32_t i0 = 0;
                if (lane == 0) i0 = atomicAdd(s_ctr, 4u);
                i0 = __shfl_sync(kFull, i0, 0);
                uint4 d = make_uint4(0u, 0u, 0u, 0u);  // {start, len, seg, slot}; len 0 = no item
                if (i0 + grp < pe) d = __ldg(items + i0 + grp);
                while (i0 < pe) {
                    uint32_t i0n = 0;
                    if (lane == 0) i0n = atomicAdd(s_ctr, 4u);
                    i0n = __shfl_sync(kFull, i0n, 0);
                    uint4 dn = make_uint4(0u, 0u, 0u, 0u);
                    if (i0n + grp < pe) dn = __ldg(items + i0n + grp);
                    group_batch<MODE>(a, d, sl, sm_new, sm_add, sm_old);
                    i0 = i0n;
                    d = dn;
                }
            } else {
                // Short-piece copies (a few entries per piece: the Yahoo-Music shape has ~6.5): the warp takes 32 items at a
                // time, one descriptor per lane (one coalesced 512-byte load), and when all 32 are at most one 32-entry step
                // long every LANE streams its own item — 32 items in flight per warp instead of 4, and with the work-list
                // order of the storage the lanes' loads are still neighbours in memory.  The arithmetic is the 8-lane
                // group's, replayed by one lane (lane_item): same partial sums, same tree, same bits.  A batch that holds a
                // longer item goes through the groups, four items at a time.
                uint32_t i0 = 0;
                if (lane == 0) i0 = atomicAdd(s_ctr, 32u);
                i0 = __shfl_sync(kFull, i0, 0);
                uint4 d = make_uint4(0u, 0u, 0u, 0u);
                if (i0 + lane < pe) d = __ldg(items + i0 + lane);
                while (i0 < pe) {
                    uint32_t i0n = 0;
                    if (lane == 0) i0n = atomicAdd(s_ctr, 32u);
                    i0n = __shfl_sync(kFull, i0n, 0);
                    uint4 dn = make_uint4(0u, 0u, 0u, 0u);
                    if (i0n + lane < pe) dn = __ldg(items + i0n + lane);
                    if (__all_sync(kFull, d.y <= 32u)) {
                        lane_item<MODE>(a, d, sm_new, sm_add, sm_old);
                    } else {
#pragma unroll 1
                        for (int sb = 0; sb < 8; ++sb) {
                            uint4 dd;
                            dd.x = __shfl_sync(kFull, d.x, 4 * sb + grp);
                            dd.y = __shfl_sync(kFull, d.y, 4 * sb + grp);
                            dd.z = __shfl_sync(kFull, d.z, 4 * sb + grp);
                            dd.w = __shfl_sync(kFull, d.w, 4 * sb + grp);
                            if (__all_sync(kFull, dd.y == 0u)) break;
                            group_batch<MODE>(a, dd, sl, sm_new, sm_add, sm_old);
                        }
                    }
                    i0 = i0n;
                    d = dn;
                }
            }
        }
        ib = pe;
        ++p;
    }
}

// Programmatic dependent launch (sweeps follow one another in the stream, each depending on the one before): the kernel
// lets its successor be scheduled at once (its CTAs take an SM as soon as one of ours leaves) and the successor runs its
// prologue — item range, first panel: layout data no sweep writes — before it waits for our completion and memory flush.
// What that removes from the 2kT dependent launches of an outer iteration is the launch gap and the prologue's dependent
// L2 round trips.  Both instructions are no-ops when the launch does not carry the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int MODE, bool SHORT>
__global__ void __launch_bounds__(kSweepThreads, 1) k_panel_sweep(PanelSweepArgs a) {
    constexpr bool SOLVE = MODE & kSolve;
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned s_ctr;
    pdl_launch_dependents();
    const bool tracing = a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (tracing) a.trace[0] = global_ns();
    const uint32_t ib = a.cta_item_ptr[blockIdx.x];
    const uint32_t ie = a.cta_item_ptr[blockIdx.x + 1];
    const int p = first_panel(a.panel_item_ptr, a.npanels, ib);
    const uint32_t pend0 = p < a.npanels ? a.panel_item_ptr[p + 1] : 0u;
    pdl_wait();  // everything below reads what the previous sweep wrote (factor vectors, residual) or writes what it read
    if (tracing) a.trace[1] = global_ns();
    if (a.trace_cta != nullptr && threadIdx.x == 0) a.trace_cta[4 * blockIdx.x + 0] = global_ns();
    sweep_cta_range<MODE, SHORT>(LaunchView{a}, smem, &s_ctr, ib, ie, p, threadIdx.x, blockDim.x, pend0);
    if (tracing) { a.trace[2] = global_ns(); a.trace[6] = (unsigned long long)MODE; }
    if (a.trace_cta != nullptr && threadIdx.x == 0) {
        a.trace_cta[4 * blockIdx.x + 1] = global_ns();
        a.trace_cta[4 * blockIdx.x + 2] = ((unsigned long long)ib << 32) | ie;
        a.trace_cta[4 * blockIdx.x + 3] = (unsigned long long)p;
    }
    if (SOLVE && a.fin.enabled) {
        // the first finalize round's segment metadata is on its way while the grid assembles at the barrier
        FinalizeMeta meta;
        if (a.fin.lanes == 32) finalize_load_meta<32>(a.fin.nseg, a.fin.slot_ptr, a.fin.seg_ptr, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, meta);
        else finalize_load_meta<1>(a.fin.nseg, a.fin.slot_ptr, a.fin.seg_ptr, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, meta);
        // grid-wide barrier (monotonic counter; every CTA of the launch is resident), then the CTAs share the segments
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = 0;
        if (!grid_barrier(a.fin.bar, a.fin.bar_target, a.fin.status, &s_abort)) return;
        if (tracing) a.trace[3] = global_ns();
        if (a.fin.lanes == 32)
            finalize_segments<32>(a.fin.nseg, a.fin.slot_ptr, a.partials, a.fin.seg_ptr, a.fin.lambda, a.fin.nmf, a.fin.out,
                                  a.fin.peer_ll, a.fin.vec_off, a.fin.rank, a.fin.nranks, a.fin.epoch, &meta);
        else
            finalize_segments<1>(a.fin.nseg, a.fin.slot_ptr, a.partials, a.fin.seg_ptr, a.fin.lambda, a.fin.nmf, a.fin.out,
                                 a.fin.peer_ll, a.fin.vec_off, a.fin.rank, a.fin.nranks, a.fin.epoch, &meta);
        if (tracing) a.trace[4] = global_ns();
        if (a.fin.ll != nullptr) ll_unpack_entries(a.fin.ll, a.fin.vec, a.fin.dim, a.fin.own_lo, a.fin.own_hi, a.fin.epoch, a.fin.status);
        if (tracing) a.trace[5] = global_ns();
    }
}

// =============================================================================================
// Persistent CCD++ kernel: ONE cooperative launch runs a whole outer iteration on this GPU — k ranks x [fused CSC sweep,
// fused CSR sweep, (T-1) x (solve CSC, solve CSR)] — what the reference drives as k*(2+2T) launches with a device
// synchronisation after each (cuda_src/CCD_CUDA.cu:339-378).  Grid = one 1024-thread CTA per SM (co-residency is
// enforced by cudaLaunchCooperativeKernel).  A phase = [stage panels, stream the CTA's items] -> grid barrier ->
// [finalize: slots added in order, g / (lambda*deg + h), multi-GPU: LL words to the peers, poll the peers' words] ->
// grid barrier.  Same sweep body and same finalize as the per-launch kernels, so the results are bit-identical; what
// goes away is the fixed cost of 2kT dependent launches (launch latency, metadata round trips, pipeline ramp-up): the
// item ranges and first panels of both copies are read once per launch, and a phase boundary costs two grid barriers.
// The copy of v_t for the next outer iteration's add-back (v_old) rides on the last finalize of the rank.
// =============================================================================================
// The launch's arguments live in __constant__ memory (written by the host right before the launch): every device
// function reads them as constant-bank operands, no registers.  That matters because the phases are OUT-OF-LINE calls:
// with the sweep bodies inlined into the phase loop the compiler hoisted their loop-invariant address arithmetic out of
// the loop and spilled it (4.5 KB of spills per thread, 2x slower sweeps); as calls, each mode keeps the register
// allocation it has as a kernel of its own.
__constant__ PersistArgs c_P;

struct PhaseVectors {
    const float *g_new, *g_add, *g_old, *s_add, *s_old;
    float* out;      // the vector this phase solves (full length)
    float* v_prev;   // CSR phases: v_old[t]
    const float* v;  // CSR phases: H[t]
    int last_inner;  // 1 on the last inner iteration of the rank
};
struct PhaseSweep {
    const uint16_t* idx16;
    float* val;
    const WorkItem* items;
    const uint32_t* panel_item_ptr;
    float2* partials;
    int64_t gdim, seg_offset;
    int npanels;
    uint32_t panel_rows;
};
struct PersistShared {
    PhaseVectors vec;
    PhaseSweep sweep;  // this phase's copy: per-side constants out of c_P
    uint32_t ib[2], ie[2];  // this CTA's item range on the CSC [0] and the CSR [1] copy
    int p0[2];              // the panel holding its first item
    int mode;
    int abort;
    unsigned ctr;
    unsigned zero;  // always 0, rewritten every phase: see persist_sweep
};
__device__ __forceinline__ PersistShared& persist_shared() {
    __shared__ PersistShared s;
    return s;
}

// What a sweep reads, as seen from the persistent kernel: a shared-memory record that thread 0 rewrites every phase
// (per-side constants copied from c_P + the phase's vectors).  Reading c_P directly inside the sweep bodies looks free
// (constant-bank operands) but makes every address computation loop-invariant with respect to the phase loop; ptxas
// then hoists eight bodies' worth of them out of that loop, spills them, and reloads them inside the streaming loops.
template <int SIDE>
struct PersistView {
    __device__ __forceinline__ const PersistSide& sw() const { return SIDE == 0 ? c_P.csc : c_P.csr; }
    __device__ __forceinline__ const uint16_t* idx16() const { return sw().idx16; }
    __device__ __forceinline__ float* val() const { return sw().val; }
    __device__ __forceinline__ const WorkItem* items() const { return sw().items; }
    __device__ __forceinline__ const uint32_t* panel_item_ptr() const { return sw().panel_item_ptr; }
    __device__ __forceinline__ int npanels() const { return sw().npanels; }
    __device__ __forceinline__ uint32_t panel_rows() const { return sw().panel_rows; }
    __device__ __forceinline__ int64_t gdim() const { return sw().gdim; }
    __device__ __forceinline__ int64_t seg_offset() const { return sw().seg_offset; }
    __device__ __forceinline__ const float* g_new() const { return persist_shared().vec.g_new; }
    __device__ __forceinline__ const float* g_add() const { return persist_shared().vec.g_add; }
    __device__ __forceinline__ const float* g_old() const { return persist_shared().vec.g_old; }
    __device__ __forceinline__ const float* s_add() const { return persist_shared().vec.s_add; }
    __device__ __forceinline__ const float* s_old() const { return persist_shared().vec.s_old; }
    __device__ __forceinline__ float2* partials() const { return sw().partials; }
    __device__ __forceinline__ uint32_t pf_dist() const { return 0u; }
    __device__ __forceinline__ uint32_t npad() const { return 0u; }
};

// Inlined into the phase loop.  Everything a sweep derives from the thread index is derived from `tid + zero`, where
// `zero` is a shared-memory word rewritten (with 0) every phase: otherwise the compiler hoists each mode's loop-invariant
// values out of the phase loop, where the eight bodies' worth of them cannot all stay in registers, and reloads the
// spilled ones inside the streaming loops (4 KB of spills per thread, sweeps 2.5x slower).
template <int MODE, int SIDE>
__device__ __forceinline__ void persist_sweep(unsigned tid, unsigned nthr, unsigned zero) {
    extern __shared__ __align__(16) float smem[];
    PersistShared& S = persist_shared();
    sweep_cta_range<MODE>(PersistView<SIDE>{}, smem + zero, &S.ctr, S.ib[SIDE] + zero, S.ie[SIDE], S.p0[SIDE], tid, nthr);
}

// finalize of one solving phase: this side's block of the solved vector from the slots, then (multi-GPU) the exchange;
// CSR side, last inner iteration: keep v_t for the next outer iteration's add-back (nobody reads v_prev any more in this
// rank, nobody writes v_t again before the next outer iteration)
template <int SIDE>
__device__ __noinline__ bool persist_finalize(unsigned epoch) {
    const PersistSide& sd = SIDE == 0 ? c_P.csc : c_P.csr;
    PersistShared& S = persist_shared();
    float* vec = S.vec.out;
    float* out = vec + sd.seg_offset;
    if (sd.lanes == 32)
        finalize_segments<32>(sd.nseg, sd.slot_ptr, sd.partials, sd.seg_ptr, c_P.lambda, c_P.nmf, out, sd.peer_ll, sd.seg_offset, c_P.rank, c_P.nranks, epoch);
    else
        finalize_segments<1>(sd.nseg, sd.slot_ptr, sd.partials, sd.seg_ptr, c_P.lambda, c_P.nmf, out, sd.peer_ll, sd.seg_offset, c_P.rank, c_P.nranks, epoch);
    bool ok = true;
    if (sd.ll != nullptr) ok = ll_unpack_entries(sd.ll, vec, sd.dim, sd.seg_offset, sd.seg_offset + sd.nseg, epoch, c_P.status);
    if (SIDE == 1 && S.vec.last_inner) {
        float* v_prev = S.vec.v_prev;
        const float* v = S.vec.v;
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < c_P.ldn; j += (int64_t)gridDim.x * blockDim.x)
            v_prev[j] = __ldcg(v + j);
    }
    return ok;
}

// one phase on copy SIDE (0: CSC, solves v_t = H[t]; 1: CSR, solves u_t = W[t]); returns false when a wait timed out
template <int SIDE>
__device__ __forceinline__ bool persist_phase(unsigned ph) {
    PersistShared& S = persist_shared();
    if (threadIdx.x == 0) {
        const int T = c_P.T;
        const int t = (int)(ph / (unsigned)(2 * T)), it = (int)((ph >> 1) % (unsigned)T);
        const int sub = t == 0 ? c_P.pending : t - 1;  // rank whose subtraction is still deferred
        float* u = c_P.W + (int64_t)t * c_P.ldm;
        float* v = c_P.H + (int64_t)t * c_P.ldn;
        float* v_prev = c_P.v_old + (int64_t)t * c_P.ldn;  // v_t as the previous outer iteration left it
        const float* u_sub = sub >= 0 ? c_P.W + (int64_t)sub * c_P.ldm : nullptr;
        const float* v_sub = sub >= 0 ? c_P.H + (int64_t)sub * c_P.ldn : nullptr;
        PhaseVectors V;
        V.g_new = SIDE == 0 ? u : v; V.g_add = nullptr; V.g_old = nullptr; V.s_add = nullptr; V.s_old = nullptr;
        V.out = SIDE == 0 ? v : u; V.v_prev = v_prev; V.v = v; V.last_inner = it == T - 1;
        int mode = kSolve;
        if (it == 0) {  // first inner iteration: the deferred subtraction and this rank's add-back ride along
            if (SIDE == 0) {
                V.g_old = u_sub; V.s_add = v; V.s_old = v_sub;
                mode |= (sub >= 0 ? kSub : 0) | (c_P.add ? kAdd : 0);
            } else {
                V.g_add = v_prev; V.s_add = u; V.s_old = u_sub;
                V.g_old = (sub == t) ? v_prev : v_sub;  // k == 1: the subtracted rank's v was just overwritten
                mode |= (sub >= 0 ? kSub : 0) | (c_P.add ? (kAdd | kAddSep) : 0);
            }
        }
        S.vec = V;
        const PersistSide& sd = SIDE == 0 ? c_P.csc : c_P.csr;
        PhaseSweep W;
        W.idx16 = sd.idx16; W.val = sd.val; W.items = sd.items; W.panel_item_ptr = sd.panel_item_ptr; W.partials = sd.partials;
        W.gdim = sd.gdim; W.seg_offset = sd.seg_offset; W.npanels = sd.npanels; W.panel_rows = sd.panel_rows;
        S.sweep = W;
        S.mode = mode;
        S.zero = ph >> 31;  // 0 — but only at run time
    }
    __syncthreads();
    constexpr int kA = SIDE == 0 ? kAdd : (kAdd | kAddSep);
    // a zero the compiler cannot see through, re-read from shared memory every phase: whatever a sweep derives from the
    // thread index is derived from `tid + zero`, so none of it is loop-invariant with respect to the phase loop
    const unsigned zero = 0u;
    const unsigned tid = threadIdx.x + zero, nthr = blockDim.x + zero;
    switch (S.mode) {
        case kSolve: persist_sweep<kSolve, SIDE>(tid, nthr, zero); break;
        case kSolve | kSub: persist_sweep<kSolve | kSub, SIDE>(tid, nthr, zero); break;
        case kSolve | kA: persist_sweep<kSolve | kA, SIDE>(tid, nthr, zero); break;
        default: persist_sweep<kSolve | kSub | kA, SIDE>(tid, nthr, zero); break;
    }
    if (!grid_barrier(c_P.bar, c_P.bar_base + (2u * ph + 1u) * gridDim.x, c_P.status, &S.abort)) return false;
    if (!persist_finalize<SIDE>(c_P.epoch_base + ph + 1u)) S.abort = 1;  // every thread that timed out says so; the barrier makes it CTA-wide
    if (!grid_barrier(c_P.bar, c_P.bar_base + (2u * ph + 2u) * gridDim.x, c_P.status, &S.abort)) return false;
    if (c_P.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) c_P.stamps[ph + 1u] = global_ns();
    return true;
}

__global__ void __launch_bounds__(kSweepThreads, 1) k_ccd_persistent() {
    PersistShared& S = persist_shared();
    {   // this CTA's item range and first panel on both copies: read once per launch
        const uint32_t ib_c = c_P.csc.cta_item_ptr[blockIdx.x], ie_c = c_P.csc.cta_item_ptr[blockIdx.x + 1];
        const uint32_t ib_r = c_P.csr.cta_item_ptr[blockIdx.x], ie_r = c_P.csr.cta_item_ptr[blockIdx.x + 1];
        const int p_c = first_panel(c_P.csc.panel_item_ptr, c_P.csc.npanels, ib_c);
        const int p_r = first_panel(c_P.csr.panel_item_ptr, c_P.csr.npanels, ib_r);
        if (threadIdx.x == 0) {
            S.ib[0] = ib_c; S.ie[0] = ie_c; S.p0[0] = p_c;
            S.ib[1] = ib_r; S.ie[1] = ie_r; S.p0[1] = p_r;
            S.abort = 0;
            if (c_P.stamps != nullptr && blockIdx.x == 0) c_P.stamps[0] = global_ns();
        }
    }
    __syncthreads();
    const unsigned nphase = 2u * (unsigned)c_P.k * (unsigned)c_P.T;
#pragma unroll 1
    for (unsigned ph = 0; ph < nphase; ph += 2) {
        if (!persist_phase<0>(ph)) return;
        if (!persist_phase<1>(ph + 1u)) return;
    }
}

// =============================================================================================
// TMA bulk-copy variant of the panel sweep (MF_PIPELINE_TMA_BULK; measured alternative, not the default —
// DESIGN.md §4): warp 0 of the CTA is a producer that moves
// every work item of the CTA's range from HBM into a ring of shared-memory slots with 1-D bulk
// async copies (cp.async.bulk, completion counted on an mbarrier per slot); warps 1..31 consume the
// slots.  The HBM stream is thereby decoupled from the arithmetic: ~40-58 items (70-100 KB) are in
// flight per SM whatever the consumers are doing, and the consumers hold no prefetch registers.
//   slot      = 64-byte header (the item descriptor) + chunk*2 bytes of indices + chunk*4 bytes of values
//   full[s]   producer -> consumers: arrive.expect_tx(bytes) by the producer lane + complete_tx by the copies
//   empty[s]  consumers -> producer: one arrive by the 8-lane group that read the slot
//   item n of the CTA (n counted from the CTA's first item, across panels) lives in slot n % nslots during
//   ring round n / nslots; barrier parities follow from n, so no state is reset between panels.
// Consumers claim batches of four consecutive items from a shared-memory counter exactly as the register
// ring kernel does, so the reduction tree of an item (lane-serial over its steps, xor-butterfly over the
// 8 lanes of its group) is identical in both kernels — they give bit-identical results.
// =============================================================================================
namespace tma {

constexpr uint32_t kChunkMax = 512;                                  // largest item the slots hold
constexpr uint32_t kSlotHeader = 64;
constexpr uint32_t kSlotBytes = kSlotHeader + kChunkMax * 2 + kChunkMax * 4;  // 3136: 784 words = 16 mod 32 banks
constexpr uint32_t kSpinLimit = 1u << 24;

// First protocol timeout of a launch is recorded here (code, block, warp, lane, item n, slot, round, observed)
// and the thread leaves the kernel: results are then wrong, but nothing hangs and the host can report it.
__device__ unsigned int g_timeout[8] = {0, 0, 0, 0, 0, 0, 0, 0};
__device__ __forceinline__ void report_timeout(unsigned code, unsigned n, unsigned slot, unsigned round, unsigned seen) {
    if (atomicCAS(&g_timeout[0], 0u, code) == 0u) {
        g_timeout[1] = blockIdx.x; g_timeout[2] = threadIdx.x >> 5; g_timeout[3] = threadIdx.x & 31;
        g_timeout[4] = n; g_timeout[5] = slot; g_timeout[6] = round; g_timeout[7] = seen;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}
// bounded wait: a protocol bug traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > kSpinLimit) __trap();
}
// bounded wait that reports and returns false instead of trapping (cp.async pipeline)
__device__ __forceinline__ bool mbar_wait_report(uint32_t bar, uint32_t parity, unsigned code, unsigned n, unsigned slot,
                                                 unsigned round) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > kSpinLimit) { report_timeout(code, n, slot, round, 0u); return false; }
    return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace tma

// One batch out of the shared-memory slot ring: the 8-lane group `sl`-lane belongs to streams the item whose
// descriptor is `d` and whose indices / values sit in slot `sb` (len 0 = no item for this group).
template <int MODE>
__device__ __forceinline__ void consume_slot(const PanelSweepArgs& a, const uint4 d, const unsigned char* sb, int sl,
                                             const float* __restrict__ sm_new, const float* __restrict__ sm_add,
                                             const float* __restrict__ sm_old) {
    using namespace tma;
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve;
    constexpr bool WRITE = SUB || ADD;
    const uint32_t len = d.y;
    const uint32_t lane_off = 4u * (uint32_t)sl;
    float s_add = 0.0f, s_old = 0.0f;
    if (len != 0u) {
        if (ADD) s_add = __ldg(a.s_add + a.seg_offset + d.z);
        if (SUB) s_old = __ldg(a.s_old + a.seg_offset + d.z);
    }
    const uint32_t maxlen = __reduce_max_sync(kFull, len);
    const uint32_t minlen = __reduce_min_sync(kFull, len);
    const uint16_t* sidx = reinterpret_cast<const uint16_t*>(sb + kSlotHeader) + lane_off;
    const float* sval = reinterpret_cast<const float*>(sb + kSlotHeader + kChunkMax * 2u) + lane_off;
    float* gval = a.val + d.x + lane_off;
    float g = 0.0f, h = 0.0f;
    uint32_t o = 0;
#pragma unroll 2
    for (; o + 32u <= minlen; o += 32u) {  // every lane of every group has 4 entries here
        Step e;
        e.i = *reinterpret_cast<const uint2*>(sidx + o);
        e.v = *reinterpret_cast<const float4*>(sval + o);
        calc4<MODE>(e, sm_new, sm_add, sm_old, s_add, s_old, g, h);
        if (WRITE) __stcs(reinterpret_cast<float4*>(gval + o), e.v);
    }
#pragma unroll 1
    for (; o < maxlen; o += 32u) {
        if (o + lane_off < len) {
            Step e;
            e.i = *reinterpret_cast<const uint2*>(sidx + o);
            e.v = *reinterpret_cast<const float4*>(sval + o);
            calc4<MODE>(e, sm_new, sm_add, sm_old, s_add, s_old, g, h);
            if (WRITE) __stcs(reinterpret_cast<float4*>(gval + o), e.v);
        }
    }
    if (SOLVE) {
#pragma unroll
        for (int q = 1; q < 8; q <<= 1) {
            g += __shfl_xor_sync(kFull, g, q);
            h += __shfl_xor_sync(kFull, h, q);
        }
        if (sl == 0 && len != 0u) a.partials[d.w] = make_float2(g, h);
    }
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_panel_sweep_tma(PanelSweepArgs a) {
    using namespace tma;
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ unsigned s_ctr;
    __shared__ volatile unsigned s_issued;  // items (counted from the CTA's first) the producer has started to fill

    const uint32_t PR = a.panel_rows;
    const uint32_t stride = PR + 8;  // 8 zeroed floats behind each panel: the padding slot
    constexpr bool NEEDNEW = SOLVE || (ADD && !ADDSEP);
    float* smem = reinterpret_cast<float*>(smraw);
    float* sm_new = smem;
    float* sm_add = smem;
    float* sm_old = smem;
    int nvec = 0;
    if (NEEDNEW) { sm_new = smem + nvec * stride; ++nvec; }
    if (ADD) { if (ADDSEP) { sm_add = smem + nvec * stride; ++nvec; } else sm_add = sm_new; }
    if (SUB) { sm_old = smem + nvec * stride; ++nvec; }
    const float* g_add = ADDSEP ? a.g_add : a.g_new;
    const uint32_t vec_bytes = ((uint32_t)nvec * stride * 4u + 127u) & ~127u;
    const uint32_t NS = a.nslots;
    unsigned char* slots = smraw + vec_bytes;
    const uint32_t slots_u32 = smem_u32(slots);
    const uint32_t full_u32 = slots_u32 + NS * kSlotBytes;   // full[s]  at full_u32 + 8*s
    const uint32_t empty_u32 = full_u32 + NS * 8u;           // empty[s] at empty_u32 + 8*s

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, sl = lane & 7;
    const uint4* __restrict__ items = reinterpret_cast<const uint4*>(a.items);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NS; ++s) {
            mbar_init(full_u32 + 8u * s, 1u);
            mbar_init(empty_u32 + 8u * s, 1u);
        }
        s_issued = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t ib0 = a.cta_item_ptr[blockIdx.x];
    uint32_t ib = ib0;
    const uint32_t ie = a.cta_item_ptr[blockIdx.x + 1];
    int p = first_panel(a.panel_item_ptr, a.npanels, ib);

    while (ib < ie && p < a.npanels) {
        const uint32_t pend = a.panel_item_ptr[p + 1];
        const uint32_t pe = ie < pend ? ie : pend;
        if (pe > ib) {
            __syncthreads();  // consumers are done with the previous panel's vectors and counter
            const int64_t base = (int64_t)p * PR;
            const uint32_t cnt = (uint32_t)((a.gdim - base) < (int64_t)PR ? (a.gdim - base) : (int64_t)PR);
            if (NEEDNEW) stage_panel(sm_new, a.g_new, base, cnt, stride, threadIdx.x, blockDim.x);
            if (ADD && ADDSEP) stage_panel(sm_add, g_add, base, cnt, stride, threadIdx.x, blockDim.x);
            if (SUB) stage_panel(sm_old, a.g_old, base, cnt, stride, threadIdx.x, blockDim.x);
            if (threadIdx.x == 0) s_ctr = ib;
            __syncthreads();

            if (warp == 0) {
                // ---------------- producer: every lane moves one item per round ----------------
                for (uint32_t rbase = ib; rbase < pe; rbase += 32) {
                    const uint32_t i = rbase + lane;
                    if (i < pe) {
                        const uint32_t n = i - ib0, slot = n % NS, round = n / NS;
                        const uint4 d = __ldg(items + i);  // {start, len, seg, slot}
                        mbar_wait(empty_u32 + 8u * slot, (round & 1u) ^ 1u);  // the slot's previous occupant was read
                        unsigned char* sb = slots + slot * kSlotBytes;
                        *reinterpret_cast<uint4*>(sb) = d;
                        const uint32_t bar = full_u32 + 8u * slot, dst = slots_u32 + slot * kSlotBytes + kSlotHeader;
                        mbar_expect_tx(bar, d.y * 6u);
                        bulk_g2s(dst, a.idx16 + d.x, d.y * 2u, bar);
                        bulk_g2s(dst + kChunkMax * 2u, a.val + d.x, d.y * 4u, bar);
                    }
                    // Publish how far the ring has been (re)armed.  A consumer may only wait on full[slot] for
                    // item n once the producer has armed that slot for n: a parity wait issued earlier would
                    // alias the slot's previous ring round.
                    __syncwarp();
                    if (lane == 0) s_issued = (rbase + 32u < pe ? rbase + 32u : pe) - ib0;
                }
            } else {
                // ---------------- consumers: four items per warp, one per 8-lane group ----------------
                for (;;) {
                    uint32_t i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&s_ctr, 4u);
                    i0 = __shfl_sync(kFull, i0, 0);
                    if (i0 >= pe) break;
                    const uint32_t mine = i0 + grp;
                    const bool have = mine < pe;
                    uint4 d = make_uint4(0u, 0u, 0u, 0u);
                    const unsigned char* sb = slots;
                    uint32_t slot = 0;
                    if (have) {
                        const uint32_t n = mine - ib0;
                        slot = n % NS;
                        uint32_t spins = 0;
                        while (s_issued <= n)  // the producer has not armed this slot for item n yet
                            if (++spins > kSpinLimit) __trap();
                        mbar_wait(full_u32 + 8u * slot, (n / NS) & 1u);  // descriptor + indices + values have landed
                        sb = slots + slot * kSlotBytes;
                        d = *reinterpret_cast<const uint4*>(sb);
                    }
                    consume_slot<MODE>(a, d, sb, sl, sm_new, sm_add, sm_old);
                    __syncwarp();  // every lane of the group has finished reading the slot
                    if (have && sl == 0) mbar_arrive(empty_u32 + 8u * slot);
                }
            }
        }
        ib = pe;
        ++p;
    }
}

// =============================================================================================
// cp.async pipeline variant (MF_PIPELINE_ASYNC; measured alternative, not the default): same slot ring and the
// same consumers as the bulk-copy
// kernel above, but the ring is fed by kProducerWarps producer WARPS using 16-byte cp.async.cg copies
// (LDGSTS, L1-bypassing) — one warp-wide instruction moves 32 x 16 bytes from arbitrary addresses.  A producer warp handles a batch of four items at a time, one per 8-lane group,
// exactly like the consumers; each lane arrives on the slot's `full` mbarrier once its own copies have landed
// (cp.async groups + wait_group), so the barrier expects 8 arrivals.
//   slot_round[s]  ring round (+1) the slot is currently armed for, written by the producer after it has
//                  waited for the slot's previous occupant to be consumed; a consumer waits for its round to
//                  show up there before it issues the parity wait (which would otherwise alias the previous
//                  round of the slot).
// =============================================================================================
namespace tma {
constexpr int kProducerWarps = 4;
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
}  // namespace tma

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_panel_sweep_async(PanelSweepArgs a) {
    using namespace tma;
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ unsigned s_ctr;

    const uint32_t PR = a.panel_rows;
    const uint32_t stride = PR + 8;  // 8 zeroed floats behind each panel: the padding slot
    constexpr bool NEEDNEW = SOLVE || (ADD && !ADDSEP);
    float* smem = reinterpret_cast<float*>(smraw);
    float* sm_new = smem;
    float* sm_add = smem;
    float* sm_old = smem;
    int nvec = 0;
    if (NEEDNEW) { sm_new = smem + nvec * stride; ++nvec; }
    if (ADD) { if (ADDSEP) { sm_add = smem + nvec * stride; ++nvec; } else sm_add = sm_new; }
    if (SUB) { sm_old = smem + nvec * stride; ++nvec; }
    const float* g_add = ADDSEP ? a.g_add : a.g_new;
    const uint32_t vec_bytes = ((uint32_t)nvec * stride * 4u + 127u) & ~127u;
    const uint32_t NS = a.nslots;
    unsigned char* slots = smraw + vec_bytes;
    const uint32_t slots_u32 = smem_u32(slots);
    const uint32_t full_u32 = slots_u32 + NS * kSlotBytes;   // full[s]  at full_u32 + 8*s
    const uint32_t empty_u32 = full_u32 + NS * 8u;           // empty[s] at empty_u32 + 8*s
    volatile uint32_t* slot_round = reinterpret_cast<volatile uint32_t*>(slots + NS * (kSlotBytes + 16u));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, sl = lane & 7;
    const uint4* __restrict__ items = reinterpret_cast<const uint4*>(a.items);

    if (threadIdx.x < NS) {
        mbar_init(full_u32 + 8u * threadIdx.x, 8u);   // the 8 lanes of the producing group
        mbar_init(empty_u32 + 8u * threadIdx.x, 1u);  // lane 0 of the consuming group
        slot_round[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t ib0 = a.cta_item_ptr[blockIdx.x];
    uint32_t ib = ib0;
    const uint32_t ie = a.cta_item_ptr[blockIdx.x + 1];
    int p = first_panel(a.panel_item_ptr, a.npanels, ib);

    while (ib < ie && p < a.npanels) {
        const uint32_t pend = a.panel_item_ptr[p + 1];
        const uint32_t pe = ie < pend ? ie : pend;
        if (pe > ib) {
            __syncthreads();  // consumers are done with the previous panel's vectors and counter
            const int64_t base = (int64_t)p * PR;
            const uint32_t cnt = (uint32_t)((a.gdim - base) < (int64_t)PR ? (a.gdim - base) : (int64_t)PR);
            if (NEEDNEW) stage_panel(sm_new, a.g_new, base, cnt, stride, threadIdx.x, blockDim.x);
            if (ADD && ADDSEP) stage_panel(sm_add, g_add, base, cnt, stride, threadIdx.x, blockDim.x);
            if (SUB) stage_panel(sm_old, a.g_old, base, cnt, stride, threadIdx.x, blockDim.x);
            if (threadIdx.x == 0) s_ctr = ib;
            __syncthreads();

            if (warp < kProducerWarps) {
                // ---------------- producers: warp w feeds batches w, w+P, w+2P, ... of this panel range ----------------
                // Each lane commits its copies of a batch as one cp.async group and signals the slot's `full`
                // barrier two batches later, after cp.async.wait_group has retired that group: up to three
                // batches per lane group are in flight.  (cp.async.mbarrier.arrive.noinc was measured to fire
                // early when a lane has tens of copies outstanding; the explicit wait is exact.)
                uint32_t i0 = ib + 4u * (uint32_t)warp;
                uint4 d = make_uint4(0u, 0u, 0u, 0u);
                if (i0 + grp < pe) d = __ldg(items + i0 + grp);
                uint32_t pb0 = 0u, pb1 = 0u;  // `full` barriers of this lane's previous two batches (0 = none)
                while (i0 < pe) {
                    const uint32_t i0n = i0 + 4u * kProducerWarps;
                    uint4 dn = make_uint4(0u, 0u, 0u, 0u);
                    if (i0n + grp < pe) dn = __ldg(items + i0n + grp);  // next batch's descriptors, one batch ahead
                    uint32_t cur = 0u;
                    if (d.y != 0u) {
                        const uint32_t n = i0 + grp - ib0, slot = n % NS, round = n / NS;
                        if (!mbar_wait_report(empty_u32 + 8u * slot, (round & 1u) ^ 1u, 1u, n, slot, round)) return;  // previous occupant read
                        const uint32_t sbase = slots_u32 + slot * kSlotBytes;
                        if (sl == 0) {
                            *reinterpret_cast<uint4*>(slots + slot * kSlotBytes) = d;
                            slot_round[slot] = round + 1u;  // armed: consumers of this round may now wait on full[slot]
                        }
                        const unsigned char* gi = reinterpret_cast<const unsigned char*>(a.idx16 + d.x);
                        const unsigned char* gv = reinterpret_cast<const unsigned char*>(a.val + d.x);
                        const uint32_t nbi = d.y * 2u, nbv = d.y * 4u;  // bytes, multiples of 16
                        for (uint32_t c = 16u * (uint32_t)sl; c < nbi; c += 128u) cp_async_16(sbase + kSlotHeader + c, gi + c);
                        for (uint32_t c = 16u * (uint32_t)sl; c < nbv; c += 128u)
                            cp_async_16(sbase + kSlotHeader + kChunkMax * 2u + c, gv + c);
                        cur = full_u32 + 8u * slot;
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 2;" ::: "memory");  // the batch committed two rounds ago has landed
                    if (pb1 != 0u) mbar_arrive(pb1);
                    pb1 = pb0; pb0 = cur;
                    i0 = i0n;
                    d = dn;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                if (pb1 != 0u) mbar_arrive(pb1);
                if (pb0 != 0u) mbar_arrive(pb0);
            } else {
                // ---------------- consumers: four items per warp, one per 8-lane group ----------------
                for (;;) {
                    uint32_t i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&s_ctr, 4u);
                    i0 = __shfl_sync(kFull, i0, 0);
                    if (i0 >= pe) break;
                    const uint32_t mine = i0 + grp;
                    const bool have = mine < pe;
                    uint4 d = make_uint4(0u, 0u, 0u, 0u);
                    const unsigned char* sb = slots;
                    uint32_t slot = 0;
                    if (have) {
                        const uint32_t n = mine - ib0, round = n / NS;
                        slot = n % NS;
                        uint32_t spins = 0;
                        while (slot_round[slot] != round + 1u)  // not armed for this round yet
                            if (++spins > kSpinLimit) { report_timeout(2u, n, slot, round, slot_round[slot]); return; }
                        if (!mbar_wait_report(full_u32 + 8u * slot, round & 1u, 3u, n, slot, round)) return;  // data has landed
                        sb = slots + slot * kSlotBytes;
                        d = *reinterpret_cast<const uint4*>(sb);
                    }
#ifdef MF_DEBUG_ASSERT
                    if (have && ((d.x & 7u) || d.y > kChunkMax || (d.y & 7u) || d.y == 0u || ((uintptr_t)sb & 15u))) {
                        report_timeout(12u, d.x, d.y, mine, slot);
                        return;
                    }
#endif
                    consume_slot<MODE>(a, d, sb, sl, sm_new, sm_add, sm_old);
                    __syncwarp();  // every lane of the group has finished reading the slot
                    if (have && sl == 0) mbar_arrive(empty_u32 + 8u * slot);
                }
            }
        }
        ib = pe;
        ++p;
    }
}

// =============================================================================================
// STREAM pipeline (MF_PIPELINE_STREAM): the ratings reach the SM as a few large TMA bulk copies instead of per-lane
// loads.  The padded entries are stored in work-list order (layout.cuh), so the items [ib, ie) a CTA walks are ONE
// contiguous stretch [cta_start_ptr[b], cta_start_ptr[b+1]) of the index and value arrays.  The first kFeeders warps
// are producers (one issuing lane each): the stretch is cut into tiles of kTile entries, and tile t goes into stage
// t % NST of a shared-memory ring with two cp.async.bulk copies (indices, values; completion counted in bytes on the
// stage's `full` mbarrier), issued by producer warp t % kFeeders.  The kConsumerWarps other warps
// consume the items (one item per warp at a time, 128-entry steps, a lane owns 4 consecutive entries) reading indices and
// values from the ring: the HBM stream is decoupled from the arithmetic, the bytes in flight per SM are bounded by the
// ring (up to 96 KB) instead of by registers x warps or by what L1 is left beside the staged panels, and an item need
// not start on a 128-byte line any more (padding granularity 8 instead of 32 entries).
//   full[s]      producer -> consumers: expect_tx + complete_tx of the two copies
//   consumed[s]  entries of the stage's current tile already consumed; the warp that completes the tile resets
//                it and arrives on empty[s] (count 1), which lets the producer refill the stage
//   s_armed[s]   1 + the tile stage s is armed for: a consumer may only issue its parity wait on full[s] for tile t once
//                t has been armed (a parity wait issued a whole ring round early would alias the previous round)
// A ring of RING = NST * kTile entries is addressed with a mask (RING is a power of two); items never straddle more than
// two tiles (chunk <= kTile) and a lane's 4 entries never straddle the wrap (everything is a multiple of 8 entries).
// =============================================================================================
// Consumer warps: few on purpose.  The entries of the items being processed (4 per warp) are pinned in the ring until
// their tile has been consumed completely; with 31 consumer warps that window (124 items x ~150 entries) is larger than
// the ring itself, nothing is prefetched and the kernel runs at half the speed of the register ring (measured, round 2).
#ifndef MF_STREAM_WARPS
#define MF_STREAM_WARPS 24
#endif
// Producer warps: several, one issuing lane each.  Measured (scripts/ubench/bulk_stream.cu, B200): one thread completes a
// [wait empty, expect_tx, cp.async.bulk] round every ~280 ns whatever the copy size (2-8 KB), i.e. 6 KB tiles from one
// producer are 21 GB/s per SM = 3.2 TB/s per GPU; the rounds of different warps overlap perfectly (4 warps: 48 GB/s per
// SM = 7.1 TB/s, HBM-bound).
// A stage must be served by the same producer warp in every ring round — a warp is never a round ahead of itself,
// while another warp could be, and its parity wait on empty[s] would then alias an older phase (measured: 3 or 6
// producers on a 16-stage ring corrupt the ring) — so the producer count is a power of two, capped by the stage count.
#ifndef MF_STREAM_PRODUCERS
#define MF_STREAM_PRODUCERS 4
#endif
namespace stream {
constexpr uint32_t kTile = 1024;  // entries per stage: 2 KB of indices + 4 KB of values
constexpr uint32_t kFeeders = MF_STREAM_PRODUCERS;
constexpr uint32_t kConsumerWarps = MF_STREAM_WARPS;
constexpr uint32_t kConsumers = 32u * kConsumerWarps;
constexpr uint32_t kThreads = kConsumers + 32u * kFeeders;
static_assert((kFeeders & (kFeeders - 1u)) == 0u && kFeeders >= 1u && kFeeders <= 16u, "producer warps: a power of two");
static_assert(kThreads <= 1024u, "too many warps");
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

// One panel of a factor vector -> shared memory as ONE bulk copy issued by consumer thread 0 (completion on `bar`);
// the copy moves the panel's valid entries rounded up to 4 (factor rows are padded to 32 entries, so it stays inside the
// allocation); stage_panel_tail then zeroes everything from `cnt` to `stride`.
__device__ __forceinline__ uint32_t stage_panel_bytes(uint32_t cnt) { return ((cnt + 3u) & ~3u) * 4u; }
__device__ __forceinline__ void stage_panel_tail(float* __restrict__ sm, uint32_t cnt, uint32_t stride, uint32_t ctid) {
    for (uint32_t i = cnt + ctid; i < stride; i += kConsumers) sm[i] = 0.0f;
}
}  // namespace stream

template <int MODE>
__global__ void __launch_bounds__(stream::kThreads, 1) k_panel_sweep_stream(PanelSweepArgs a) {
    using namespace tma;
    using namespace stream;
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    constexpr bool WRITE = SUB || ADD;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ volatile unsigned s_armed[16];  // per stage: 1 + the tile the stage is armed for (0: none yet)

    const uint32_t RING = a.ring_entries, NST = RING / kTile, rmask = RING - 1u;
    float* ring_v = reinterpret_cast<float*>(smraw);
    uint16_t* ring_i = reinterpret_cast<uint16_t*>(smraw + (size_t)RING * 4u);
    const uint32_t ring_v_u32 = smem_u32(ring_v), ring_i_u32 = smem_u32(ring_i);
    const uint32_t full_u32 = ring_i_u32 + RING * 2u;     // full[s]  at full_u32 + 8*s
    const uint32_t empty_u32 = full_u32 + NST * 8u;       // empty[s] at empty_u32 + 8*s
    unsigned* consumed = reinterpret_cast<unsigned*>(smraw + (size_t)RING * 6u + (size_t)NST * 16u);
    const uint32_t vec_bar = full_u32 + NST * 20u;        // completion of the staged factor panels (one phase per panel)
    float* smem = reinterpret_cast<float*>(smraw + (((size_t)RING * 6u + (size_t)NST * 20u + 8u + 127u) & ~(size_t)127u));

    const uint32_t PR = a.panel_rows;
    const uint32_t stride = PR + 8;  // 8 zeroed floats behind each panel: the padding slot
    constexpr bool NEEDNEW = SOLVE || (ADD && !ADDSEP);
    float* sm_new = smem;
    float* sm_add = smem;
    float* sm_old = smem;
    {
        int n = 0;
        if (NEEDNEW) { sm_new = smem + n * stride; ++n; }
        if (ADD) { if (ADDSEP) { sm_add = smem + n * stride; ++n; } else sm_add = sm_new; }
        if (SUB) { sm_old = smem + n * stride; ++n; }
    }
    const float* g_add = ADDSEP ? a.g_add : a.g_new;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* __restrict__ items = reinterpret_cast<const uint4*>(a.items);

    if (threadIdx.x == 0) {
        for (uint32_t st = 0; st < NST; ++st) {
            mbar_init(full_u32 + 8u * st, 1u);
            mbar_init(empty_u32 + 8u * st, 1u);
            consumed[st] = 0u;
            s_armed[st] = 0u;
        }
        mbar_init(vec_bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ib = a.cta_item_ptr[blockIdx.x];
    const uint32_t ie = a.cta_item_ptr[blockIdx.x + 1];
    const uint32_t stream0 = a.cta_start_ptr[blockIdx.x], stream1 = a.cta_start_ptr[blockIdx.x + 1];
    int p = first_panel(a.panel_item_ptr, a.npanels, ib);  // (contains a CTA-wide barrier: the mbarriers are initialised)
    const uint32_t total = stream1 - stream0;

    if (warp < (int)kFeeders) {
        // ---------------- producers: warp w moves tiles w, w + kFeeders, ... ----------------
        const uint32_t nfeed = kFeeders < NST ? kFeeders : NST;
        if (lane == 0 && (uint32_t)warp < nfeed) {
            const uint32_t ntiles = (total + kTile - 1u) / kTile;
            for (uint32_t t = (uint32_t)warp; t < ntiles; t += nfeed) {
                const uint32_t st = t & (NST - 1u), round = t / NST;
                if (round > 0u) mbar_wait(empty_u32 + 8u * st, (round - 1u) & 1u);  // the stage's previous tile has been consumed
                const uint32_t n = total - t * kTile < kTile ? total - t * kTile : kTile;
                const uint32_t bar = full_u32 + 8u * st;
                mbar_expect_tx(bar, n * 6u);
                bulk_g2s(ring_i_u32 + st * kTile * 2u, a.idx16 + stream0 + t * kTile, n * 2u, bar);
                bulk_g2s(ring_v_u32 + st * kTile * 4u, a.val + stream0 + t * kTile, n * 4u, bar);
                __threadfence_block();
                s_armed[st] = t + 1u;
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers ----------------
        const unsigned ctid = threadIdx.x - 32u * kFeeders;
        uint32_t vec_phase = 0;
        while (ib < ie && p < a.npanels) {
            const uint32_t pend = a.panel_item_ptr[p + 1];
            const uint32_t pe = ie < pend ? ie : pend;
            if (pe > ib) {
                consumer_sync();  // every consumer warp is done with the previous panel and counter
                const int64_t base = (int64_t)p * PR;
                const uint32_t cnt = (uint32_t)((a.gdim - base) < (int64_t)PR ? (a.gdim - base) : (int64_t)PR);
                if (ctid == 0) {
                    constexpr uint32_t nvec = (NEEDNEW ? 1u : 0u) + ((ADD && ADDSEP) ? 1u : 0u) + (SUB ? 1u : 0u);
                    const uint32_t nb = stage_panel_bytes(cnt);
                    // the generic-proxy reads of the previous panel (ordered by the barrier above) precede these async-proxy writes
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(vec_bar, nvec * nb);
                    if (NEEDNEW) bulk_g2s(smem_u32(sm_new), a.g_new + base, nb, vec_bar);
                    if (ADD && ADDSEP) bulk_g2s(smem_u32(sm_add), g_add + base, nb, vec_bar);
                    if (SUB) bulk_g2s(smem_u32(sm_old), a.g_old + base, nb, vec_bar);
                }
                mbar_wait(vec_bar, vec_phase & 1u);
                ++vec_phase;
                if (NEEDNEW) stage_panel_tail(sm_new, cnt, stride, ctid);
                if (ADD && ADDSEP) stage_panel_tail(sm_add, cnt, stride, ctid);
                if (SUB) stage_panel_tail(sm_old, cnt, stride, ctid);
                consumer_sync();

                // One item per warp at a time (static round-robin over the consumer warps: neighbours in the list have
                // nearly the same length).  What the consumers hold pinned in the ring is then kConsumerWarps items
                // (~150 entries each) instead of four per warp: with 8-lane groups the pinned window was as large as the
                // ring and nothing was prefetched (measured: the warps spent 40 % of their time waiting for tiles).
                // Lane l owns entries 128 s + 4 l .. + 3 of step s; the item's sums are lane-serial over the steps, then an
                // xor-butterfly over the 32 lanes: a fixed tree of the item alone (not the 8-lane tree of the other
                // pipelines: results agree with them to rounding, not bit for bit).
                // Descriptors are fetched kDepth items ahead, the per-segment scalars of the fused modes one item ahead; the
                // item loop is unrolled kDepth times so that every load in flight has a register of its own (rotating the
                // descriptors through moves makes each move wait for the load it copies: prefetch distance zero, measured).
                constexpr uint32_t kDepth = 4;
                const uint32_t cw = (uint32_t)warp - kFeeders;
                const uint32_t lane_off = 4u * (uint32_t)lane;
                uint4 dq[kDepth];  // {start, len, seg, slot}; len 0 = no item
                float sa[kDepth], so[kDepth];
                uint32_t i0 = ib + cw;
#pragma unroll
                for (uint32_t u = 0; u < kDepth; ++u) {
                    dq[u] = make_uint4(0u, 0u, 0u, 0u);
                    sa[u] = 0.0f; so[u] = 0.0f;
                    if (i0 + u * kConsumerWarps < pe) dq[u] = __ldg(items + i0 + u * kConsumerWarps);
                }
                if (dq[0].y != 0u) {
                    if (ADD) sa[0] = __ldcg(a.s_add + a.seg_offset + dq[0].z);
                    if (SUB) so[0] = __ldcg(a.s_old + a.seg_offset + dq[0].z);
                }
                for (; i0 < pe; i0 += kDepth * kConsumerWarps) {
#pragma unroll
                    for (uint32_t u = 0; u < kDepth; ++u) {
                        const uint4 d = dq[u];
                        if (d.y == 0u) break;  // past the end of the range (warp-uniform)
                        const float s_add = sa[u], s_old = so[u];
                        dq[u] = make_uint4(0u, 0u, 0u, 0u);
                        if (i0 + (u + kDepth) * kConsumerWarps < pe) dq[u] = __ldg(items + i0 + (u + kDepth) * kConsumerWarps);
                        {
                            const uint32_t n = (u + 1u) % kDepth;  // the next item's scalars (its descriptor arrived long ago)
                            sa[n] = 0.0f; so[n] = 0.0f;
                            if (dq[n].y != 0u) {
                                if (ADD) sa[n] = __ldcg(a.s_add + a.seg_offset + dq[n].z);
                                if (SUB) so[n] = __ldcg(a.s_old + a.seg_offset + dq[n].z);
                            }
                        }
                        const uint32_t len = d.y;
                        const uint32_t rel = d.x - stream0;
                        const uint32_t t0 = rel / kTile, t1 = (rel + len - 1u) / kTile;
                        {   // (a stage stays armed for tile t until t has been consumed completely — which includes this item)
                            uint32_t spins = 0;
                            while (s_armed[t0 & (NST - 1u)] != t0 + 1u)  // the producer has not armed the stage for this tile yet
                                if (++spins > kSpinLimit) __trap();
                            mbar_wait(full_u32 + 8u * (t0 & (NST - 1u)), (t0 / NST) & 1u);
                            if (t1 != t0) {
                                while (s_armed[t1 & (NST - 1u)] != t1 + 1u)
                                    if (++spins > kSpinLimit) __trap();
                                mbar_wait(full_u32 + 8u * (t1 & (NST - 1u)), (t1 / NST) & 1u);
                            }
                        }
                        float* gval = a.val + d.x + lane_off;
                        const uint32_t r0 = rel + lane_off;
                        float g = 0.0f, h = 0.0f;
#pragma unroll 1
                        for (uint32_t o = 0; o < len; o += 128u) {
                            if (o + lane_off < len) {
                                const uint32_t pos = (r0 + o) & rmask;
                                Step e;
                                e.i = *reinterpret_cast<const uint2*>(ring_i + pos);
                                e.v = *reinterpret_cast<const float4*>(ring_v + pos);
                                calc4<MODE>(e, sm_new, sm_add, sm_old, s_add, s_old, g, h);
                                if (WRITE) __stcs(reinterpret_cast<float4*>(gval + o), e.v);
                            }
                        }
                        __syncwarp();  // every lane has read its entries out of the ring
                        if (lane == 0) {
                            // hand the entries back: the warp that completes a tile lets the producer refill the stage
                            // (no fence: a warp's shared-memory loads and this atomic are processed in order, the __syncwarp
                            // covers the other lanes, and a fence here would also wait for the descriptor loads in flight)
                            const uint32_t in0 = (t0 + 1u) * kTile - rel < len ? (t0 + 1u) * kTile - rel : len;
                            {
                                const uint32_t st = t0 & (NST - 1u);
                                const uint32_t tile_total = total - t0 * kTile < kTile ? total - t0 * kTile : kTile;
                                if (atomicAdd(&consumed[st], in0) + in0 == tile_total) {
                                    consumed[st] = 0u;
                                    mbar_arrive(empty_u32 + 8u * st);
                                }
                            }
                            if (t1 != t0) {
                                const uint32_t st = t1 & (NST - 1u), in1 = len - in0;
                                const uint32_t tile_total = total - t1 * kTile < kTile ? total - t1 * kTile : kTile;
                                if (atomicAdd(&consumed[st], in1) + in1 == tile_total) {
                                    consumed[st] = 0u;
                                    mbar_arrive(empty_u32 + 8u * st);
                                }
                            }
                        }
                        if (SOLVE) {
#pragma unroll
                            for (int q = 1; q < 32; q <<= 1) {
                                g += __shfl_xor_sync(kFull, g, q);
                                h += __shfl_xor_sync(kFull, h, q);
                            }
                            if (lane == 0) a.partials[d.w] = make_float2(g, h);
                        }
                    }
                }
            }
            ib = pe;
            ++p;
        }
    }
    if (SOLVE && a.fin.enabled) {
        // grid-wide barrier (monotonic counter; every CTA of the launch is resident), then the CTAs share the segments
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = 0;
        if (!grid_barrier(a.fin.bar, a.fin.bar_target, a.fin.status, &s_abort)) return;
        if (a.fin.lanes == 32)
            finalize_segments<32>(a.fin.nseg, a.fin.slot_ptr, a.partials, a.fin.seg_ptr, a.fin.lambda, a.fin.nmf, a.fin.out,
                                  a.fin.peer_ll, a.fin.vec_off, a.fin.rank, a.fin.nranks, a.fin.epoch);
        else
            finalize_segments<1>(a.fin.nseg, a.fin.slot_ptr, a.partials, a.fin.seg_ptr, a.fin.lambda, a.fin.nmf, a.fin.out,
                                 a.fin.peer_ll, a.fin.vec_off, a.fin.rank, a.fin.nranks, a.fin.epoch);
        if (a.fin.ll != nullptr) ll_unpack_entries(a.fin.ll, a.fin.vec, a.fin.dim, a.fin.own_lo, a.fin.own_hi, a.fin.epoch, a.fin.status);
    }
}

template <int LANES>
__global__ void __launch_bounds__(256) k_finalize(int64_t nseg, const uint32_t* __restrict__ slot_ptr,
                                                  const float2* __restrict__ partials, const uint32_t* __restrict__ seg_ptr,
                                                  float lambda, int nmf, float* __restrict__ out, PushArgs push) {
    // grid-stride over segments (a push launch uses few, large CTAs so that few system-scope fences are needed)
    finalize_segments<LANES>(nseg, slot_ptr, partials, seg_ptr, lambda, nmf, out, push.peer_ll, push.vec_off, push.rank,
                             push.nranks, push.epoch);
    if (push.barrier) {
        // rare (once per outer iteration): everything this rank did before is visible, then publish the epoch
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(push.ticket, 1u);
            if (t == gridDim.x - 1) {
                *push.ticket = 0u;
                __threadfence_system();
                for (int p = 0; p < push.nranks; ++p)
                    if (p != push.rank) {
                        unsigned* f = push.peer_flags[p] + push.rank;
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(push.epoch) : "memory");
                    }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_ll_unpack(const unsigned long long* ll, float* __restrict__ vec, int64_t dim,
                                                   int64_t own_lo, int64_t own_hi, unsigned epoch) {
    ll_unpack_entries(ll, vec, dim, own_lo, own_hi, epoch);
}

// waits until every peer has published `epoch` (or later) in this rank's flag words
__global__ void k_exchange_wait(const unsigned* flags, int rank, int nranks, unsigned epoch) {
    const int p = threadIdx.x;
    if (p < nranks && p != rank) {
        unsigned spins = 0, v = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (++spins > (1u << 27)) __trap();
        }
    }
}

// DIRECT layout: one warp per segment on the caller's arrays (uint32 indices, gathers through L1/L2).
// The simple path: used when a copy is not index-sorted, and as the A/B partner of the panel kernels.
template <int MODE>
__global__ void __launch_bounds__(256) k_direct_sweep(DirectSweepArgs a) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    const int lane = threadIdx.x & 31;
    const float* g_add = ADDSEP ? a.g_add : a.g_new;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < a.nseg; s += nwarps) {
        const uint32_t lo = a.ptr[s], hi = a.ptr[s + 1];
        float g = 0.0f, h = 0.0f, s_add = 0.0f, s_old = 0.0f;
        if (ADD) s_add = a.s_add[a.seg_offset + s];
        if (SUB) s_old = a.s_old[a.seg_offset + s];
#pragma unroll 4
        for (uint32_t e = lo + lane; e < hi; e += 32) {
            const uint32_t i = a.idx[e];
            float x = a.val[e];
            if (SUB) x = __fsub_rn(x, __fmul_rn(a.g_old[i], s_old));
            if (ADD) x = __fadd_rn(x, __fmul_rn(g_add[i], s_add));
            if (SOLVE) {
                const float un = a.g_new[i];
                g = fmaf(un, x, g);
                h = fmaf(un, un, h);
            }
            if (SUB || ADD) a.val[e] = x;
        }
        if (SOLVE) {
            g = warp_sum(g);
            h = warp_sum(h);
            if (lane == 0) {
                const uint32_t deg = hi - lo;
                float r = deg ? g / (a.lambda * deg + h) : 0.0f;
                if (a.nmf) r = fmaxf(r, 0.0f);
                a.out[s] = r;
            }
        }
    }
}

// kernel attributes are per device: remember per (template instance, device) whether they have been set
struct PerDeviceOnce {
    unsigned long long mask = 0;  // devices 0..63
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        if (mask >> dev & 1ull) return false;
        mask |= 1ull << dev;
        return true;
    }
};

template <int MODE, bool SHORT>
int launch_panel_v(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;  // per template instance
    if (once.need()) {
        MF_CUDA(cudaFuncSetAttribute(k_panel_sweep<MODE, SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    }
    static const bool pdl = getenv("MF_NO_PDL") == nullptr;
    // MF_COOP_LAUNCH=1: the sweeps that end in the in-kernel finalize (grid barrier) are launched cooperatively, so that the
    // driver itself guarantees — or refuses — the co-residency of the grid instead of the occupancy check at session
    // creation plus the bounded wait.  Opt-in: see profiles/README.md for what it costs.
    static const bool coop = getenv("MF_COOP_LAUNCH") != nullptr && atoi(getenv("MF_COOP_LAUNCH")) != 0;
    const bool want_coop = coop && a.fin.enabled;
    if (pdl || want_coop) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[2];
        unsigned n = 0;
        if (pdl && !want_coop) {  // (a cooperative grid starts only when all of it fits: nothing to overlap with its predecessor)
            at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[n].val.programmaticStreamSerializationAllowed = 1;
            ++n;
        }
        if (want_coop) {
            at[n].id = cudaLaunchAttributeCooperative;
            at[n].val.cooperative = 1;
            ++n;
        }
        cfg.attrs = at; cfg.numAttrs = n;
        MF_CUDA(cudaLaunchKernelEx(&cfg, k_panel_sweep<MODE, SHORT>, a));
    } else {
        k_panel_sweep<MODE, SHORT><<<ncta, threads, smem, st>>>(a);
    }
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int MODE>
int launch_panel(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    return a.short_items ? launch_panel_v<MODE, true>(a, ncta, threads, smem, st) : launch_panel_v<MODE, false>(a, ncta, threads, smem, st);
}

template <int MODE>
int launch_panel_tma(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.need()) {
        MF_CUDA(cudaFuncSetAttribute(k_panel_sweep_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    }
    k_panel_sweep_tma<MODE><<<ncta, threads, smem, st>>>(a);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int MODE>
int launch_panel_async(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.need()) {
        MF_CUDA(cudaFuncSetAttribute(k_panel_sweep_async<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    }
    k_panel_sweep_async<MODE><<<ncta, threads, smem, st>>>(a);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int MODE>
int launch_panel_stream(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.need()) {
        MF_CUDA(cudaFuncSetAttribute(k_panel_sweep_stream<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    }
    k_panel_sweep_stream<MODE><<<ncta, stream::kThreads, smem, st>>>(a);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int MODE>
int launch_direct(const DirectSweepArgs& a, int sm_count, cudaStream_t st) {
    int64_t blocks = (a.nseg + 7) / 8;
    int64_t cap = (int64_t)sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_direct_sweep<MODE><<<(unsigned)blocks, 256, 0, st>>>(a);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

}  // namespace

bool ccd_persistent_supported(int ncta, size_t smem, int device) {
    int coop = 0, sms = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) { cudaGetLastError(); return false; }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); return false; }
    if (smem > 227 * 1024 - 256) return false;
    if (cudaFuncSetAttribute(k_ccd_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256) != cudaSuccess) { cudaGetLastError(); return false; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ccd_persistent, kSweepThreads, smem) != cudaSuccess) { cudaGetLastError(); return false; }
    return (int64_t)per_sm * sms >= ncta;
}

int ccd_persistent_launch(const PersistArgs& a, int ncta, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.need()) MF_CUDA(cudaFuncSetAttribute(k_ccd_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    // the arguments go to __constant__ memory, stream-ordered before the launch (the caller synchronises the stream before
    // its next launch and the library is not re-entrant, so the symbol is never rewritten under a running kernel)
    MF_CUDA(cudaMemcpyToSymbolAsync(c_P, &a, sizeof(PersistArgs), 0, cudaMemcpyHostToDevice, st));
    // cooperative launch: the driver guarantees that all `ncta` CTAs are resident together (or fails the launch), which
    // is what the in-kernel grid barrier needs
    MF_CUDA(cudaLaunchCooperativeKernel((const void*)k_ccd_persistent, dim3((unsigned)ncta), dim3(kSweepThreads), nullptr, smem, st));
    return MF_OK;
}

int panel_timeout_report(char* buf, size_t n) {
    unsigned int h[8] = {0};
    if (cudaMemcpyFromSymbol(h, tma::g_timeout, sizeof(h)) != cudaSuccess) return 0;
    if (h[0] == 0) return 0;
    snprintf(buf, n, "pipeline timeout: code %u (1 producer/empty, 2 consumer/armed, 3 consumer/full) block %u warp %u lane %u item %u slot %u round %u seen %u",
             h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    unsigned int z[8] = {0};
    cudaMemcpyToSymbol(tma::g_timeout, z, sizeof(z));
    return 1;
}

int panel_sweep_threads() { return kSweepThreads; }

int panel_sweep_vectors(int mode) {
    int n = 0;
    if ((mode & kSolve) || ((mode & kAdd) && !(mode & kAddSep))) ++n;
    if ((mode & kAdd) && (mode & kAddSep)) ++n;
    if (mode & kSub) ++n;
    return n;
}

size_t panel_sweep_smem(int mode, int panel_rows) {
    return (size_t)panel_sweep_vectors(mode) * (size_t)(panel_rows + 8) * sizeof(float);
}

// STREAM pipeline: the ring gets what the staged vectors leave: the largest power of two of entries (6 bytes each), at
// most 16 tiles, at least 2 tiles and at least two work items
static size_t stream_overhead(uint32_t ring) { return ((size_t)ring * 6u + (size_t)(ring / stream::kTile) * 20u + 8u + 127u) & ~(size_t)127u; }
uint32_t panel_stream_ring(int mode, int panel_rows, int chunk) {
    const size_t cap = 227 * 1024 - 256, vec = panel_sweep_smem(mode, panel_rows);
    if (chunk > (int)stream::kTile) return 0;
    for (uint32_t ring = 16u * stream::kTile; ring >= 2u * stream::kTile; ring >>= 1)
        if (stream_overhead(ring) + vec <= cap) return ring;
    return 0;
}
size_t panel_stream_smem(int mode, int panel_rows, uint32_t ring) { return stream_overhead(ring) + panel_sweep_smem(mode, panel_rows); }

int panel_sweep(int mode, const PanelSweepArgs& a_in, int ncta, int threads, int chunk, int pipeline, cudaStream_t st) {
    PanelSweepArgs a = a_in;
    const size_t vec = panel_sweep_smem(mode, (int)a.panel_rows);
    const size_t cap = 227 * 1024 - 256;
    if (vec > cap) {
        set_error("panel sweep mode %d needs %zu bytes of shared memory (panel_rows=%u)", mode, vec, a.panel_rows);
        return MF_ERR_ARG;
    }
#define MF_DISPATCH(LAUNCH)                                                                                         \
    switch (mode) {                                                                                                 \
        case kSolve: return LAUNCH<kSolve>(a, ncta, threads, smem, st);                                             \
        case kSub: return LAUNCH<kSub>(a, ncta, threads, smem, st);                                                 \
        case kAdd: return LAUNCH<kAdd>(a, ncta, threads, smem, st);                                                 \
        case kSub | kSolve: return LAUNCH<kSub | kSolve>(a, ncta, threads, smem, st);                               \
        case kAdd | kSolve: return LAUNCH<kAdd | kSolve>(a, ncta, threads, smem, st);                               \
        case kSub | kAdd | kSolve: return LAUNCH<kSub | kAdd | kSolve>(a, ncta, threads, smem, st);                 \
        case kAdd | kAddSep | kSolve: return LAUNCH<kAdd | kAddSep | kSolve>(a, ncta, threads, smem, st);           \
        case kSub | kAdd | kAddSep | kSolve: return LAUNCH<kSub | kAdd | kAddSep | kSolve>(a, ncta, threads, smem, st); \
        default: set_error("panel sweep: unsupported mode %d", mode); return MF_ERR_ARG;                            \
    }
    if (pipeline == MF_PIPELINE_STREAM) {
        const uint32_t ring = panel_stream_ring(mode, (int)a.panel_rows, chunk);
        if (ring == 0) {
            set_error("stream pipeline: no room for the ring beside %d staged vectors of %u entries (chunk %d)", panel_sweep_vectors(mode), a.panel_rows, chunk);
            return MF_ERR_ARG;
        }
        a.ring_entries = ring;
        const size_t smem = panel_stream_smem(mode, (int)a.panel_rows, ring);
        MF_DISPATCH(launch_panel_stream)
    }
    if (pipeline != MF_PIPELINE_REGISTERS && chunk <= (int)tma::kChunkMax) {  // experimental slot-ring pipelines
        // the slot ring gets whatever shared memory the panel vectors leave (at most 64 slots); per slot: data, two
        // mbarriers, one round word
        const size_t vec_al = (vec + 127) & ~(size_t)127;
        const size_t per_slot = tma::kSlotBytes + 16 + 4;
        size_t ns = (cap - vec_al) / per_slot;
        if (ns > 64) ns = 64;
        if (ns >= 40) {  // fewer slots could deadlock the producers against their own un-signalled batches
            a.nslots = (uint32_t)ns;
            a.fin.enabled = 0;
            const size_t smem = vec_al + ns * per_slot;
            if (pipeline == MF_PIPELINE_TMA_BULK) { MF_DISPATCH(launch_panel_tma) }
            MF_DISPATCH(launch_panel_async)
        }
    }
#undef MF_DISPATCH
    const size_t smem = vec;
    switch (mode) {
        case kSolve: return launch_panel<kSolve>(a, ncta, threads, smem, st);
        case kSub: return launch_panel<kSub>(a, ncta, threads, smem, st);
        case kAdd: return launch_panel<kAdd>(a, ncta, threads, smem, st);
        case kSub | kSolve: return launch_panel<kSub | kSolve>(a, ncta, threads, smem, st);
        case kAdd | kSolve: return launch_panel<kAdd | kSolve>(a, ncta, threads, smem, st);
        case kSub | kAdd | kSolve: return launch_panel<kSub | kAdd | kSolve>(a, ncta, threads, smem, st);
        case kAdd | kAddSep | kSolve: return launch_panel<kAdd | kAddSep | kSolve>(a, ncta, threads, smem, st);
        case kSub | kAdd | kAddSep | kSolve: return launch_panel<kSub | kAdd | kAddSep | kSolve>(a, ncta, threads, smem, st);
        default: set_error("panel sweep: unsupported mode %d", mode); return MF_ERR_ARG;
    }
}

// Can `ncta` CTAs of the register-ring sweep be resident at once (the in-kernel finalize needs a grid barrier)?
bool panel_sweep_grid_resident(int ncta, int threads, int panel_rows, int sm_count) {
    const size_t smem = panel_sweep_smem(kSolve | kSub | kAdd | kAddSep, panel_rows);  // the largest footprint of any mode
    if (smem > 227 * 1024 - 256) return false;
    if (cudaFuncSetAttribute(k_panel_sweep<kSolve | kSub | kAdd | kAddSep, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_panel_sweep<kSolve | kSub | kAdd | kAddSep, false>, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return (int64_t)per_sm * sm_count >= ncta;
}

int panel_finalize_lanes(int64_t nseg, int64_t nslots) { return nslots > 4 * nseg ? 32 : 1; }

int panel_finalize(int64_t nseg, int64_t nslots, const uint32_t* slot_ptr, const float2* partials, const uint32_t* seg_ptr,
                   float lambda, int nmf, float* out, const FinalizePush* fp, cudaStream_t st) {
    PushArgs push;
    push.peer_ll = nullptr; push.peer_flags = nullptr; push.ticket = nullptr; push.vec_off = 0; push.rank = 0; push.nranks = 1;
    push.epoch = 0; push.barrier = 0;
    if (fp) {
        push.peer_ll = fp->barrier ? nullptr : fp->peer_ll; push.peer_flags = fp->peer_flags; push.ticket = fp->ticket;
        push.vec_off = fp->vec_off; push.rank = fp->rank; push.nranks = fp->nranks; push.epoch = fp->epoch; push.barrier = fp->barrier;
    }
    int64_t nthreads = nslots > 4 * nseg ? nseg * 32 : nseg;
    if (nthreads < 1) nthreads = 1;
    if (nseg <= 0 && !push.barrier) return MF_OK;
    const int64_t blocks = (nthreads + 255) / 256;
    if (nslots > 4 * nseg)  // many slots per segment (long columns cut by panels and chunks): a warp per segment
        k_finalize<32><<<(unsigned)blocks, 256, 0, st>>>(nseg, slot_ptr, partials, seg_ptr, lambda, nmf, out, push);
    else
        k_finalize<1><<<(unsigned)blocks, 256, 0, st>>>(nseg, slot_ptr, partials, seg_ptr, lambda, nmf, out, push);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int exchange_unpack(const unsigned long long* ll, float* vec, int64_t dim, int64_t own_lo, int64_t own_hi, unsigned epoch,
                    cudaStream_t st) {
    const int64_t nother = dim - (own_hi - own_lo);
    if (nother <= 0) return MF_OK;
    int64_t blocks = (nother + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_ll_unpack<<<(unsigned)blocks, 256, 0, st>>>(ll, vec, dim, own_lo, own_hi, epoch);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int exchange_wait(const unsigned* flags, int rank, int nranks, unsigned epoch, cudaStream_t st) {
    k_exchange_wait<<<1, 32 * ((nranks + 31) / 32), 0, st>>>(flags, rank, nranks, epoch);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int direct_sweep(int mode, const DirectSweepArgs& a, int sm_count, cudaStream_t st) {
    if (a.nseg <= 0) return MF_OK;
    switch (mode) {
        case kSolve: return launch_direct<kSolve>(a, sm_count, st);
        case kSub: return launch_direct<kSub>(a, sm_count, st);
        case kAdd: return launch_direct<kAdd>(a, sm_count, st);
        case kSub | kSolve: return launch_direct<kSub | kSolve>(a, sm_count, st);
        case kAdd | kSolve: return launch_direct<kAdd | kSolve>(a, sm_count, st);
        case kSub | kAdd | kSolve: return launch_direct<kSub | kAdd | kSolve>(a, sm_count, st);
        case kAdd | kAddSep | kSolve: return launch_direct<kAdd | kAddSep | kSolve>(a, sm_count, st);
        case kSub | kAdd | kAddSep | kSolve: return launch_direct<kSub | kAdd | kAddSep | kSolve>(a, sm_count, st);
        default: set_error("direct sweep: unsupported mode %d", mode); return MF_ERR_ARG;
    }
}

}  // namespace mf
