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
// vector(s) in shared memory, then its warps pull batches of four items from a shared-memory counter.
// Four short items (<= 64 entries) are handled at once by the four 8-lane groups of the warp;
// otherwise the warp walks the items one after another with all 32 lanes, 256 entries per step,
// two steps in flight.  Ratings are streamed with 16-byte loads (8 x uint16 indices, 2 x float4
// values) and 16-byte stores; the factor gathers never leave shared memory.
#include "ccd_kernels.cuh"

namespace mf {
namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Elems {
    uint4 i;        // 8 x uint16 panel-local indices
    float4 a, b;    // 8 values
};

__device__ __forceinline__ Elems load8(const uint16_t* __restrict__ idx16, const float* __restrict__ val, uint32_t pos) {
    Elems e;
    e.i = __ldcs(reinterpret_cast<const uint4*>(idx16 + pos));
    e.a = __ldcs(reinterpret_cast<const float4*>(val + pos));
    e.b = __ldcs(reinterpret_cast<const float4*>(val + pos + 4));
    return e;
}

template <int MODE>
__device__ __forceinline__ void calc8(Elems& e, const float* __restrict__ sm_new, const float* __restrict__ sm_add,
                                      const float* __restrict__ sm_old, float s_add, float s_old, float& g, float& h) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve;
    float v[8] = {e.a.x, e.a.y, e.a.z, e.a.w, e.b.x, e.b.y, e.b.z, e.b.w};
    const uint32_t w[4] = {e.i.x, e.i.y, e.i.z, e.i.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t i = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
        float x = v[j];
        if (SUB) x = __fsub_rn(x, __fmul_rn(sm_old[i], s_old));
        if (ADD) x = __fadd_rn(x, __fmul_rn(sm_add[i], s_add));
        if (SOLVE) {
            const float un = sm_new[i];
            g = fmaf(un, x, g);
            h = fmaf(un, un, h);
        }
        v[j] = x;
    }
    if (SUB || ADD) {
        e.a = make_float4(v[0], v[1], v[2], v[3]);
        e.b = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__device__ __forceinline__ void store8(float* __restrict__ val, uint32_t pos, const Elems& e) {
    __stcs(reinterpret_cast<float4*>(val + pos), e.a);
    __stcs(reinterpret_cast<float4*>(val + pos + 4), e.b);
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_panel_sweep(PanelSweepArgs a) {
    constexpr bool SUB = MODE & kSub, ADD = MODE & kAdd, SOLVE = MODE & kSolve, ADDSEP = MODE & kAddSep;
    constexpr bool WRITE = SUB || ADD;
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned s_ctr;

    const uint32_t PR = a.panel_rows;
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
    const float* g_add = ADDSEP ? a.g_add : a.g_new;

    const int lane = threadIdx.x & 31;
    const int grp = lane >> 3, sl = lane & 7;
    const uint4* __restrict__ items = reinterpret_cast<const uint4*>(a.items);

    uint32_t ib = a.cta_item_ptr[blockIdx.x];
    const uint32_t ie = a.cta_item_ptr[blockIdx.x + 1];
    int p = 0;
    while (p < a.npanels && a.panel_item_ptr[p + 1] <= ib) ++p;

    while (ib < ie && p < a.npanels) {
        const uint32_t pend = a.panel_item_ptr[p + 1];
        const uint32_t pe = ie < pend ? ie : pend;
        if (pe > ib) {
            __syncthreads();  // every warp is done with the previous panel and counter
            const int64_t base = (int64_t)p * PR;
            const uint32_t cnt = (uint32_t)((a.gdim - base) < (int64_t)PR ? (a.gdim - base) : (int64_t)PR);
            for (uint32_t i = threadIdx.x; i < stride; i += blockDim.x) {
                const bool in = i < cnt;
                if (NEEDNEW) sm_new[i] = in ? a.g_new[base + i] : 0.0f;
                if (ADD && ADDSEP) sm_add[i] = in ? g_add[base + i] : 0.0f;
                if (SUB) sm_old[i] = in ? a.g_old[base + i] : 0.0f;
            }
            if (threadIdx.x == 0) s_ctr = ib;
            __syncthreads();

            for (;;) {
                uint32_t i0 = 0;
                if (lane == 0) i0 = atomicAdd(&s_ctr, 4u);
                i0 = __shfl_sync(kFull, i0, 0);
                if (i0 >= pe) break;
                const uint32_t mine = i0 + grp;
                uint4 d = make_uint4(0u, 0u, 0u, 0u);  // {start, len, seg, slot}; len 0 = no item
                if (mine < pe) d = __ldg(items + mine);
                const bool all_short = __all_sync(kFull, d.y <= 64u);
                if (all_short) {
                    // four items at once, 8 lanes x 8 entries each
                    float g = 0.0f, h = 0.0f, s_add = 0.0f, s_old = 0.0f;
                    const uint32_t off = (uint32_t)sl * 8u;
                    if (off < d.y) {
                        if (ADD) s_add = __ldg(a.s_add + a.seg_offset + d.z);
                        if (SUB) s_old = __ldg(a.s_old + a.seg_offset + d.z);
                        Elems e = load8(a.idx16, a.val, d.x + off);
                        calc8<MODE>(e, sm_new, sm_add, sm_old, s_add, s_old, g, h);
                        if (WRITE) store8(a.val, d.x + off, e);
                    }
                    if (SOLVE) {
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) {
                            g += __shfl_xor_sync(kFull, g, o);
                            h += __shfl_xor_sync(kFull, h, o);
                        }
                        if (sl == 0 && d.y != 0u) a.partials[d.w] = make_float2(g, h);
                    }
                } else {
                    // one item at a time with the whole warp
#pragma unroll 1
                    for (int gi = 0; gi < 4; ++gi) {
                        const uint32_t start = __shfl_sync(kFull, d.x, gi * 8);
                        const uint32_t len = __shfl_sync(kFull, d.y, gi * 8);
                        const uint32_t seg = __shfl_sync(kFull, d.z, gi * 8);
                        const uint32_t slot = __shfl_sync(kFull, d.w, gi * 8);
                        if (len == 0u) continue;
                        float g = 0.0f, h = 0.0f, s_add = 0.0f, s_old = 0.0f;
                        if (ADD) s_add = __ldg(a.s_add + a.seg_offset + seg);
                        if (SUB) s_old = __ldg(a.s_old + a.seg_offset + seg);
#pragma unroll 1
                        for (uint32_t off = (uint32_t)lane * 8u; off < len; off += 512u) {
                            const bool two = off + 256u < len;
                            Elems e0 = load8(a.idx16, a.val, start + off);
                            Elems e1;
                            if (two) e1 = load8(a.idx16, a.val, start + off + 256u);
                            calc8<MODE>(e0, sm_new, sm_add, sm_old, s_add, s_old, g, h);
                            if (WRITE) store8(a.val, start + off, e0);
                            if (two) {
                                calc8<MODE>(e1, sm_new, sm_add, sm_old, s_add, s_old, g, h);
                                if (WRITE) store8(a.val, start + off + 256u, e1);
                            }
                        }
                        if (SOLVE) {
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                g += __shfl_xor_sync(kFull, g, o);
                                h += __shfl_xor_sync(kFull, h, o);
                            }
                            if (lane == 0) a.partials[slot] = make_float2(g, h);
                        }
                    }
                }
            }
        }
        ib = pe;
        ++p;
    }
}

// one thread per segment: add the segment's slots in order, apply the regulariser, divide.
// h = lambda*deg + sum u^2 with lambda*deg a float*unsigned product as at src/CCD.cpp:112,120;
// empty segment -> 0 (src/CCD.cpp:8).
__global__ void k_finalize(int64_t nseg, const uint32_t* __restrict__ slot_ptr, const float2* __restrict__ partials,
                           const uint32_t* __restrict__ seg_ptr, float lambda, int nmf, float* __restrict__ out) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t deg = seg_ptr[s + 1] - seg_ptr[s];
    float r = 0.0f;
    if (deg != 0u) {
        float g = 0.0f, h = 0.0f;
        for (uint32_t q = slot_ptr[s]; q < slot_ptr[s + 1]; ++q) {
            const float2 pr = partials[q];
            g += pr.x;
            h += pr.y;
        }
        r = g / (lambda * deg + h);
        if (nmf) r = fmaxf(r, 0.0f);
    }
    out[s] = r;
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

template <int MODE>
int launch_panel(const PanelSweepArgs& a, int ncta, int threads, size_t smem, cudaStream_t st) {
    static bool attr_set = false;  // per template instance
    if (!attr_set) {
        MF_CUDA(cudaFuncSetAttribute(k_panel_sweep<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
        attr_set = true;
    }
    k_panel_sweep<MODE><<<ncta, threads, smem, st>>>(a);
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

int panel_sweep(int mode, const PanelSweepArgs& a, int ncta, int threads, cudaStream_t st) {
    const size_t smem = panel_sweep_smem(mode, (int)a.panel_rows);
    if (smem > 227 * 1024 - 64) {
        set_error("panel sweep mode %d needs %zu bytes of shared memory (panel_rows=%u)", mode, smem, a.panel_rows);
        return MF_ERR_ARG;
    }
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

int panel_finalize(int64_t nseg, const uint32_t* slot_ptr, const float2* partials, const uint32_t* seg_ptr, float lambda,
                   int nmf, float* out, cudaStream_t st) {
    if (nseg <= 0) return MF_OK;
    k_finalize<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(nseg, slot_ptr, partials, seg_ptr, lambda, nmf, out);
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
