// prep.cu — integer tier: builds the PANEL layout (layout.cuh) from a caller-order compressed
// copy, converts values between the two orders, and the small index ops (degree bins, nnz-balanced
// partition).  Everything here is exact integer / copy work and is checked bit-for-bit against a
// numpy restatement in tests/ (the reference itself has no builder: it loads ready-made CSR and CSC,
// src/tools.cpp:58-59, src/pmf_util.h:108-136).
#include "layout.cuh"

#include <algorithm>
#include <thread>
#include <vector>

namespace mf {
namespace {

constexpr int kPad = 8;        // unit of lengths in the bins / cost model; every padded length is a multiple of it
constexpr uint32_t kCost0 = 8; // per-item fixed cost in units of 8 entries (descriptor, reduce, store)

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t lo, uint32_t hi,
                                                    uint32_t key) {
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per piece q = p*nseg + s
__global__ void k_piece_count(int64_t nseg, int npanels, uint32_t panel_rows, uint32_t chunk, uint32_t pad,
                              const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx,
                              uint32_t* __restrict__ piece_first, uint32_t* __restrict__ padded,
                              uint32_t* __restrict__ nitem) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t Q = nseg * npanels;
    if (q >= Q) return;
    int p = (int)(q / nseg);
    int64_t s = q - (int64_t)p * nseg;
    uint32_t lo = ptr[s], hi = ptr[s + 1];
    uint64_t k0 = (uint64_t)p * panel_rows, k1 = k0 + panel_rows;
    uint32_t first = lower_bound_u32(idx, lo, hi, (uint32_t)k0);
    uint32_t last = k1 > 0xffffffffull ? hi : lower_bound_u32(idx, first, hi, (uint32_t)k1);
    uint32_t cnt = last - first;
    uint32_t pd = (cnt + pad - 1) / pad * pad;
    piece_first[q] = first;
    padded[q] = pd;
    nitem[q] = (pd + chunk - 1) / chunk;
}

// one thread per segment: how many work items (= partial slots) the segment owns
__global__ void k_seg_items(int64_t nseg, int npanels, const uint32_t* __restrict__ item_ptr,
                            uint32_t* __restrict__ seg_items) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint32_t n = 0;
    for (int p = 0; p < npanels; ++p) {
        int64_t q = (int64_t)p * nseg + s;
        n += item_ptr[q + 1] - item_ptr[q];
    }
    seg_items[s] = n;
}

// one warp per piece: emit the piece's work items (provisional start = the piece-order position; the storage
// position is assigned after the items have been put in work-list order, k_item_set_start)
__global__ void k_items(int64_t nseg, int npanels, uint32_t chunk, const uint32_t* __restrict__ piece_ptr,
                        const uint32_t* __restrict__ item_ptr, const uint32_t* __restrict__ slot_ptr,
                        WorkItem* __restrict__ items, uint32_t* __restrict__ item_cost) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t Q = nseg * npanels;
    for (int64_t q = warp; q < Q; q += nwarps) {
        uint32_t dst0 = piece_ptr[q];
        uint32_t pd = piece_ptr[q + 1] - dst0;
        if (pd == 0) continue;
        int p = (int)(q / nseg);
        int64_t s = q - (int64_t)p * nseg;
        // slots of segment s are ordered (panel, chunk): count the items of the earlier panels
        uint32_t before = 0;
        for (int pp = lane; pp < p; pp += 32) {
            int64_t qq = (int64_t)pp * nseg + s;
            before += item_ptr[qq + 1] - item_ptr[qq];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        uint32_t it0 = item_ptr[q], nit = item_ptr[q + 1] - it0;
        uint32_t slot0 = slot_ptr[s] + before;
        for (uint32_t j = lane; j < nit; j += 32) {
            WorkItem w;
            w.start = dst0 + j * chunk;
            uint32_t rem = pd - j * chunk;
            w.len = rem < chunk ? rem : chunk;
            w.seg = (uint32_t)s;
            w.slot = slot0 + j;
            items[it0 + j] = w;
            item_cost[it0 + j] = w.len / kPad + kCost0;
        }
    }
}

// STREAM order (layout.cuh): the padded entries are stored in WORK-LIST order — item i of the final list starts at
// the sum of the lengths of the items before it — so that the range of items a CTA walks is one contiguous stretch of
// the index and value arrays (fetched with a few large bulk copies instead of one gather per item).
__global__ void k_item_len(int64_t nitems, const WorkItem* __restrict__ items, uint32_t* __restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nitems) len[i] = items[i].len;
}
__global__ void k_item_set_start(int64_t nitems, const uint32_t* __restrict__ start, WorkItem* __restrict__ items) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nitems) items[i].start = start[i];
}

// one warp per piece: copy + pad the entries of each of its work items to the item's place in the stream
// (perm[original item] = position in the final list)
__global__ void k_fill_entries(int64_t nseg, int npanels, uint32_t panel_rows, uint32_t chunk,
                               const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx,
                               const float* __restrict__ val, const uint32_t* __restrict__ piece_first,
                               const uint32_t* __restrict__ piece_ptr, const uint32_t* __restrict__ item_ptr,
                               const uint32_t* __restrict__ perm, const WorkItem* __restrict__ items,
                               uint16_t* __restrict__ idx16, float* __restrict__ pval) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t Q = nseg * npanels;
    for (int64_t q = warp; q < Q; q += nwarps) {
        uint32_t pd = piece_ptr[q + 1] - piece_ptr[q];
        if (pd == 0) continue;
        int p = (int)(q / nseg);
        int64_t s = q - (int64_t)p * nseg;
        uint32_t src0 = piece_first[q];
        uint32_t src1 = (p == npanels - 1) ? ptr[s + 1] : piece_first[q + nseg];
        uint32_t cnt = src1 - src0;
        uint32_t base = (uint32_t)p * panel_rows;
        uint32_t it0 = item_ptr[q], nit = item_ptr[q + 1] - it0;
        for (uint32_t j = 0; j < nit; ++j) {
            const uint32_t dst0 = items[perm[it0 + j]].start;
            const uint32_t e0 = j * chunk, e1 = e0 + chunk < pd ? e0 + chunk : pd;
            for (uint32_t e = e0 + lane; e < e1; e += 32) {
                bool real = e < cnt;
                // stored pre-multiplied by 4: the byte offset of the factor entry inside the shared-memory panel
                idx16[dst0 + e - e0] = (uint16_t)((real ? idx[src0 + e] - base : panel_rows) << 2);
                pval[dst0 + e - e0] = real ? val[src0 + e] : 0.0f;
            }
        }
    }
}

// Degree-binned order: inside every panel the work items are ranked longest-first (bin = len/8, counting
// sort with per-bin cursors; the order inside one bin is arbitrary) and then DEALT round-robin into
// kDealLanes lanes that are stored one after the other: item of rank k goes to lane k % kDealLanes.  So
//   * consecutive items of a lane are neighbours-at-distance-kDealLanes in the length ranking: the four
//     items a warp processes together have nearly the same length;
//   * every contiguous stretch of the list — in particular every CTA's equal-cost range — holds the same
//     mix of long and short items, so errors of the cost model cancel instead of piling up on some CTAs.
// The order only affects which warp picks an item up, never a result (every item writes its own slot).
constexpr uint32_t kDealLanes = 148;
__device__ __forceinline__ uint32_t deal_position(uint32_t k, uint32_t n) {
    const uint32_t lane = k % kDealLanes, r = k / kDealLanes;
    const uint32_t q = n / kDealLanes, rem = n % kDealLanes;
    return lane * q + (lane < rem ? lane : rem) + r;
}
__device__ __forceinline__ int panel_of_item(const uint32_t* __restrict__ panel_item_ptr, int npanels, uint32_t i) {
    int lo = 0, hi = npanels;  // last p with panel_item_ptr[p] <= i
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (panel_item_ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}
__global__ void k_item_bin_count(int64_t nitems, int npanels, uint32_t nbins, const uint32_t* __restrict__ panel_item_ptr,
                                 const WorkItem* __restrict__ items, uint32_t* __restrict__ bin_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const int p = panel_of_item(panel_item_ptr, npanels, (uint32_t)i);
    atomicAdd(&bin_count[(uint32_t)p * nbins + (nbins - 1u - items[i].len / kPad)], 1u);
}
__global__ void k_item_bin_scatter(int64_t nitems, int npanels, uint32_t nbins, const uint32_t* __restrict__ panel_item_ptr,
                                   const WorkItem* __restrict__ items, const uint32_t* __restrict__ bin_ptr,
                                   uint32_t* __restrict__ bin_cursor, WorkItem* __restrict__ sorted, uint32_t* __restrict__ cost,
                                   uint32_t* __restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const WorkItem w = items[i];
    const int p = panel_of_item(panel_item_ptr, npanels, (uint32_t)i);
    const uint32_t b = (uint32_t)p * nbins + (nbins - 1u - w.len / kPad);
    const uint32_t rank = bin_ptr[b] + atomicAdd(&bin_cursor[b], 1u);  // longest-first rank, global index
    const uint32_t pbeg = panel_item_ptr[p];
    const uint32_t dst = pbeg + deal_position(rank - pbeg, panel_item_ptr[p + 1] - pbeg);
    sorted[dst] = w;
    cost[dst] = w.len / kPad + kCost0;
    perm[i] = dst;
}

__global__ void k_panel_item_ptr(int64_t nseg, int npanels, const uint32_t* __restrict__ item_ptr,
                                 uint32_t* __restrict__ panel_item_ptr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= npanels) panel_item_ptr[p] = item_ptr[(int64_t)p * nseg];
}

// What a CTA pays when its range runs into another panel (all warps drain, the panel's vectors are staged, the pipeline
// refills) is charged to the first item of every panel but the first, so that the equal-cost cut hands the CTAs that
// span a panel boundary fewer items.  Measured on the Netflix shape (per-CTA time stamps, profiles/README.md round 2):
// those CTAs — one in five on the 30-panel CSC side — reached the grid barrier 7 us (solve sweep) to 22 us (fused sweep)
// after the others: 8-17 % of a single-GPU sweep and half of an 8-GPU one.
__global__ void k_panel_entry_cost(int npanels, const uint32_t* __restrict__ panel_item_ptr, uint32_t extra, uint32_t* __restrict__ cost) {
    int p = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (p < npanels && panel_item_ptr[p + 1] > panel_item_ptr[p]) cost[panel_item_ptr[p]] += extra;
}

// equal-cost contiguous item ranges: cta_item_ptr[j] = first item whose cost prefix >= j*total/ncta
__global__ void k_cta_ranges(int ncta, int64_t nitems, const uint32_t* __restrict__ cost_prefix,
                             uint32_t* __restrict__ cta_item_ptr) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ncta) return;
    if (j == ncta) { cta_item_ptr[j] = (uint32_t)nitems; return; }
    uint64_t total = cost_prefix[nitems];
    uint32_t target = (uint32_t)(total * (uint64_t)j / (uint64_t)ncta);
    cta_item_ptr[j] = lower_bound_u32(cost_prefix, 0, (uint32_t)nitems, target);
}

// first padded entry of every CTA's item range (STREAM pipeline: a CTA's items are one contiguous stretch)
__global__ void k_cta_starts(int ncta, int64_t nitems, uint32_t npad, const uint32_t* __restrict__ cta_item_ptr,
                             const WorkItem* __restrict__ items, uint32_t* __restrict__ cta_start_ptr) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ncta) return;
    const uint32_t i = cta_item_ptr[j];
    cta_start_ptr[j] = i < (uint32_t)nitems ? items[i].start : npad;
}

// one warp per piece: value copy between the stream order and the caller's order
template <bool TO_RAW>
__global__ void k_copy_values(int64_t nseg, int npanels, uint32_t chunk, const uint32_t* __restrict__ ptr,
                              const uint32_t* __restrict__ piece_first, const uint32_t* __restrict__ piece_ptr,
                              const uint32_t* __restrict__ item_ptr, const uint32_t* __restrict__ perm,
                              const WorkItem* __restrict__ items, float* __restrict__ pval, float* __restrict__ raw) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t Q = nseg * npanels;
    for (int64_t q = warp; q < Q; q += nwarps) {
        if (piece_ptr[q + 1] == piece_ptr[q]) continue;
        int p = (int)(q / nseg);
        int64_t s = q - (int64_t)p * nseg;
        uint32_t src0 = piece_first[q];
        uint32_t src1 = (p == npanels - 1) ? ptr[s + 1] : piece_first[q + nseg];
        uint32_t cnt = src1 - src0;
        uint32_t it0 = item_ptr[q], nit = item_ptr[q + 1] - it0;
        for (uint32_t j = 0; j < nit; ++j) {
            const uint32_t e0 = j * chunk;
            if (e0 >= cnt) break;
            const uint32_t dst0 = items[perm[it0 + j]].start;
            const uint32_t e1 = e0 + chunk < cnt ? e0 + chunk : cnt;
            for (uint32_t e = e0 + lane; e < e1; e += 32) {
                if (TO_RAW) raw[src0 + e] = pval[dst0 + e - e0];
                else pval[dst0 + e - e0] = raw[src0 + e];
            }
        }
    }
}

__global__ void k_check_sorted(int64_t nseg, const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx,
                               uint32_t gdim, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nseg; s += nwarps) {
        uint32_t lo = ptr[s], hi = ptr[s + 1];
        bool b = false;
        for (uint32_t e = lo + lane; e < hi; e += 32) {
            uint32_t v = idx[e];
            if (v >= gdim) atomicOr(bad, 2);                       // not an index at all
            if (e + 1 < hi && v >= idx[e + 1]) b = true;           // out of order (or a duplicate)
        }
        if (b) atomicOr(bad, 1);
    }
}

// bad |= 1 when any a[i] >= bound (test pairs, prediction pairs, COO indices)
__global__ void k_check_below(int64_t n, const uint32_t* __restrict__ a, uint32_t bound, int* __restrict__ bad) {
    bool b = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (a[i] >= bound) b = true;
    if (b) *bad = 1;
}

// dst[perm[i]] = src[i]
__global__ void k_scatter_perm(int64_t n, const uint32_t* __restrict__ perm, const float* __restrict__ src, float* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[perm[i]] = src[i];
}

__global__ void k_degree_bins(int64_t nseg, const uint32_t* __restrict__ ptr, unsigned long long* __restrict__ seg_in_bin,
                              unsigned long long* __restrict__ nnz_in_bin) {
    __shared__ unsigned long long sa[33], sb[33];
    if (threadIdx.x < 33) { sa[threadIdx.x] = 0; sb[threadIdx.x] = 0; }
    __syncthreads();
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += (int64_t)gridDim.x * blockDim.x) {
        uint32_t d = ptr[s + 1] - ptr[s];
        int b = 32 - __clz(d);  // bit length: 0 for empty, 1 for deg 1, 2 for 2..3, ...
        atomicAdd(&sa[b], 1ull);
        atomicAdd(&sb[b], (unsigned long long)d);
    }
    __syncthreads();
    if (threadIdx.x < 33) {
        if (sa[threadIdx.x]) atomicAdd(&seg_in_bin[threadIdx.x], sa[threadIdx.x]);
        if (sb[threadIdx.x]) atomicAdd(&nnz_in_bin[threadIdx.x], sb[threadIdx.x]);
    }
}

__global__ void k_partition(int64_t nseg, const uint32_t* __restrict__ ptr, int P, long long* __restrict__ bound) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    if (p == 0) { bound[0] = 0; return; }
    if (p == P) { bound[P] = nseg; return; }
    uint64_t nnz = ptr[nseg];
    uint64_t target = (nnz * (uint64_t)p + (uint64_t)P - 1) / (uint64_t)P;
    int64_t lo = 0, hi = nseg;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if ((uint64_t)ptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    bound[p] = lo;
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block > 0 ? (n + block - 1) / block : 1); }

}  // namespace

int side_free(Side& s) {
    void* ptrs[] = {s.ptr, s.idx, s.val, s.piece_ptr, s.piece_first, s.item_ptr, s.idx16, s.pval, s.items,
                    s.slot_ptr, s.partials, s.cta_item_ptr, s.cta_start_ptr, s.panel_item_ptr, s.item_perm, s.unsort_perm, s.als_items, s.als_queue, s.als_counters, s.als_partial};
    for (void* p : ptrs)
        if (p) dev_free(p);
    s = Side();
    return MF_OK;
}

int check_below(const uint32_t* d_a, int64_t n, uint64_t bound, bool* ok, cudaStream_t st) {
    *ok = true;
    if (n <= 0 || bound > 0xffffffffull) return MF_OK;
    int* d_bad = nullptr;
    MF_TRY(dev_alloc(&d_bad, 1));
    MF_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    k_check_below<<<148 * 4, 256, 0, st>>>(n, d_a, (uint32_t)bound, d_bad);
    int h = 0;
    MF_CUDA(cudaMemcpyAsync(&h, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    dev_free(d_bad);
    *ok = h == 0;
    return MF_OK;
}

int scatter_by_perm(const uint32_t* perm, const float* src, float* dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return MF_OK;
    k_scatter_perm<<<148 * 8, 256, 0, st>>>(n, perm, src, dst);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

// Sorts every segment of a caller-order copy by index (the reference never asks for sorted segments, the panel cut does):
// done on the host, one thread per stretch of segments — an input that needs it is rare (the reference's files and this
// repo's builder are sorted) and it is paid once per session.  s.unsort_perm[i] = caller position of sorted entry i, so
// that values can still be returned in the caller's order.
int side_sort_segments(Side& s, cudaStream_t st) {
    const size_t n = (size_t)s.nnz;
    std::vector<uint32_t> ptr((size_t)s.nseg + 1), idx(n), perm(n);
    std::vector<float> val(n);
    MF_CUDA(cudaMemcpyAsync(ptr.data(), s.ptr, sizeof(uint32_t) * ptr.size(), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaMemcpyAsync(idx.data(), s.idx, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaMemcpyAsync(val.data(), s.val, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> idx2(n);
    std::vector<float> val2(n);
    const unsigned nthreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nthreads; ++w)
        pool.emplace_back([&, w]() {
            std::vector<uint32_t> order;
            for (int64_t sg = w; sg < s.nseg; sg += nthreads) {
                const uint32_t lo = ptr[sg], hi = ptr[sg + 1];
                order.resize(hi - lo);
                for (uint32_t e = lo; e < hi; ++e) order[e - lo] = e;
                std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return idx[a] < idx[b]; });
                for (uint32_t e = lo; e < hi; ++e) { idx2[e] = idx[order[e - lo]]; val2[e] = val[order[e - lo]]; perm[e] = order[e - lo]; }
            }
        });
    for (auto& t : pool) t.join();
    MF_TRY(dev_alloc(&s.unsort_perm, n));
    MF_CUDA(cudaMemcpyAsync(s.idx, idx2.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    MF_CUDA(cudaMemcpyAsync(s.val, val2.data(), sizeof(float) * n, cudaMemcpyHostToDevice, st));
    MF_CUDA(cudaMemcpyAsync(s.unsort_perm, perm.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    MF_CUDA(cudaStreamSynchronize(st));
    return MF_OK;
}

int side_check_sorted(const Side& s, bool* sorted, bool* in_range, cudaStream_t st) {
    int* d_bad = nullptr;
    MF_TRY(dev_alloc(&d_bad, 1));
    MF_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (s.nseg > 0 && s.nnz > 0) {
        int64_t warps = s.nseg < 65536 * 8 ? s.nseg : 65536 * 8;
        k_check_sorted<<<grid_for(warps * 32, 256), 256, 0, st>>>(s.nseg, s.ptr, s.idx, (uint32_t)s.gdim, d_bad);
    }
    int h = 0;
    MF_CUDA(cudaMemcpyAsync(&h, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    dev_free(d_bad);
    *sorted = (h & 1) == 0;
    *in_range = (h & 2) == 0;
    return MF_OK;
}

int side_build_panels(Side& s, int panel_rows, int chunk, int ncta, cudaStream_t st) {
    MF_REQUIRE(panel_rows > 0 && panel_rows <= 16376 && panel_rows % 8 == 0, "panel_rows must be a multiple of 8 in (0, 16376]");
    MF_REQUIRE(chunk >= 8 && chunk % 8 == 0, "chunk must be a positive multiple of 8");
    if (s.pad < kPad || s.pad % kPad != 0 || chunk % s.pad != 0) s.pad = kPad;
    MF_REQUIRE(ncta > 0, "ncta must be positive");
    s.panel_rows = panel_rows;
    s.chunk = chunk;
    s.ncta = ncta;
    s.npanels = (int)((s.gdim + panel_rows - 1) / panel_rows);
    if (s.npanels < 1) s.npanels = 1;
    const int64_t Q = s.nseg * s.npanels;
    MF_REQUIRE(Q < (int64_t)1 << 31, "too many (panel, segment) pieces: raise panel_rows");

    uint32_t *padded = nullptr, *nitem = nullptr, *seg_items = nullptr, *tmp = nullptr, *cost = nullptr, *cost_prefix = nullptr;
    MF_TRY(tmp_alloc(&padded, (size_t)Q, st));
    MF_TRY(tmp_alloc(&nitem, (size_t)Q, st));
    MF_TRY(dev_alloc(&s.piece_first, (size_t)Q));
    MF_TRY(dev_alloc(&s.piece_ptr, (size_t)Q + 1));
    MF_TRY(dev_alloc(&s.item_ptr, (size_t)Q + 1));
    MF_TRY(tmp_alloc(&seg_items, (size_t)s.nseg, st));
    MF_TRY(dev_alloc(&s.slot_ptr, (size_t)s.nseg + 1));
    MF_TRY(dev_alloc(&s.panel_item_ptr, (size_t)s.npanels + 1));
    MF_TRY(dev_alloc(&s.cta_item_ptr, (size_t)ncta + 1));
    MF_TRY(dev_alloc(&s.cta_start_ptr, (size_t)ncta + 1));
    size_t tmp_n = scan_tmp_elems((size_t)(Q > s.nseg ? Q : s.nseg));
    MF_TRY(tmp_alloc(&tmp, tmp_n, st));

    trace_mark("    layout: index allocations");
    if (Q > 0)
        k_piece_count<<<grid_for(Q, 256), 256, 0, st>>>(s.nseg, s.npanels, (uint32_t)panel_rows, (uint32_t)chunk, (uint32_t)s.pad, s.ptr,
                                                        s.idx, s.piece_first, padded, nitem);
    MF_CUDA(cudaGetLastError());
    MF_TRY(exclusive_scan_u32(padded, s.piece_ptr, (size_t)Q, tmp, st));
    MF_TRY(exclusive_scan_u32(nitem, s.item_ptr, (size_t)Q, tmp, st));
    if (s.nseg > 0)
        k_seg_items<<<grid_for(s.nseg, 256), 256, 0, st>>>(s.nseg, s.npanels, s.item_ptr, seg_items);
    MF_CUDA(cudaGetLastError());
    MF_TRY(exclusive_scan_u32(seg_items, s.slot_ptr, (size_t)s.nseg, tmp, st));

    uint32_t h_npad = 0, h_nitems = 0, h_nslots = 0;
    MF_CUDA(cudaMemcpyAsync(&h_npad, s.piece_ptr + Q, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaMemcpyAsync(&h_nitems, s.item_ptr + Q, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaMemcpyAsync(&h_nslots, s.slot_ptr + s.nseg, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    s.npad = h_npad;
    s.nitems = h_nitems;
    s.nslots = h_nslots;
    MF_REQUIRE(s.nitems == s.nslots, "internal: item/slot count mismatch (%lld vs %lld)", (long long)s.nitems, (long long)s.nslots);

    trace_mark("    layout: count + scans");
    MF_TRY(dev_alloc(&s.idx16, (size_t)s.npad + 8));
    MF_TRY(dev_alloc(&s.pval, (size_t)s.npad + 8));
    MF_TRY(dev_alloc(&s.items, (size_t)s.nitems));
    MF_TRY(dev_alloc(&s.partials, (size_t)s.nslots));
    MF_TRY(tmp_alloc(&cost, (size_t)s.nitems, st));
    MF_TRY(tmp_alloc(&cost_prefix, (size_t)s.nitems + 1, st));
    trace_mark("    layout: big allocations");
    const int64_t fill_warps = Q < 148 * 64 * 8 ? Q : 148 * 64 * 8;
    if (Q > 0)
        k_items<<<grid_for(fill_warps * 32, 256), 256, 0, st>>>(s.nseg, s.npanels, (uint32_t)chunk, s.piece_ptr, s.item_ptr, s.slot_ptr, s.items, cost);
    MF_CUDA(cudaGetLastError());
    k_panel_item_ptr<<<grid_for(s.npanels + 1, 128), 128, 0, st>>>(s.nseg, s.npanels, s.item_ptr, s.panel_item_ptr);
    MF_CUDA(cudaGetLastError());
    // longest-first inside each panel
    if (s.nitems > 0) {
        const uint32_t nbins = (uint32_t)chunk / kPad + 1u;
        const size_t nb = (size_t)s.npanels * nbins;
        uint32_t *bin_count = nullptr, *bin_ptr = nullptr, *tmp2 = nullptr;
        WorkItem* sorted = nullptr;
        MF_TRY(tmp_alloc(&bin_count, nb, st));
        MF_TRY(tmp_alloc(&bin_ptr, nb + 1, st));
        MF_TRY(tmp_alloc(&tmp2, scan_tmp_elems(nb), st));
        MF_TRY(dev_alloc(&sorted, (size_t)s.nitems));
        MF_TRY(dev_alloc(&s.item_perm, (size_t)s.nitems));
        MF_CUDA(cudaMemsetAsync(bin_count, 0, sizeof(uint32_t) * nb, st));
        k_item_bin_count<<<grid_for(s.nitems, 256), 256, 0, st>>>(s.nitems, s.npanels, nbins, s.panel_item_ptr, s.items, bin_count);
        MF_CUDA(cudaGetLastError());
        MF_TRY(exclusive_scan_u32(bin_count, bin_ptr, nb, tmp2, st));
        MF_CUDA(cudaMemsetAsync(bin_count, 0, sizeof(uint32_t) * nb, st));
        k_item_bin_scatter<<<grid_for(s.nitems, 256), 256, 0, st>>>(s.nitems, s.npanels, nbins, s.panel_item_ptr, s.items, bin_ptr,
                                                                   bin_count, sorted, cost, s.item_perm);
        MF_CUDA(cudaGetLastError());
        MF_CUDA(cudaStreamSynchronize(st));
        dev_free(s.items);
        s.items = sorted;
        tmp_free(bin_count, st); tmp_free(bin_ptr, st); tmp_free(tmp2, st);
        // stream order: an item's entries start where the items before it in the list end
        uint32_t *len = nullptr, *start = nullptr, *tmp3 = nullptr;
        MF_TRY(tmp_alloc(&len, (size_t)s.nitems, st));
        MF_TRY(tmp_alloc(&start, (size_t)s.nitems + 1, st));
        MF_TRY(tmp_alloc(&tmp3, scan_tmp_elems((size_t)s.nitems), st));
        k_item_len<<<grid_for(s.nitems, 256), 256, 0, st>>>(s.nitems, s.items, len);
        MF_CUDA(cudaGetLastError());
        MF_TRY(exclusive_scan_u32(len, start, (size_t)s.nitems, tmp3, st));
        k_item_set_start<<<grid_for(s.nitems, 256), 256, 0, st>>>(s.nitems, start, s.items);
        MF_CUDA(cudaGetLastError());
        k_fill_entries<<<grid_for(fill_warps * 32, 256), 256, 0, st>>>(s.nseg, s.npanels, (uint32_t)panel_rows, (uint32_t)chunk, s.ptr, s.idx,
                                                                      s.val, s.piece_first, s.piece_ptr, s.item_ptr, s.item_perm, s.items,
                                                                      s.idx16, s.pval);
        MF_CUDA(cudaGetLastError());
        tmp_free(len, st); tmp_free(start, st); tmp_free(tmp3, st);
    }
    if (s.nitems > 0 && s.npanels > 1 && s.panel_cost > 0) {
        k_panel_entry_cost<<<grid_for(s.npanels, 128), 128, 0, st>>>(s.npanels, s.panel_item_ptr, (uint32_t)s.panel_cost, cost);
        MF_CUDA(cudaGetLastError());
    }
    {   // (the scratch of the piece scans is sized for Q elements; a copy with long pieces has more items than pieces)
        uint32_t* tmp_cost = nullptr;
        MF_TRY(tmp_alloc(&tmp_cost, scan_tmp_elems((size_t)s.nitems), st));
        MF_TRY(exclusive_scan_u32(cost, cost_prefix, (size_t)s.nitems, tmp_cost, st));
        tmp_free(tmp_cost, st);
    }
    k_cta_ranges<<<grid_for(ncta + 1, 128), 128, 0, st>>>(ncta, s.nitems, cost_prefix, s.cta_item_ptr);
    MF_CUDA(cudaGetLastError());
    k_cta_starts<<<grid_for(ncta + 1, 128), 128, 0, st>>>(ncta, s.nitems, (uint32_t)s.npad, s.cta_item_ptr, s.items, s.cta_start_ptr);
    MF_CUDA(cudaGetLastError());
    MF_CUDA(cudaStreamSynchronize(st));
    trace_mark("    layout: fill + order + ranges");
    tmp_free(padded, st); tmp_free(nitem, st); tmp_free(seg_items, st); tmp_free(tmp, st); tmp_free(cost, st); tmp_free(cost_prefix, st);
    return MF_OK;
}

int side_panel_to_raw(const Side& s, float* dst_raw, cudaStream_t st) {
    int64_t Q = s.nseg * s.npanels;
    if (Q == 0 || s.nnz == 0) return MF_OK;
    int64_t warps = Q < 148 * 64 * 8 ? Q : 148 * 64 * 8;
    k_copy_values<true><<<grid_for(warps * 32, 256), 256, 0, st>>>(s.nseg, s.npanels, (uint32_t)s.chunk, s.ptr, s.piece_first, s.piece_ptr,
                                                                  s.item_ptr, s.item_perm, s.items, s.pval, dst_raw);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

int side_raw_to_panel(Side& s, const float* src_raw, cudaStream_t st) {
    int64_t Q = s.nseg * s.npanels;
    if (Q == 0 || s.nnz == 0) return MF_OK;
    int64_t warps = Q < 148 * 64 * 8 ? Q : 148 * 64 * 8;
    k_copy_values<false><<<grid_for(warps * 32, 256), 256, 0, st>>>(s.nseg, s.npanels, (uint32_t)s.chunk, s.ptr, s.piece_first, s.piece_ptr,
                                                                   s.item_ptr, s.item_perm, s.items, s.pval, const_cast<float*>(src_raw));
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

// ---- C-ABI: small integer ops -------------------------------------------------------------
static int stage_ptr(const uint32_t* ptr, int64_t n, uint32_t** d_ptr, int device) {
    MF_CUDA(cudaSetDevice(device));
    MF_TRY(dev_alloc(d_ptr, (size_t)n));
    MF_CUDA(cudaMemcpy(*d_ptr, ptr, sizeof(uint32_t) * (size_t)n, cudaMemcpyDefault));
    return MF_OK;
}

}  // namespace mf

extern "C" int mf_degree_bins(int64_t nseg, const uint32_t* ptr, uint64_t* seg_in_bin, uint64_t* nnz_in_bin, int device) {
    using namespace mf;
    MF_REQUIRE(nseg >= 0 && ptr && seg_in_bin && nnz_in_bin, "mf_degree_bins: bad argument");
    uint32_t* d_ptr = nullptr;
    MF_TRY(stage_ptr(ptr, nseg + 1, &d_ptr, device));
    unsigned long long* d_bins = nullptr;
    MF_TRY(dev_alloc(&d_bins, 66));
    MF_CUDA(cudaMemset(d_bins, 0, 66 * sizeof(unsigned long long)));
    if (nseg > 0) k_degree_bins<<<148 * 2, 256>>>(nseg, d_ptr, d_bins, d_bins + 33);
    MF_CUDA(cudaGetLastError());
    MF_CUDA(cudaMemcpy(seg_in_bin, d_bins, 33 * sizeof(uint64_t), cudaMemcpyDefault));
    MF_CUDA(cudaMemcpy(nnz_in_bin, d_bins + 33, 33 * sizeof(uint64_t), cudaMemcpyDefault));
    dev_free(d_bins);
    dev_free(d_ptr);
    return MF_OK;
}

extern "C" int mf_partition(int64_t nseg, const uint32_t* ptr, int P, int64_t* bound, int device) {
    using namespace mf;
    MF_REQUIRE(nseg >= 0 && ptr && bound && P >= 1, "mf_partition: bad argument");
    uint32_t* d_ptr = nullptr;
    MF_TRY(stage_ptr(ptr, nseg + 1, &d_ptr, device));
    long long* d_b = nullptr;
    MF_TRY(dev_alloc(&d_b, (size_t)P + 1));
    k_partition<<<(P + 1 + 127) / 128, 128>>>(nseg, d_ptr, P, d_b);
    MF_CUDA(cudaGetLastError());
    MF_CUDA(cudaMemcpy(bound, d_b, sizeof(int64_t) * ((size_t)P + 1), cudaMemcpyDefault));
    dev_free(d_b);
    dev_free(d_ptr);
    return MF_OK;
}
