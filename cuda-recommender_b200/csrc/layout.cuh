// layout.cuh — how one compressed-sparse copy of the ratings lives in HBM.
//
// A "side" is one copy with its solve direction:
//   CSC side: segments = item columns, gathers the user factor u (length rows), solves v
//   CSR side: segments = user rows,    gathers the item factor v (length cols), solves u
// (the reference keeps the same two copies, src/pmf_util.h:34-149, and reaches the CSR one through
//  the pointer-swap transpose, pmf_util.h:66-81).
//
// DIRECT layout: the caller's arrays as they are (ptr / uint32 idx / val).
//
// PANEL layout (the B200 layout): the gather dimension is cut into panels of
// `panel_rows` factor entries so that one panel of the factor vector sits in shared memory while
// the ratings that reference it stream past.  Every segment is cut at the panel boundaries into
// pieces (panel p, segment s), each piece padded to a multiple of `pad` entries and cut into work items of at
// most `chunk` entries; storage is in WORK-LIST order ("stream order"): panel-major, and inside a panel in the
// degree-binned, dealt order of the work items (below), item i starting where items 0..i-1 end — so the items
// a CTA walks are one contiguous stretch of the arrays, which the STREAM pipeline fetches with large TMA bulk
// copies.  `pad`: 8 for the STREAM pipeline (alignment no longer matters: the bulk copies are contiguous);
// 32 for the register-ring pipeline (every item then starts on a 64-byte boundary of the index array and a
// 128-byte boundary of the value array — measured −13 % there).  Indices are stored panel-local in 16 bits, pre-multiplied
// by 4 (the byte offset of the factor entry inside the staged panel, so panel_rows <= 16376); padding
// entries carry the offset of a zeroed shared-memory slot (index panel_rows) and val = 0, so they add
// nothing to g, h or the residual.
// Pieces are cut into work items of at most `chunk` entries; item j of segment s writes its partial
// (g, h) to a fixed slot, slots of one segment are contiguous and ordered (panel, chunk), and a
// finalize pass adds them in that order — the reduction tree of a segment depends only on its own
// entries and the global panel grid, never on scheduling or on the multi-GPU shard it sits in.
// Inside a panel the work items are ranked longest-first (degree bins) and dealt round-robin into 148
// lanes stored back to back (prep.cu): four consecutive items have nearly the same length, and every
// contiguous range of the list holds the same mix of lengths.
#pragma once
#include "common.cuh"

namespace mf {

struct WorkItem {        // 16 bytes, loaded as one uint4
    uint32_t start;      // first entry in the padded arrays (multiple of 8)
    uint32_t len;        // entries including padding (multiple of 8, <= chunk)
    uint32_t seg;        // local segment id
    uint32_t slot;       // partial-sum slot
};

struct Side {
    // shape
    int64_t nseg = 0;        // local segments (this shard)
    int64_t seg_offset = 0;  // global id of local segment 0 (multi-GPU)
    int64_t gdim = 0;        // length of the gathered factor vector (global)
    int64_t nnz = 0;         // local entries
    // direct layout (always present for ptr; idx/val only while needed)
    uint32_t* ptr = nullptr;  // [nseg+1] local, rebased to 0
    uint32_t* idx = nullptr;  // [nnz]
    float* val = nullptr;     // [nnz]   (DIRECT: the live residual)
    // panel layout
    int panel_rows = 0, chunk = 0, npanels = 0;
    int pad = 8;  // pieces are padded to a multiple of this many entries (multiple of 8, divides chunk)
    bool short_items = false;  // pieces average only a few entries: the sweeps run one item per LANE where they can (ccd_kernels.cu)
    int panel_cost = 0;  // cost-model charge (units of 8 entries) for entering another panel inside a CTA's range (prep.cu)
    int64_t npad = 0, nitems = 0, nslots = 0;
    uint32_t* piece_ptr = nullptr;    // [npanels*nseg + 1] start of piece (p,s) in the padded arrays
    uint32_t* piece_first = nullptr;  // [npanels*nseg]     raw offset of the piece's first entry
    uint32_t* item_ptr = nullptr;     // [npanels*nseg + 1] first work item of piece (p,s)
    uint16_t* idx16 = nullptr;        // [npad]
    float* pval = nullptr;            // [npad]  (PANEL: the live residual)
    WorkItem* items = nullptr;        // [nitems]
    uint32_t* slot_ptr = nullptr;     // [nseg+1]
    float2* partials = nullptr;       // [nslots]
    uint32_t* cta_item_ptr = nullptr; // [ncta+1] equal-cost contiguous item ranges
    uint32_t* cta_start_ptr = nullptr; // [ncta+1] first padded entry of each range (= items[cta_item_ptr[j]].start; npad at the end)
    uint32_t* panel_item_ptr = nullptr; // [npanels+1]
    uint32_t* item_perm = nullptr;      // [nitems] position in `items` of the j-th item in piece order (item_ptr numbering)
    int ncta = 0;
    bool sorted = true;
    uint32_t* unsort_perm = nullptr;  // [nnz] only when the caller's segments were not index-sorted: caller position of entry i
    // ALS work list (als.cu, built on first use): items sorted longest-first, long segments split into parts
    void* als_items = nullptr;        // AlsItem[als_nitems]
    int64_t als_nitems = 0;
    unsigned* als_queue = nullptr;    // queue head
    unsigned* als_counters = nullptr; // arrival counter per split segment
    float* als_partial = nullptr;     // partial tiles of the parts of split segments
};

int side_free(Side& s);
// raw device arrays must be set (ptr, idx, val, nseg, gdim, nnz); builds everything else.
int side_build_panels(Side& s, int panel_rows, int chunk, int ncta, cudaStream_t st);
// dst_raw[nnz] <- current panel values in the caller's order; and the inverse
int side_panel_to_raw(const Side& s, float* dst_raw, cudaStream_t st);
int side_raw_to_panel(Side& s, const float* src_raw, cudaStream_t st);
// *sorted: every segment's indices are strictly ascending; *in_range: every index is below gdim
int side_check_sorted(const Side& s, bool* sorted, bool* in_range, cudaStream_t st);
// sorts the segments of the caller-order arrays by index (host side; rare path), keeps the permutation in s.unsort_perm
int side_sort_segments(Side& s, cudaStream_t st);
// *ok = every d_a[i] < bound (device array)
int check_below(const uint32_t* d_a, int64_t n, uint64_t bound, bool* ok, cudaStream_t st);
// dst[perm[i]] = src[i]
int scatter_by_perm(const uint32_t* perm, const float* src, float* dst, int64_t n, cudaStream_t st);

}  // namespace mf
