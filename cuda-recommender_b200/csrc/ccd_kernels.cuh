// ccd_kernels.cuh — launch interface of the CCD++ sweeps (ccd_kernels.cu).
#pragma once
#include "layout.cuh"

namespace mf {

// sweep mode bits
enum : int {
    kSub = 1,     // residual -= g_old[idx] * s_old[seg]
    kAdd = 2,     // residual += g_add[idx] * s_add[seg]
    kSolve = 4,   // accumulate (g, h) against g_new
    kAddSep = 8,  // the add-back gathers a vector other than g_new (CSR side of the fused schedule)
};

// In-kernel finalize of a solving sweep (register-ring pipeline): after its items every CTA arrives at a grid-wide
// barrier (all CTAs are co-resident: one persistent CTA per SM), then the CTAs share the segments: add the slots in
// order, out = g / (lambda*deg + h), and — multi-GPU — store the block into the peers' LL receive buffers.  Same
// reduction trees as the stand-alone finalize kernel, so both give identical bits.
struct SweepFinalize {
    int enabled;                         // 0: partial sums only (a separate panel_finalize launch follows)
    int lanes;                           // 1: a thread per segment, 32: a warp per segment
    int64_t nseg;
    const uint32_t* slot_ptr;
    const uint32_t* seg_ptr;
    float lambda;
    int nmf;
    float* out;                          // this shard's block of the factor vector
    unsigned* bar;                       // grid barrier counter (monotonic)
    unsigned bar_target;                 // value of *bar once every CTA of this launch has arrived
    unsigned* status;                    // session status word: a wait that times out records a code here and the kernel returns
    unsigned long long* const* peer_ll;  // multi-GPU push (nullptr: none), as FinalizePush
    int64_t vec_off;
    int rank, nranks;
    unsigned epoch;
    // multi-GPU receive side, in the same launch: after its own block is sent, the kernel polls this rank's LL receive
    // buffer for the blocks of the peers and completes the factor vector (what exchange_unpack does as a launch)
    const unsigned long long* ll;        // nullptr: none
    float* vec;                          // the full-length factor vector
    int64_t dim, own_lo, own_hi;
};

struct PanelSweepArgs {
    const uint16_t* idx16;
    float* val;
    const WorkItem* items;
    const uint32_t* cta_item_ptr;
    const uint32_t* panel_item_ptr;
    int npanels;
    uint32_t panel_rows;
    int64_t gdim;
    int64_t seg_offset;   // global id of local segment 0 (indexes s_add / s_old)
    const float* g_new;   // gathered factor for the solve (and the add-back unless kAddSep)
    const float* g_add;   // gathered factor for the add-back when kAddSep
    const float* g_old;   // gathered factor of the rank being subtracted
    const float* s_add;   // per-segment factor for the add-back (full-length vector)
    const float* s_old;   // per-segment factor of the rank being subtracted
    float2* partials;
    uint32_t nslots;      // TMA pipeline: shared-memory slots of the ring (set by panel_sweep)
    const uint32_t* cta_start_ptr;  // STREAM pipeline: [ncta+1] first padded entry of every CTA's item range
    uint32_t ring_entries;          // STREAM pipeline: entries of the shared-memory ring (set by panel_sweep)
    uint32_t pf_dist;               // register ring: L2 prefetch distance in entries (0: off)
    uint32_t npad;                  // padded entries of the copy
    int short_items;                // register ring: 1 = short-piece copy: warps take 32 items at a time, a lane per item when all are <= 32 entries
    unsigned long long* trace_cta;  // nullptr, or 4 words per CTA: after the dependency wait / after its items / item range
    unsigned long long* trace;      // nullptr, or 8 words: %globaltimer of CTA 0 at entry / after the dependency wait / after
                                    // its items / after the grid barrier / after its finalize share / at exit (MF_SWEEP_TRACE)
    SweepFinalize fin;    // register-ring pipeline only
};

// ---- persistent kernel (one cooperative launch per outer iteration, ccd_kernels.cu) ----
struct PersistSide {
    const uint16_t* idx16;
    float* val;
    const WorkItem* items;
    const uint32_t* cta_item_ptr;
    const uint32_t* panel_item_ptr;
    int npanels;
    uint32_t panel_rows;
    int64_t gdim, seg_offset, nseg;
    const uint32_t* slot_ptr;
    const uint32_t* seg_ptr;
    float2* partials;
    int lanes;                           // finalize: 1 thread or 32 lanes per segment
    // multi-GPU exchange of the vector this side solves (nullptr: single GPU)
    unsigned long long* const* peer_ll;  // [nranks] LL receive buffers of every rank for the solved factor matrix
    const unsigned long long* ll;        // this rank's receive buffer
    int64_t dim;                         // full length of the solved vector
};
struct PersistArgs {
    PersistSide csc, csr;
    float *W, *H, *v_old;
    int64_t ldm, ldn;
    int k, T;
    int add;        // 1 from the second outer iteration on (src/CCD.cpp:100)
    int pending;    // rank whose subtraction is deferred at entry (-1: none)
    float lambda;
    int nmf;
    unsigned* bar;        // grid barrier counter (monotonic)
    unsigned bar_base;    // its value before this launch
    unsigned* status;     // 0 = ok; set by a wait that timed out
    int rank, nranks;
    unsigned epoch_base;  // exchange epoch before this launch (phase n uses epoch_base + n)
    unsigned long long* stamps;  // nullptr, or [1 + 2kT] %globaltimer at kernel start and after every phase
};
// smem: dynamic shared memory (the largest panel footprint of any phase).  Returns MF_ERR_UNSUPPORTED when the grid
// cannot be co-resident / cooperative launch is unavailable (the caller then uses the per-launch path).
int ccd_persistent_launch(const PersistArgs& a, int ncta, size_t smem, cudaStream_t st);
bool ccd_persistent_supported(int ncta, size_t smem, int device);

struct DirectSweepArgs {
    int64_t nseg;
    const uint32_t* ptr;
    const uint32_t* idx;
    float* val;
    int64_t seg_offset;
    const float* g_new;
    const float* g_add;
    const float* g_old;
    const float* s_add;
    const float* s_old;
    float lambda;
    int nmf;
    float* out;  // already offset to this shard's first segment
};

int panel_timeout_report(char* buf, size_t n);  // 1 + message when a pipeline wait timed out since the last call
int panel_sweep_threads();  // threads per CTA of the register-ring sweep kernels (per-launch and persistent)
int panel_sweep_vectors(int mode);
size_t panel_sweep_smem(int mode, int panel_rows);
uint32_t panel_stream_ring(int mode, int panel_rows, int chunk);  // STREAM pipeline: ring entries that fit beside the staged vectors (0: none)
size_t panel_stream_smem(int mode, int panel_rows, uint32_t ring);
int panel_sweep(int mode, const PanelSweepArgs& a, int ncta, int threads, int chunk, int pipeline, cudaStream_t st);
// multi-GPU: the finalize kernel also sends the solved block to every peer (LL protocol) or, as a barrier, only
// publishes the epoch in the peers' flag words
struct FinalizePush {
    unsigned long long* const* peer_ll;  // [nranks] LL receive buffer of every rank for this factor matrix (device array)
    unsigned* const* peer_flags;         // [nranks] flag words of every rank (device array)
    unsigned* ticket;                    // local CTA ticket counter
    int64_t vec_off;                     // index of out[0] inside the factor vector
    int rank, nranks;
    unsigned epoch;
    int barrier;                         // 1 = send no values, publish the epoch (exchange_wait on the other side)
};
bool panel_sweep_grid_resident(int ncta, int threads, int panel_rows, int sm_count);  // grid barrier possible?
int panel_finalize_lanes(int64_t nseg, int64_t nslots);  // 32: many slots per segment -> a warp per segment, else 1
int panel_finalize(int64_t nseg, int64_t nslots, const uint32_t* slot_ptr, const float2* partials, const uint32_t* seg_ptr,
                   float lambda, int nmf, float* out, const FinalizePush* push, cudaStream_t st);
int exchange_wait(const unsigned* flags, int rank, int nranks, unsigned epoch, cudaStream_t st);
// polls the LL receive buffer for the entries of `vec` owned by peers (everything outside [own_lo, own_hi)) and stores them
int exchange_unpack(const unsigned long long* ll, float* vec, int64_t dim, int64_t own_lo, int64_t own_hi, unsigned epoch,
                    cudaStream_t st);
int direct_sweep(int mode, const DirectSweepArgs& a, int sm_count, cudaStream_t st);

}  // namespace mf
