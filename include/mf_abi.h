/*
 * mf_abi.h — C-ABI boundary of the B200-native matrix-factorization training path.
 *
 * This is the drop-in boundary for the GPU path of Zialus/CUDA-Recommender.  The
 * reference has no FFI: its "operator API" is two C++ functions over C++ containers,
 *
 *   void kernel_wrapper_ccdpp_NV(SparseMatrix& R, TestData& T, MatData& W, MatData& H,
 *                                parameter& parameters);      cuda_src/CCD_CUDA.h:49, CCD_CUDA.cu:164
 *   void kernel_wrapper_als_NV  (SparseMatrix& R, TestData& T, MatData& W, MatData& H,
 *                                parameter& parameters);      cuda_src/ALS_CUDA.h:40, ALS_CUDA.cu:183
 *
 * called from src/main.cpp:11-17.  cuda-recommender_b200/host/shim.cpp keeps those two
 * names and signatures (so the reference's own main.cpp links unchanged)
 * and forwards to mf_ccdpp_train / mf_als_train below: plain pointers and sizes, no C++
 * or torch types.  Everything else in this header is the device-resident "session" form
 * of the same path (upload once, iterate, read back), the step-level entry points the
 * parity tests drive, and the integer-tier ingest ops.
 *
 * Conventions
 *   - all indices uint32, all values/factors float32 (src/pmf_util.h:26,146-148)
 *   - CCD++ factor layout  W[t*rows + i], H[t*cols + j]   (MatData W[t][i], main.cpp:92-93)
 *   - ALS   factor layout  W[i*k + t],    H[j*k + t]      (MatData W[i][t], main.cpp:86-87)
 *   - every function returns 0 on success, else an MF_ERR_* code; mf_last_error() gives
 *     the message.  Nothing throws, nothing falls back to the CPU: without a usable
 *     CUDA device every compute entry point fails with MF_ERR_CUDA.
 *   - pointers in mf_ratings / mf_testset / factor arguments may be host OR device
 *     pointers (copies use cudaMemcpyDefault); the one-shot trainers are the host-buffer
 *     path, sessions keep everything resident in HBM.
 */
#ifndef MF_ABI_H
#define MF_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_ABI_VERSION 3

enum {
    MF_OK = 0,
    MF_ERR_ARG = 1,      /* bad argument / inconsistent sizes */
    MF_ERR_CUDA = 2,     /* CUDA runtime error (message has cudaGetErrorString) */
    MF_ERR_NCCL = 3,     /* NCCL error */
    MF_ERR_STATE = 4,    /* call not valid for this session (e.g. ALS step on a CCD++ session) */
    MF_ERR_UNSUPPORTED = 5
};

enum { MF_SOLVER_CCD = 0, MF_SOLVER_ALS = 1 };               /* src/pmf.h:6 solvertype */
enum { MF_SCHEDULE_FUSED = 0, MF_SCHEDULE_REFERENCE = 1 };   /* REFERENCE = CCD_CUDA.cu:339-378 launch order */
enum { MF_LAYOUT_PANEL = 0, MF_LAYOUT_DIRECT = 1 };          /* HBM layout of the rating copies */
/* how the panel sweep feeds its warps: a per-lane register ring (default, fastest measured), or —
 * kept for comparison — producer warps with cp.async / one bulk-copy (TMA) descriptor per work item into a
 * shared-memory slot ring guarded by mbarriers */
enum { MF_PIPELINE_REGISTERS = 0, MF_PIPELINE_ASYNC = 1, MF_PIPELINE_TMA_BULK = 2, MF_PIPELINE_STREAM = 3 };
enum { MF_SIDE_CSC = 0, MF_SIDE_CSR = 1 };                   /* CSC: columns solve v / H;  CSR: rows solve u / W */

/* Paired CSR + CSC of the same ratings — src/pmf_util.h:34-149 (SparseMatrix). */
typedef struct mf_ratings {
    int64_t rows, cols, nnz;
    const uint32_t* csr_row_ptr; /* [rows+1] */
    const uint32_t* csr_col_idx; /* [nnz]    */
    const float* csr_val;        /* [nnz]    */
    const uint32_t* csc_col_ptr; /* [cols+1] */
    const uint32_t* csc_row_idx; /* [nnz]    */
    const float* csc_val;        /* [nnz]    */
} mf_ratings;

/* Held-out COO triples — src/pmf_util.h:151-211 (TestData). nnz may be 0. */
typedef struct mf_testset {
    int64_t nnz;
    const uint32_t* row;
    const uint32_t* col;
    const float* val;
} mf_testset;

/* Mirrors class parameter (src/pmf.h:8-43) field for field, then the extensions. */
typedef struct mf_params {
    int32_t solver_type;       /* MF_SOLVER_*            pmf.h:10 */
    uint32_t k;                /*                        pmf.h:11 */
    int32_t threads;           /* ignored by the GPU path pmf.h:12 */
    int32_t maxiter;           /* outer iterations       pmf.h:13 */
    int32_t maxinneriter;      /* CCD++ inner iterations pmf.h:14 */
    float lambda;              /*                        pmf.h:15 */
    float eps;                 /* used only with early_stop = 1 (inert in the reference) pmf.h:16 */
    int32_t do_predict;        /* with verbose: per-rank incremental test RMSE (the reference's commented-out block, CCD.cpp:141-148) pmf.h:17 */
    int32_t verbose;           /* per-rank report lines "iter %d rank %d time %f [rmse %f]"  pmf.h:18 */
    int32_t do_nmf;            /* 1: solved coordinates are clamped at 0 (inert in the reference; same as nmf_project) pmf.h:19 */
    uint32_t nBlocks;          /* accepted, ignored: geometry comes from the work partition  pmf.h:22 */
    uint32_t nThreadsPerBlock; /* accepted, ignored      pmf.h:23 */
    /* ---- extensions (zero = default) ---- */
    int32_t device;            /* CUDA device ordinal */
    int32_t schedule;          /* MF_SCHEDULE_* (CCD++) */
    int32_t layout;            /* MF_LAYOUT_*   (CCD++) */
    int32_t quiet;             /* 1: do not print the per-iteration "[-INFO-] iteration num" line */
    int32_t panel_rows;        /* 0 = default and max (16376); factor entries per shared-memory panel */
    int32_t chunk;             /* 0 = default (512); max rating entries per work item */
    int32_t nmf_project;       /* 1: clamp solved coordinates at 0 (extension; the reference never does) */
    int32_t no_launch_timing;  /* 1: skip the per-launch CUDA events (mf_kernel_times stays zero) */
    int32_t pipeline;          /* MF_PIPELINE_* */
    int32_t timing_stride;     /* CCD++: per-launch events only on every timing_stride-th rank (0/1 = every rank) */
    int32_t pad_entries;       /* 0 = default (32): pieces are padded to a multiple of this many entries (8, 16, 32, 64) */
    int32_t early_stop;        /* 1: the -e rule of CCDR1 is active (single GPU, panel layout): the inner iterations of a rank end
                                * after the one whose function decrease is below eps * (largest decrease seen so far); 0 (default):
                                * eps is inert, exactly like the reference */
    int32_t reserved[4];
} mf_params;

/* One line of the reference's per-iteration report (CCD_CUDA.cu:405, ALS_CUDA.cu:360). */
typedef struct mf_iter_stats {
    double rank_time;   /* seconds in coordinate solves (CCD++); 0 for ALS            */
    double update_time; /* seconds in residual updates (CCD++) / the two half-steps (ALS) */
    double rmse;        /* test RMSE after this outer iteration (NaN when no test set)  */
    double rmse_time;   /* seconds                                                      */
} mf_iter_stats;

/* Per-kernel-family device times accumulated by the last iterate call (CUDA events on the session stream
 * around the launches of that family; with timing_stride > 1 only the launches of every n-th rank are timed,
 * *_launches counts the timed ones). */
typedef struct mf_kernel_times {
    double solve_s;        int64_t solve_launches;        /* read-only solve sweeps               */
    double fused_s;        int64_t fused_launches;        /* update(s)+solve sweeps               */
    double update_s;       int64_t update_launches;       /* stand-alone residual updates         */
    double finalize_s;     int64_t finalize_launches;     /* per-segment partial reduction + g/h  */
    double als_s;          int64_t als_launches;          /* ALS half-steps                       */
    double rmse_s;         int64_t rmse_launches;
    double collective_s;   int64_t collective_launches;   /* NCCL all-gathers (multi-GPU)         */
    int64_t solve_bytes, fused_bytes, update_bytes;       /* HBM bytes one launch of the family must move */
    int64_t total_launches;  /* every kernel launched by the last iterate call, timed or not (sweeps, finalize, ALS, RMSE) */
    /* CCD++ persistent kernel (one cooperative launch per outer iteration; solve_* / fused_* then hold the in-kernel
     * phase times, %globaltimer stamps at the grid barriers, and *_launches count phases) */
    double persistent_s;   int64_t persistent_launches;   /* CUDA events around the launches                           */
    int64_t persistent_bytes;                              /* HBM bytes one launch must move: all 2kT phases             */
} mf_kernel_times;

typedef struct mf_session mf_session;

/* ---- library ---- */
int mf_abi_version(void);
const char* mf_last_error(void);
int mf_device_count(int* count);
void mf_params_default(mf_params* p); /* the reference defaults, src/pmf.h:26-42 */
/* host-side factor seeding exactly as the reference does it (initial_col, src/tools.cpp:165-173):
 * srand(0); for i<n, for j<k: X[j*n + i] = 0.1f*(float(rand())/RAND_MAX) + 0.001f                 */
void mf_host_initial_col(float* X, int64_t k, int64_t n);

/* ---- one-shot trainers: the drop-in for kernel_wrapper_ccdpp_NV / kernel_wrapper_als_NV ----
 * W/H in: initial factors (CCD++ ignores H and starts it at zero, CCD_CUDA.cu:263-269,287);
 * W/H out: final factors, same layout.  stats: NULL or maxiter entries.  Prints the
 * reference's per-iteration line unless params->quiet.                                     */
int mf_ccdpp_train(const mf_ratings* R, const mf_testset* T, float* W, float* H, const mf_params* params,
                   mf_iter_stats* stats);
int mf_als_train(const mf_ratings* R, const mf_testset* T, float* W, float* H, const mf_params* params,
                 mf_iter_stats* stats);

/* Device memory the library keeps cached between sessions of one process — the stream-ordered pool that holds a session's
 * rating / layout arena and its scratch, so that a second training call of the process skips a multi-GB cudaMalloc and a
 * cudaFree that was measured to stall for up to 0.9 s — goes back to the driver.  The reference leaves no device state
 * behind (cudaDeviceReset, cuda_src/CCD_CUDA.cu:177, ALS_CUDA.cu:196): the two kernel_wrapper_* shims call this after
 * training; a host that trains repeatedly does not.  Call it only while no session is open on `device` (it also destroys
 * the NCCL communicators cached for that device).                                                                      */
int mf_release_cached_memory(int device);

/* ---- sessions: ratings, residual, factors resident in HBM ---- */
int mf_session_create(const mf_ratings* R, const mf_testset* T, const mf_params* params, mf_session** out);
int mf_session_destroy(mf_session* s);
/* multi-GPU: one session per rank/GPU; this rank keeps CSR row block + CSC column block `rank` of
 * `nranks` (nnz-balanced split) and all-gathers fresh factor blocks over NCCL.  nccl_unique_id is the
 * 128-byte ncclUniqueId every rank must share (rank 0: mf_dist_unique_id, then broadcast it).        */
/* Sessions created with the SAME unique id (same rank, size, device) in one process share one NCCL communicator:
 * ncclCommInitRank (0.5-1 s) is paid once; mf_release_cached_memory() destroys the cached communicators.        */
int mf_dist_unique_id(void* id128);
int mf_session_create_dist(const mf_ratings* R, const mf_testset* T, const mf_params* params, int rank, int nranks,
                           const void* nccl_unique_id, mf_session** out);
int mf_session_set_factors(mf_session* s, const float* W, const float* H); /* H==NULL: zero (CCD++ start) */
int mf_session_get_factors(mf_session* s, float* W, float* H);
/* the value arrays as the solver currently holds them, in the caller's original CSR / CSC order
 * (CCD++: the residual, which the reference keeps in R, CCD.cpp:25,36).  Pending deferred updates of
 * the fused schedule are flushed first.  Either pointer may be NULL.  (multi-GPU: local block only)    */
int mf_session_get_values(mf_session* s, float* csr_val, float* csc_val);
/* n_outer outer iterations from the current state; stats: NULL or n_outer entries.            */
int mf_session_ccdpp_iterate(mf_session* s, int n_outer, mf_iter_stats* stats);
int mf_session_als_iterate(mf_session* s, int n_iter, mf_iter_stats* stats);
int mf_session_rmse(mf_session* s, double* rmse);
/* Predictions w_i . h_j of the CURRENT factors for n arbitrary (row, col) pairs (0-based; host or device pointers):
 * FP32 products summed in rank order in FP64 — the loop body of calculate_rmse_from_file, src/extras.cpp:165-168
 * (and of dot(), src/tools.cpp:184-198) — so the doubles are bit-identical to the CPU path's on the same factors. */
int mf_session_predict(mf_session* s, int64_t n, const uint32_t* row, const uint32_t* col, double* out);
/* CCD++, last outer iteration, k entries each (any pointer may be NULL): device seconds per rank and the incremental test
 * RMSE after each rank — the reference's verbose block, src/CCD.cpp:141-148 with calrmse_r1, src/tools.cpp:260-270; both
 * need verbose != 0, do_predict != 0 and a test set — and the inner iterations run per rank (< maxinneriter only with
 * early_stop). */
int mf_session_rank_stats(mf_session* s, double* seconds, double* rmse, int32_t* inner_iters);
/* Predict-only path (no session): out[e] = w_row[e] . h_col[e] for a SAVED model, W [rows][k] and H [cols][k] row-major
 * as save_mat_t / load_mat_t hold them (src/tools.cpp:90-153); host pointers; same arithmetic as mf_session_predict —
 * the loop body of calculate_rmse_from_file, src/extras.cpp:143-180. */
int mf_predict_pairs(const float* W, const float* H, int64_t rows, int64_t cols, int64_t k, int64_t n, const uint32_t* row,
                     const uint32_t* col, double* out, int device);
int mf_session_kernel_times(mf_session* s, mf_kernel_times* out);
/* device seconds of the last iterate call, CUDA events on the session stream (RMSE excluded) */
int mf_session_last_seconds(mf_session* s, double* seconds);

/* ---- step-level entry points (what the parity tests drive) ----
 * mf_session_ccd_solve : one coordinate sweep for rank t — side MF_SIDE_CSC solves v=H[t] from the CSC
 *                        copy and u=W[t] (CCD.cpp:110-113); MF_SIDE_CSR solves u from v (CCD.cpp:118-121).
 * mf_session_ccd_update: residual update of BOTH copies with rank t, add!=0 adds (CCD.cpp:100-103) else
 *                        subtracts (CCD.cpp:133-134).
 * mf_session_als_half  : one ALS half-step — MF_SIDE_CSR updates W from H (ALS.cpp:98-158), MF_SIDE_CSC
 *                        updates H from W (ALS.cpp:161-219).                                            */
int mf_session_ccd_solve(mf_session* s, int t, int side);
int mf_session_ccd_update(mf_session* s, int t, int add);
int mf_session_als_half(mf_session* s, int side);

/* ---- integer tier (bit-exact ops) ---- */
/* COO (unique pairs, any order; host or device pointers) -> CSR sorted by (row,col) + CSC sorted by
 * (col,row), written to caller buffers (host or device).                                              */
int mf_build_csr_csc(int64_t rows, int64_t cols, int64_t nnz, const uint32_t* coo_row, const uint32_t* coo_col,
                     const float* coo_val, uint32_t* csr_row_ptr, uint32_t* csr_col_idx, float* csr_val,
                     uint32_t* csc_col_ptr, uint32_t* csc_row_idx, float* csc_val, int device);
/* 33-bin degree histogram: bin b = segments with bit-length(deg) == b (0 = empty, 1 = deg 1, 2 = 2..3 …). */
int mf_degree_bins(int64_t nseg, const uint32_t* ptr, uint64_t* seg_in_bin, uint64_t* nnz_in_bin, int device);
/* nnz-balanced contiguous partition into P blocks: bound[p] = first segment with ptr[s] >= ceil(p*nnz/P). */
int mf_partition(int64_t nseg, const uint32_t* ptr, int P, int64_t* bound, int device);
/* The work list of an ALS half-step, as the session plans it from a HOST pointer array (host-only, no device needed): one
 * item {segment, part, nparts, slot} per segment, or nparts items for a segment with more than `split` entries (parts of
 * equal length, a multiple of 32; the library's default split is 8192), sorted longest-first.  items: NULL or room for
 * 4 x (*n_items) uint32 (call twice); *n_slots = partial-tile slots the split segments need.                           */
int mf_als_plan(int64_t nseg, const uint32_t* ptr, uint32_t split, uint32_t* items, int64_t* n_items, uint32_t* n_slots);
/* the panel layout this session built for one side, copied out for inspection by the tests:
 * sizes first (any pointer NULL -> only *n_padded / *n_items / *n_panels are written).                 */
int mf_session_panel_layout(mf_session* s, int side, int64_t* n_padded, int64_t* n_items, int64_t* n_panels,
                            uint16_t* idx16, float* val, uint32_t* items /* 4 x n_items */);

#ifdef __cplusplus
}
#endif
#endif /* MF_ABI_H */
