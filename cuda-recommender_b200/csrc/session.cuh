// session.cuh — device-resident training state behind the C-ABI (include/mf_abi.h).
#pragma once
#include <vector>

#include "ccd_kernels.cuh"
#include "layout.cuh"

namespace mf {

enum Family { F_SOLVE = 0, F_FUSED, F_UPDATE, F_FINALIZE, F_ALS, F_RMSE, F_COLLECTIVE, F_PERSIST, F_COUNT };

// CUDA-event stopwatch per kernel family, on the session stream.  Events are pooled; durations are
// read back after the stream has been synchronised (collect()).
struct FamilyTimer {
    struct Span { cudaEvent_t a, b; int fam; };
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<Span> spans;
    bool enabled = true;
    int64_t launched = 0;  // every start() call, timed or not
    cudaStream_t st = nullptr;
    int open_fam = -1;
    cudaEvent_t open_ev = nullptr;

    cudaEvent_t take();
    void start(int fam);
    void stop();
    void collect(double* seconds, int64_t* launches);  // adds into seconds[F_COUNT], launches[F_COUNT]
    void destroy();
};

struct Dist;  // dist.cu

}  // namespace mf

struct mf_session {
    mf_params prm;
    int device = 0, sm_count = 148;
    cudaStream_t st = nullptr;
    mf::DeviceArena* arena = nullptr;  // ratings, layout arrays (everything long-lived that is not exported through IPC)
    int64_t rows = 0, cols = 0, nnz = 0;
    int rank = 0, nranks = 1;
    std::vector<int64_t> row_bound, col_bound;  // nranks+1 each
    mf::Side csc, csr;
    bool panel = false;
    int k = 0;
    int64_t ldm = 0, ldn = 0;  // CCD++: leading dimensions of W[k][ldm], H[k][ldn]
    float *W = nullptr, *H = nullptr;
    float* v_old = nullptr;  // CCD++: H as the previous outer iteration left it, [k][ldn] (add-back of the CSR copy)
    int64_t nt = 0;
    uint32_t *trow = nullptr, *tcol = nullptr;
    float* tval = nullptr;
    double* d_acc = nullptr;
    unsigned* d_gridbar = nullptr;  // [0] grid barrier counter of the in-kernel finalize (monotonic; gridbar_total = expected value), [1] status word
    unsigned gridbar_total = 0;
    bool fin_in_kernel = true;
    bool persistent = false;        // one cooperative launch per outer iteration (ccd_kernels.cu: k_ccd_persistent)
    size_t persist_smem = 0;
    uint32_t pf_dist = 0;  // register-ring sweeps: L2 prefetch distance in entries (0: off)
    unsigned long long* d_stamps = nullptr;  // [1 + 2kT] phase time stamps of the last persistent launch
    std::vector<unsigned long long> h_stamps;
    bool broken = false;            // a device-side wait timed out: the session refuses further work
    // ---- optional per-rank reporting and the -e stop rule (CCD++) ----
    bool rank_report = false;       // verbose && do_predict: per-rank time and incremental test RMSE (src/CCD.cpp:141-148)
    bool early_stop = false;        // mf_params.early_stop: inner iterations end when the function decrease falls under eps * max
    float* tres = nullptr;          // [nt] test residual of calrmse_r1 (src/tools.cpp:260-270)
    float* u_prev = nullptr;        // [ldm] u_t as it was when the rank started
    float* vec_prev = nullptr;      // [max(ldm, ldn)] the vector a solve sweep is about to overwrite (stop rule)
    double* d_rank_acc = nullptr;   // [k] sum of squared test residuals after each rank; [k .. k+1] function decrease of the two sweeps
    std::vector<cudaEvent_t> rank_ev;          // k+1 events (rank_report)
    std::vector<double> rank_seconds, rank_rmse;  // last outer iteration
    std::vector<int> rank_inner;               // inner iterations run per rank in the last outer iteration
    double fundec_max = 0.0;
    // MF_SWEEP_TRACE=<file>: in-kernel time stamps of the first launches of the session (debug timeline, profiles/)
    unsigned long long* d_trace = nullptr;
    int trace_cap = 0, trace_n = 0;
    unsigned long long* d_trace_cta = nullptr;  // [kTraceCtaLaunches][ncta][4]: per-CTA stamps of launches trace_cta_from .. +kTraceCtaLaunches
    int trace_cta_from = 0;
    int outer_done = 0;
    int pending = -1;  // rank whose subtraction from the residual is still deferred (fused schedule)
    mf::FamilyTimer timer;
    double fam_seconds[mf::F_COUNT] = {0};
    int64_t fam_launches[mf::F_COUNT] = {0};
    double last_seconds = 0.0;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    mf::Dist* dist = nullptr;
};

namespace mf {
// dist.cu
int dist_unique_id(void* id128);
int dist_create(Dist** out, int rank, int nranks, const void* id128, int device);
int dist_destroy(Dist* d);
void dist_release_cached(int device);  // destroys the communicators kept between sessions
// optional NVLink peer-to-peer exchange (CUDA IPC); all ranks agree on it, and fall back to NCCL together when unavailable
int dist_setup_p2p(Dist* d, float* W, float* H, int64_t ldm, int64_t ldn, cudaStream_t st);
unsigned long long* const* dist_peer_ll(const Dist* d, bool h);
unsigned long long* dist_ll(const Dist* d, bool h);
bool dist_p2p(const Dist* d);
int dist_rank(const Dist* d);
unsigned* const* dist_peer_flags(const Dist* d);
unsigned* dist_flags(const Dist* d);
unsigned dist_next_epoch(Dist* d);
unsigned dist_advance_epoch(Dist* d, unsigned n);  // returns the epoch before; the next n epochs belong to the caller
// in-place all-gather of a full-length vector whose block r = [bound[r], bound[r+1]) was produced by rank r
int dist_allgather_blocks(Dist* d, float* vec, const std::vector<int64_t>& bound, int64_t elems_per_unit, cudaStream_t st);
int dist_allreduce_sum_double(Dist* d, double* dev_value, cudaStream_t st);
// als.cu
int als_half_step(Side& s, const float* Y, float* X, int k, float lambda, int sm_count, cudaStream_t st);
int als_plan_host(const uint32_t* ptr, int64_t nseg, uint32_t split, uint32_t* items4, int64_t* n_items, uint32_t* n_slots);
}  // namespace mf
