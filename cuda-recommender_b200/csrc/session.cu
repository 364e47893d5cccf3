// session.cu — the C-ABI (include/mf_abi.h): device residency, the CCD++ and ALS drivers, the
// one-shot trainers that replace kernel_wrapper_ccdpp_NV / kernel_wrapper_als_NV.
//
// Reference orchestration being replaced: ccdpp_NV (cuda_src/CCD_CUDA.cu:224-451) and als_NV
// (cuda_src/ALS_CUDA.cu:200-406); schedule semantics follow the CPU path ccdr1_OMP (src/CCD.cpp:45-163)
// and ALS_OMP (src/ALS.cpp:81-233), which is also what the oracle restates.
#include "session.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "rmse.cuh"

namespace mf {

// ------------------------------------------------------------------------------------------
// FamilyTimer
// ------------------------------------------------------------------------------------------
cudaEvent_t FamilyTimer::take() {
    if (used == pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        pool.push_back(e);
    }
    return pool[used++];
}
void FamilyTimer::start(int fam) {
    if (fam != F_COLLECTIVE) ++launched;
    if (!enabled) return;
    open_fam = fam;
    open_ev = take();
    cudaEventRecord(open_ev, st);
}
void FamilyTimer::stop() {
    if (!enabled || open_fam < 0) return;
    cudaEvent_t b = take();
    cudaEventRecord(b, st);
    spans.push_back({open_ev, b, open_fam});
    open_fam = -1;
}
void FamilyTimer::collect(double* seconds, int64_t* launches) {
    for (const Span& s : spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            seconds[s.fam] += ms * 1e-3;
            launches[s.fam] += 1;
        }
    }
    spans.clear();
    used = 0;
}
void FamilyTimer::destroy() {
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
    spans.clear();
    used = 0;
}

namespace {

constexpr size_t kSmemMax = 227 * 1024 - 256;

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// largest panel (multiple of 8) such that nvec vectors of (panel+8) floats fit in shared memory
int panel_cap(int nvec) {
    int64_t cap = (int64_t)(kSmemMax / (sizeof(float) * nvec)) - 8;
    return (int)(cap / 8 * 8);
}

// panel size for a gather dimension: one panel if it fits, else the fewest equal panels
int choose_panel_rows(int64_t gdim, int cap) {
    if (gdim <= cap) return (int)std::max<int64_t>(8, round_up(gdim, 8));
    int64_t np = (gdim + cap - 1) / cap;
    return (int)round_up((gdim + np - 1) / np, 8);
}

// nnz-balanced contiguous partition (same rule as oracle orc_partition / mf_partition)
void partition_host(const std::vector<uint32_t>& ptr, int P, std::vector<int64_t>& bound) {
    const int64_t nseg = (int64_t)ptr.size() - 1;
    const uint64_t nnz = ptr[nseg];
    bound.assign(P + 1, 0);
    for (int p = 1; p < P; ++p) {
        uint64_t target = (nnz * (uint64_t)p + (uint64_t)P - 1) / (uint64_t)P;
        bound[p] = std::lower_bound(ptr.begin(), ptr.end(), target,
                                    [](uint32_t a, uint64_t t) { return (uint64_t)a < t; }) - ptr.begin();
        if (bound[p] > nseg) bound[p] = nseg;
    }
    bound[P] = nseg;
}

// upload segments [s0, s1) of one compressed copy (host or device source pointers).  With `keep` the copies are only
// enqueued on `st` (the rebased pointer array lives in *keep until the caller has synchronised the stream).
int upload_side(Side& sd, const std::vector<uint32_t>& ptr_full, int64_t s0, int64_t s1, const uint32_t* idx,
                const float* val, int64_t gdim, cudaStream_t st, std::vector<uint32_t>* keep = nullptr) {
    sd.nseg = s1 - s0;
    sd.seg_offset = s0;
    sd.gdim = gdim;
    const uint32_t e0 = ptr_full[s0], e1 = ptr_full[s1];
    sd.nnz = (int64_t)e1 - (int64_t)e0;
    std::vector<uint32_t> own;
    std::vector<uint32_t>& local = keep ? *keep : own;
    local.resize((size_t)sd.nseg + 1);
    for (int64_t s = 0; s <= sd.nseg; ++s) local[s] = ptr_full[s0 + s] - e0;
    MF_TRY(dev_alloc(&sd.ptr, (size_t)sd.nseg + 1));
    MF_TRY(dev_alloc(&sd.idx, (size_t)sd.nnz));
    MF_TRY(dev_alloc(&sd.val, (size_t)sd.nnz));
    MF_CUDA(cudaMemcpyAsync(sd.ptr, local.data(), sizeof(uint32_t) * local.size(), cudaMemcpyHostToDevice, st));
    if (sd.nnz > 0) {
        MF_TRY(upload_bytes(sd.idx, idx + e0, sizeof(uint32_t) * (size_t)sd.nnz, st));
        MF_TRY(upload_bytes(sd.val, val + e0, sizeof(float) * (size_t)sd.nnz, st));
    }
    if (!keep) MF_CUDA(cudaStreamSynchronize(st));  // `local` goes out of scope
    return MF_OK;
}

int fetch_ptr(const uint32_t* src, int64_t n, std::vector<uint32_t>& out) {
    out.resize((size_t)n);
    MF_CUDA(cudaMemcpy(out.data(), src, sizeof(uint32_t) * (size_t)n, cudaMemcpyDefault));
    return MF_OK;
}

// ------------------------------------------------------------------------------------------
// CCD++ sweeps on a session
// ------------------------------------------------------------------------------------------
struct SweepVectors {
    const float* g_new = nullptr;
    const float* g_add = nullptr;
    const float* g_old = nullptr;
    const float* s_add = nullptr;
    const float* s_old = nullptr;
};

int family_of(int mode) {
    if (!(mode & kSolve)) return F_UPDATE;
    return (mode & (kSub | kAdd)) ? F_FUSED : F_SOLVE;
}

// one sweep of `mode` over side sd; when it solves, `out` (full-length vector) receives this shard's block
int run_sweep(mf_session* s, Side& sd, int mode, const SweepVectors& v, float* out, bool push = false) {
    if (sd.nnz == 0 && !(mode & kSolve)) return MF_OK;
    const int nmf = s->prm.nmf_project;
    if (s->panel) {
        PanelSweepArgs a;
        a.idx16 = sd.idx16; a.val = sd.pval; a.items = sd.items;
        a.cta_item_ptr = sd.cta_item_ptr; a.panel_item_ptr = sd.panel_item_ptr;
        a.npanels = sd.npanels; a.panel_rows = (uint32_t)sd.panel_rows; a.gdim = sd.gdim;
        a.seg_offset = sd.seg_offset;
        a.g_new = v.g_new; a.g_add = v.g_add; a.g_old = v.g_old; a.s_add = v.s_add; a.s_old = v.s_old;
        a.partials = sd.partials;
        a.nslots = 0;
        a.cta_start_ptr = sd.cta_start_ptr; a.ring_entries = 0;
        a.pf_dist = s->pf_dist; a.npad = (uint32_t)sd.npad;
        a.short_items = sd.short_items ? 1 : 0;
        a.trace_cta = (s->d_trace_cta && s->trace_n >= s->trace_cta_from && s->trace_n < s->trace_cta_from + 12) ? s->d_trace_cta + 4 * 256 * (size_t)(s->trace_n - s->trace_cta_from) : nullptr;
        a.trace = (s->d_trace && s->trace_n < s->trace_cap) ? s->d_trace + 8 * (size_t)s->trace_n++ : nullptr;
        a.fin.enabled = 0;
        const bool solve = (mode & kSolve) != 0;
        const bool is_h = push && out >= s->H && out < s->H + (int64_t)s->k * s->ldn;
        FinalizePush fp;
        if (solve && push) {  // fused solve -> exchange: the finalize sends the block to every peer (LL words over NVLink)
            fp.peer_ll = dist_peer_ll(s->dist, is_h);
            fp.peer_flags = dist_peer_flags(s->dist);
            fp.ticket = dist_flags(s->dist) + s->nranks;
            fp.vec_off = sd.seg_offset;
            fp.rank = s->rank; fp.nranks = s->nranks; fp.epoch = dist_next_epoch(s->dist); fp.barrier = 0;
        }
        // finalize inside the sweep kernel (grid barrier, register-ring pipeline) unless switched off
        const bool in_kernel = solve && sd.nitems > 0 && s->fin_in_kernel &&
                               (s->prm.pipeline == MF_PIPELINE_REGISTERS || s->prm.pipeline == MF_PIPELINE_STREAM);
        if (in_kernel) {
            SweepFinalize& f = a.fin;
            f.enabled = 1;
            f.lanes = panel_finalize_lanes(sd.nseg, sd.nslots);
            f.nseg = sd.nseg; f.slot_ptr = sd.slot_ptr; f.seg_ptr = sd.ptr;
            f.lambda = s->prm.lambda; f.nmf = nmf; f.out = out + sd.seg_offset;
            f.bar = s->d_gridbar; f.bar_target = s->gridbar_total + (unsigned)sd.ncta; f.status = s->d_gridbar + 1;
            f.peer_ll = push ? fp.peer_ll : nullptr;
            f.vec_off = sd.seg_offset; f.rank = s->rank; f.nranks = s->nranks; f.epoch = push ? fp.epoch : 0u;
            f.ll = push ? dist_ll(s->dist, is_h) : nullptr;  // receive the peers' blocks in the same launch
            f.vec = out; f.dim = is_h ? s->cols : s->rows; f.own_lo = sd.seg_offset; f.own_hi = sd.seg_offset + sd.nseg;
        }
        if (sd.nitems > 0) {
            s->timer.start(family_of(mode));
            MF_TRY(panel_sweep(mode, a, sd.ncta, s->prm.pipeline == MF_PIPELINE_REGISTERS ? panel_sweep_threads() : 1024, sd.chunk, s->prm.pipeline, s->st));
            if (in_kernel) s->gridbar_total += (unsigned)sd.ncta;  // only once the launch has been accepted
            s->timer.stop();
        }
        if (solve) {
            if (!in_kernel) {
                s->timer.start(F_FINALIZE);
                MF_TRY(panel_finalize(sd.nseg, sd.nslots, sd.slot_ptr, sd.partials, sd.ptr, s->prm.lambda, nmf, out + sd.seg_offset,
                                      push ? &fp : nullptr, s->st));
                s->timer.stop();
            }
            if (push && !in_kernel) {
                s->timer.start(F_COLLECTIVE);
                MF_TRY(exchange_unpack(dist_ll(s->dist, is_h), out, is_h ? s->cols : s->rows, sd.seg_offset, sd.seg_offset + sd.nseg,
                                       fp.epoch, s->st));
                s->timer.stop();
            }
        }
    } else {
        DirectSweepArgs a;
        a.nseg = sd.nseg; a.ptr = sd.ptr; a.idx = sd.idx; a.val = sd.val; a.seg_offset = sd.seg_offset;
        a.g_new = v.g_new; a.g_add = v.g_add; a.g_old = v.g_old; a.s_add = v.s_add; a.s_old = v.s_old;
        a.lambda = s->prm.lambda; a.nmf = nmf; a.out = out ? out + sd.seg_offset : nullptr;
        s->timer.start(family_of(mode));
        MF_TRY(direct_sweep(mode, a, s->sm_count, s->st));
        s->timer.stop();
    }
    return MF_OK;
}

int gather_blocks(mf_session* s, float* vec, const std::vector<int64_t>& bound, int64_t unit) {
    if (s->nranks <= 1) return MF_OK;
    s->timer.start(F_COLLECTIVE);
    MF_TRY(dist_allgather_blocks(s->dist, vec, bound, unit, s->st));
    s->timer.stop();
    return MF_OK;
}

// multi-GPU with peer mappings and the panel layout: the exchange is fused into the finalize kernel
bool fused_exchange(const mf_session* s) { return s->nranks > 1 && s->panel && dist_p2p(s->dist); }

// v = H[t] from the CSC copy and u = W[t]
int solve_v(mf_session* s, int t, int mode, const SweepVectors& v) {
    float* out = s->H + (int64_t)t * s->ldn;
    if (fused_exchange(s)) return run_sweep(s, s->csc, mode, v, out, true);
    MF_TRY(run_sweep(s, s->csc, mode, v, out));
    return gather_blocks(s, out, s->col_bound, 1);
}
int solve_u(mf_session* s, int t, int mode, const SweepVectors& v) {
    float* out = s->W + (int64_t)t * s->ldm;
    if (fused_exchange(s)) return run_sweep(s, s->csr, mode, v, out, true);
    MF_TRY(run_sweep(s, s->csr, mode, v, out));
    return gather_blocks(s, out, s->row_bound, 1);
}

// all ranks have finished everything they enqueued before this point (peer-to-peer path only): keeps a fast
// rank's next pushes from landing in a factor row a slower peer is still reading (RMSE, downloads)
int exchange_barrier(mf_session* s) {
    if (!fused_exchange(s)) return MF_OK;
    FinalizePush fp;
    fp.peer_ll = nullptr; fp.peer_flags = dist_peer_flags(s->dist); fp.ticket = dist_flags(s->dist) + s->nranks;
    fp.vec_off = 0; fp.rank = s->rank; fp.nranks = s->nranks; fp.epoch = dist_next_epoch(s->dist); fp.barrier = 1;
    s->timer.start(F_COLLECTIVE);
    MF_TRY(panel_finalize(0, 0, nullptr, nullptr, nullptr, 0.f, 0, nullptr, &fp, s->st));
    MF_TRY(exchange_wait(dist_flags(s->dist), s->rank, s->nranks, fp.epoch, s->st));
    s->timer.stop();
    return MF_OK;
}

// residual (+|-)= u_t v_t^T on both copies — UpdateRating on R then Rt, src/CCD.cpp:100-103 / :133-134
int update_both(mf_session* s, int t, bool add) {
    const float* u = s->W + (int64_t)t * s->ldm;
    const float* v = s->H + (int64_t)t * s->ldn;
    SweepVectors a, b;
    if (add) { a.g_new = u; a.s_add = v; b.g_new = v; b.s_add = u; }
    else     { a.g_old = u; a.s_old = v; b.g_old = v; b.s_old = u; }
    MF_TRY(run_sweep(s, s->csc, add ? kAdd : kSub, a, nullptr));
    MF_TRY(run_sweep(s, s->csr, add ? kAdd : kSub, b, nullptr));
    return MF_OK;
}

int flush_pending(mf_session* s) {
    if (s->pending < 0) return MF_OK;
    MF_TRY(update_both(s, s->pending, false));
    s->pending = -1;
    return MF_OK;
}

// ---- optional per-rank work (both inert in the reference) ----
// rank start: remember u_t (and v_t, which the fused schedule keeps anyway) for the incremental test RMSE
int rank_begin(mf_session* s, int t) {
    if (!s->rank_report) return MF_OK;
    MF_CUDA(cudaEventRecord(s->rank_ev[t], s->st));
    MF_CUDA(cudaMemcpyAsync(s->u_prev, s->W + (int64_t)t * s->ldm, sizeof(float) * (size_t)s->ldm, cudaMemcpyDeviceToDevice, s->st));
    MF_CUDA(cudaMemcpyAsync(s->v_old + (int64_t)t * s->ldn, s->H + (int64_t)t * s->ldn, sizeof(float) * (size_t)s->ldn, cudaMemcpyDeviceToDevice, s->st));
    return MF_OK;
}
// rank end (before v_old[t] is refreshed): calrmse_r1 with the rank's old and new vectors — src/CCD.cpp:144-146
int rank_end(mf_session* s, int t) {
    if (!s->rank_report) return MF_OK;
    s->timer.start(F_RMSE);
    MF_TRY(rmse_r1_accumulate(s->nt, s->trow, s->tcol, s->tres, s->W + (int64_t)t * s->ldm, s->H + (int64_t)t * s->ldn, s->u_prev,
                              s->v_old + (int64_t)t * s->ldn, s->d_acc, s->sm_count, s->st));
    s->timer.stop();
    MF_CUDA(cudaMemcpyAsync(s->d_rank_acc + t, s->d_acc, sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    MF_CUDA(cudaEventRecord(s->rank_ev[t + 1], s->st));
    return MF_OK;
}
// -e stop rule of CCDR1 (restated in oracle/mf_oracle.c, orc_ccdpp_ex): the function decrease of an inner iteration is the
// sum over both sweeps of (lambda*deg + h)(old - new)^2; the rank's inner iterations end after the one whose decrease is
// below eps * (largest decrease seen so far); the very first inner iteration of a run never enters the maximum.
int stop_rule_before(mf_session* s, const float* vec, int64_t n) {
    if (!s->early_stop) return MF_OK;
    MF_CUDA(cudaMemcpyAsync(s->vec_prev, vec, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s->st));
    return MF_OK;
}
int stop_rule_after(mf_session* s, const Side& sd, const float* vec, int which) {
    if (!s->early_stop) return MF_OK;
    MF_TRY(fundec_accumulate(sd.nseg, sd.ptr, sd.slot_ptr, sd.partials, s->prm.lambda, s->vec_prev + sd.seg_offset, vec + sd.seg_offset,
                             s->d_acc, s->sm_count, s->st));
    MF_CUDA(cudaMemcpyAsync(s->d_rank_acc + s->k + which, s->d_acc, sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    return MF_OK;
}
// true: stop the rank's inner iterations now (one host synchronisation per inner iteration: the rule is opt-in)
int stop_rule_decide(mf_session* s, int t, int it, bool* stop) {
    *stop = false;
    if (!s->early_stop) return MF_OK;
    double f[2] = {0, 0};
    MF_CUDA(cudaMemcpyAsync(f, s->d_rank_acc + s->k, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    MF_CUDA(cudaStreamSynchronize(s->st));
    const double cur = f[0] + f[1];
    if (cur < s->fundec_max * (double)s->prm.eps) { *stop = true; return MF_OK; }
    if (!(s->outer_done == 0 && t == 0 && it == 0)) s->fundec_max = std::max(s->fundec_max, cur);
    return MF_OK;
}

// one rank of one outer iteration, fused schedule: the deferred subtraction of the
// previous rank and this rank's add-back ride on the first solve sweep of each copy.
int ccd_rank_fused(mf_session* s, int t, bool add) {
    const int T = s->prm.maxinneriter;
    if (T <= 0) return MF_OK;
    float* u = s->W + (int64_t)t * s->ldm;
    float* v = s->H + (int64_t)t * s->ldn;
    const int sub = s->pending;
    const float* u_sub = sub >= 0 ? s->W + (int64_t)sub * s->ldm : nullptr;
    const float* v_sub = sub >= 0 ? s->H + (int64_t)sub * s->ldn : nullptr;
    const float* v_prev_iter = s->v_old + (int64_t)t * s->ldn;  // v_t as the previous outer iteration left it
    MF_TRY(rank_begin(s, t));
    bool stop = false;
    {   // CSC: [subtract `sub`] [add back t with the old (u_t, v_t)] solve v_t against u_t
        SweepVectors a;
        a.g_new = u; a.g_old = u_sub; a.s_add = v; a.s_old = v_sub;
        MF_TRY(stop_rule_before(s, v, s->ldn));
        MF_TRY(solve_v(s, t, kSolve | (sub >= 0 ? kSub : 0) | (add ? kAdd : 0), a));
        MF_TRY(stop_rule_after(s, s->csc, v, 0));
    }
    {   // CSR: [subtract `sub`] [add back t with the old v_t (saved) and old u_t] solve u_t against the new v_t
        SweepVectors a;
        a.g_new = v; a.g_add = v_prev_iter; a.s_add = u; a.s_old = u_sub;
        a.g_old = (sub == t) ? v_prev_iter : v_sub;  // k == 1: the subtracted rank's v was just overwritten
        MF_TRY(stop_rule_before(s, u, s->ldm));
        MF_TRY(solve_u(s, t, kSolve | (sub >= 0 ? kSub : 0) | (add ? (kAdd | kAddSep) : 0), a));
        MF_TRY(stop_rule_after(s, s->csr, u, 1));
    }
    MF_TRY(stop_rule_decide(s, t, 0, &stop));
    int done = 1;
    for (int it = 1; it < T && !stop; ++it, ++done) {
        SweepVectors a, b;
        a.g_new = u;
        MF_TRY(stop_rule_before(s, v, s->ldn));
        MF_TRY(solve_v(s, t, kSolve, a));
        MF_TRY(stop_rule_after(s, s->csc, v, 0));
        b.g_new = v;
        MF_TRY(stop_rule_before(s, u, s->ldm));
        MF_TRY(solve_u(s, t, kSolve, b));
        MF_TRY(stop_rule_after(s, s->csr, u, 1));
        MF_TRY(stop_rule_decide(s, t, it, &stop));
    }
    if (!s->rank_inner.empty()) s->rank_inner[t] = done;
    MF_TRY(rank_end(s, t));
    // keep v_t for the next outer iteration's add-back on the CSR copy (taken now: no rank writes H[t] again
    // before that, so the copy can never race with a peer's push)
    MF_CUDA(cudaMemcpyAsync(s->v_old + (int64_t)t * s->ldn, v, sizeof(float) * (size_t)s->ldn, cudaMemcpyDeviceToDevice, s->st));
    s->pending = t;
    return MF_OK;
}

// Device view of one side for the persistent kernel.
PersistSide persist_side(mf_session* s, const Side& sd, bool solves_h) {
    PersistSide p;
    p.idx16 = sd.idx16; p.val = sd.pval; p.items = sd.items; p.cta_item_ptr = sd.cta_item_ptr; p.panel_item_ptr = sd.panel_item_ptr;
    p.npanels = sd.npanels; p.panel_rows = (uint32_t)sd.panel_rows; p.gdim = sd.gdim; p.seg_offset = sd.seg_offset; p.nseg = sd.nseg;
    p.slot_ptr = sd.slot_ptr; p.seg_ptr = sd.ptr; p.partials = sd.partials;
    p.lanes = panel_finalize_lanes(sd.nseg, sd.nslots);
    const bool push = fused_exchange(s);
    p.peer_ll = push ? dist_peer_ll(s->dist, solves_h) : nullptr;
    p.ll = push ? dist_ll(s->dist, solves_h) : nullptr;
    p.dim = solves_h ? s->cols : s->rows;
    return p;
}

// May this session run an outer iteration as one persistent launch?  (panel layout, fused schedule, register-ring
// pipeline, co-resident grid with cooperative launch, and — multi-GPU — the peer-to-peer exchange)
bool use_persistent(const mf_session* s) {
    return s->persistent && !s->rank_report && !s->early_stop && s->panel && s->prm.schedule == MF_SCHEDULE_FUSED && s->prm.pipeline == MF_PIPELINE_REGISTERS &&
           s->prm.maxinneriter >= 1 && (s->nranks == 1 || fused_exchange(s));
}

// one whole outer iteration (k ranks, fused schedule) as ONE cooperative launch — replaces the loop of
// cuda_src/CCD_CUDA.cu:339-378.  Phase time stamps (%globaltimer at the grid barriers) feed the per-family times.
int ccd_outer_persistent(mf_session* s, bool add, bool timing) {
    const int k = s->k, T = s->prm.maxinneriter;
    const unsigned nphase = (unsigned)(2 * k * T);
    PersistArgs a;
    a.csc = persist_side(s, s->csc, true);
    a.csr = persist_side(s, s->csr, false);
    a.W = s->W; a.H = s->H; a.v_old = s->v_old; a.ldm = s->ldm; a.ldn = s->ldn;
    a.k = k; a.T = T; a.add = add ? 1 : 0; a.pending = s->pending;
    a.lambda = s->prm.lambda; a.nmf = s->prm.nmf_project;
    a.bar = s->d_gridbar; a.bar_base = s->gridbar_total; a.status = s->d_gridbar + 1;
    a.rank = s->rank; a.nranks = s->nranks;
    a.epoch_base = fused_exchange(s) ? dist_advance_epoch(s->dist, nphase) : 0u;
    a.stamps = nullptr;
    if (timing) {
        if (!s->d_stamps) MF_TRY(dev_alloc(&s->d_stamps, (size_t)nphase + 1));
        a.stamps = s->d_stamps;
    }
    const int ncta = s->csc.ncta;
    s->timer.start(F_PERSIST);
    MF_TRY(ccd_persistent_launch(a, ncta, s->persist_smem, s->st));
    s->timer.stop();
    s->gridbar_total += 2u * nphase * (unsigned)ncta;  // two grid barriers per phase (advanced only after a successful launch)
    s->pending = k - 1;
    if (timing) {
        s->h_stamps.resize((size_t)nphase + 1);
        MF_CUDA(cudaMemcpyAsync(s->h_stamps.data(), s->d_stamps, sizeof(unsigned long long) * ((size_t)nphase + 1), cudaMemcpyDeviceToHost, s->st));
    }
    return MF_OK;
}

// folds the phase stamps of the last persistent launch into the per-family times (call after the stream is synchronised)
void fold_stamps(mf_session* s, bool add, bool had_pending) {
    const int k = s->k, T = s->prm.maxinneriter;
    if (s->h_stamps.size() != (size_t)(2 * k * T + 1)) return;
    size_t n = 0;
    for (int t = 0; t < k; ++t)
        for (int it = 0; it < T; ++it)
            for (int side = 0; side < 2; ++side, ++n) {
                const bool fused = it == 0 && (add || t > 0 || had_pending);
                const int fam = fused ? F_FUSED : F_SOLVE;
                s->fam_seconds[fam] += (double)(s->h_stamps[n + 1] - s->h_stamps[n]) * 1e-9;
                s->fam_launches[fam] += 1;
            }
    s->h_stamps.clear();
}

// after a synchronisation: did a device-side wait time out?  (persistent kernel; status word next to the barrier counter)
int check_device_status(mf_session* s) {
    if (!s->panel || !s->d_gridbar) return MF_OK;
    unsigned st = 0;
    MF_CUDA(cudaMemcpy(&st, s->d_gridbar + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (st == 0) return MF_OK;
    s->broken = true;
    set_error("device-side wait timed out (%s): another kernel held SMs of this device, or a peer rank stopped; the session is unusable",
              st == 1 ? "grid barrier" : "multi-GPU exchange");
    return MF_ERR_STATE;
}

// the reference's launch order: add-back, T x (v-solve, u-solve), subtract — CCD_CUDA.cu:347-378
int ccd_rank_reference(mf_session* s, int t, bool add) {
    const int T = s->prm.maxinneriter;
    float* u = s->W + (int64_t)t * s->ldm;
    float* v = s->H + (int64_t)t * s->ldn;
    MF_TRY(rank_begin(s, t));
    if (add) MF_TRY(update_both(s, t, true));
    bool stop = false;
    int done = 0;
    for (int it = 0; it < T && !stop; ++it, ++done) {
        SweepVectors a, b;
        a.g_new = u;
        MF_TRY(stop_rule_before(s, v, s->ldn));
        MF_TRY(solve_v(s, t, kSolve, a));
        MF_TRY(stop_rule_after(s, s->csc, v, 0));
        b.g_new = v;
        MF_TRY(stop_rule_before(s, u, s->ldm));
        MF_TRY(solve_u(s, t, kSolve, b));
        MF_TRY(stop_rule_after(s, s->csr, u, 1));
        MF_TRY(stop_rule_decide(s, t, it, &stop));
    }
    if (!s->rank_inner.empty()) s->rank_inner[t] = done;
    MF_TRY(rank_end(s, t));
    return update_both(s, t, false);
}

int session_rmse(mf_session* s, double* out, double* seconds) {
    if (s->nt <= 0) { *out = NAN; if (seconds) *seconds = 0; return MF_OK; }
    MF_CUDA(cudaEventRecord(s->ev_c, s->st));
    s->timer.start(F_RMSE);
    if (s->prm.solver_type == MF_SOLVER_ALS)
        MF_TRY(rmse_accumulate(s->nt, s->trow, s->tcol, s->tval, s->W, s->H, s->k, 1, s->k, 1, s->k, s->d_acc, s->sm_count, s->st));
    else
        MF_TRY(rmse_accumulate(s->nt, s->trow, s->tcol, s->tval, s->W, s->H, s->k, s->ldm, 1, s->ldn, 1, s->d_acc, s->sm_count, s->st));
    s->timer.stop();
    double acc = 0.0;
    MF_CUDA(cudaMemcpyAsync(&acc, s->d_acc, sizeof(double), cudaMemcpyDeviceToHost, s->st));
    MF_CUDA(cudaEventRecord(s->ev_b, s->st));
    MF_CUDA(cudaStreamSynchronize(s->st));
    *out = sqrt(acc / (double)s->nt);
    if (seconds) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->ev_c, s->ev_b);
        *seconds = ms * 1e-3;
    }
    return MF_OK;
}

int check_ratings(const mf_ratings* R) {
    MF_REQUIRE(R != nullptr, "ratings pointer is NULL");
    MF_REQUIRE(R->rows > 0 && R->cols > 0 && R->nnz >= 0, "bad shape %lld x %lld, nnz %lld", (long long)R->rows,
               (long long)R->cols, (long long)R->nnz);
    MF_REQUIRE(R->rows < ((int64_t)1 << 32) && R->cols < ((int64_t)1 << 32) && R->nnz < ((int64_t)1 << 32),
               "indices are uint32 (src/pmf_util.h:146-148): shape out of range");
    MF_REQUIRE(R->csr_row_ptr && R->csc_col_ptr, "ptr arrays are NULL");
    MF_REQUIRE(R->nnz == 0 || (R->csr_col_idx && R->csr_val && R->csc_row_idx && R->csc_val), "index/value arrays are NULL");
    return MF_OK;
}

int create_impl(const mf_ratings* R, const mf_testset* T, const mf_params* params, int rank, int nranks,
                const void* nccl_id, mf_session** out) {
    trace_mark("(enter session create)");
    MF_REQUIRE(out != nullptr && params != nullptr, "NULL argument");
    *out = nullptr;
    MF_TRY(check_ratings(R));
    MF_REQUIRE(params->k >= 1, "k must be >= 1");
    MF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    MF_REQUIRE(params->solver_type == MF_SOLVER_CCD || params->solver_type == MF_SOLVER_ALS, "bad solver_type");
    int ndev = 0;
    MF_CUDA(cudaGetDeviceCount(&ndev));
    MF_REQUIRE(params->device >= 0 && params->device < ndev, "device %d not present (%d CUDA devices)", params->device, ndev);
    MF_CUDA(cudaSetDevice(params->device));

    mf_session* s = new mf_session();
    s->prm = *params;
    if (const char* e = getenv("MF_L2_PREFETCH")) s->pf_dist = (uint32_t)(atoi(e) / 8 * 8);
    if (getenv("MF_SWEEP_TRACE")) {
        s->trace_cap = 4096;
        if (cudaMalloc(&s->d_trace, sizeof(unsigned long long) * 8 * (size_t)s->trace_cap) != cudaSuccess) { s->d_trace = nullptr; s->trace_cap = 0; cudaGetLastError(); }
        else cudaMemset(s->d_trace, 0, sizeof(unsigned long long) * 8 * (size_t)s->trace_cap);
        if (const char* e = getenv("MF_SWEEP_TRACE_CTA")) {  // per-CTA stamps of 12 launches starting at this launch number
            s->trace_cta_from = atoi(e);
            if (cudaMalloc(&s->d_trace_cta, sizeof(unsigned long long) * 4 * 256 * 12) != cudaSuccess) { s->d_trace_cta = nullptr; cudaGetLastError(); }
            else cudaMemset(s->d_trace_cta, 0, sizeof(unsigned long long) * 4 * 256 * 12);
        }
    }
    if (const char* e = getenv("MF_PIPELINE")) {  // A/B switch for runs that do not set mf_params.pipeline themselves
        if (!strcmp(e, "stream")) s->prm.pipeline = MF_PIPELINE_STREAM;
        else if (!strcmp(e, "registers")) s->prm.pipeline = MF_PIPELINE_REGISTERS;
    }
    s->device = params->device;
    s->rank = rank;
    s->nranks = nranks;
    s->rows = R->rows; s->cols = R->cols; s->nnz = R->nnz;
    s->k = (int)params->k;
    int rc = MF_OK;
    auto fail = [&](int code) { mf_session_destroy(s); return code; };

    if (cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device) != cudaSuccess) { set_error("cudaDeviceGetAttribute failed"); return fail(MF_ERR_CUDA); }
    {   // keep freed scratch in the stream-ordered pool instead of returning it to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(MF_ERR_CUDA); }
    cudaEventCreate(&s->ev_a); cudaEventCreate(&s->ev_b); cudaEventCreate(&s->ev_c);
    s->timer.st = s->st;
    s->timer.enabled = true;

    trace_mark("device/stream setup");
    std::vector<uint32_t> rp, cp;
    if ((rc = fetch_ptr(R->csr_row_ptr, R->rows + 1, rp)) != MF_OK) return fail(rc);
    if ((rc = fetch_ptr(R->csc_col_ptr, R->cols + 1, cp)) != MF_OK) return fail(rc);
    if (rp[0] != 0 || cp[0] != 0 || rp[R->rows] != (uint32_t)R->nnz || cp[R->cols] != (uint32_t)R->nnz) {
        set_error("ptr arrays inconsistent with nnz (ptr[0]=%u/%u, ptr[last]=%u/%u, nnz=%lld)", rp[0], cp[0], rp[R->rows],
                  cp[R->cols], (long long)R->nnz);
        return fail(MF_ERR_ARG);
    }
    if (!std::is_sorted(rp.begin(), rp.end()) || !std::is_sorted(cp.begin(), cp.end())) {
        set_error("ptr arrays must be non-decreasing");
        return fail(MF_ERR_ARG);
    }
    partition_host(rp, nranks, s->row_bound);
    partition_host(cp, nranks, s->col_bound);
    {   // one device arena for this shard's ratings and layout arrays: ~36 B per rating (two raw copies, two panel
        // copies with padding) + 12 B per (panel, segment) piece + slack; it grows by chunks if the estimate is short
        const int64_t nnz_r = (int64_t)rp[s->row_bound[rank + 1]] - (int64_t)rp[s->row_bound[rank]];
        const int64_t nnz_c = (int64_t)cp[s->col_bound[rank + 1]] - (int64_t)cp[s->col_bound[rank]];
        const int64_t seg_r = s->row_bound[rank + 1] - s->row_bound[rank], seg_c = s->col_bound[rank + 1] - s->col_bound[rank];
        const int64_t pieces = seg_r * (R->cols / 16376 + 1) + seg_c * (R->rows / 16376 + 1);
        const int64_t factors = nranks > 1 ? 0 : 4 * (int64_t)params->k * (R->rows + 2 * R->cols + 96);
        const size_t hint = (size_t)(18 * (nnz_r + nnz_c) + 12 * pieces + 24 * (seg_r + seg_c) + factors + 12 * (T ? T->nnz : 0)) + ((size_t)64 << 20);
        s->arena = arena_create(hint, s->st);
        trace_mark("  arena chunk");
    }
    ArenaScope arena_scope(s->arena);
    // The CSC copy goes up first; the CSR copy follows on a second stream while the CSC copy is checked and its panel
    // layout is built (the first sweep of a rank runs on the CSC copy, and the upload is the longest part of a short call).
    cudaStream_t st_up = nullptr;
    std::vector<uint32_t> csr_local;
    if (cudaStreamCreateWithFlags(&st_up, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(MF_ERR_CUDA); }
    auto fail_up = [&](int code) { cudaStreamSynchronize(st_up); cudaStreamDestroy(st_up); return fail(code); };
    auto csr_arrived = [&]() {
        const cudaError_t e = cudaStreamSynchronize(st_up);
        if (e != cudaSuccess) { set_error("upload of the CSR copy failed: %s", cudaGetErrorString(e)); return (int)MF_ERR_CUDA; }
        return (int)MF_OK;
    };
    if ((rc = upload_side(s->csc, cp, s->col_bound[rank], s->col_bound[rank + 1], R->csc_row_idx, R->csc_val, R->rows, s->st)) != MF_OK) return fail_up(rc);
    if ((rc = upload_side(s->csr, rp, s->row_bound[rank], s->row_bound[rank + 1], R->csr_col_idx, R->csr_val, R->cols, st_up, &csr_local)) != MF_OK) return fail_up(rc);

    trace_mark("upload CSC (CSR in flight)");
    const bool ccd = params->solver_type == MF_SOLVER_CCD;
    if (ccd) {
        s->panel = params->layout == MF_LAYOUT_PANEL;
        bool ok_r = true, ok_c = true, range_r = true, range_c = true;
        // every index must be an index (both layouts gather with it); the panel cut also needs them ascending inside a
        // segment — the reference does not, so a copy that is not sorted is sorted here, once, instead of being refused
        if ((rc = side_check_sorted(s->csc, &ok_c, &range_c, s->st)) != MF_OK) return fail_up(rc);
        if (!range_c) { set_error("CSC copy: a row index is >= rows (%lld)", (long long)R->rows); return fail_up(MF_ERR_ARG); }
        if (s->panel && !ok_c && (rc = side_sort_segments(s->csc, s->st)) != MF_OK) return fail_up(rc);
        trace_mark("  index check (CSC)");
        if (s->panel) {
            // Panel size.  Shared memory a sweep needs: CSC side up to 2 vectors (u_new, u_old), CSR side up to 3
            // (v_new, v_add, v_old).  Measured (profiles/README.md): panels that fill shared memory leave the SM
            // almost no L1 for the rating streams and the updating sweeps slow down by 30-60 %; 12-16 K entries
            // (48-64 KB per vector) is the sweet spot on B200.  Hard cap 16376: idx16 stores index*4.
            int cap_c = std::min(panel_cap(2), 16376), cap_r = std::min(panel_cap(3), 16376);
            // STREAM pipeline: the ring of rating tiles (16 K entries, 96 KB) lives beside the staged vectors
            if (s->prm.pipeline == MF_PIPELINE_STREAM) cap_r = std::min(cap_r, 11112);
            const int want = params->panel_rows > 0 ? params->panel_rows / 8 * 8 : 16376;
            cap_c = std::min(cap_c, want);
            cap_r = std::min(cap_r, want);
            const int chunk = params->chunk > 0 ? std::max(8, params->chunk / 8 * 8) : 512;
            const int pr_c = choose_panel_rows(s->csc.gdim, std::max(cap_c, 8)), pr_r = choose_panel_rows(s->csr.gdim, std::max(cap_r, 8));
            // Padding granularity of a piece: 32 entries (whole 64/128-byte lines per 8-lane group) unless the pieces are
            // so short that the padding would dominate the traffic (very sparse shapes cut by many panels, e.g. the
            // Yahoo-Music shape: ~6 entries per piece; measured 2.55 s -> 2.04 s per outer iteration with 16)
            auto pick_pad = [&](const Side& sd, int pr) {
                if (params->pad_entries > 0) return (int)params->pad_entries;
                if (s->prm.pipeline == MF_PIPELINE_STREAM) return 8;  // contiguous bulk copies: items need no line alignment
                const int64_t npan = (sd.gdim + pr - 1) / pr;
                const int64_t pieces = std::max<int64_t>(1, std::min<int64_t>(sd.nseg * npan, std::max<int64_t>(sd.nnz, 1)));
                return sd.nnz / pieces < 24 ? 16 : 32;
            };
            s->csc.pad = pick_pad(s->csc, pr_c);
            s->csr.pad = pick_pad(s->csr, pr_r);
            {   // short-piece copies (same criterion as the 16-entry padding: fewer than 24 entries per piece on average): one item
                // per lane, and the padding can drop to 8 entries — an item's loads no longer belong to an 8-lane group
                auto is_short = [&](const Side& sd, int pr) {
                    const int64_t npan = (sd.gdim + pr - 1) / pr;
                    const int64_t pieces = std::max<int64_t>(1, std::min<int64_t>(sd.nseg * npan, std::max<int64_t>(sd.nnz, 1)));
                    return sd.nnz / pieces < 24;
                };
                const char* e = getenv("MF_SHORT_ITEMS");  // A/B switch: 0 = off, 1 = forced on
                for (Side* sd : {&s->csc, &s->csr}) {
                    const int pr = sd == &s->csc ? pr_c : pr_r;
                    sd->short_items = e ? atoi(e) != 0 : is_short(*sd, pr);
                    if (sd->short_items && params->pad_entries <= 0 && s->prm.pipeline != MF_PIPELINE_STREAM) sd->pad = getenv("MF_SHORT_PAD") ? atoi(getenv("MF_SHORT_PAD")) : 8;
                }
            }
            {   // panel-entry charge of the work partition (prep.cu), in units of 8 entries; measured optimum on the Netflix shape between
                // 4000 and 16000 (flat): the CTAs that span a panel boundary then finish with the others instead of 7-22 us later
                int pc = 8000;
                if (const char* e = getenv("MF_PANEL_COST")) pc = atoi(e);
                s->csc.panel_cost = pc;
                s->csr.panel_cost = pc;
            }
            if ((rc = side_build_panels(s->csc, pr_c, chunk, s->sm_count, s->st)) != MF_OK) return fail_up(rc);
            if ((rc = csr_arrived()) != MF_OK) return fail_up(rc);
            if ((rc = side_check_sorted(s->csr, &ok_r, &range_r, s->st)) != MF_OK) return fail_up(rc);
            if (!range_r) { set_error("CSR copy: a column index is >= cols (%lld)", (long long)R->cols); return fail_up(MF_ERR_ARG); }
            if (!ok_r && (rc = side_sort_segments(s->csr, s->st)) != MF_OK) return fail_up(rc);
            if ((rc = side_build_panels(s->csr, pr_r, chunk, s->sm_count, s->st)) != MF_OK) return fail_up(rc);
            trace_mark("  build both panel layouts");
            // the caller-order index/value arrays are no longer needed: the residual lives in the panel arrays
            dev_free(s->csc.idx); s->csc.idx = nullptr; dev_free(s->csc.val); s->csc.val = nullptr;
            dev_free(s->csr.idx); s->csr.idx = nullptr; dev_free(s->csr.val); s->csr.val = nullptr;
        } else {
            if ((rc = csr_arrived()) != MF_OK) return fail_up(rc);
            if ((rc = side_check_sorted(s->csr, &ok_r, &range_r, s->st)) != MF_OK) return fail_up(rc);
            if (!range_r) { set_error("CSR copy: a column index is >= cols (%lld)", (long long)R->cols); return fail_up(MF_ERR_ARG); }
        }
        if ((rc = csr_arrived()) != MF_OK) return fail_up(rc);
        trace_mark("sortedness + panel layout");
        if (nranks > 1) arena_bind(nullptr);  // multi-GPU: the factor matrices are exported to the peers through CUDA IPC — allocations of their own
        s->ldm = round_up(s->rows, 32);
        s->ldn = round_up(s->cols, 32);
        if ((rc = dev_alloc(&s->W, (size_t)s->k * s->ldm)) != MF_OK) return fail_up(rc);
        if ((rc = dev_alloc(&s->H, (size_t)s->k * s->ldn)) != MF_OK) return fail_up(rc);
        if ((rc = dev_alloc(&s->v_old, (size_t)s->k * s->ldn)) != MF_OK) return fail_up(rc);
        if (cudaMemsetAsync(s->W, 0, sizeof(float) * (size_t)s->k * s->ldm, s->st) != cudaSuccess ||
            cudaMemsetAsync(s->H, 0, sizeof(float) * (size_t)s->k * s->ldn, s->st) != cudaSuccess ||
            cudaMemsetAsync(s->v_old, 0, sizeof(float) * (size_t)s->k * s->ldn, s->st) != cudaSuccess) {
            set_error("cudaMemsetAsync of the factor matrices failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail_up(MF_ERR_CUDA);
        }
    } else {
        if ((rc = csr_arrived()) != MF_OK) return fail_up(rc);
        {   // ALS gathers factor rows with the indices: they must be indices
            bool sorted = true, in_c = true, in_r = true;
            if ((rc = side_check_sorted(s->csc, &sorted, &in_c, s->st)) != MF_OK) return fail_up(rc);
            if ((rc = side_check_sorted(s->csr, &sorted, &in_r, s->st)) != MF_OK) return fail_up(rc);
            if (!in_c || !in_r) { set_error("%s copy: an index is outside the %lld x %lld matrix", in_c ? "CSR" : "CSC", (long long)R->rows, (long long)R->cols); return fail_up(MF_ERR_ARG); }
        }
        if (nranks > 1) arena_bind(nullptr);
        s->ldm = s->k; s->ldn = s->k;
        if ((rc = dev_alloc(&s->W, (size_t)s->rows * s->k)) != MF_OK) return fail_up(rc);
        if ((rc = dev_alloc(&s->H, (size_t)s->cols * s->k)) != MF_OK) return fail_up(rc);
        if (cudaMemsetAsync(s->W, 0, sizeof(float) * (size_t)s->rows * s->k, s->st) != cudaSuccess ||
            cudaMemsetAsync(s->H, 0, sizeof(float) * (size_t)s->cols * s->k, s->st) != cudaSuccess) {
            set_error("cudaMemsetAsync of the factor matrices failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail_up(MF_ERR_CUDA);
        }
    }

    cudaStreamDestroy(st_up);
    trace_mark("  factor allocations");
    arena_bind(s->arena);
    s->nt = T ? T->nnz : 0;
    if (s->nt > 0) {
        if (!(T->row && T->col && T->val)) { set_error("test arrays are NULL"); return fail(MF_ERR_ARG); }
        if ((rc = dev_alloc(&s->trow, (size_t)s->nt)) != MF_OK) return fail(rc);
        if ((rc = dev_alloc(&s->tcol, (size_t)s->nt)) != MF_OK) return fail(rc);
        if ((rc = dev_alloc(&s->tval, (size_t)s->nt)) != MF_OK) return fail(rc);
        if ((rc = upload_bytes(s->trow, T->row, sizeof(uint32_t) * (size_t)s->nt, s->st)) != MF_OK) return fail(rc);
        if ((rc = upload_bytes(s->tcol, T->col, sizeof(uint32_t) * (size_t)s->nt, s->st)) != MF_OK) return fail(rc);
        if ((rc = upload_bytes(s->tval, T->val, sizeof(float) * (size_t)s->nt, s->st)) != MF_OK) return fail(rc);
        bool ok_row = true, ok_col = true;
        if ((rc = check_below(s->trow, s->nt, (uint64_t)s->rows, &ok_row, s->st)) != MF_OK) return fail(rc);
        if ((rc = check_below(s->tcol, s->nt, (uint64_t)s->cols, &ok_col, s->st)) != MF_OK) return fail(rc);
        if (!ok_row || !ok_col) { set_error("test set: a %s index is outside the %lld x %lld matrix", ok_row ? "column" : "row", (long long)s->rows, (long long)s->cols); return fail(MF_ERR_ARG); }
    }
    if (ccd) {
        if (s->prm.do_nmf) s->prm.nmf_project = 1;  // -N: the clamp of the finalize (inert in the reference, src/pmf.h:36)
        s->rank_inner.assign((size_t)s->k, 0);
        s->early_stop = s->prm.early_stop != 0;
        if (s->early_stop && (nranks > 1 || !s->panel)) {
            set_error("early_stop needs a single-GPU session with the panel layout");
            return fail(MF_ERR_UNSUPPORTED);
        }
        s->rank_report = s->prm.verbose != 0 && s->prm.do_predict != 0 && s->nt > 0;
        if (s->rank_report) {
            if ((rc = dev_alloc(&s->tres, (size_t)s->nt)) != MF_OK) return fail(rc);
            if ((rc = dev_alloc(&s->u_prev, (size_t)s->ldm)) != MF_OK) return fail(rc);
            // the reference's calrmse_r1 starts from the raw test ratings: training starts at H = 0 (src/CCD.cpp:63-67)
            if (cudaMemcpyAsync(s->tres, s->tval, sizeof(float) * (size_t)s->nt, cudaMemcpyDeviceToDevice, s->st) != cudaSuccess) { set_error("cudaMemcpyAsync failed"); return fail(MF_ERR_CUDA); }
            s->rank_ev.resize((size_t)s->k + 1);
            for (auto& e : s->rank_ev) cudaEventCreate(&e);
            s->rank_seconds.assign((size_t)s->k, 0.0);
            s->rank_rmse.assign((size_t)s->k, 0.0);
        }
        if (s->early_stop && (rc = dev_alloc(&s->vec_prev, (size_t)std::max(s->ldm, s->ldn))) != MF_OK) return fail(rc);
        if ((s->rank_report || s->early_stop) && (rc = dev_alloc(&s->d_rank_acc, (size_t)s->k + 2)) != MF_OK) return fail(rc);
    }
    trace_mark("  test set");
    if ((rc = dev_alloc(&s->d_acc, rmse_scratch_doubles(s->sm_count))) != MF_OK) return fail(rc);
    if ((rc = dev_alloc(&s->d_gridbar, 2)) != MF_OK) return fail(rc);
    if (cudaMemsetAsync(s->d_gridbar, 0, 2 * sizeof(unsigned), s->st) != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(MF_ERR_CUDA); }
    // finalize inside the sweep kernel needs every CTA of a sweep resident at once (grid barrier): checked, not assumed
    s->fin_in_kernel = getenv("MF_SEPARATE_FINALIZE") == nullptr && s->panel &&
                       panel_sweep_grid_resident(std::max(s->csc.ncta, s->csr.ncta), panel_sweep_threads(), std::max(s->csc.panel_rows, s->csr.panel_rows), s->sm_count);
    // opt-in (MF_PERSISTENT=1): on one GPU the per-launch kernels are faster, because every phase of the persistent kernel
    // runs with the shared-memory carve-out of the largest one (profiles/README.md, round 2)
    if (ccd && s->panel && getenv("MF_PERSISTENT") != nullptr && atoi(getenv("MF_PERSISTENT")) != 0 && s->csc.ncta == s->csr.ncta) {
        // the persistent kernel holds the largest panel footprint of any phase: CSC side up to 2 vectors, CSR side up to 3
        s->persist_smem = std::max(panel_sweep_smem(kSolve | kSub | kAdd, s->csc.panel_rows), panel_sweep_smem(kSolve | kSub | kAdd | kAddSep, s->csr.panel_rows));
        s->persistent = ccd_persistent_supported(s->csc.ncta, s->persist_smem, s->device);
    }
    arena_bind(nullptr);
    if (nranks > 1) {
        if (!nccl_id) { set_error("multi-GPU session needs the shared ncclUniqueId"); return fail(MF_ERR_ARG); }
        if ((rc = dist_create(&s->dist, rank, nranks, nccl_id, s->device)) != MF_OK) return fail(rc);
        if (ccd && s->panel && (rc = dist_setup_p2p(s->dist, s->W, s->H, s->ldm, s->ldn, s->st)) != MF_OK) return fail(rc);
    }
    trace_mark("  scratch + comm");
    cudaError_t e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) { set_error("session setup failed: %s", cudaGetErrorString(e)); return fail(MF_ERR_CUDA); }
    trace_mark("factors + test set + comm");
    *out = s;
    return MF_OK;
}

// Algorithmic HBM bytes of one sweep launch on this session: per rating entry 2 B of index + 4 B of value
// (+ 4 B written back by an updating sweep) in the panel layout — padding entries are NOT counted, they are
// overhead of the layout — plus the work-item descriptors, the partial-sum slots, the segment pointers and the
// staged factor vectors.  DIRECT layout: 4 B indices.
int64_t sweep_bytes(const mf_session* s, const Side& sd, int mode) {
    const bool write = mode & (kSub | kAdd);
    if (s->panel) {
        int64_t b = sd.nnz * (2 + 4 + (write ? 4 : 0)) + sd.nitems * 16;
        if (mode & kSolve) b += sd.nslots * 8 * 2 + sd.nseg * (4 + 4 + 4);
        return b + (int64_t)panel_sweep_vectors(mode) * sd.gdim * 4;
    }
    int64_t b = sd.nnz * (4 + 4 + (write ? 4 : 0)) + sd.nseg * 4;
    if (mode & kSolve) b += sd.nseg * 4;
    return b + sd.gdim * 4;
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" {

int mf_abi_version(void) { return MF_ABI_VERSION; }
const char* mf_last_error(void) { return mf::get_error(); }

int mf_device_count(int* count) {
    MF_REQUIRE(count != nullptr, "NULL argument");
    *count = 0;
    MF_CUDA(cudaGetDeviceCount(count));
    return MF_OK;
}

void mf_params_default(mf_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->solver_type = MF_SOLVER_CCD;  // src/pmf.h:27-40
    p->k = 10;
    p->threads = 4;
    p->maxiter = 5;
    p->maxinneriter = 1;
    p->lambda = 0.1f;
    p->eps = 1e-3f;
    p->nBlocks = 32;
    p->nThreadsPerBlock = 256;
}

void mf_host_initial_col(float* X, int64_t k, int64_t n) {
    if (!X) return;
    srand(0L);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < k; ++j) X[j * n + i] = 0.1f * (float(rand()) / RAND_MAX) + 0.001f;
}

int mf_release_cached_memory(int device) {
    int ndev = 0;
    MF_CUDA(cudaGetDeviceCount(&ndev));
    MF_REQUIRE(device >= 0 && device < ndev, "device %d not present (%d CUDA devices)", device, ndev);
    MF_CUDA(cudaSetDevice(device));
    MF_CUDA(cudaDeviceSynchronize());
    dist_release_cached(device);
    upload_release_cached();
    cudaMemPool_t pool;
    MF_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    MF_CUDA(cudaMemPoolTrimTo(pool, 0));
    return MF_OK;
}

int mf_session_create(const mf_ratings* R, const mf_testset* T, const mf_params* params, mf_session** out) {
    return create_impl(R, T, params, 0, 1, nullptr, out);
}

int mf_dist_unique_id(void* id128) { return dist_unique_id(id128); }

int mf_session_create_dist(const mf_ratings* R, const mf_testset* T, const mf_params* params, int rank, int nranks,
                           const void* nccl_unique_id, mf_session** out) {
    return create_impl(R, T, params, rank, nranks, nccl_unique_id, out);
}

int mf_session_destroy(mf_session* s) {
    if (!s) return MF_OK;
    cudaSetDevice(s->device);
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->d_trace) {  // dump the sweep timeline: one line per launch, seven numbers (ns relative to the launch's entry stamp; mode last)
        if (const char* path = getenv("MF_SWEEP_TRACE")) {
            std::vector<unsigned long long> h(8 * (size_t)s->trace_n);
            char name[1200];
            snprintf(name, sizeof(name), "%s.rank%d", path, s->rank);
            FILE* fp = s->trace_n > 0 && cudaMemcpy(h.data(), s->d_trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) == cudaSuccess ? fopen(name, "a") : nullptr;
            if (fp) {
                fprintf(fp, "# session of %d launches: entry_abs_ns after_wait after_items after_barrier after_finalize after_unpack mode\n", s->trace_n);
                for (int i = 0; i < s->trace_n; ++i) {
                    const unsigned long long* r = h.data() + 8 * (size_t)i;
                    fprintf(fp, "%llu %lld %lld %lld %lld %lld %llu\n", r[0], (long long)(r[1] - r[0]), (long long)(r[2] - r[0]), r[3] ? (long long)(r[3] - r[0]) : -1,
                            r[4] ? (long long)(r[4] - r[0]) : -1, r[5] ? (long long)(r[5] - r[0]) : -1, r[6]);
                }
                fclose(fp);
            }
        }
        cudaFree(s->d_trace);
        if (s->d_trace_cta) {
            const char* path = getenv("MF_SWEEP_TRACE");
            std::vector<unsigned long long> h(4 * 256 * 12);
            char name[1200];
            snprintf(name, sizeof(name), "%s.cta.rank%d", path ? path : "trace", s->rank);
            FILE* fp = cudaMemcpy(h.data(), s->d_trace_cta, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) == cudaSuccess ? fopen(name, "w") : nullptr;
            if (fp) {
                fprintf(fp, "# launch cta start_ns(rel to launch min) end_ns ib ie first_panel\n");
                for (int l = 0; l < 12; ++l) {
                    unsigned long long t0 = ~0ull;
                    for (int c = 0; c < 256; ++c) if (h[4 * (256 * l + c)] && h[4 * (256 * l + c)] < t0) t0 = h[4 * (256 * l + c)];
                    for (int c = 0; c < 256; ++c) {
                        const unsigned long long* r = h.data() + 4 * (256 * (size_t)l + c);
                        if (!r[0]) continue;
                        fprintf(fp, "%d %d %llu %llu %llu %llu %llu\n", s->trace_cta_from + l, c, r[0] - t0, r[1] - t0, r[2] >> 32, r[2] & 0xffffffffull, r[3]);
                    }
                }
                fclose(fp);
            }
            cudaFree(s->d_trace_cta);
        }
    }
    if (s->dist) dist_destroy(s->dist);
    side_free(s->csc);
    side_free(s->csr);
    for (auto& e : s->rank_ev) cudaEventDestroy(e);
    void* ptrs[] = {s->W, s->H, s->v_old, s->trow, s->tcol, s->tval, s->d_acc, s->d_gridbar, s->d_stamps, s->tres, s->u_prev, s->vec_prev, s->d_rank_acc};
    for (void* p : ptrs)
        if (p) dev_free(p);
    arena_destroy(s->arena);
    s->arena = nullptr;
    s->timer.destroy();
    if (s->ev_a) cudaEventDestroy(s->ev_a);
    if (s->ev_b) cudaEventDestroy(s->ev_b);
    if (s->ev_c) cudaEventDestroy(s->ev_c);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
    return MF_OK;
}

int mf_session_set_factors(mf_session* s, const float* W, const float* H) {
    MF_REQUIRE(s && W, "NULL argument");
    MF_CUDA(cudaSetDevice(s->device));
    if (s->prm.solver_type == MF_SOLVER_CCD) {
        MF_CUDA(cudaMemcpy2DAsync(s->W, sizeof(float) * s->ldm, W, sizeof(float) * s->rows, sizeof(float) * s->rows, s->k, cudaMemcpyDefault, s->st));
        if (H) MF_CUDA(cudaMemcpy2DAsync(s->H, sizeof(float) * s->ldn, H, sizeof(float) * s->cols, sizeof(float) * s->cols, s->k, cudaMemcpyDefault, s->st));
        else MF_CUDA(cudaMemsetAsync(s->H, 0, sizeof(float) * (size_t)s->k * s->ldn, s->st));  // CCD_CUDA.cu:287
        MF_CUDA(cudaMemcpyAsync(s->v_old, s->H, sizeof(float) * (size_t)s->k * s->ldn, cudaMemcpyDeviceToDevice, s->st));
    } else {
        MF_REQUIRE(H != nullptr, "ALS needs initial H (ALS_CUDA.cu:237-243)");
        MF_CUDA(cudaMemcpyAsync(s->W, W, sizeof(float) * (size_t)s->rows * s->k, cudaMemcpyDefault, s->st));
        MF_CUDA(cudaMemcpyAsync(s->H, H, sizeof(float) * (size_t)s->cols * s->k, cudaMemcpyDefault, s->st));
    }
    MF_CUDA(cudaStreamSynchronize(s->st));
    return MF_OK;
}

int mf_session_get_factors(mf_session* s, float* W, float* H) {
    MF_REQUIRE(s, "NULL argument");
    MF_CUDA(cudaSetDevice(s->device));
    if (s->prm.solver_type == MF_SOLVER_CCD) {
        if (W) MF_CUDA(cudaMemcpy2DAsync(W, sizeof(float) * s->rows, s->W, sizeof(float) * s->ldm, sizeof(float) * s->rows, s->k, cudaMemcpyDefault, s->st));
        if (H) MF_CUDA(cudaMemcpy2DAsync(H, sizeof(float) * s->cols, s->H, sizeof(float) * s->ldn, sizeof(float) * s->cols, s->k, cudaMemcpyDefault, s->st));
    } else {
        if (W) MF_CUDA(cudaMemcpyAsync(W, s->W, sizeof(float) * (size_t)s->rows * s->k, cudaMemcpyDefault, s->st));
        if (H) MF_CUDA(cudaMemcpyAsync(H, s->H, sizeof(float) * (size_t)s->cols * s->k, cudaMemcpyDefault, s->st));
    }
    MF_CUDA(cudaStreamSynchronize(s->st));
    return MF_OK;
}

int mf_session_get_values(mf_session* s, float* csr_val, float* csc_val) {
    MF_REQUIRE(s, "NULL argument");
    MF_CUDA(cudaSetDevice(s->device));
    if (s->prm.solver_type == MF_SOLVER_CCD) MF_TRY(flush_pending(s));
    Side* sides[2] = {&s->csr, &s->csc};
    float* dsts[2] = {csr_val, csc_val};
    for (int i = 0; i < 2; ++i) {
        Side& sd = *sides[i];
        if (!dsts[i] || sd.nnz == 0) continue;
        if (s->panel) {
            float* tmp = nullptr;
            MF_TRY(dev_alloc(&tmp, (size_t)sd.nnz));
            int rc = side_panel_to_raw(sd, tmp, s->st);
            float* tmp2 = nullptr;
            if (rc == MF_OK && sd.unsort_perm) {  // the segments were sorted on upload: back to the caller's order
                rc = dev_alloc(&tmp2, (size_t)sd.nnz);
                if (rc == MF_OK) rc = scatter_by_perm(sd.unsort_perm, tmp, tmp2, sd.nnz, s->st);
            }
            if (rc == MF_OK && cudaMemcpyAsync(dsts[i], tmp2 ? tmp2 : tmp, sizeof(float) * (size_t)sd.nnz, cudaMemcpyDefault, s->st) != cudaSuccess) rc = MF_ERR_CUDA;
            cudaStreamSynchronize(s->st);
            dev_free(tmp);
            if (tmp2) dev_free(tmp2);
            MF_TRY(rc);
        } else {
            MF_CUDA(cudaMemcpyAsync(dsts[i], sd.val, sizeof(float) * (size_t)sd.nnz, cudaMemcpyDefault, s->st));
        }
    }
    MF_CUDA(cudaStreamSynchronize(s->st));
    return MF_OK;
}

int mf_session_rmse(mf_session* s, double* rmse) {
    MF_REQUIRE(s && rmse, "NULL argument");
    MF_CUDA(cudaSetDevice(s->device));
    int rc = session_rmse(s, rmse, nullptr);
    s->timer.collect(s->fam_seconds, s->fam_launches);
    return rc;
}

int mf_session_predict(mf_session* s, int64_t n, const uint32_t* row, const uint32_t* col, double* out) {
    MF_REQUIRE(s && n >= 0 && (n == 0 || (row && col && out)), "bad argument");
    if (n == 0) return MF_OK;
    MF_CUDA(cudaSetDevice(s->device));
    // pairs and results may live on the host or on the device: stage through device scratch
    uint32_t *d_row = nullptr, *d_col = nullptr;
    double* d_out = nullptr;
    MF_TRY(dev_alloc(&d_row, (size_t)n));
    int rc = dev_alloc(&d_col, (size_t)n);
    if (rc == MF_OK) rc = dev_alloc(&d_out, (size_t)n);
    auto cuda_ok = [&](cudaError_t e) {
        if (e != cudaSuccess && rc == MF_OK) { set_error("mf_session_predict: %s", cudaGetErrorString(e)); rc = MF_ERR_CUDA; }
    };
    if (rc == MF_OK) {
        cuda_ok(cudaMemcpyAsync(d_row, row, sizeof(uint32_t) * (size_t)n, cudaMemcpyDefault, s->st));
        cuda_ok(cudaMemcpyAsync(d_col, col, sizeof(uint32_t) * (size_t)n, cudaMemcpyDefault, s->st));
    }
    if (rc == MF_OK) {
        bool ok_row = true, ok_col = true;
        rc = check_below(d_row, n, (uint64_t)s->rows, &ok_row, s->st);
        if (rc == MF_OK) rc = check_below(d_col, n, (uint64_t)s->cols, &ok_col, s->st);
        if (rc == MF_OK && (!ok_row || !ok_col)) { set_error("mf_session_predict: a pair is outside the %lld x %lld matrix", (long long)s->rows, (long long)s->cols); rc = MF_ERR_ARG; }
    }
    if (rc == MF_OK) {
        if (s->prm.solver_type == MF_SOLVER_ALS)
            rc = predict_pairs(n, d_row, d_col, s->W, s->H, s->k, 1, s->k, 1, s->k, d_out, s->sm_count, s->st);
        else
            rc = predict_pairs(n, d_row, d_col, s->W, s->H, s->k, s->ldm, 1, s->ldn, 1, d_out, s->sm_count, s->st);
    }
    if (rc == MF_OK) cuda_ok(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDefault, s->st));
    cuda_ok(cudaStreamSynchronize(s->st));
    dev_free(d_row); dev_free(d_col); dev_free(d_out);
    return rc;
}

int mf_session_rank_stats(mf_session* s, double* seconds, double* rmse, int32_t* inner_iters) {
    MF_REQUIRE(s, "NULL argument");
    if (s->prm.solver_type != MF_SOLVER_CCD) { set_error("not a CCD++ session"); return MF_ERR_STATE; }
    if ((seconds || rmse) && !s->rank_report) { set_error("per-rank report is off (needs verbose, do_predict and a test set)"); return MF_ERR_STATE; }
    for (int t = 0; t < s->k; ++t) {
        if (seconds) seconds[t] = s->rank_seconds[t];
        if (rmse) rmse[t] = s->rank_rmse[t];
        if (inner_iters) inner_iters[t] = s->rank_inner[t];
    }
    return MF_OK;
}

int mf_predict_pairs(const float* W, const float* H, int64_t rows, int64_t cols, int64_t k, int64_t n, const uint32_t* row,
                     const uint32_t* col, double* out, int device) {
    MF_REQUIRE(W && H && rows > 0 && cols > 0 && k > 0 && n >= 0 && (n == 0 || (row && col && out)), "mf_predict_pairs: bad argument");
    if (n == 0) return MF_OK;
    int ndev = 0, sms = 0;
    MF_CUDA(cudaGetDeviceCount(&ndev));
    MF_REQUIRE(device >= 0 && device < ndev, "device %d not present (%d CUDA devices)", device, ndev);
    MF_CUDA(cudaSetDevice(device));
    MF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    for (int64_t e = 0; e < n; ++e)  // host pairs: the model's shape bounds them (a bad pair would read past the factors)
        MF_REQUIRE(row[e] < (uint64_t)rows && col[e] < (uint64_t)cols, "pair %lld (%u, %u) is outside the %lld x %lld model", (long long)e, row[e], col[e], (long long)rows, (long long)cols);
    float *dW = nullptr, *dH = nullptr;
    uint32_t *d_row = nullptr, *d_col = nullptr;
    double* d_out = nullptr;
    int rc = dev_alloc(&dW, (size_t)rows * k);
    if (rc == MF_OK) rc = dev_alloc(&dH, (size_t)cols * k);
    if (rc == MF_OK) rc = dev_alloc(&d_row, (size_t)n);
    if (rc == MF_OK) rc = dev_alloc(&d_col, (size_t)n);
    if (rc == MF_OK) rc = dev_alloc(&d_out, (size_t)n);
    auto cuda_ok = [&](cudaError_t e) {
        if (e != cudaSuccess && rc == MF_OK) { set_error("mf_predict_pairs: %s", cudaGetErrorString(e)); rc = MF_ERR_CUDA; }
    };
    if (rc == MF_OK) {
        cuda_ok(cudaMemcpy(dW, W, sizeof(float) * (size_t)rows * k, cudaMemcpyHostToDevice));
        cuda_ok(cudaMemcpy(dH, H, sizeof(float) * (size_t)cols * k, cudaMemcpyHostToDevice));
        cuda_ok(cudaMemcpy(d_row, row, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        cuda_ok(cudaMemcpy(d_col, col, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    if (rc == MF_OK) rc = predict_pairs(n, d_row, d_col, dW, dH, (int)k, 1, k, 1, k, d_out, sms, nullptr);
    if (rc == MF_OK) cuda_ok(cudaMemcpy(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    void* ptrs[] = {dW, dH, d_row, d_col, d_out};
    for (void* p : ptrs)
        if (p) dev_free(p);
    return rc;
}

int mf_session_ccdpp_iterate(mf_session* s, int n_outer, mf_iter_stats* stats) {
    MF_REQUIRE(s && n_outer >= 0, "bad argument");
    if (s->prm.solver_type != MF_SOLVER_CCD) { set_error("not a CCD++ session"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    const bool timing_on = s->prm.no_launch_timing == 0;
    const int tstride = s->prm.timing_stride > 1 ? s->prm.timing_stride : 1;
    s->timer.enabled = timing_on;
    for (int f = 0; f < F_COUNT; ++f) { s->fam_seconds[f] = 0; s->fam_launches[f] = 0; }
    s->timer.launched = 0;
    double total = 0.0;
    if (s->broken) { set_error("session is unusable after a device-side timeout"); return MF_ERR_STATE; }
    const bool persist = use_persistent(s);
    auto one_outer = [&](bool add) -> int {
        MF_TRY(exchange_barrier(s));
        if (persist) return ccd_outer_persistent(s, add, timing_on);
        for (int t = 0; t < s->k; ++t) {
            s->timer.enabled = timing_on && (t % tstride == 0);
            if (s->prm.schedule == MF_SCHEDULE_REFERENCE) MF_TRY(ccd_rank_reference(s, t, add));
            else MF_TRY(ccd_rank_fused(s, t, add));
        }
        s->timer.enabled = timing_on;
        return MF_OK;
    };
    if (!stats && !persist && !s->rank_report) {
        // no per-iteration report wanted: one event pair around all n_outer iterations, one sync
        MF_CUDA(cudaEventRecord(s->ev_a, s->st));
        for (int it = 0; it < n_outer; ++it) {
            MF_TRY(one_outer(s->outer_done > 0));
            s->outer_done++;
        }
        MF_CUDA(cudaEventRecord(s->ev_b, s->st));
        MF_CUDA(cudaStreamSynchronize(s->st));
        float ms = 0.f;
        MF_CUDA(cudaEventElapsedTime(&ms, s->ev_a, s->ev_b));
        total = ms * 1e-3;
        s->timer.collect(s->fam_seconds, s->fam_launches);
        MF_TRY(check_device_status(s));
        s->last_seconds = total;
        {
            char msg[256];
            if (s->panel && panel_timeout_report(msg, sizeof(msg))) { set_error("%s", msg); return MF_ERR_STATE; }
        }
        return MF_OK;
    }
    for (int it = 0; it < n_outer; ++it) {
        const bool add = s->outer_done > 0;  // src/CCD.cpp:100: add-back only from the second outer iteration on
        const bool had_pending = s->pending >= 0;
        double before[F_COUNT];
        memcpy(before, s->fam_seconds, sizeof(before));
        MF_CUDA(cudaEventRecord(s->ev_a, s->st));
        MF_TRY(one_outer(add));
        MF_CUDA(cudaEventRecord(s->ev_b, s->st));
        MF_CUDA(cudaStreamSynchronize(s->st));
        float ms = 0.f;
        MF_CUDA(cudaEventElapsedTime(&ms, s->ev_a, s->ev_b));
        total += ms * 1e-3;
        s->outer_done++;
        s->timer.collect(s->fam_seconds, s->fam_launches);
        if (s->rank_report) {
            std::vector<double> acc((size_t)s->k);
            MF_CUDA(cudaMemcpy(acc.data(), s->d_rank_acc, sizeof(double) * (size_t)s->k, cudaMemcpyDeviceToHost));
            for (int t = 0; t < s->k; ++t) {
                float rms = 0.f;
                cudaEventElapsedTime(&rms, s->rank_ev[t], s->rank_ev[t + 1]);
                s->rank_seconds[t] = rms * 1e-3;
                s->rank_rmse[t] = sqrt(acc[t] / (double)s->nt);
            }
        }
        MF_TRY(check_device_status(s));
        if (persist && timing_on) fold_stamps(s, add, had_pending);
        if (stats) {
            mf_iter_stats& o = stats[it];
            const double upd = s->fam_seconds[F_UPDATE] - before[F_UPDATE];
            o.update_time = upd;
            o.rank_time = ms * 1e-3 - upd;
            MF_TRY(session_rmse(s, &o.rmse, &o.rmse_time));
            s->timer.collect(s->fam_seconds, s->fam_launches);
        }
    }
    s->last_seconds = total;
    return MF_OK;
}

int mf_session_als_iterate(mf_session* s, int n_iter, mf_iter_stats* stats) {
    MF_REQUIRE(s && n_iter >= 0, "bad argument");
    if (s->prm.solver_type != MF_SOLVER_ALS) { set_error("not an ALS session"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    s->timer.enabled = s->prm.no_launch_timing == 0;
    for (int f = 0; f < F_COUNT; ++f) { s->fam_seconds[f] = 0; s->fam_launches[f] = 0; }
    double total = 0.0;
    for (int it = 0; it < n_iter; ++it) {
        MF_CUDA(cudaEventRecord(s->ev_a, s->st));
        MF_TRY(mf_session_als_half(s, MF_SIDE_CSR));
        MF_TRY(mf_session_als_half(s, MF_SIDE_CSC));
        MF_CUDA(cudaEventRecord(s->ev_b, s->st));
        MF_CUDA(cudaStreamSynchronize(s->st));
        float ms = 0.f;
        MF_CUDA(cudaEventElapsedTime(&ms, s->ev_a, s->ev_b));
        total += ms * 1e-3;
        s->outer_done++;
        s->timer.collect(s->fam_seconds, s->fam_launches);
        if (stats) {
            mf_iter_stats& o = stats[it];
            o.rank_time = 0.0;
            o.update_time = ms * 1e-3;
            MF_TRY(session_rmse(s, &o.rmse, &o.rmse_time));
            s->timer.collect(s->fam_seconds, s->fam_launches);
        }
    }
    s->last_seconds = total;
    return MF_OK;
}

int mf_session_als_half(mf_session* s, int side) {
    MF_REQUIRE(s && (side == MF_SIDE_CSR || side == MF_SIDE_CSC), "bad argument");
    if (s->prm.solver_type != MF_SOLVER_ALS) { set_error("not an ALS session"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    s->timer.start(F_ALS);
    if (side == MF_SIDE_CSR) {  // W from H over the rows — src/ALS.cpp:98-158
        MF_TRY(als_half_step(s->csr, s->H, s->W + s->csr.seg_offset * s->k, s->k, s->prm.lambda, s->sm_count, s->st));
        s->timer.stop();
        MF_TRY(gather_blocks(s, s->W, s->row_bound, s->k));
    } else {                    // H from W over the columns — src/ALS.cpp:161-219
        MF_TRY(als_half_step(s->csc, s->W, s->H + s->csc.seg_offset * s->k, s->k, s->prm.lambda, s->sm_count, s->st));
        s->timer.stop();
        MF_TRY(gather_blocks(s, s->H, s->col_bound, s->k));
    }
    return MF_OK;
}

int mf_session_ccd_solve(mf_session* s, int t, int side) {
    MF_REQUIRE(s && t >= 0 && t < s->k && (side == MF_SIDE_CSR || side == MF_SIDE_CSC), "bad argument");
    if (s->prm.solver_type != MF_SOLVER_CCD) { set_error("not a CCD++ session"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    MF_TRY(flush_pending(s));
    // multi-GPU: a step-level call may solve the same side twice in a row; the receive words of that side must not be
    // overwritten while a slower peer still polls them for the previous call
    MF_TRY(exchange_barrier(s));
    SweepVectors a;
    if (side == MF_SIDE_CSC) { a.g_new = s->W + (int64_t)t * s->ldm; MF_TRY(solve_v(s, t, kSolve, a)); }
    else                     { a.g_new = s->H + (int64_t)t * s->ldn; MF_TRY(solve_u(s, t, kSolve, a)); }
    MF_CUDA(cudaStreamSynchronize(s->st));
    s->timer.collect(s->fam_seconds, s->fam_launches);
    return MF_OK;
}

int mf_session_ccd_update(mf_session* s, int t, int add) {
    MF_REQUIRE(s && t >= 0 && t < s->k, "bad argument");
    if (s->prm.solver_type != MF_SOLVER_CCD) { set_error("not a CCD++ session"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    MF_TRY(flush_pending(s));
    MF_TRY(update_both(s, t, add != 0));
    MF_CUDA(cudaStreamSynchronize(s->st));
    s->timer.collect(s->fam_seconds, s->fam_launches);
    return MF_OK;
}

int mf_als_plan(int64_t nseg, const uint32_t* ptr, uint32_t split, uint32_t* items, int64_t* n_items, uint32_t* n_slots) {
    MF_REQUIRE(nseg >= 0 && ptr != nullptr, "bad argument");
    return als_plan_host(ptr, nseg, split, items, n_items, n_slots);
}

int mf_session_kernel_times(mf_session* s, mf_kernel_times* out) {
    MF_REQUIRE(s && out, "NULL argument");
    memset(out, 0, sizeof(*out));
    out->solve_s = s->fam_seconds[F_SOLVE];       out->solve_launches = s->fam_launches[F_SOLVE];
    out->fused_s = s->fam_seconds[F_FUSED];       out->fused_launches = s->fam_launches[F_FUSED];
    out->update_s = s->fam_seconds[F_UPDATE];     out->update_launches = s->fam_launches[F_UPDATE];
    out->finalize_s = s->fam_seconds[F_FINALIZE]; out->finalize_launches = s->fam_launches[F_FINALIZE];
    out->als_s = s->fam_seconds[F_ALS];           out->als_launches = s->fam_launches[F_ALS];
    out->rmse_s = s->fam_seconds[F_RMSE];         out->rmse_launches = s->fam_launches[F_RMSE];
    out->collective_s = s->fam_seconds[F_COLLECTIVE]; out->collective_launches = s->fam_launches[F_COLLECTIVE];
    if (s->prm.solver_type == MF_SOLVER_CCD) {
        // average over the two copies: a "launch" of a family alternates between the CSC and the CSR side
        out->solve_bytes = (sweep_bytes(s, s->csc, kSolve) + sweep_bytes(s, s->csr, kSolve)) / 2;
        out->fused_bytes = (sweep_bytes(s, s->csc, kSolve | kSub | kAdd) + sweep_bytes(s, s->csr, kSolve | kSub | kAdd | kAddSep)) / 2;
        out->update_bytes = (sweep_bytes(s, s->csc, kSub) + sweep_bytes(s, s->csr, kSub)) / 2;
    }
    out->total_launches = s->timer.launched;
    out->persistent_s = s->fam_seconds[F_PERSIST]; out->persistent_launches = s->fam_launches[F_PERSIST];
    if (s->prm.solver_type == MF_SOLVER_CCD && s->prm.maxinneriter >= 1) {
        // one persistent launch = k ranks x [fused CSC, fused CSR, (T-1) x (solve CSC, solve CSR)]
        const int64_t T = s->prm.maxinneriter;
        out->persistent_bytes = (int64_t)s->k * (sweep_bytes(s, s->csc, kSolve | kSub | kAdd) + sweep_bytes(s, s->csr, kSolve | kSub | kAdd | kAddSep) +
                                                (T - 1) * (sweep_bytes(s, s->csc, kSolve) + sweep_bytes(s, s->csr, kSolve)));
    }
    return MF_OK;
}

int mf_session_last_seconds(mf_session* s, double* seconds) {
    MF_REQUIRE(s && seconds, "NULL argument");
    *seconds = s->last_seconds;
    return MF_OK;
}

int mf_session_panel_layout(mf_session* s, int side, int64_t* n_padded, int64_t* n_items, int64_t* n_panels,
                            uint16_t* idx16, float* val, uint32_t* items) {
    MF_REQUIRE(s && (side == MF_SIDE_CSR || side == MF_SIDE_CSC), "bad argument");
    if (!s->panel) { set_error("session does not use the panel layout"); return MF_ERR_STATE; }
    MF_CUDA(cudaSetDevice(s->device));
    const Side& sd = side == MF_SIDE_CSR ? s->csr : s->csc;
    if (n_padded) *n_padded = sd.npad;
    if (n_items) *n_items = sd.nitems;
    if (n_panels) *n_panels = sd.npanels;
    if (idx16 && sd.npad) MF_CUDA(cudaMemcpy(idx16, sd.idx16, sizeof(uint16_t) * (size_t)sd.npad, cudaMemcpyDefault));
    if (val && sd.npad) MF_CUDA(cudaMemcpy(val, sd.pval, sizeof(float) * (size_t)sd.npad, cudaMemcpyDefault));
    if (items && sd.nitems) MF_CUDA(cudaMemcpy(items, sd.items, sizeof(WorkItem) * (size_t)sd.nitems, cudaMemcpyDefault));
    return MF_OK;
}

// ---- one-shot trainers ---------------------------------------------------------------------
static int train_impl(const mf_ratings* R, const mf_testset* T, float* W, float* H, const mf_params* params,
                      mf_iter_stats* stats, int solver) {
    MF_REQUIRE(R && W && H && params, "NULL argument");
    mf_params p = *params;
    p.solver_type = solver;
    mf_session* s = nullptr;
    MF_TRY(mf_session_create(R, T, &p, &s));
    int rc = mf_session_set_factors(s, W, solver == MF_SOLVER_CCD ? nullptr : H);
    trace_mark("train: set_factors");
    double rank_acc = 0.0, upd_acc = 0.0;
    for (int it = 0; rc == MF_OK && it < p.maxiter; ++it) {
        mf_iter_stats st;
        rc = solver == MF_SOLVER_CCD ? mf_session_ccdpp_iterate(s, 1, &st) : mf_session_als_iterate(s, 1, &st);
        if (rc != MF_OK) break;
        trace_mark("train: one iteration + rmse");
        rank_acc += st.rank_time;
        upd_acc += st.update_time;
        if (stats) stats[it] = st;
        if (!p.quiet && solver == MF_SOLVER_CCD && p.verbose) {
            // the per-rank lines of the reference's (commented-out) verbose block, src/CCD.cpp:141-148
            std::vector<double> sec((size_t)p.k), rm((size_t)p.k);
            const bool have = mf_session_rank_stats(s, sec.data(), rm.data(), nullptr) == MF_OK;
            for (unsigned t = 0; have && t < p.k; ++t) {
                printf("iter %d rank %d time %f", it + 1, (int)t + 1, sec[t]);
                if (p.do_predict) printf(" rmse %f", rm[t]);
                printf("\n");
            }
        }
        if (!p.quiet) {
            if (solver == MF_SOLVER_CCD)  // line format of CCD_CUDA.cu:405
                printf("[-INFO-] iteration num %d \trank_time %.4lf|%.4lf s \tupdate_time %.4lf|%.4lfs \tRMSE=%lf time:%fs\n",
                       it + 1, st.rank_time, rank_acc, st.update_time, upd_acc, st.rmse, st.rmse_time);
            else                          // line format of ALS_CUDA.cu:360
                printf("[-INFO-] iteration num %d \tupdate_time %.4lf|%.4lfs \tRMSE=%lf time:%fs\n", it + 1,
                       st.update_time, upd_acc, st.rmse, st.rmse_time);
        }
    }
    if (rc == MF_OK) rc = mf_session_get_factors(s, W, H);
    trace_mark("train: get_factors");
    mf_session_destroy(s);
    trace_mark("train: destroy");
    return rc;
}

int mf_ccdpp_train(const mf_ratings* R, const mf_testset* T, float* W, float* H, const mf_params* params,
                   mf_iter_stats* stats) {
    return train_impl(R, T, W, H, params, stats, MF_SOLVER_CCD);
}

int mf_als_train(const mf_ratings* R, const mf_testset* T, float* W, float* H, const mf_params* params,
                 mf_iter_stats* stats) {
    return train_impl(R, T, W, H, params, stats, MF_SOLVER_ALS);
}

}  // extern "C"
