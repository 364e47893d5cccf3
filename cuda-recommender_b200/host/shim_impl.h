// shim_impl.h — the reference's two GPU entry points, re-implemented on the C-ABI.
//
//   void kernel_wrapper_ccdpp_NV(SparseMatrix&, TestData&, MatData& W, MatData& H, parameter&)   cuda_src/CCD_CUDA.h:49
//   void kernel_wrapper_als_NV  (SparseMatrix&, TestData&, MatData& W, MatData& H, parameter&)   cuda_src/ALS_CUDA.h:40
//
// Same names, signatures, in/out contract and error behaviour (print "<solver> FAILED: ..." and
// return, cuda_src/CCD_CUDA.cu:174-176): the reference's own src/main.cpp:11-17 links against this
// file unchanged.  It only uses the public accessors that exist both in the
// reference's headers and in this build's host/pmf_util.h, so it compiles against either.
// The flatten / unflatten the reference does inside ccdpp_NV / als_NV (CCD_CUDA.cu:250-269,409-427;
// ALS_CUDA.cu:224-243,364-382) happens here; everything else is behind mf_ccdpp_train / mf_als_train.
// (Included after a pmf.h — this build's host/pmf.h or the reference's src/pmf.h — by shim.cpp and by
// oracle/dropin_shim.cpp respectively.)
#include <algorithm>
#include <cstdio>
#include <vector>

#include "../../include/mf_abi.h"

namespace {

void fill_views(SparseMatrix& R, TestData& T, mf_ratings& r, mf_testset& t) {
    r.rows = R.rows; r.cols = R.cols; r.nnz = R.nnz;
    r.csr_row_ptr = R.get_csr_row_ptr(); r.csr_col_idx = R.get_csr_col_indx(); r.csr_val = R.get_csr_val();
    r.csc_col_ptr = R.get_csc_col_ptr(); r.csc_row_idx = R.get_csc_row_indx(); r.csc_val = R.get_csc_val();
    t.nnz = T.nnz; t.row = T.getTestRow(); t.col = T.getTestCol(); t.val = T.getTestVal();
}

template <typename P>
auto ext_device(const P& p, int) -> decltype(p.device) { return p.device; }
template <typename P>
int ext_device(const P&, long) { return 0; }
template <typename P>
auto ext_schedule(const P& p, int) -> decltype(p.schedule) { return p.schedule; }
template <typename P>
int ext_schedule(const P&, long) { return 0; }
template <typename P>
auto ext_layout(const P& p, int) -> decltype(p.layout) { return p.layout; }
template <typename P>
int ext_layout(const P&, long) { return 0; }
template <typename P>
auto ext_early_stop(const P& p, int) -> decltype(p.early_stop) { return p.early_stop; }
template <typename P>
int ext_early_stop(const P&, long) { return 0; }

mf_params to_abi(const parameter& p, int solver) {
    mf_params q;
    mf_params_default(&q);
    q.solver_type = solver;
    q.k = p.k; q.threads = p.threads; q.maxiter = p.maxiter; q.maxinneriter = p.maxinneriter;
    q.lambda = p.lambda; q.eps = p.eps; q.do_predict = p.do_predict; q.verbose = p.verbose; q.do_nmf = p.do_nmf;
    q.nBlocks = p.nBlocks; q.nThreadsPerBlock = p.nThreadsPerBlock;
    // the reference's parameter class has no such members; this build's host/pmf.h does
    q.device = ext_device(p, 0); q.schedule = ext_schedule(p, 0); q.layout = ext_layout(p, 0);
    q.early_stop = ext_early_stop(p, 0);  // the reference's -e stays inert (its main.cpp cannot set this)
    return q;
}

// MatData <-> one flat array; `outer` vectors of `inner` floats each, concatenated
void flatten(const MatData& M, size_t outer, size_t inner, std::vector<float>& flat) {
    flat.resize(outer * inner);
    for (size_t a = 0; a < outer; ++a) std::copy(M[a].begin(), M[a].begin() + inner, flat.begin() + a * inner);
}
void unflatten(const std::vector<float>& flat, size_t outer, size_t inner, MatData& M) {
    for (size_t a = 0; a < outer; ++a) std::copy(flat.begin() + a * inner, flat.begin() + (a + 1) * inner, M[a].begin());
}

}  // namespace

void kernel_wrapper_ccdpp_NV(SparseMatrix& R, TestData& T, MatData& W, MatData& H, parameter& parameters) {
    mf_ratings r; mf_testset t;
    fill_views(R, T, r, t);
    mf_params q = to_abi(parameters, MF_SOLVER_CCD);
    std::vector<float> w, h;  // W[t][i] -> w[t*rows+i]
    flatten(W, q.k, (size_t)R.rows, w);
    flatten(H, q.k, (size_t)R.cols, h);
    const int rc = mf_ccdpp_train(&r, &t, w.data(), h.data(), &q, nullptr);
    mf_release_cached_memory(q.device);  // no device state left behind, like the reference's cudaDeviceReset (CCD_CUDA.cu:177)
    if (rc != MF_OK) {
        std::fprintf(stderr, "CCD FAILED: %s\n", mf_last_error());
        return;
    }
    unflatten(w, q.k, (size_t)R.rows, W);
    unflatten(h, q.k, (size_t)R.cols, H);
}

void kernel_wrapper_als_NV(SparseMatrix& R, TestData& T, MatData& W, MatData& H, parameter& parameters) {
    mf_ratings r; mf_testset t;
    fill_views(R, T, r, t);
    mf_params q = to_abi(parameters, MF_SOLVER_ALS);
    std::vector<float> w, h;  // W[i][t] -> w[i*k+t]
    flatten(W, (size_t)R.rows, q.k, w);
    flatten(H, (size_t)R.cols, q.k, h);
    const int rc = mf_als_train(&r, &t, w.data(), h.data(), &q, nullptr);
    mf_release_cached_memory(q.device);  // as above (ALS_CUDA.cu:196)
    if (rc != MF_OK) {
        std::fprintf(stderr, "ALS FAILED: %s\n", mf_last_error());
        return;
    }
    unflatten(w, (size_t)R.rows, q.k, W);
    unflatten(h, (size_t)R.cols, q.k, H);
}
