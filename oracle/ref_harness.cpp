// ref_harness.cpp — thin C entry points around the UNMODIFIED reference CPU path.
//
// TEST INFRASTRUCTURE ONLY (see oracle/mf_oracle.c header).  This file contains no
// arithmetic of its own: it is compiled by oracle/build_ref.sh together with
// /root/reference/src/{CCD,ALS,tools,extras}.cpp (taken where they lie, never copied
// into this repo) into oracle/_ref/libmfref.so, and only forwards to the reference's
// public functions:
//   load()            src/tools.cpp:3      initial_col()  src/tools.cpp:165
//   ccdr1_OMP()       src/CCD.cpp:45       ALS_OMP()      src/ALS.cpp:81
//   calrmse()         src/tools.cpp:235
// exactly as src/main.cpp:63-129 drives them.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#include "extras.h"
#include "CCD.h"
#include "ALS.h"

namespace {
// the reference reports its timings only on stdout (src/CCD.cpp:158, src/ALS.cpp:229);
// capture that stream into a file for the duration of the solver call.
struct StdoutCapture {
    int saved = -1;
    std::string path;
    explicit StdoutCapture(const char* p) : path(p ? p : "") {
        if (path.empty()) return;
        fflush(stdout);
        saved = dup(1);
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd >= 0) { dup2(fd, 1); close(fd); }
    }
    ~StdoutCapture() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};
}  // namespace

extern "C" {

// same libc sequence as the reference (src/tools.cpp:165-173); X[j*n+i]
void ref_initial_col(float* X, long k, long n) {
    MatData M;
    initial_col(M, k, n);
    for (long j = 0; j < k; ++j)
        for (long i = 0; i < n; ++i) X[j * n + i] = M[j][i];
}

// Reads <dir>/meta_modified_all with the reference loader and reports the shape.
int ref_probe(const char* dir, long* rows, long* cols, long* nnz, long* nnz_test,
              long* max_row_nnz, long* max_col_nnz) {
    SparseMatrix R;
    TestData T;
    StdoutCapture cap("/dev/null");
    load(dir, R, T);
    *rows = R.rows; *cols = R.cols; *nnz = R.nnz; *nnz_test = T.nnz;
    *max_row_nnz = R.max_row_nnz_; *max_col_nnz = R.max_col_nnz_;
    return 0;
}

// Runs the reference CPU solver on the dataset directory `dir` the way main.cpp does.
//   als        0 = CCD++ (W[t*rows+i], H[t*cols+j]); 1 = ALS (W[i*k+t], H[j*k+t])
//   W_in/H_in  optional initial factors in the same flat layout; NULL -> initial_col
//   W_out/H_out final factors (flat); csr_val_out/csc_val_out optional: the value arrays
//              as the solver leaves them (CCD++ leaves the residual there, CCD.cpp:25,36)
//   rmse_out   calrmse after the run;  log_path: where the solver's stdout goes (NULL=keep)
//   seconds_out wall time of the solver call alone
int ref_train(const char* dir, int als, int k, float lambda, int maxiter, int maxinner, int threads,
              const float* W_in, const float* H_in, float* W_out, float* H_out,
              float* csr_val_out, float* csc_val_out, double* rmse_out, const char* log_path,
              double* seconds_out) {
    SparseMatrix R;
    TestData T;
    {
        StdoutCapture cap("/dev/null");
        load(dir, R, T);
    }
    parameter param;
    param.k = k; param.lambda = lambda; param.maxiter = maxiter; param.maxinneriter = maxinner;
    param.threads = threads; param.enable_omp = true;
    param.solver_type = als ? solvertype::ALS : solvertype::CCD;
    MatData W, H;
    const long m = R.rows, n = R.cols;
    if (als) { initial_col(W, m, k); initial_col(H, n, k); }     // main.cpp:86-87
    else     { initial_col(W, k, m); initial_col(H, k, n); }     // main.cpp:92-93
    if (W_in) {
        if (als) for (long i = 0; i < m; ++i) for (int t = 0; t < k; ++t) W[i][t] = W_in[i * k + t];
        else     for (int t = 0; t < k; ++t) for (long i = 0; i < m; ++i) W[t][i] = W_in[t * m + i];
    }
    if (H_in) {
        if (als) for (long j = 0; j < n; ++j) for (int t = 0; t < k; ++t) H[j][t] = H_in[j * k + t];
        else     for (int t = 0; t < k; ++t) for (long j = 0; j < n; ++j) H[t][j] = H_in[t * n + j];
    }
    double t0 = omp_get_wtime();
    {
        StdoutCapture cap(log_path);
        if (als) ALS_OMP(R, W, H, T, param); else ccdr1_OMP(R, W, H, T, param);
    }
    double t1 = omp_get_wtime();
    if (seconds_out) *seconds_out = t1 - t0;
    if (W_out) {
        if (als) for (long i = 0; i < m; ++i) for (int t = 0; t < k; ++t) W_out[i * k + t] = W[i][t];
        else     for (int t = 0; t < k; ++t) for (long i = 0; i < m; ++i) W_out[t * m + i] = W[t][i];
    }
    if (H_out) {
        if (als) for (long j = 0; j < n; ++j) for (int t = 0; t < k; ++t) H_out[j * k + t] = H[j][t];
        else     for (int t = 0; t < k; ++t) for (long j = 0; j < n; ++j) H_out[t * n + j] = H[t][j];
    }
    if (csr_val_out) memcpy(csr_val_out, R.get_csr_val(), sizeof(float) * R.nnz);
    if (csc_val_out) memcpy(csc_val_out, R.get_csc_val(), sizeof(float) * R.nnz);
    if (rmse_out) *rmse_out = T.nnz > 0 ? calrmse(T, W, H, als != 0, true) : 0.0;
    return 0;
}

}  // extern "C"
