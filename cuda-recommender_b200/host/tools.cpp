// tools.cpp — see tools.h.  Behaviour follows the reference; the code is this build's own.
#include "tools.h"

#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/mf_abi.h"

namespace {
std::string join(const std::string& dir, const std::string& name) { return dir + "/" + name; }
double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}
}  // namespace

// <srcdir>/meta_modified_all: "m n nnz", three COO file names (never opened), CSR ptr/idx/val names,
// CSC ptr/idx/val names, nnz_test, test val/row/col names — reference parser: tools.cpp:3-85.
void load(const char* srcdir, SparseMatrix& R, TestData& T) {
    const std::string dir(srcdir);
    std::ifstream meta(join(dir, "meta_modified_all"));
    if (!meta) {
        std::printf("Can't open meta input file.\n");
        std::exit(EXIT_FAILURE);
    }
    long m = 0, n = 0, nnz = 0;
    std::string name[12];
    unsigned long nnz_test = 0;
    if (!(meta >> m >> n >> nnz)) { std::fprintf(stderr, "meta_modified_all: bad header\n"); std::abort(); }
    for (int i = 0; i < 9; ++i)
        if (!(meta >> name[i])) { std::fprintf(stderr, "meta_modified_all: missing file name %d\n", i); std::abort(); }
    if (!(meta >> nnz_test >> name[9] >> name[10] >> name[11])) { std::fprintf(stderr, "meta_modified_all: bad test block\n"); std::abort(); }

    auto t0 = std::chrono::high_resolution_clock::now();
    R.initialize_matrix(m, n, nnz);
    std::cout << "[info] Alloc TIMER: " << seconds_since(t0) << "s.\n";
    t0 = std::chrono::high_resolution_clock::now();
    R.read_binary_file(join(dir, name[3]), join(dir, name[4]), join(dir, name[5]), join(dir, name[6]), join(dir, name[7]),
                       join(dir, name[8]));
    std::cout << "[info] Train TIMER: " << seconds_since(t0) << "s.\n";
    t0 = std::chrono::high_resolution_clock::now();
    T.read_binary_file(m, n, (long)nnz_test, join(dir, name[9]), join(dir, name[10]), join(dir, name[11]));
    std::cout << "[info] Tests TIMER: " << seconds_since(t0) << "s.\n";
}

// Same libc sequence and fill order as the reference (tools.cpp:165-173); the arithmetic lives in the
// library (mf_host_initial_col) so that every host of the C-ABI seeds identically.
void initial_col(MatData& X, long k, long n) {
    std::vector<float> flat((size_t)k * (size_t)n);
    mf_host_initial_col(flat.data(), k, n);
    X.assign(k, VecData(n));
    for (long j = 0; j < k; ++j) std::copy(flat.begin() + j * n, flat.begin() + (j + 1) * n, X[j].begin());
}

double dot(const MatData& W, long i, const MatData& H, long j, bool ifALS) {
    double acc = 0;
    if (ifALS) {
        const size_t k = W.empty() ? 0 : W[0].size();
        for (size_t t = 0; t < k; ++t) acc += W[i][t] * H[j][t];
    } else {
        for (size_t t = 0; t < W.size(); ++t) acc += W[t][i] * H[t][j];
    }
    return acc;
}

double calrmse(TestData& T, const MatData& W, const MatData& H, bool ifALS, bool) {
    double sq = 0;
#pragma omp parallel for reduction(+ : sq)
    for (long e = 0; e < T.nnz; ++e) {
        double err = dot(W, T.getTestRow()[e], H, T.getTestCol()[e], ifALS) - T.getTestVal()[e];
        sq += err * err;
    }
    return T.nnz ? std::sqrt(sq / T.nnz) : 0.0;
}

// model file: two longs (rows, cols) then row-major floats — format of the reference's save_mat_t /
// load_mat_t (tools.cpp:90-153), which its main never calls (main.cpp:146-149 commented out).
void save_mat_t(const MatData& A, FILE* fp, bool row_major) {
    if (!fp || A.empty()) { std::fprintf(stderr, "save_mat_t: nothing to write\n"); std::exit(EXIT_FAILURE); }
    const long m = row_major ? (long)A.size() : (long)A[0].size();
    const long n = row_major ? (long)A[0].size() : (long)A.size();
    std::fwrite(&m, sizeof(long), 1, fp);
    std::fwrite(&n, sizeof(long), 1, fp);
    std::vector<float> buf((size_t)m * n);
    for (long i = 0; i < m; ++i)
        for (long j = 0; j < n; ++j) buf[(size_t)i * n + j] = row_major ? A[i][j] : A[j][i];
    std::fwrite(buf.data(), sizeof(float), buf.size(), fp);
}

MatData load_mat_t(FILE* fp, bool row_major) {
    long m = 0, n = 0;
    if (!fp || std::fread(&m, sizeof(long), 1, fp) != 1 || std::fread(&n, sizeof(long), 1, fp) != 1 || m <= 0 || n <= 0) {
        std::fprintf(stderr, "load_mat_t: bad header\n");
        std::exit(EXIT_FAILURE);
    }
    std::vector<float> buf((size_t)m * n);
    if (std::fread(buf.data(), sizeof(float), buf.size(), fp) != buf.size()) { std::fprintf(stderr, "load_mat_t: short read\n"); std::abort(); }
    MatData A = row_major ? MatData(m, VecData(n)) : MatData(n, VecData(m));
    for (long i = 0; i < m; ++i)
        for (long j = 0; j < n; ++j) (row_major ? A[i][j] : A[j][i]) = buf[(size_t)i * n + j];
    return A;
}
