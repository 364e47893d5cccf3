// pmf_util.h — host containers for the ratings: paired CSR + CSC (`SparseMatrix`) and the
// held-out COO triples (`TestData`), plus their binary readers.
// Same public surface as the reference's containers (/root/reference/src/pmf_util.h:34-211:
// rows/cols/nnz/max_*_nnz_ members, get_csr_*/get_csc_* accessors, get_shallow_transpose,
// getTestRow/Col/Val) and the same on-disk format (headerless little-endian files: ptr as int32,
// idx as uint32, val as float32).  Storage here is std::vector behind
// shared_ptr, whole-file reads, and every short read or inconsistent ptr array is reported.
#ifndef B200_PMF_UTIL_H
#define B200_PMF_UTIL_H

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <omp.h>

#define DTYPE float

using VecData = std::vector<DTYPE>;
using MatData = std::vector<VecData>;

namespace b200io {
template <typename T>
inline void read_exact(const std::string& path, T* dst, size_t count) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(EXIT_FAILURE); }
    size_t got = count ? std::fread(dst, sizeof(T), count, f) : 0;
    std::fclose(f);
    if (got != count) {
        std::fprintf(stderr, "short read on %s: %zu of %zu elements\n", path.c_str(), got, count);
        std::abort();  // same policy as CHECK_FREAD (src/util.h:10-13)
    }
}
}  // namespace b200io

class SparseMatrix {
public:
    long rows = 0, cols = 0, nnz = 0, max_row_nnz_ = 0, max_col_nnz_ = 0;

    void initialize_matrix(long rows_, long cols_, long nnz_) {
        rows = rows_; cols = cols_; nnz = nnz_;
        csr_ptr_ = std::make_shared<std::vector<unsigned>>(rows + 1);
        csc_ptr_ = std::make_shared<std::vector<unsigned>>(cols + 1);
        csr_idx_ = std::make_shared<std::vector<unsigned>>(nnz);
        csc_idx_ = std::make_shared<std::vector<unsigned>>(nnz);
        csr_val_ = std::make_shared<std::vector<DTYPE>>(nnz);
        csc_val_ = std::make_shared<std::vector<DTYPE>>(nnz);
    }

    void read_binary_file(const std::string& f_csr_ptr, const std::string& f_csr_idx, const std::string& f_csr_val,
                          const std::string& f_csc_ptr, const std::string& f_csc_idx, const std::string& f_csc_val) {
        max_row_nnz_ = read_one(f_csr_ptr, f_csr_idx, f_csr_val, *csr_ptr_, *csr_idx_, *csr_val_, cols);
        max_col_nnz_ = read_one(f_csc_ptr, f_csc_idx, f_csc_val, *csc_ptr_, *csc_idx_, *csc_val_, rows);
    }

    // the transpose shares storage: its CSC is this matrix's CSR and vice versa
    SparseMatrix get_shallow_transpose() const {
        SparseMatrix t;
        t.rows = cols; t.cols = rows; t.nnz = nnz;
        t.max_row_nnz_ = max_col_nnz_; t.max_col_nnz_ = max_row_nnz_;
        t.csr_ptr_ = csc_ptr_; t.csr_idx_ = csc_idx_; t.csr_val_ = csc_val_;
        t.csc_ptr_ = csr_ptr_; t.csc_idx_ = csr_idx_; t.csc_val_ = csr_val_;
        return t;
    }

    unsigned* get_csc_col_ptr() const { return csc_ptr_->data(); }
    unsigned* get_csc_row_indx() const { return csc_idx_->data(); }
    DTYPE* get_csc_val() const { return csc_val_->data(); }
    unsigned* get_csr_row_ptr() const { return csr_ptr_->data(); }
    unsigned* get_csr_col_indx() const { return csr_idx_->data(); }
    DTYPE* get_csr_val() const { return csr_val_->data(); }

private:
    long read_one(const std::string& f_ptr, const std::string& f_idx, const std::string& f_val, std::vector<unsigned>& ptr,
                  std::vector<unsigned>& idx, std::vector<DTYPE>& val, long minor_dim) {
        b200io::read_exact(f_ptr, reinterpret_cast<int32_t*>(ptr.data()), ptr.size());
        b200io::read_exact(f_idx, idx.data(), idx.size());
        b200io::read_exact(f_val, val.data(), val.size());
        long widest = 0;
        if (ptr.empty() || ptr.front() != 0 || (long)ptr.back() != nnz) {
            std::fprintf(stderr, "%s: ptr array does not span [0, nnz=%ld]\n", f_ptr.c_str(), nnz);
            std::exit(EXIT_FAILURE);
        }
        for (size_t s = 0; s + 1 < ptr.size(); ++s) {
            if (ptr[s + 1] < ptr[s]) { std::fprintf(stderr, "%s: ptr array decreases at %zu\n", f_ptr.c_str(), s); std::exit(EXIT_FAILURE); }
            widest = std::max<long>(widest, (long)ptr[s + 1] - (long)ptr[s]);
        }
        for (unsigned v : idx)
            if ((long)v >= minor_dim) { std::fprintf(stderr, "%s: index %u out of range\n", f_idx.c_str(), v); std::exit(EXIT_FAILURE); }
        return widest;
    }

    std::shared_ptr<std::vector<unsigned>> csr_ptr_, csc_ptr_, csr_idx_, csc_idx_;
    std::shared_ptr<std::vector<DTYPE>> csr_val_, csc_val_;
};

class TestData {
public:
    long rows = 0, cols = 0, nnz = 0;

    void read_binary_file(long rows_, long cols_, long nnz_, const std::string& f_val, const std::string& f_row,
                          const std::string& f_col) {
        rows = rows_; cols = cols_; nnz = nnz_;
        row_.resize(nnz); col_.resize(nnz); val_.resize(nnz);
        b200io::read_exact(f_val, val_.data(), val_.size());
        b200io::read_exact(f_row, row_.data(), row_.size());
        b200io::read_exact(f_col, col_.data(), col_.size());
    }

    unsigned* getTestRow() const { return const_cast<unsigned*>(row_.data()); }
    unsigned* getTestCol() const { return const_cast<unsigned*>(col_.data()); }
    DTYPE* getTestVal() const { return const_cast<DTYPE*>(val_.data()); }

private:
    std::vector<unsigned> row_, col_;
    std::vector<DTYPE> val_;
};

#endif  // B200_PMF_UTIL_H
