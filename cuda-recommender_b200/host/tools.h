// tools.h — dataset loader, factor seeding, host-side RMSE (the host pieces the north star keeps:
// /root/reference/src/tools.cpp:3-85 load, :165-173 initial_col, :184-198 dot, :235-248 calrmse).
#ifndef B200_TOOLS_H
#define B200_TOOLS_H

#include "pmf_util.h"

void load(const char* srcdir, SparseMatrix& R, TestData& T);
void initial_col(MatData& X, long k, long n);
double dot(const MatData& W, long i, const MatData& H, long j, bool ifALS);
double calrmse(TestData& T, const MatData& W, const MatData& H, bool ifALS, bool iscol = true);
void save_mat_t(const MatData& A, FILE* fp, bool row_major = true);
MatData load_mat_t(FILE* fp, bool row_major = true);

#endif  // B200_TOOLS_H
