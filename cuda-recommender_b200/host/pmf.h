// pmf.h — run parameters of the trainer.
// Field names, types and defaults are those of the reference's `class parameter`
// (/root/reference/src/pmf.h:8-43) so that code written against it — notably the two
// kernel_wrapper_*_NV entry points in shim.cpp — compiles against either header.
#ifndef B200_PMF_H
#define B200_PMF_H

#include <cstdio>

#include "pmf_util.h"

enum class solvertype { CCD, ALS };

class parameter {
public:
    solvertype solver_type = solvertype::CCD;
    unsigned k = 10;                 // rank                               (-k)
    int threads = 4;                 // OpenMP threads of the CPU path     (-n)
    int maxiter = 5;                 // outer iterations                   (-t)
    int maxinneriter = 1;            // CCD++ inner iterations             (-T)
    float lambda = 0.1f;             // regulariser                        (-l)
    float eps = 1e-3f;               // stop rule of CCDR1; acts only when -e is given (early_stop), inert in the reference (-e)
    int do_predict = 0;              // prediction output + per-rank RMSE  (-p)
    int verbose = 0;                 // per-rank report lines              (-q)
    int do_nmf = 0;                  // non-negative factors: solved coordinates clamped at 0; inert in the reference (-N)
    bool enable_cuda = false;        //                                    (-CUDA)
    bool enable_omp = false;         //                                    (-OMP)
    unsigned nBlocks = 32;           // accepted for compatibility; launch geometry is chosen by the library
    unsigned nThreadsPerBlock = 256; // accepted for compatibility
    char src_dir[1024];

    // extensions of this build (not in the reference)
    int device = 0;                  // CUDA device ordinal                (-device)
    int schedule = 0;                // 0 fused, 1 reference launch order  (-schedule)
    int layout = 0;                  // 0 panel, 1 direct                  (-layout)
    int early_stop = 0;              // 1 when -e was given on the command line: the eps rule is active
    bool load_model = false;         // predict-only: read <dir>/model, write <dir>/output       (-load)

    parameter() { std::snprintf(src_dir, sizeof(src_dir), "../data/simple"); }
};

#endif  // B200_PMF_H
