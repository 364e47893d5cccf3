// extras.cpp — flag parser and the post-run checks of the CLI.
// Flag set, defaults and messages follow the reference (src/extras.cpp:46-141 parser and help text,
// :182-216 calculate_rmse_directly, :218-238 golden_compare); the implementation is table-driven and
// adds -device / -schedule / -layout.  Known reference quirk kept on purpose: the help text says
// "-T ... (default 5)" while the struct default is 1 (src/extras.cpp:54 vs src/pmf.h:31) — here the
// help states the real default.
#include "extras.h"

#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>

void exit_with_help() {
    std::printf(
        "Usage: b200_recommender [options] data_dir [model_filename]\n"
        "options:\n"
        "    -k rank : set the rank (default 10)\n"
        "    -n threads : set the number of threads (default 4)\n"
        "    -l lambda : set the regularization parameter lambda (default 0.1)\n"
        "    -t max_iter: set the number of iterations (default 5)\n"
        "    -T max_inner_iter: set the number of inner iterations used in CCDR1 (default 1)\n"
        "    -e epsilon : set inner termination criterion epsilon of CCDR1 (default 1e-3)\n"
        "    -p do_predict: do prediction or not (default 0)\n"
        "    -q verbose: show information or not (default 0)\n"
        "    -N do_nmf: do nmf (default 0)\n"
        "    -CUDA: Flag to enable CUDA\n"
        "    -OMP: accepted for compatibility (the OpenMP solvers live in the reference, not in this build)\n"
        "    -nBlocks: Number of blocks on CUDA (default 32; ignored, geometry is chosen by the library)\n"
        "    -nThreadsPerBlock: Number of threads per block on CUDA (default 256; ignored)\n"
        "    -ALS: Flag to enable ALS algorithm, if not present CCD++ is used\n"
        "    -device id : CUDA device ordinal (default 0)\n"
        "    -schedule s : 0 fused sweeps (default), 1 the reference's launch order\n"
        "    -layout l : 0 shared-memory panel layout (default), 1 caller-order arrays\n"
        "    -save : write W then H to <data_dir>/model (row-major, the reference's save_mat_t format)\n"
        "    -load : predict only: read <data_dir>/model, write one prediction per test rating to <data_dir>/output\n"
        "            and print the test RMSE (the reference's calculate_rmse_from_file)\n"
        "  (-e switches the stop rule on: a rank's inner iterations end once the function decrease is below eps x the\n"
        "   largest seen; -N 1 clamps factors at 0; -q 1 -p 1 prints time and test RMSE after every rank)\n");
    std::exit(EXIT_FAILURE);
}

extern bool g_save_model;
bool g_save_model = false;

parameter parse_command_line(int argc, char** argv) {
    parameter param;
    const std::map<std::string, std::function<void(parameter&)>> switches = {
        {"-CUDA", [](parameter& p) { p.enable_cuda = true; }},
        {"-OMP", [](parameter& p) { p.enable_omp = true; }},
        {"-ALS", [](parameter& p) { p.solver_type = solvertype::ALS; }},
        {"-save", [](parameter&) { g_save_model = true; }},
        {"-load", [](parameter& p) { p.load_model = true; }},
    };
    const std::map<std::string, std::function<void(parameter&, const char*)>> valued = {
        {"-nBlocks", [](parameter& p, const char* v) { p.nBlocks = std::atoi(v); }},
        {"-nThreadsPerBlock", [](parameter& p, const char* v) { p.nThreadsPerBlock = std::atoi(v); }},
        {"-device", [](parameter& p, const char* v) { p.device = std::atoi(v); }},
        {"-schedule", [](parameter& p, const char* v) { p.schedule = std::atoi(v); }},
        {"-layout", [](parameter& p, const char* v) { p.layout = std::atoi(v); }},
        {"-k", [](parameter& p, const char* v) { p.k = std::atoi(v); }},
        {"-n", [](parameter& p, const char* v) { p.threads = std::atoi(v); }},
        {"-l", [](parameter& p, const char* v) { p.lambda = (float)std::atof(v); }},
        {"-t", [](parameter& p, const char* v) { p.maxiter = std::atoi(v); }},
        {"-T", [](parameter& p, const char* v) { p.maxinneriter = std::atoi(v); }},
        {"-e", [](parameter& p, const char* v) { p.eps = (float)std::atof(v); p.early_stop = 1; }},
        {"-p", [](parameter& p, const char* v) { p.do_predict = std::atoi(v); }},
        {"-q", [](parameter& p, const char* v) { p.verbose = std::atoi(v); }},
        {"-N", [](parameter& p, const char* v) { p.do_nmf = (std::atoi(v) == 1); }},
    };
    int i = 1;
    for (; i < argc && argv[i][0] == '-'; ++i) {
        const std::string flag(argv[i]);
        auto sw = switches.find(flag);
        if (sw != switches.end()) { sw->second(param); continue; }
        auto va = valued.find(flag);
        if (va == valued.end()) {
            std::fprintf(stderr, "unknown option: %s\n", argv[i]);
            exit_with_help();
        }
        if (++i >= argc) exit_with_help();
        va->second(param, argv[i]);
    }
    if (param.do_predict != 0) param.verbose = 1;
    if (i >= argc) exit_with_help();
    std::snprintf(param.src_dir, sizeof(param.src_dir), "%s", argv[i]);
    return param;
}

// final test RMSE from the factors (double accumulation over float products, like the reference)
void calculate_rmse_directly(MatData& W, MatData& H, TestData& T, int rank, bool ifALS) {
    const double t0 = omp_get_wtime();
    if (T.nnz == 0) std::exit(EXIT_FAILURE);
    (void)rank;
    const double rmse = calrmse(T, W, H, ifALS, true);
    std::printf("Test RMSE = %lf. Calculated in %lfs\n", rmse, omp_get_wtime() - t0);
}

// counts entries that differ from the reference factors by more than 10 % (src/extras.cpp:218-238)
void golden_compare(const MatData& W, const MatData& W_ref, unsigned k, unsigned m) {
    unsigned long bad = 0;
    for (unsigned i = 0; i < k; ++i)
        for (unsigned j = 0; j < m; ++j)
            if (std::fabs((double)W[i][j] - W_ref[i][j]) > 0.1 * std::fabs((double)W_ref[i][j])) ++bad;
    if (bad == 0) {
        std::cout << "Check... PASS!" << std::endl;
    } else {
        const unsigned long entries = (unsigned long)k * m;
        std::printf("Check... NO PASS! [%.4f%%] #Error = %lu out of %lu entries.\n", 100.0 * bad / entries, bad, entries);
    }
}
