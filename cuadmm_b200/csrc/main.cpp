// main.cpp — cuadmm_exe <dir>: the reference's command-line front end (src/main.cu:8-43) over the
// C ABI.  Same hard-coded run parameters: eig_stream_num_per_gpu = 15, cpu_eig_thread_num = 30,
// sig = 1, solve(1e6, 1e-3, sig_update_threshold = 0, 50, 100, switch_admm = 5000), then
// X_opt.txt written with "%.32f\n" (include/cuadmm/memory.h:278-293).  Optional extra arguments
// (not in the reference): --max-iter N --tol T --switch-admm K --quiet --no-output.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/cuadmm_b200.h"

static int fail(const char* what) {
    fprintf(stderr, "ERROR: %s: %s\n", what, cuadmm_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <problem-dir/> [--max-iter N] [--tol T] [--switch-admm K] [--quiet] [--no-output]\n", argv[0]);
        return 2;
    }
    std::string prefix = argv[1];
    int max_iter = (int)1e6, switch_admm = 5000, verbose = 1, write_out = 1;
    double tol = 1e-3;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--max-iter") && i + 1 < argc) max_iter = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--tol") && i + 1 < argc) tol = atof(argv[++i]);
        else if (!strcmp(argv[i], "--switch-admm") && i + 1 < argc) switch_admm = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) verbose = 0;
        else if (!strcmp(argv[i], "--no-output")) write_out = 0;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    const int eig_stream_num_per_gpu = 15, cpu_eig_thread_num = 30;

    cuadmm_problem_t* prob = nullptr;
    if (cuadmm_problem_from_txt(prefix.c_str(), 0, &prob)) return fail("loading the problem");
    int64_t d[8];
    cuadmm_problem_dims(prob, d);
    printf("Loaded problem from %s\n", prefix.c_str());
    printf("              vector length: %lld\n", (long long)d[0]);
    printf("      number of constraints: %lld\n", (long long)d[1]);
    printf("           number of blocks: %lld\n", (long long)d[2]);
    printf("  number of non-zeros in At: %lld\n", (long long)d[3]);
    printf("   number of non-zeros in b: %lld\n", (long long)d[4]);
    printf("   number of non-zeros in C: %lld\n", (long long)d[5]);

    cuadmm_solver_t* solver = nullptr;
    if (cuadmm_solver_create(&solver)) return fail("creating the solver");
    cuadmm_solver_set_verbose(solver, verbose);
    const double sig = 1e0;
    if (cuadmm_solver_init_from_problem(solver, prob, eig_stream_num_per_gpu, cpu_eig_thread_num, sig)) return fail("init");
    // sGS-ADMM, as src/main.cu:39
    if (cuadmm_solver_solve(solver, max_iter, tol, 0, 50, 100, switch_admm, 1.05, 1)) return fail("solve");

    if (write_out) {
        std::vector<double> X((size_t)d[0]);
        if (cuadmm_solver_get_X(solver, X.data())) return fail("reading X");
        if (prefix.empty() || prefix.back() != '/') prefix += '/';
        const std::string out = prefix + "X_opt.txt";
        FILE* f = fopen(out.c_str(), "w");
        if (!f) { fprintf(stderr, "Unable to open file: %s\n", out.c_str()); }
        else {
            for (size_t i = 0; i < X.size(); ++i) fprintf(f, "%.32f\n", X[i]);
            fclose(f);
        }
    }
    double t[8];
    cuadmm_solver_times(solver, t);
    printf("init %.3fs, solve loop %.3fs, %lld iterations, %lld kernel launches\n", t[1], t[2],
           (long long)cuadmm_solver_iter_num(solver), (long long)cuadmm_solver_launches(solver));
    cuadmm_solver_destroy(solver);
    cuadmm_problem_destroy(prob);
    return 0;
}
