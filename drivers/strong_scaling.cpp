// strong_scaling -- C++ twin of the reference driver tests/strong_scaling.f90 (:22-55): ONE grid (argv[1]),
// gmres_mgsr_omp with cbpr2, max_iter = 50 per restart, tol = 1.d-15, ntests = 6 solves of the same system, one
// table line each.  The reference sweeps OMP threads 1..6 over the six solves; one GPU has no such axis, so the six
// lines are six solves on one handle (the first pays the set-up: workspace, tensor maps, graph capture; the multi-GPU
// strong-scaling sweep is `bench.py --gpus N`).  argv: <grid size> [ntests] [tol].
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "krylov_b200.hpp"
#include "krylov_report.hpp"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::printf(" usage ./strong_scaling <grid size>\n");        // strong_scaling.f90:29
        return 0;
    }
    const int nsize = std::atoi(argv[1]);
    const int ntests = argc > 2 ? std::atoi(argv[2]) : 6;          // :23
    const double tol = argc > 3 ? std::atof(argv[3]) : 1e-15;      // :22
    const int max_iter = 50;                                       // :24
    const std::vector<double> params{8.2, 0.2};                    // :35
    krylov::Handle h(0);
    std::vector<double> ones((size_t)nsize * nsize, 1.0), b, x, errn, verr;
    krylov::apply(h, krylov::stvec, ones, b, nsize);               // :38-39
    std::printf(" GMRES Strong Scaling Test (MGSR with Chebyshev precond)\n");
    char header[128];
    std::snprintf(header, sizeof header, "%25s%4d", "Number of Tests:", ntests);
    krylov::print_header(header);
    for (int i = 1; i <= ntests; ++i) {
        char desc[32];
        std::snprintf(desc, sizeof desc, "%15s%2d", "B200 solve=", i);          // reference: "threads=" i  (:49)
        int n_iter = 0, n_stages = 0;
        auto t0 = std::chrono::steady_clock::now();
        krylov::gmres_mgsr_omp(h, krylov::stvec, b, x, max_iter, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double l2, linf;
        krylov::error_norms(x, l2, linf);
        krylov::print_line(i, nsize * nsize, secs, (n_stages - 1) * max_iter + n_iter, n_stages, max_iter, tol,
                           errn[n_iter - 1], verr[n_iter - 1], l2, linf, desc);            // :53-54
    }
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::putchar('\n');
    return 0;
}
