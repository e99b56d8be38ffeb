// test1 -- C++ twin of the reference driver tests/test1.f90 (:21-45): grid sweep nsize = 200, 230, ... (ntests = 10
// grids), gmres_mgsr_omp with cbpr2, max_iter = 90 per restart, tol = 1.d-15, one table line per grid
// (utils.f90 print_line).  One handle serves the whole sweep (workspace, tensor maps and captured graphs are
// reused; the reference reallocates everything per call).  argv: [ntests] [first grid] [step] [tol].
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "krylov_b200.hpp"
#include "krylov_report.hpp"

int main(int argc, char **argv) {
    double tol = 1e-15;                                            // test1.f90:21
    int nsize = 200, ntests = 10, step = 30;                       // :22-23, :42
    const int max_iter = 90;                                       // :24
    if (argc > 1) ntests = std::atoi(argv[1]);
    if (argc > 2) nsize = std::atoi(argv[2]);
    if (argc > 3) step = std::atoi(argv[3]);
    if (argc > 4) tol = std::atof(argv[4]);
    const std::vector<double> params{8.2, 0.2};                    // :29
    krylov::Handle h(0);
    std::printf(" GMRES Convergence Test (MGSR with Chebyshev precond)\n");
    char header[128];
    std::snprintf(header, sizeof header, "%25s%4d%25s%s", "Number of Tests:", ntests, "Device: ", "B200");
    krylov::print_header(header);
    for (int i = 1; i <= ntests; ++i) {
        std::vector<double> ones((size_t)nsize * nsize, 1.0), b, x, errn, verr;
        krylov::apply(h, krylov::stvec, ones, b, nsize);           // :36-37  b = A*1
        int n_iter = 0, n_stages = 0;
        auto t0 = std::chrono::steady_clock::now();
        krylov::gmres_mgsr_omp(h, krylov::stvec, b, x, max_iter, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double l2, linf;
        krylov::error_norms(x, l2, linf);
        krylov::print_line(i, nsize * nsize, secs, (n_stages - 1) * max_iter + n_iter, n_stages, max_iter, tol,
                           errn[n_iter - 1], verr[n_iter - 1], l2, linf, " ");             // :40-41
        nsize += step;                                             // :42
    }
    for (int i = 0; i < 125; ++i) std::putchar('-');
    std::putchar('\n');
    return 0;
}
