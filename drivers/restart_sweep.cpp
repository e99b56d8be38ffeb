// restart_sweep -- C++ twin of tests/weak_scaling.f90 (:47-62; despite its name a restart-size
// sweep m = 20, 25, ... at a fixed grid with gmres_hh_prec_omp) and of tests/strong_scaling.f90's
// table format (utils.f90 print_header / print_line).  argv = <grid size> <tests> [mgsr|hh].
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "krylov_b200.hpp"

int main(int argc, char **argv) {
    if (argc < 3) {
        std::printf(" usage ./restart_sweep <grid size> <tests> [mgsr|hh]\n");
        return 0;
    }
    const int nsize = std::atoi(argv[1]), ntests = std::atoi(argv[2]);
    const bool hh = argc > 3 && std::strcmp(argv[3], "hh") == 0;
    const double tol = 1e-8;
    const std::vector<double> params{8.2, 0.2};
    krylov::Handle h(0);
    std::vector<double> ones((size_t)nsize * nsize, 1.0), b, x, errn, verr;
    krylov::apply(h, krylov::stvec, ones, b, nsize);
    std::printf(" GMRES restart-size sweep (%s with Chebyshev precond)\n", hh ? "Householder" : "MGSR");
    std::printf("%3s%10s%10s%10s%10s%14s%14s%14s%14s%14s%10s%15s\n", "#", "Vars", "Iters", "Restarts", "gmres(n)", "Tol.",
                "L2 Norm", "L_inf Norm", "Residual", "||I-V.t*V||", "Time", "Info");
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::putchar('\n');
    int m = 20;                                                            // weak_scaling.f90:41
    for (int t = 1; t <= ntests; ++t, m += 5) {                            // :61
        int n_iter = 0, n_stages = 0;
        auto t0 = std::chrono::steady_clock::now();
        if (hh) krylov::gmres_hh_prec_omp(h, krylov::stvec, b, x, m, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
        else krylov::gmres_mgsr_omp(h, krylov::stvec, b, x, m, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double l2 = 0, linf = 0;
        for (double v : x) { l2 += (v - 1) * (v - 1); linf = std::fmax(linf, std::fabs(v - 1)); }
        std::printf("%3d%10d%10d%10d%10d%14.2E%14.4E%14.4E%14.4E%14.4E%10.4f%20s\n", t, nsize * nsize,
                    (n_stages - 1) * m + n_iter, n_stages, m, tol, std::sqrt(l2), linf, errn[n_iter - 1], verr[n_iter - 1],
                    secs, "B200");
    }
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::putchar('\n');
    return 0;
}
