// test_poisson_mf -- C++ twin of the reference driver tests/test_poisson_mf.f90 (:1-87):
// Householder+Chebyshev then MGSR+Chebyshev GMRES on the nsize^2 Poisson grid, argv =
// <grid size> <iterations per restart> [tol], same report lines.  Runs on the GPU library.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "krylov_b200.hpp"

static double norm2_minus1(const std::vector<double> &x) {
    double s = 0;
    for (double v : x) s += (v - 1.0) * (v - 1.0);
    return std::sqrt(s);
}
static double linf_minus1(const std::vector<double> &x) {
    double m = 0;
    for (double v : x) m = std::fmax(m, std::fabs(v - 1.0));
    return m;
}

static void report(const char *title, int nsize, int max_iter, double tol, int n_iter, int n_stages,
                   const std::vector<double> &verr, const std::vector<double> &errn, const std::vector<double> &x,
                   double secs) {
    std::printf("%s\n", title);
    std::printf("N VARS=%8d MAX ITERS/STAGE=%5d    TOL=%10.2E  DEVICE= B200\n", nsize * nsize, max_iter, tol);
    std::printf("%30s%8d%10s%4d\n", "Iterations until convergence:", (n_stages - 1) * max_iter + n_iter, " Stages=", n_stages);
    std::printf("%30s%12.4E\n", "Final ||I - V.t * V||:", verr[n_iter - 1]);     // verr(n_iter)
    std::printf("%30s%12.4E\n", "Final residual:", errn[n_iter - 1]);            // errn(n_iter)
    std::printf("%30s%12.4E\n", "Max error L_max:", linf_minus1(x));
    std::printf("%30s%12.4E\n", "L2 norm:", norm2_minus1(x));
    std::printf("%30s", "First 10 solution elements");
    for (int i = 0; i < 10 && i < (int)x.size(); ++i) std::printf("%10.4f", x[i]);
    std::printf("\n%30s%12.4f secs.\n", "Elapsed time:", secs);
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::printf(" usage ./test_poisson <grid size> <iterations per restart> [tol]\n");
        return 0;
    }
    const int nsize = std::atoi(argv[1]), max_iter = std::atoi(argv[2]);
    const double tol = argc > 3 ? std::atof(argv[3]) : 1e-15;              // test_poisson_mf.f90:35
    const std::vector<double> params{8.2, 0.2};                            // :38
    krylov::Handle h(0);
    std::vector<double> ones((size_t)nsize * nsize, 1.0), b, x, errn, verr;
    krylov::apply(h, krylov::stvec, ones, b, nsize);                       // call stvec(x,b,nsize)  :40
    int n_iter = 0, n_stages = 0;
    const char *dash = "------------------------------------------------------------";
    std::printf("%s\n", dash);
    auto t0 = std::chrono::steady_clock::now();
    krylov::gmres_hh_prec_omp(h, krylov::stvec, b, x, max_iter, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    report("GMRES Poisson 2D Test Matrix Free (Householder Chebyshev)", nsize, max_iter, tol, n_iter, n_stages, verr, errn, x, secs);
    std::printf("%s\n", dash);
    t0 = std::chrono::steady_clock::now();
    krylov::gmres_mgsr_omp(h, krylov::stvec, b, x, max_iter, tol, errn, verr, n_iter, n_stages, krylov::cbpr2, params);
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    report("GMRES Poisson 2D Test Matrix Free (MGSR B200 Chebyshev version)", nsize, max_iter, tol, n_iter, n_stages, verr, errn, x, secs);
    std::printf("%s\n", dash);
    return 0;
}
