// test_hilbert -- C++ twin of the reference driver tests/test_hilbert.f90:
//   ./test_hilbert <size> <max_iterations>
// Hilbert matrix (single-precision reciprocals, hilbert.f90:15), b = matmul(A, 1), tol 1e-15,
// gmres_mgsr_dense then gmres_hh_dense, same report lines.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "krylov_b200.hpp"

static void report(const char *title, int nsize, int max_iter, double tol, bool hh) {
    std::printf("%s\n", title);
    std::printf("N=%5d MAX ITER=%6d TOL=%10.2E\n", nsize, max_iter, tol);
    krylov::Handle h(0);
    std::vector<double> A, b, x, errn, verr, ones((size_t)nsize, 1.0);
    krylov::generate_matrix(h, A, nsize);                       // test_hilbert.f90:41
    krylov::matmul(h, A, ones, b);                              // :43-45
    int n_iter = 0, n_stages = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (hh) krylov::gmres_hh_dense(h, A, b, x, max_iter, tol, errn, verr, n_iter, n_stages);      // :48
    else krylov::gmres_mgsr_dense(h, A, b, x, max_iter, tol, errn, verr, n_iter, n_stages);       // :83
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double l2 = 0, linf = 0;
    for (double v : x) { l2 += (v - 1) * (v - 1); linf = std::fmax(linf, std::fabs(v - 1)); }
    std::printf("%30s%5d MAX=%5d\n", "Ierations until convergence:", n_iter, max_iter);
    std::printf("%30s%12.4E\n", "Final ||I - V.t * V||:", verr[n_iter - 1]);    // verr(n_iter), 1-based
    std::printf("%30s%12.4E\n", "Final residual:", errn[n_iter - 1]);
    std::printf("%30s%12.4E\n", "Max error L_max:", linf);
    std::printf("%30s%12.4E\n", "L2 norm:", std::sqrt(l2));
    std::printf("%30s", "Solution (first 10):");
    for (int i = 0; i < 10 && i < nsize; ++i) std::printf("%10.4f", x[i]);
    std::printf("\n%30s%10.6f secs.\n", "Elapsed time:", secs);
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::printf(" usage ./test_hilbert <size> <max_iterations>\n");
        return 0;
    }
    const int nsize = std::atoi(argv[1]), max_iter = std::atoi(argv[2]);
    const double tol = 1e-15;                                   // :33 / :71
    for (int i = 0; i < 60; ++i) std::putchar('-');
    std::putchar('\n');
    report("GMRES Hilbert Matrix Test (MGS with reorthogonalization)", nsize, max_iter, tol, false);
    for (int i = 0; i < 60; ++i) std::putchar('-');
    std::putchar('\n');
    report("GMRES Hilbert Matrix Test (Householder version)", nsize, max_iter, tol, true);
    for (int i = 0; i < 60; ++i) std::putchar('-');
    std::putchar('\n');
    return 0;
}
