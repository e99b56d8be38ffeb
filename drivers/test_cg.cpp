// test_cg / test_bicgstab -- C++ twins of the reference drivers tests/test_cg.f90 and
// tests/test_bicgstab.f90: 15 grids 300, 350, ... 1000, tol 1e-9 (absolute), iteration cap
// 10000, pcg_omp / pbicgstab_omp with cbpr2, same table columns.  argv[1] = cg | bicgstab.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "krylov_b200.hpp"

int main(int argc, char **argv) {
    const bool bicg = argc > 1 && std::strcmp(argv[1], "bicgstab") == 0;
    const double tol = 1e-9;                         // test_cg.f90:20
    int nsize = 300;                                 // :21
    const int ntests = argc > 2 ? std::atoi(argv[2]) : 15;   // :22
    const std::vector<double> params{8.2, 0.2};      // :30
    krylov::Handle h(0);
    std::printf(" %s\n", bicg ? "BICGSTAB Convergence Test" : "Conjugate Gradient Convergence Test");
    std::printf("%25s%4d%25s%s\n", "Number of Tests:", ntests, "Device: ", "B200");
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::printf("\n%14s%14s%14s%14s%14s%14s%14s%14s\n", "# Test", "Grid Size", "Num Iters", "Tol.", "Error", "L2", "LINF", "Time");
    for (int i = 1; i <= ntests; ++i) {
        int iter = 10000;                            // :38
        double err = 0;
        std::vector<double> ones((size_t)nsize * nsize, 1.0), b, x;
        krylov::apply(h, krylov::stvec, ones, b, nsize);
        auto t0 = std::chrono::steady_clock::now();
        if (bicg) krylov::pbicgstab_omp(h, krylov::stvec, b, x, tol, iter, err, krylov::cbpr2, params);
        else krylov::pcg_omp(h, krylov::stvec, b, x, tol, iter, err, krylov::cbpr2, params);
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double l2 = 0, linf = 0;
        for (double v : x) { l2 += (v - 1) * (v - 1); linf = std::fmax(linf, std::fabs(v - 1)); }
        std::printf("%14d%14d%14d%14.2E%14.4E%14.4E%14.4E%14.4f%14.8f\n", i, nsize * nsize, iter, tol, err,
                    std::sqrt(l2), linf, secs, x[0]);
        nsize += 50;                                 // :51
    }
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::putchar('\n');
    return 0;
}
