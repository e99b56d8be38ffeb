// krylov_report.hpp -- the reference's table printers (src/utils/utils.f90:37-51 print_header / print_line) for
// the C++ driver twins: same columns, widths and number formats (A3 A10 ... / I3 I10 I10 I10 I10 ES14.2 ES14.4 x4
// F10.4 A20), so that the twins' output lines up with the Fortran drivers'.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>

namespace krylov {

inline void print_header(const char *header) {                      // utils.f90:37-43
    std::printf(" %s\n", header);
    std::printf("%3s%10s%10s%10s%10s%14s%14s%14s%14s%14s%10s%15s\n", "#", "Vars", "Iters", "Restarts", "gmres(n)", "Tol.",
                "L2 Norm", "L_inf Norm", "Residual", "||I-V.t*V||", "Time", "Info");
    for (int i = 0; i < 150; ++i) std::putchar('-');
    std::putchar('\n');
}

inline void print_line(int test, int nvars, double cpu_time, int iterations, int restarts, int max_iters, double tol,
                       double errn, double verr, double l2, double linf, const char *desc) {   // utils.f90:45-51
    std::printf("%3d%10d%10d%10d%10d%14.2E%14.4E%14.4E%14.4E%14.4E%10.4f%20s\n", test, nvars, iterations, restarts,
                max_iters, tol, l2, linf, errn, verr, cpu_time, desc);
}

inline void error_norms(const std::vector<double> &x, double &l2, double &linf) {   // norm2(x-1), maxval(abs(x-1))
    l2 = 0;
    linf = 0;
    for (double v : x) {
        l2 += (v - 1) * (v - 1);
        linf = std::fmax(linf, std::fabs(v - 1));
    }
    l2 = std::sqrt(l2);
}

}  // namespace krylov
