// Host accuracy check of the FAST policy's exp / log / reciprocal (decipher_b200/csrc/dcp_fastmath.cuh compiles on the host):
// prints the worst error in ulps against long-double libm; tests/test_fastmath_cpu.py builds and runs it.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "../../decipher_b200/csrc/dcp_fastmath.cuh"

static double ulps(double got, long double ref) {
    const double r = std::fabs((double)ref);
    return std::fabs((double)((long double)got - ref)) / (std::nextafter(r, INFINITY) - r);
}

int main() {
    std::vector<double> et(FM_EXP_TAB), lt(2 * FM_LOG_TAB);
    fm_build_tables(et.data(), lt.data());
    std::mt19937_64 g(7);
    std::uniform_real_distribution<double> u(-700, 700), u1(-1, 1), ul(-690, 690), u2(0.5, 2.0), u3(0.999, 1.001);
    double e_tab = 0, e_poly = 0, l_tab = 0, l_poly = 0, l_tab_abs = 0, l_poly_abs = 0, rc = 0;
    for (int i = 0; i < 3000000; ++i) {
        const double x = (i & 1) ? u(g) : u1(g);
        const long double ref = expl((long double)x);
        e_tab = std::fmax(e_tab, ulps(fm_exp_tab(x, et.data()), ref));
        e_poly = std::fmax(e_poly, ulps(fm_exp(x), ref));
        const double y = (i % 3 == 0) ? std::exp(ul(g)) : ((i % 3 == 1) ? u2(g) : u3(g));
        const long double rl = logl((long double)y);
        const double gt = fm_log_tab(y, lt.data()), gp = fm_log(y);
        if (std::fabs((double)rl) > 1e-3) {
            l_tab = std::fmax(l_tab, ulps(gt, rl));
            l_poly = std::fmax(l_poly, ulps(gp, rl));
        } else {
            l_tab_abs = std::fmax(l_tab_abs, std::fabs((double)((long double)gt - rl)));
            l_poly_abs = std::fmax(l_poly_abs, std::fabs((double)((long double)gp - rl)));
        }
        rc = std::fmax(rc, ulps(fm_rcp(y), 1.0L / (long double)y));
    }
    printf("exp_tab %.3f exp_poly %.3f log_tab %.3f log_poly %.3f log_tab_abs %.3e log_poly_abs %.3e rcp %.3f\n", e_tab, e_poly, l_tab, l_poly,
           l_tab_abs, l_poly_abs, rc);
    return 0;
}
