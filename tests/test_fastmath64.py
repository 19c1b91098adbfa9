"""Accuracy of the short-chain float64 exp/log/log1p used by the EM kernel (csrc/fastmath64.cuh),
compiled for the host and compared with long-double libm.  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "fastmath64.cuh"
static double ulp_err(double got, long double want) {
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
    double w = (double)want;
    double u = std::nextafter(std::fabs(w), INFINITY) - std::fabs(w);
    return (double)(fabsl((long double)got - want) / u);
}
int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double e_exp = 0, e_l1p = 0, e_log = 0, e_small = 0;
    for (int i = 0; i < 2000000; ++i) {
        double d = -std::pow(10.0, -8.0 + 10.85 * U(rng));          // -1e-8 .. -700
        if (d > -700.0) e_exp = std::fmax(e_exp, ulp_err(paa::exp_nonpos(d), expl((long double)d)));
        double s = std::pow(10.0, -18.0 * U(rng));                    // 1e-18 .. 1
        e_l1p = std::fmax(e_l1p, ulp_err(paa::log1p_unit(s), log1pl((long double)s)));
        double s2 = U(rng);
        e_l1p = std::fmax(e_l1p, ulp_err(paa::log1p_unit(s2), log1pl((long double)s2)));
        double x = std::pow(10.0, -17.0 + 20.0 * U(rng));             // weights down to 1e-17, precisions to 1e3
        e_log = std::fmax(e_log, ulp_err(paa::log_pos(x), logl((long double)x)));
        double y = 1.0 + (U(rng) - 0.5) * 1e-3;                       // near 1: relative accuracy of a tiny log
        e_small = std::fmax(e_small, ulp_err(paa::log_pos(y), logl((long double)y)));
    }
    printf("%.3f %.3f %.3f %.3f\n", e_exp, e_l1p, e_log, e_small);
    printf("%d %d\n", paa::exp_nonpos(-800.0) == 0.0, paa::exp_nonpos(0.0) == 1.0);
    printf("%d\n", paa::log1p_unit(0.0) == 0.0);
    return 0;
}
'''


def test_fastmath64_accuracy(tmp_path):
    src = tmp_path / "fm.cpp"
    src.write_text(SRC)
    exe = tmp_path / "fm"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I",
                           os.path.join(ROOT, "paa_b200", "csrc"), str(src), "-o", str(exe), "-lm"])
    out = subprocess.check_output([str(exe)], text=True).split()
    e_exp, e_l1p, e_log, e_small = [float(v) for v in out[:4]]
    # a few ulp is what the EM loop needs (its float32 rounding points hide float64 noise)
    assert e_exp <= 3.0, e_exp
    assert e_l1p <= 6.0, e_l1p      # <= 7e-16 relative: the quotient s/(s+2) carries two roundings
    assert e_log <= 6.0, e_log
    assert e_small <= 6.0, e_small
    assert out[4:7] == ["1", "1", "1"]
