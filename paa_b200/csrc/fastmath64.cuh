// Short-dependency-chain float64 exp / log / log1p for the EM loop of select_gmm_kernel.
//
// The mixture fit is a serial chain of ~10^2 dependent float64 operations per iteration (one warp per
// GT, nothing to overlap it with), so what matters is the *depth* of each transcendental, not its
// instruction count.  These versions use Estrin evaluation (depth ~4-5 FMAs instead of 12-17 for
// Horner) and no slow-path branches; accuracy is a few ulp (checked against long-double libm in
// tests/test_fastmath64.py), which is what the fit needs: its float32 rounding points make it
// insensitive to float64 noise at the 1e-15 level (oracle/gmm_oracle.py, SURVEY.md 8c sensitivity).
//
// Compiles for host and device so that the accuracy test can run without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PAA_HD __host__ __device__ __forceinline__
#else
#define PAA_HD static inline
#endif

namespace paa {

PAA_HD double f64_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof(d));
    return d;
#endif
}
PAA_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof(u));
    return u;
#endif
}

// a / b for positive normal b: reciprocal seed (MUFU.RCP64H, ~20 bits) + two Newton steps + one residual
// correction.  ~9 dependent instructions instead of the ~25 (plus slow-path branch) of the IEEE
// division routine; the quotient is within 1 ulp of a/b.
PAA_HD double div_fast(double a, double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
#else
    return a / b;
#endif
}

// 1 / b for positive normal b: reciprocal seed + two Newton steps (5 dependent instructions), ~1 ulp.
PAA_HD double rcp_fast(double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    return fma(fma(-b, r, 1.0), r, r);
#else
    return 1.0 / b;
#endif
}

// exp(d) for d <= 0.  Returns 0 below -708 (the result would be < 3e-308 and only ever feeds sums of O(1)).
// Round-to-nearest of d*log2(e) by adding 1.5 * 2^52: the integer lands in the low mantissa bits.
PAA_HD double exp_nonpos(double d) {
    if (!(d > -708.0)) return 0.0;
    const double kMagic = 6755399441055744.0;
    const double t = fma(d, 1.4426950408889634074, kMagic);
    const double nf = t - kMagic;
    const int n = (int)(uint32_t)f64_bits(t);                 // two's complement low word, in [-1022, 0]
    double r = fma(nf, -6.93147180369123816490e-01, d);
    r = fma(nf, -1.90821492927058770002e-10, r);
    // e^r = sum_{k<=13} r^k / k!, |r| <= 0.3466  (remainder < 5e-18), Estrin
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(r, 1.0, 1.0);
    const double a1 = fma(r, 1.6666666666666666e-01, 0.5);
    const double a2 = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double a3 = fma(r, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double a4 = fma(r, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double a5 = fma(r, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double a6 = fma(r, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    return p * f64_from_bits((uint64_t)(n + 1023) << 52);
}

// exp(d) for d <= 0 without a branch: arguments below -708 are clamped (the result, ~3e-308 instead of 0, only ever
// feeds sums of O(1)).  Same polynomial as exp_nonpos; for the EM loop, where a branch splits the instruction stream
// of the single warp that runs it.
PAA_HD double exp_nonpos_clamped(double d) {
    d = d > -708.0 ? d : -708.0;
    const double kMagic = 6755399441055744.0;
    const double t = fma(d, 1.4426950408889634074, kMagic);
    const double nf = t - kMagic;
    const int n = (int)(uint32_t)f64_bits(t);
    double r = fma(nf, -6.93147180369123816490e-01, d);
    r = fma(nf, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(r, 1.0, 1.0);
    const double a1 = fma(r, 1.6666666666666666e-01, 0.5);
    const double a2 = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double a3 = fma(r, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double a4 = fma(r, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double a5 = fma(r, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double a6 = fma(r, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    return p * f64_from_bits((uint64_t)(n + 1023) << 52);
}

// 2 * atanh(z) = log((1+z)/(1-z)) for |z| <= 0.1716 (w = z^2 <= 0.02944): 2z * sum_{k<=11} w^k/(2k+1)
PAA_HD double two_atanh_small(double z) {
    const double w = z * z, w2 = w * w, w4 = w2 * w2, w8 = w4 * w4;
    const double a0 = fma(w, 3.3333333333333331e-01, 1.0);
    const double a1 = fma(w, 1.4285714285714285e-01, 2.0000000000000001e-01);
    const double a2 = fma(w, 9.0909090909090912e-02, 1.1111111111111110e-01);
    const double a3 = fma(w, 6.6666666666666666e-02, 7.6923076923076927e-02);
    const double a4 = fma(w, 5.2631578947368418e-02, 5.8823529411764705e-02);
    const double a5 = fma(w, 4.3478260869565216e-02, 4.7619047619047616e-02);
    const double b0 = fma(a1, w2, a0), b1 = fma(a3, w2, a2), b2 = fma(a5, w2, a4);
    const double c0 = fma(b1, w4, b0);
    const double q = fma(b2, w8, c0);
    return (z + z) * q;
}

// log1p(s) for 0 <= s <= 1 (s = exp(lo - hi) in the two-component log-sum-exp):
// 2 atanh(z) with z = s / (2 + s) <= 1/3, no range split (no cancellation), w = z^2 <= 1/9,
// 2z * sum_{k<=17} w^k/(2k+1)  (remainder < 2e-19)
PAA_HD double log1p_unit(double s) {
    const double z = div_fast(s, s + 2.0);
    const double w = z * z, w2 = w * w, w4 = w2 * w2, w8 = w4 * w4, w16 = w8 * w8;
    const double a0 = fma(w, 3.3333333333333331e-01, 1.0);
    const double a1 = fma(w, 1.4285714285714285e-01, 2.0000000000000001e-01);
    const double a2 = fma(w, 9.0909090909090912e-02, 1.1111111111111110e-01);
    const double a3 = fma(w, 6.6666666666666666e-02, 7.6923076923076927e-02);
    const double a4 = fma(w, 5.2631578947368418e-02, 5.8823529411764705e-02);
    const double a5 = fma(w, 4.3478260869565216e-02, 4.7619047619047616e-02);
    const double a6 = fma(w, 3.7037037037037035e-02, 4.0000000000000001e-02);
    const double a7 = fma(w, 3.2258064516129031e-02, 3.4482758620689655e-02);
    const double a8 = fma(w, 2.8571428571428571e-02, 3.0303030303030304e-02);
    const double b0 = fma(a1, w2, a0), b1 = fma(a3, w2, a2), b2 = fma(a5, w2, a4), b3 = fma(a7, w2, a6);
    const double c0 = fma(b1, w4, b0), c1 = fma(b3, w4, b2);
    const double q = fma(a8, w16, fma(c1, w8, c0));
    return (z + z) * q;
}

// log(x) for positive normal x
PAA_HD double log_pos(double x) {
    uint64_t u = f64_bits(x);
    int e = (int)(u >> 52) - 1023;
    double m = f64_from_bits((u & 0x000fffffffffffffull) | 0x3ff0000000000000ull);   // [1, 2)
    const bool big = m > 1.4142135623730951;          // selects, not a branch: lanes carry different arguments
    m = big ? m * 0.5 : m;
    e = big ? e + 1 : e;
    const double t = two_atanh_small(div_fast(m - 1.0, m + 1.0));
    const double ef = (double)e;
    return fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, t));
}

}  // namespace paa
