// dcp_fastmath.cuh — branch-free FP64 exp / log / reciprocal for the FAST arithmetic policy.
//
// CUDA's exp()/log()/operator/ are IEEE-grade but expand to ~60-90 SASS instructions each with
// special-case branches (the streamed kernel spends 2/3 of its issue slots outside the FP64 pipe).
// On this path the arguments are known to be finite and in range on the regular branch of the
// saturated-zone solver, so straight-line versions suffice:
//   fm_exp : Cody-Waite reduction with the 1.5*2^52 rounding trick, degree-11 polynomial
//            (1 + r + r^2 P9(r), approximation error 0.15 ulp), exponent patched in with integer adds.
//            Valid for |x| <= 700 (callers route anything else to the slow path).
//   fm_log : fdlibm-style log(1+f) with s = f/(2+f), 7-term even polynomial (public fdlibm constants
//            Lg1..Lg7), valid for finite normal x > 0.
//   fm_rcp : rcp.approx.ftz.f64 seed (MUFU.RCP64H) + two Newton steps, ~1 ulp, valid for normal |y|.
// Polynomial coefficients are immediate operands of DFMA via the constant bank (no register moves).
// The file also compiles on the host (tests/test_fastmath_cpu.py builds it with g++) so the
// accuracy claims are checked against glibc without a GPU.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define FM_HD __device__ __forceinline__
#define FM_FMA(a, b, c) fma((a), (b), (c))
#else
#include <cmath>
#define FM_HD inline
#define FM_FMA(a, b, c) std::fma((a), (b), (c))
#endif

FM_HD int32_t fm_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int32_t)(u >> 32);
#endif
}
FM_HD int32_t fm_lo(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int32_t)(u & 0xffffffffu);
#endif
}
FM_HD double fm_make(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x;
#endif
}

// 1/y for normal y, ~1 ulp
FM_HD double fm_rcp(double y) {
    double r;
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
#else
    r = fm_make(fm_hi(1.0 / y), 0);                   // host stand-in for the ~20-bit hardware seed (MUFU.RCP64H)
#endif
    double e = FM_FMA(-y, r, 1.0);
    r = FM_FMA(r, e, r);
    e = FM_FMA(-y, r, 1.0);
    r = FM_FMA(r, e, r);
    return r;
}

// exp(x), |x| <= 700
FM_HD double fm_exp(double x) {
    const double L2E = 1.4426950408889634074;         // 1/ln2
    const double LN2_HI = 6.93147180369123816490e-01; // fdlibm split of ln2
    const double LN2_LO = 1.90821492927058770002e-10;
    const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52: adding it rounds to nearest integer
    const double t = FM_FMA(x, L2E, MAGIC);
    const int32_t k = fm_lo(t);
    const double kf = t - MAGIC;
    double r = FM_FMA(kf, -LN2_HI, x);
    r = FM_FMA(kf, -LN2_LO, r);
    double p = 2.5100569275813683e-08;
    p = FM_FMA(p, r, 2.762032742826824e-07);
    p = FM_FMA(p, r, 2.75572680728901e-06);
    p = FM_FMA(p, r, 2.4801520792572694e-05);
    p = FM_FMA(p, r, 0.00019841269863303223);
    p = FM_FMA(p, r, 0.0013888888917538296);
    p = FM_FMA(p, r, 0.008333333333330011);
    p = FM_FMA(p, r, 0.04166666666662348);
    p = FM_FMA(p, r, 0.16666666666666669);
    p = FM_FMA(p, r, 0.5000000000000001);
    const double r2 = r * r;
    const double y = 1.0 + FM_FMA(r2, p, r);          // in (0.70, 1.42)
    return fm_make(fm_hi(y) + (k << 20), fm_lo(y));   // * 2^k, result stays normal for |x| <= 700
}

// log(x), x finite, normal, > 0
FM_HD double fm_log(double x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int32_t hi = fm_hi(x);
    int32_t k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int32_t up = (hi + 0x95f64) & 0x100000;     // mantissa above sqrt(2): halve it
    k += up >> 20;
    const double m = fm_make(hi | (up ^ 0x3ff00000), fm_lo(x));   // in [sqrt(2)/2, sqrt(2))
    const double f = m - 1.0;
    const double s = f * fm_rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * FM_FMA(w, FM_FMA(w, Lg6, Lg4), Lg2);
    const double t2 = z * FM_FMA(w, FM_FMA(w, FM_FMA(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    // log(x) = k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f)
    return FM_FMA(dk, LN2_HI, -((hfsq - FM_FMA(s, hfsq + R, dk * LN2_LO)) - f));
}
