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

#include <cmath>
#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define FM_FMA(a, b, c) fma((a), (b), (c))
#else
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

// Coefficient table: on the device it lives in constant memory so that every DFMA takes its
// coefficient straight from the constant bank (c[3][..]) instead of two register moves per use.
#define FM_NC 32
#define FM_TABLE                                                                                                   \
    /* 0  L2E    */ 1.4426950408889634074, /* 1  LN2_HI */ 6.93147180369123816490e-01,                              \
    /* 2  LN2_LO */ 1.90821492927058770002e-10, /* 3  MAGIC  */ 6755399441055744.0,                                 \
    /* 4..13 exp: (exp(r)-1-r)/r^2 ~ sum c[4+j] r^j, j = 0..9 */                                                    \
    0.5000000000000001, 0.16666666666666669, 0.04166666666662348, 0.008333333333330011, 0.0013888888917538296,      \
    0.00019841269863303223, 2.4801520792572694e-05, 2.75572680728901e-06, 2.762032742826824e-07,                    \
    2.5100569275813683e-08, /* 14..20 log: fdlibm Lg1..Lg7 */                                                       \
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,         \
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0
static const double FM_CH[FM_NC] = {FM_TABLE};
#if defined(__CUDACC__)
static __constant__ double FM_CD[FM_NC] = {FM_TABLE};
#endif
#if defined(__CUDA_ARCH__)
#define FMC(i) FM_CD[i]
#else
#define FMC(i) FM_CH[i]
#endif

// Selects without the NaN bookkeeping of fmax()/fmin(): the compiler turns `a > b ? a : b` into
// DSETP.MAX + fix-up moves (~7 instructions, two of them on the FP64 pipe); a PTX setp/selp pair is 3, and
// max(x, 0) needs no FP64 instruction at all (mask by the sign bit on the integer pipe).
// Callers pass ordered finite operands, or accept either operand on NaN.
#if defined(__CUDA_ARCH__)
FM_HD double fm_sel_gt(double a, double b, double x, double y) {      // a > b ? x : y
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %3, %4, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
    return r;
}
FM_HD double fm_sel_lt(double a, double b, double x, double y) {      // a < b ? x : y
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %3, %4, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
    return r;
}
#else
FM_HD double fm_sel_gt(double a, double b, double x, double y) { return a > b ? x : y; }
FM_HD double fm_sel_lt(double a, double b, double x, double y) { return a < b ? x : y; }
#endif
FM_HD double fm_max(double a, double b) { return fm_sel_gt(a, b, a, b); }
FM_HD double fm_min(double a, double b) { return fm_sel_lt(a, b, a, b); }
FM_HD double fm_relu(double x) {                                       // max(x, 0) for non-NaN x
    const int32_t hi = fm_hi(x);
    const int32_t m = ~(hi >> 31);
    return fm_make(hi & m, fm_lo(x) & m);
}

// 1/y for normal y, <= 1 ulp: ~20-bit seed r0, e = 1 - y r0, r = r0 + r0 (e + e^2)  [error e^3 ~ 2^-60]
FM_HD double fm_rcp(double y) {
    double r;
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
#else
    r = fm_make(fm_hi(1.0 / y), 0);                   // host stand-in for the ~20-bit hardware seed (MUFU.RCP64H)
#endif
    const double e = FM_FMA(-y, r, 1.0);
    const double t = FM_FMA(e, e, e);
    return FM_FMA(r, t, r);
}

// exp(x), |x| <= 700.
FM_HD double fm_exp(double x) {
    const double t = FM_FMA(x, FMC(0), FMC(3));
    const int32_t k = fm_lo(t);
    const double kf = t - FMC(3);
    double r = FM_FMA(kf, -FMC(1), x);
    r = FM_FMA(kf, -FMC(2), r);
    const double r2 = r * r;
#ifdef FM_EXP_ESTRIN
    const double p01 = FM_FMA(FMC(5), r, FMC(4));
    const double p23 = FM_FMA(FMC(7), r, FMC(6));
    const double p45 = FM_FMA(FMC(9), r, FMC(8));
    const double p67 = FM_FMA(FMC(11), r, FMC(10));
    const double p89 = FM_FMA(FMC(13), r, FMC(12));
    const double r4 = r2 * r2;
    const double q0 = FM_FMA(p23, r2, p01);
    const double q1 = FM_FMA(p67, r2, p45);
    const double r8 = r4 * r4;
    const double s0 = FM_FMA(q1, r4, q0);
    const double p = FM_FMA(p89, r8, s0);
#else
    // even / odd Horner chains in r^2: two independent chains of depth 4, ONE coefficient per DFMA (the coefficient
    // comes straight from the constant bank / a uniform register; Estrin's c1*r + c0 pairs need a vector register
    // per coefficient, ~20 registers held across the whole time loop)
    double pe = FM_FMA(FMC(12), r2, FMC(10));
    double po = FM_FMA(FMC(13), r2, FMC(11));
    pe = FM_FMA(pe, r2, FMC(8));
    po = FM_FMA(po, r2, FMC(9));
    pe = FM_FMA(pe, r2, FMC(6));
    po = FM_FMA(po, r2, FMC(7));
    pe = FM_FMA(pe, r2, FMC(4));
    po = FM_FMA(po, r2, FMC(5));
    const double p = FM_FMA(po, r, pe);
#endif
    const double y = 1.0 + FM_FMA(r2, p, r);          // in (0.70, 1.42)
    return fm_make(fm_hi(y) + (k << 20), fm_lo(y));   // * 2^k, result stays normal for |x| <= 700
}

// log(x), x finite, normal, > 0
FM_HD double fm_log(double x) {
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
    const double t1 = w * FM_FMA(w, FM_FMA(w, FMC(19), FMC(17)), FMC(15));
    const double t2 = z * FM_FMA(w, FM_FMA(w, FM_FMA(w, FMC(20), FMC(18)), FMC(16)), FMC(14));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    // log(x) = k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f)
    return FM_FMA(dk, FMC(1), -((hfsq - FM_FMA(s, hfsq + R, dk * FMC(2))) - f));
}

// ------------------------------------------------------------------------------------------------
// Table-driven exp / log for the second-generation resident kernel: a short table in SHARED memory replaces most of
// the polynomial -- 10 FP64 instructions per exp (16 above) and 12 per log (25 above).  The physics blocks of that kernel
// are FP64-pipe bound whenever both teams of a sub-partition are inside one.
//   fm_exp_tab : x = (64 k2 + j) ln2/64 + r, |r| <= ln2/128:  exp(x) = 2^k2 T[j] (1 + r + r^2/2 + ... + r^5/120),
//                T[j] = 2^(j/64) (64 doubles); truncation r^6/720 < 4e-17.  Valid for |x| <= 700.
//   fm_log_tab : x = 2^k m, m in [1, 2), j = top 7 mantissa bits, c_j ~ 1/m: r = m c_j - 1 in ONE rounding (|r| < 2^-7.9),
//                log(x) = (k + a_j) ln2 + L[j] + log1p(r) with log1p by its series to r^7; for m above sqrt(2), a_j = 1 and
//                L[j] = -log(c_j) - ln2, so that x near 1 takes k + a_j = 0 and a small L[j] (no cancellation).
//                128 pairs {c_j, L[j]}.  Valid for finite normal x > 0.
// fm_build_tables fills both tables on the host in long double; the kernels copy them to shared memory once per CTA.
// ------------------------------------------------------------------------------------------------
#define FM_EXP_TAB 64
#define FM_LOG_TAB 128
#define FM_LOG_SPLIT 53               // first j whose interval midpoint 1 + (j + 0.5)/128 lies above sqrt(2)
inline void fm_build_tables(double *etab /*[FM_EXP_TAB]*/, double *ltab /*[2 * FM_LOG_TAB]: c_j, L_j*/) {      // host only
    for (int j = 0; j < FM_EXP_TAB; ++j) etab[j] = (double)exp2l((long double)j / FM_EXP_TAB);
    for (int j = 0; j < FM_LOG_TAB; ++j) {
        const long double mid = 1.0L + ((long double)j + 0.5L) / FM_LOG_TAB;
        const double c = (double)(1.0L / mid);
        long double L = -logl((long double)c);
        if (j >= FM_LOG_SPLIT) L -= 0.693147180559945309417232121458L;
        ltab[2 * j] = c;
        ltab[2 * j + 1] = (double)L;
    }
}

FM_HD double fm_exp_tab(double x, const double *etab) {
    const double t = FM_FMA(x, 92.332482616893656768, FMC(3));            // 64 / ln2; magic 1.5 * 2^52: k = nearest integer
    const int32_t k = fm_lo(t);
    const double kf = t - FMC(3);
    double r = FM_FMA(kf, -0x1.62e42fee00000p-7, x);                      // ln2/64 hi: 21 trailing zero bits, kf * hi is exact
    r = FM_FMA(kf, -0x1.a39ef35793c76p-39, r);                            // ln2/64 lo
    const double T = etab[k & (FM_EXP_TAB - 1)];
    double q = FM_FMA(r, 8.3333333333333332e-03, 4.1666666666666664e-02);  // 1/120, 1/24
    q = FM_FMA(q, r, 1.6666666666666666e-01);
    q = FM_FMA(q, r, 0.5);
    q = FM_FMA(q, r, 1.0);
    const double y = FM_FMA(T * r, q, T);                                 // T (1 + r q)
    return fm_make(fm_hi(y) + ((k >> 6) << 20), fm_lo(y));                // * 2^(k >> 6); y in [1, 2.01), stays normal for |x| <= 700
}

FM_HD double fm_log_tab(double x, const double *ltab) {
    const int32_t hi = fm_hi(x);
    const int32_t j = (hi >> 13) & (FM_LOG_TAB - 1);
    const int32_t k = (hi >> 20) - 1023 + (j >= FM_LOG_SPLIT ? 1 : 0);
    const double m = fm_make((hi & 0x000fffff) | 0x3ff00000, fm_lo(x));  // [1, 2)
#if defined(__CUDA_ARCH__)
    const double2 cl = reinterpret_cast<const double2 *>(ltab)[j];      // one 16-byte load (the table is 16-byte aligned)
    const double c = cl.x, L = cl.y;
#else
    const double c = ltab[2 * j], L = ltab[2 * j + 1];
#endif
    const double r = FM_FMA(m, c, -1.0);
    // log1p(r) = r - r^2/2 + r^3/3 - ... + r^7/7, |r| < 2^-7.9: next term r^8/8 < 2^-66
    double p = FM_FMA(r, 1.4285714285714285e-01, -1.6666666666666666e-01);
    p = FM_FMA(p, r, 0.2);
    p = FM_FMA(p, r, -0.25);
    p = FM_FMA(p, r, 3.3333333333333331e-01);
    p = FM_FMA(p, r, -0.5);
    const double dk = (double)k;
    const double r2 = r * r;
    // (k ln2_hi + L) carries the large part; the series and k ln2_lo are added last
    return FM_FMA(dk, FMC(1), L) + (FM_FMA(r2, p, r) + dk * FMC(2));
}
