// dx_fastmath.cuh -- branch-free fp64 reciprocal / exp / log / log1p shared by the K_fit kernels that evaluate them per
// design group (dx_fit_lane.cu) or per cell (dx_fit_tile.cu).  Every multiply-add that may be fused is written as fma(), so
// translation units compiled with and without implicit contraction produce the same bits.
#pragma once
#include <cuda_runtime.h>

// 1 / d for d in the normal range, well inside the exponent limits (r, r + mu, count bins): the 20-bit reciprocal seed and
// two Newton steps (relative error 2^-80 before the final rounding)
__device__ __forceinline__ double fl_rcp_n(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
// pivots: anything may come in
__device__ __forceinline__ double fl_rcp(double d) {
    const double ad = fabs(d);
    return (ad > 1e-280 && ad < 1e280) ? fl_rcp_n(d) : 1.0 / d;
}

// Branch-free fp64 exp / log / log1p for the argument ranges of this kernel (<= 1 ulp against glibc on 2e6 random
// arguments each, tools/check_lane_math.py).  Without branches the compiler interleaves the evaluations of several design
// groups, which is where the instruction-level parallelism of a thread-per-gene kernel comes from; the library versions
// carry special-case branches that serialise them.
// exp(x), |x| <= 700 (linear predictors are clipped to [-297.8, 283.9]): 2^n * P(r), r = x - n ln2, |r| <= 0.347
// (the coefficients sit in constant memory: a 64-bit literal costs two move instructions per use, a constant-bank pair one
// uniform load per two coefficients)
static __constant__ double FL_EXP_C[16] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,          // 1/13! .. 1/10!
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
    1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10};
__device__ __forceinline__ double fl_exp(double x) {
    const double t = fma(x, FL_EXP_C[12], FL_EXP_C[13]);                         // n = rint(x / ln 2) in the low word
    const int n = __double2loint(t);
    const double nf = t - FL_EXP_C[13];
    double r = fma(nf, FL_EXP_C[14], x);
    r = fma(nf, FL_EXP_C[15], r);
    double q = FL_EXP_C[0];
#pragma unroll
    for (int j = 1; j < 12; ++j) q = fma(q, r, FL_EXP_C[j]);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    return __hiloint2double(__double2hiint(q) + (n << 20), __double2loint(q));
}
// log(x), x positive and normal: x = 2^e m, m in [sqrt(1/2), sqrt(2)); f = m - 1, s = f / (2 + f);
// log(m) = f - f^2/2 + s (f^2/2 + R(s^2))  (the fdlibm scheme and its minimax coefficients)
static __constant__ double FL_LOG_C[10] = {
    1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01, 6.666666666666735130e-01,
    1.479819860511658591e-01, 1.818357216161805012e-01, 2.857142874366239149e-01, 4503601774854144.0,
    6.93147180369123816490e-01, 1.90821492927058770002e-10};
__device__ __forceinline__ double fl_log(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const int up = (hi >= 0x3ff6a09f) ? 1 : 0;                                   // m > sqrt(2): halve it
    hi -= up << 20;
    e += up;
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double de = __hiloint2double(0x43300000, e ^ 0x80000000) - FL_LOG_C[7];             // (double)e without a conversion
    const double s = f * fl_rcp_n(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, FL_LOG_C[0], FL_LOG_C[1]), FL_LOG_C[2]);
    const double t2 = z * fma(w, fma(w, fma(w, FL_LOG_C[4], FL_LOG_C[5]), FL_LOG_C[6]), FL_LOG_C[3]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return fma(de, FL_LOG_C[8], -((hfsq - fma(s, hfsq + R, de * FL_LOG_C[9])) - f));
}
// log1p(x), x >= 0 (mu / r): log(u) + (x - (u - 1)) / u with u = 1 + x; +inf and NaN pass through
__device__ __forceinline__ double fl_log1p(double x) {
    const double u = 1.0 + x;
    const double c = x - (u - 1.0);
    double ru;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ru) : "d"(u));                       // the correction is below ulp(u): 20 bits do
    const double l = fma(c, ru, fl_log(u));
    return (x < 1e300) ? l : x;
}
