// dx_common.cuh -- device helpers shared by the dxb200 kernels (sm_100a).
// Warp primitives, fp64 special functions (digamma, trigamma, stable lgamma ratio) and the tail
// probabilities of scipy.stats used by /root/reference/diffxpy/stats/stats.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#define DX_FULL_MASK 0xffffffffu

// batchglm param_bounds (float64): log(nextafter(0,1)) / 2.5 and nextafter(log(DBL_MAX), -inf) / 2.5
#define DX_ETA_MIN (-297.77580822352261)
#define DX_ETA_MAX (283.91308089306964)
#define DX_TINY 4.9406564584124654e-324
#define DX_ETA_SCALE_FROZEN 34.5
#define DX_CHOL_PIVOT_RTOL 8.8817841970012523e-16 /* 2^-50 */
#define DX_NEWTON_MAX_STEP 5.0
#define DX_NEWTON_MAX_HALVINGS 8
#define DX_NEWTON_GAIN_RTOL 1e-14

__device__ __forceinline__ double dx_clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__device__ __forceinline__ double dx_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DX_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double dx_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(DX_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ long long dx_warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DX_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ unsigned dx_warp_sum_u32(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DX_FULL_MASK, v, o);
    return v;
}
// inclusive prefix sum across the warp
__device__ __forceinline__ long long dx_warp_scan_ll(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(DX_FULL_MASK, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// fp64 library routines behind one switch: out-of-line single copies (each inlined copy is 1-6 KB of SASS) or inlined
#ifndef DX_MATH_OUTLINE
#define DX_MATH_OUTLINE 0      // measured on B200: the call overhead costs more than the instruction-cache misses it saves
#endif
#if DX_MATH_OUTLINE
#define DX_MATH_FN static __device__ __noinline__
#else
#define DX_MATH_FN static __device__ __forceinline__
#endif
DX_MATH_FN double dx_log(double x) { return log(x); }
DX_MATH_FN double dx_exp(double x) { return exp(x); }
DX_MATH_FN double dx_log1p(double x) { return log1p(x); }
DX_MATH_FN double dx_lgamma(double x) { return lgamma(x); }

// ---------------------------------------------------------------------------------------------
// digamma / trigamma for x > 0 (recurrence up to x >= 10, then the asymptotic series)
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ double dx_digamma(double x) {
    double res = 0.0;
    while (x < 10.0) { res -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 +
                     f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return res + dx_log(x) - 0.5 / x + t;
}
static __device__ __noinline__ double dx_trigamma(double x) {
    double res = 0.0;
    while (x < 10.0) { res += 1.0 / (x * x); x += 1.0; }
    const double f = 1.0 / (x * x);
    const double t = 1.0 / 6.0 + f * (-1.0 / 30.0 + f * (1.0 / 42.0 + f * (-1.0 / 30.0 + f * (5.0 / 66.0 +
                     f * (-691.0 / 2730.0 + f * (7.0 / 6.0))))));
    return res + 1.0 / x + 0.5 * f + t * f / x;
}
// G(r, y) = lgamma(r + y) - lgamma(r) - y log r, stable for large r (second-order Stirling beyond 1e7)
static __device__ __noinline__ double dx_lgamma_ratio(double r, double y) {
    if (r >= 1e7) return y * (y - 1.0) / (2.0 * r) - y * (y - 1.0) * (2.0 * y - 1.0) / (12.0 * r * r);
    return dx_lgamma(r + y) - dx_lgamma(r) - y * dx_log(r);
}

// ---------------------------------------------------------------------------------------------
// tail probabilities
// ---------------------------------------------------------------------------------------------
// scipy.stats.norm.cdf == special.ndtr (cephes): 0.5 erfc(-x/sqrt2) split at |x| = 1/sqrt2
__device__ __forceinline__ double dx_ndtr(double a) {
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 0.70710678118654752440) return fma(0.5, erf(x), 0.5);
    double y = 0.5 * erfc(z);
    return x > 0 ? 1.0 - y : y;
}
// two-sided normal p-value the way stats.py:217 writes it, and its accurate logarithm
__device__ __forceinline__ void dx_two_sided_normal(double z, double* p_ref, double* logp) {
    z = fabs(z);
    *p_ref = 2.0 * (1.0 - dx_ndtr(z));
    const double x = z * 0.70710678118654752440;
    *logp = isnan(z) ? z : (isinf(z) ? -INFINITY : fma(-x, x, log(erfcx(x))));   // log(erfc(x)) = log(2 sf(z))
    // (explicit fma: the same bits from translation units compiled with and without multiply-add contraction)
}

// regularized incomplete gamma: P (series) / Q (modified Lentz continued fraction)
__device__ __forceinline__ double dx_igam_series(double a, double x) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 2000; ++n) {
        ap += 1.0;
        del *= x / ap;
        sum += del;
        if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - lgamma(a));
}
__device__ __forceinline__ double dx_igamc_cf(double a, double x, double* log_out) {
    const double fpmin = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / fpmin, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < fpmin) d = fpmin;
        c = b + an / c; if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    const double lg = -x + a * log(x) - lgamma(a) + log(h);
    if (log_out) *log_out = lg;
    return exp(lg);
}
// chi^2 survival: reference form 1 - cdf (stats.py:38, 281) and the accurate log of the upper tail
__device__ __forceinline__ void dx_chi2_sf(double stat, double df, double* p_ref, double* logp) {
    if (isnan(stat)) { *p_ref = stat; *logp = stat; return; }
    const double a = 0.5 * df, x = 0.5 * stat;
    if (!(x > 0.0)) { *p_ref = 1.0; *logp = 0.0; return; }
    if (isinf(x)) { *p_ref = 0.0; *logp = -INFINITY; return; }
    if (x > 1.0 && x > a) {
        double lq;
        const double q = dx_igamc_cf(a, x, &lq);
        *p_ref = 1.0 - (1.0 - q);      // cdf = 1 - Q, then 1 - cdf
        *logp = lq;
    } else {
        const double p = dx_igam_series(a, x);
        *p_ref = 1.0 - p;
        *logp = log1p(-p);
    }
}

// regularized incomplete beta I_x(a, b): the factor F = 2F1(a + b, 1; a + 1; x) of I_x = x^a (1-x)^b / (a B(a,b)) * F.
// Power series when it is short (all terms positive: sum_n (a+b)_n / (a+1)_n x^n, x < 1/2 and (a+b) x < 200: at most a few hundred terms -- the Welch
// test at a million degrees of freedom: a = 1/2, b = df/2, x = t^2 / (df + t^2), i.e. ~t^2/2 + 40 terms instead of ~sqrt(df)
// rounds of the continued fraction with four divisions each); else the continued fraction (Numerical-Recipes style Lentz).
__device__ __forceinline__ double dx_betacf(double a, double b, double x) {
    if (x < 0.5 && (a + b) * x < 200.0) {
        double term = 1.0, sum = 1.0;
        const double ab = a + b, a1 = a + 1.0;
        for (int n = 0; n < 4000; ++n) {
            term *= (ab + (double)n) / (a1 + (double)n) * x;
            sum += term;
            if (term < sum * 1e-17) return sum;
        }
    }
    const double fpmin = 1e-300;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    // (reciprocals instead of divisions: the recurrence is a serial chain of up to thousands of rounds when a is large)
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = __drcp_rn(d);
    double h = d;
    for (int m = 1; m < 5000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x * __drcp_rn((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa * __drcp_rn(c); if (fabs(c) < fpmin) c = fpmin;
        d = __drcp_rn(d);
        h *= d * c;
        aa = -(a + m) * (qab + m) * x * __drcp_rn((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa * __drcp_rn(c); if (fabs(c) < fpmin) c = fpmin;
        d = __drcp_rn(d);
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h;
}
// log Beta(a, b) with the large-argument cancellation removed
__device__ __forceinline__ double dx_lbeta(double a, double b) {
    if (a < b) { const double t = a; a = b; b = t; }
    if (a > 1e6) return lgamma(b) - (dx_lgamma_ratio(a, b) + b * log(a));
    return lgamma(a) + lgamma(b) - lgamma(a + b);
}
// I_x(a, b) with both x and y = 1 - x supplied (no cancellation in either tail)
__device__ __forceinline__ double dx_betainc_xy(double a, double b, double x, double y) {
    if (!(x > 0.0)) return 0.0;
    if (!(y > 0.0)) return 1.0;
    const double lx = (x < 0.5) ? log(x) : log1p(-y);
    const double ly = (y < 0.5) ? log(y) : log1p(-x);
    const double lbt = a * lx + b * ly - dx_lbeta(a, b);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lbt) * dx_betacf(a, b, x) / a;
    return 1.0 - exp(lbt) * dx_betacf(b, a, y) / b;
}
// 2 * scipy.stats.t.sf(t, df) for t >= 0  (stats.py:177) = I_{df/(df+t^2)}(df/2, 1/2)
__device__ __forceinline__ double dx_two_sided_t(double t, double df) {
    if (isnan(t) || isnan(df)) return NAN;
    if (isinf(t)) return 0.0;
    if (df > 1e15) { return erfc(fabs(t) * 0.70710678118654752440); }
    const double t2 = t * t;
    if (t2 > df * 1e300) return 0.0;
    const double x = df / (df + t2), y = t2 / (df + t2);
    return dx_betainc_xy(0.5 * df, 0.5, x, y);
}

// ---------------------------------------------------------------------------------------------
// small SPD linear algebra in shared memory, one warp per matrix (n <= 32 in practice)
// ---------------------------------------------------------------------------------------------
// guarded Cholesky of the n x n SPD matrix M (row-major, smem) into Lm (lower); warp cooperative.
static __device__ __noinline__ bool chol_warp(const double* M, double* Lm, int n, int lane) {
#pragma unroll 1
    for (int e = lane; e < n * n; e += 32) Lm[e] = M[e];
    __syncwarp();
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        const double v = Lm[j * n + j];
        const double mjj = M[j * n + j];
        if (!(isfinite(v) && v > DX_CHOL_PIVOT_RTOL * mjj && mjj > 0.0)) { ok = false; break; }
        const double d = sqrt(v);
        __syncwarp();
#pragma unroll 1
        for (int i = j + lane; i < n; i += 32) Lm[i * n + j] = (i == j) ? d : Lm[i * n + j] / d;
        __syncwarp();
#pragma unroll 1
        for (int i = j + 1 + lane; i < n; i += 32) {
            const double lij = Lm[i * n + j];
#pragma unroll 1
            for (int cc = j + 1; cc <= i; ++cc) Lm[i * n + cc] -= lij * Lm[cc * n + j];
        }
        __syncwarp();
    }
    return ok;
}

// solve L L^T x = rhs serially (one lane); x may alias rhs
static __device__ __noinline__ void chol_solve_serial(const double* Lm, int n, const double* rhs, double* x) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        double s = rhs[i];
#pragma unroll 1
        for (int k = 0; k < i; ++k) s -= Lm[i * n + k] * x[k];
        x[i] = s / Lm[i * n + i];
    }
#pragma unroll 1
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll 1
        for (int k = i + 1; k < n; ++k) s -= Lm[k * n + i] * x[k];
        x[i] = s / Lm[i * n + i];
    }
}

