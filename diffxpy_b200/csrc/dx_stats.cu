// dx_stats.cu -- K_stats and K_twogroup.
//  * /root/reference/diffxpy/stats/stats.py restated on device: wald_test (:181-218), wald_test_chisq
//    (:221-282), likelihood_ratio_test (:9-39), two_coef_z_test (:285-326), t_test_moments (:116-178);
//    every kernel writes the p-value in the reference's own floating-point form (1 - cdf, which
//    saturates at 0 for |z| > 8.3) and, where asked, an accurate log p from the survival function.
//  * /root/reference/diffxpy/testing/det.py:1545-1620 / 1672-1743: per-gene group moments, Welch t and
//    the Mann-Whitney U test (scipy.stats.mannwhitneyu asymptotic method, tie + continuity corrected)
//    computed from the 2-group tail counts of K_ingest with exact integer rank sums.
#include <cstdlib>
#include <cooperative_groups.h>
#include "dx_common.cuh"
#include "dx_internal.h"
#include "dx_fit_args.cuh"

namespace {

__global__ void dx_wald_kernel(int64_t n, const double* __restrict__ theta, const double* __restrict__ sd,
                               double* __restrict__ pval, double* __restrict__ logp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // stats.py:215 floors sd at the smallest subnormal with np.nextafter(..., where=sd < tiny): a NaN sd (singular Fisher
    // information) fails that comparison and stays NaN, so the gene gets p = NaN and is left out of the correction.
    // (fmax would return the non-NaN operand and turn such a gene into p = 0.)
    const double s0 = sd[i];
    const double s = (s0 < DX_TINY) ? DX_TINY : s0;
    double p, lp;
    dx_two_sided_normal(theta[i] / s, &p, &lp);
    if (isnan(s0) || isnan(theta[i])) { p = NAN; lp = NAN; }
    pval[i] = p;
    if (logp) logp[i] = lp;
}

__global__ void dx_wald_from_fit_kernel(int64_t G, int64_t stride, int p, int coef, const double* __restrict__ a_var,
                                        const double* __restrict__ fim_inv, double* __restrict__ theta,
                                        double* __restrict__ sd, double* __restrict__ pval, double* __restrict__ logp) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const double t = a_var[(int64_t)coef * stride + g];
    double v = fim_inv[g * (int64_t)p * p + coef * p + coef];
    if (v < DX_TINY) v = DX_TINY;                // det.py:796 (NaN stays NaN)
    const double s = sqrt(v);
    double pv, lp;
    dx_two_sided_normal(t / ((s < DX_TINY) ? DX_TINY : s), &pv, &lp);
    if (isnan(s) || isnan(t)) { pv = NAN; lp = NAN; }            // singular Fisher information: no test, not p = 0
    if (theta) theta[g] = t;
    if (sd) sd[g] = s;
    pval[g] = pv;
    if (logp) logp[g] = lp;
}

// one warp per gene: stat = d^T cov^-1 d through the guarded Cholesky of cov
__global__ void dx_wald_chisq_kernel(int64_t G, int k, const double* __restrict__ theta, const double* __restrict__ cov,
                                     double* __restrict__ stat, double* __restrict__ pval, double* __restrict__ logp) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (g >= G) return;
    double* M = sm + (size_t)warp * (2 * k * k + k);
    double* L = M + k * k;
    double* y = L + k * k;
    bool fin = true;
    for (int e = lane; e < k * k; e += 32) { M[e] = cov[g * (int64_t)k * k + e]; fin = fin && isfinite(M[e]); }
    fin = __all_sync(DX_FULL_MASK, fin);
    __syncwarp();
    bool ok = fin && chol_warp(M, L, k, lane);
    __syncwarp();
    if (lane == 0) {
        double w = NAN;
        if (ok) {
            w = 0.0;
            for (int i = 0; i < k; ++i) {
                double s = theta[(int64_t)i * G + g];
                for (int j = 0; j < i; ++j) s -= L[i * k + j] * y[j];
                y[i] = s / L[i * k + i];
                w += y[i] * y[i];
            }
        }
        double pv, lp;
        dx_chi2_sf(w, (double)k, &pv, &lp);
        if (stat) stat[g] = w;
        pval[g] = pv;
        if (logp) logp[g] = lp;
    }
}

__global__ void dx_lrt_kernel(int64_t n, const double* __restrict__ llf, const double* __restrict__ llr, int ddf,
                              double* __restrict__ pval, double* __restrict__ logp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dev = 2.0 * (llf[i] - llr[i]);
    double pv, lp;
    dx_chi2_sf(dev, (double)ddf, &pv, &lp);
    pval[i] = pv;
    if (logp) logp[i] = lp;
}

__global__ void dx_two_coef_z_kernel(int64_t n, const double* __restrict__ t0, const double* __restrict__ t1,
                                     const double* __restrict__ sd0, const double* __restrict__ sd1,
                                     double* __restrict__ pval) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = sd0[i] * sd0[i] + sd1[i] * sd1[i];
    if (d < DX_TINY) d = DX_TINY;
    double pv, lp;
    dx_two_sided_normal(fabs(t0[i] - t1[i]) / sqrt(d), &pv, &lp);
    pval[i] = pv;
}

__device__ __forceinline__ double dx_welch_p(double mu0, double mu1, double var0, double var1, double n0, double n1) {
    const double a = var0 / n0, b = var1 / n1;
    double s = sqrt(a + b);
    s = dx_clip(s, DX_TINY, DBL_MAX);
    const double t = fabs(mu0 - mu1) / s;
    double div = a * a / (n0 - 1.0) + b * b / (n1 - 1.0);
    div = dx_clip(div, DX_TINY, DBL_MAX);
    double df = (a + b) * (a + b) / div;
    df = dx_clip(df, DX_TINY, DBL_MAX);
    return dx_two_sided_t(t, df);
}

__global__ void dx_t_moments_kernel(int64_t n, const double* __restrict__ mu0, const double* __restrict__ mu1,
                                    const double* __restrict__ var0, const double* __restrict__ var1,
                                    double n0, double n1, double* __restrict__ pval) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pval[i] = dx_welch_p(mu0[i], mu1[i], var0[i], var1[i], n0, n1);
}

// ------------------------------------------------------------------------------------------------
// two-group tests from tail counts: one CTA of 128 threads per gene
// ------------------------------------------------------------------------------------------------
constexpr int TG_T = 128;
constexpr int TG_CAP = 4096;      // distinct-value blocks the on-chip sort holds (65 count blocks + overflow)

__device__ __forceinline__ unsigned dx_float_key(float v) {   // order-preserving map float -> uint32
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(TG_T)
dx_two_group_kernel(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* __restrict__ tail_all,
                    const int32_t* __restrict__ ovf_count, const int64_t* __restrict__ indptr,
                    const float* __restrict__ ovf_val, const uint8_t* __restrict__ ovf_group, int do_rank,
                    double* __restrict__ out, int64_t* __restrict__ u1x2_out, int64_t* __restrict__ tie_out,
                    int32_t* __restrict__ flags_out) {
    __shared__ unsigned s_cnt[2][DX_HIST_BINS + 1];   // [group][0] = zeros, [group][y] = #{count == y}
    __shared__ double s_red[TG_T / 32][6];
    __shared__ unsigned s_key[TG_CAP];
    __shared__ unsigned s_idx[TG_CAP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t g = blockIdx.x;
    if (g >= G) return;
    // sample 0 = group ga; sample 1 = group gb, or every other group when gb < 0 ("versus rest").  Tail counts are
    // additive over groups, so the rest is the sum over k != ga.
    __shared__ unsigned s_tail[2][DX_HIST_BINS + 1];
    const uint32_t* tail_g = tail_all + g * (int64_t)K * DX_HIST_BINS;
    const int novf = ovf_count[g];
    const float* oval = ovf_val + indptr[g];
    const uint8_t* ogrp = ovf_group + indptr[g];
    if (tid < 2 * DX_HIST_BINS) {
        const int k = tid / DX_HIST_BINS, j = tid % DX_HIST_BINS;
        unsigned c = 0;
        if (k == 0) c = tail_g[ga * DX_HIST_BINS + j];
        else if (gb >= 0) c = tail_g[gb * DX_HIST_BINS + j];
        else { for (int kk = 0; kk < K; ++kk) if (kk != ga) c += tail_g[kk * DX_HIST_BINS + j]; }
        s_tail[k][j] = c;
    }
    if (tid < 2) s_tail[tid][DX_HIST_BINS] = 0u;
    __syncthreads();
    auto member = [&](int t) -> int {      // 0 / 1: sample of overflow entry t; -1: not part of this comparison
        const int og = ogrp[t];
        if (og == ga) return 0;
        if (gb >= 0) return (og == gb) ? 1 : -1;
        return (og < K) ? 1 : -1;
    };
    // bins from tail counts
    if (tid < 2 * DX_HIST_BINS) {
        const int k = tid / DX_HIST_BINS, j = tid % DX_HIST_BINS;
        s_cnt[k][j + 1] = s_tail[k][j] - s_tail[k][j + 1];
    }
    // moments: sum y, sum y^2, overflow counts per group
    double s0 = 0, q0 = 0, s1 = 0, q1 = 0, c0 = 0, c1 = 0;
    if (tid < 2 * DX_HIST_BINS) {
        const int k = tid / DX_HIST_BINS, j = tid % DX_HIST_BINS;
        const double cj = (double)s_tail[k][j];
        if (k == 0) { s0 += cj; q0 += cj * (2.0 * j + 1.0); } else { s1 += cj; q1 += cj * (2.0 * j + 1.0); }
    }
    for (int t = tid; t < novf; t += TG_T) {
        const double y = (double)oval[t];
        const int mb = member(t);
        if (mb == 0) { s0 += y; q0 += y * y; c0 += 1.0; } else if (mb == 1) { s1 += y; q1 += y * y; c1 += 1.0; }
    }
    s0 = dx_warp_sum(s0); q0 = dx_warp_sum(q0); s1 = dx_warp_sum(s1); q1 = dx_warp_sum(q1);
    c0 = dx_warp_sum(c0); c1 = dx_warp_sum(c1);
    if (lane == 0) { s_red[warp][0] = s0; s_red[warp][1] = q0; s_red[warp][2] = s1; s_red[warp][3] = q1; s_red[warp][4] = c0; s_red[warp][5] = c1; }
    __syncthreads();
    s0 = q0 = s1 = q1 = c0 = c1 = 0;
    for (int w = 0; w < TG_T / 32; ++w) { s0 += s_red[w][0]; q0 += s_red[w][1]; s1 += s_red[w][2]; q1 += s_red[w][3]; c0 += s_red[w][4]; c1 += s_red[w][5]; }
    const double mean0 = s0 / n0, mean1 = s1 / n1;
    const double var0 = q0 / n0 - mean0 * mean0, var1 = q1 / n1 - mean1 * mean1;
    if (tid == 0) {
        s_cnt[0][0] = (unsigned)(n0 - (double)s_tail[0][0] - c0);
        s_cnt[1][0] = (unsigned)(n1 - (double)s_tail[1][0] - c1);
    }
    int flags = 0;
    double p_rank = NAN;
    long long r1x2 = 0, tie = 0;
    if (do_rank) {
        const int m = novf + DX_HIST_BINS + 1;
        int mp = 1;
        while (mp < m) mp <<= 1;
        if (novf > 0 && mp > TG_CAP) {
            flags |= DX_FLAG_RANK_OVERFLOW;
        } else {
            if (novf > 0) {
                // blocks: idx 0..64 = count value idx; idx 65+t = overflow entry t; bitonic sort by value
                for (int i = tid; i < mp; i += TG_T) {
                    unsigned key = 0xffffffffu;
                    if (i <= DX_HIST_BINS) key = dx_float_key((float)i);
                    else if (i < m && member(i - DX_HIST_BINS - 1) >= 0) key = dx_float_key(oval[i - DX_HIST_BINS - 1]);
                    s_key[i] = key;
                    s_idx[i] = (unsigned)i;
                }
                __syncthreads();
                for (int k2 = 2; k2 <= mp; k2 <<= 1) {
                    for (int j = k2 >> 1; j > 0; j >>= 1) {
                        for (int i = tid; i < mp; i += TG_T) {
                            const int l = i ^ j;
                            if (l > i) {
                                const bool up = ((i & k2) == 0);
                                const unsigned ka = s_key[i], kb = s_key[l];
                                if ((ka > kb) == up) {
                                    s_key[i] = kb; s_key[l] = ka;
                                    const unsigned t2 = s_idx[i]; s_idx[i] = s_idx[l]; s_idx[l] = t2;
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
            } else {
                __syncthreads();
            }
            if (novf == 0) {
                // counts only (the usual case): the tie blocks are the 65 count values in ascending order -- block b holds
                // cnt0[b] + cnt1[b] cells -- so rank sum and tie term are a prefix sum and two integer reductions over the
                // bins (exact: the same integers as the serial merge below, in any order)
                __shared__ long long s_scan[TG_T];
                __shared__ long long s_r[TG_T / 32], s_t[TG_T / 32];
                const long long ta = (tid <= DX_HIST_BINS) ? (long long)s_cnt[0][tid] : 0;
                const long long tt = (tid <= DX_HIST_BINS) ? ta + (long long)s_cnt[1][tid] : 0;
                s_scan[tid] = tt;
                __syncthreads();
                for (int o = 1; o < TG_T; o <<= 1) {
                    const long long add = (tid >= o) ? s_scan[tid - o] : 0;
                    __syncthreads();
                    s_scan[tid] += add;
                    __syncthreads();
                }
                const long long cum = s_scan[tid] - tt;          // cells in the blocks below
                long long rr = ta * (2 * cum + tt + 1), ti = tt * tt * tt - tt;
                rr = dx_warp_sum_ll(rr); ti = dx_warp_sum_ll(ti);
                if (lane == 0) { s_r[warp] = rr; s_t[warp] = ti; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 0; w < TG_T / 32; ++w) { r1x2 += s_r[w]; tie += s_t[w]; }
                    const double n = n0 + n1;
                    const double u1 = 0.5 * (double)r1x2 - n0 * (n0 + 1.0) * 0.5;
                    const double mu = n0 * n1 * 0.5;
                    const double s = sqrt(n0 * n1 / 12.0 * ((n + 1.0) - (double)tie / (n * (n - 1.0))));
                    double num = u1 - mu;
                    num -= 0.5 * (double)((num > 0.0) - (num < 0.0));
                    const double z = num / s;
                    p_rank = dx_clip(erfc(fabs(z) * 0.70710678118654752440), 0.0, 1.0);
                    if (isnan(z)) p_rank = NAN;
                }
            } else if (tid == 0) {
                // serial merge of tie blocks in ascending value order (exact integers)
                long long cum = 0;
                int i = 0;
                while (i < m) {
                    long long ta = 0, tb = 0;
                    const unsigned key = (novf > 0) ? s_key[i] : (unsigned)i;
                    int e = i;
                    while (e < m && ((novf > 0) ? s_key[e] : (unsigned)e) == key) {
                        const unsigned id = (novf > 0) ? s_idx[e] : (unsigned)e;
                        if (id <= DX_HIST_BINS) { ta += s_cnt[0][id]; tb += s_cnt[1][id]; }
                        else if (id < (unsigned)m) {        // (padding of the sort shares the largest key with skipped entries)
                            const int mb = member((int)id - DX_HIST_BINS - 1);
                            if (mb == 0) ta += 1; else if (mb == 1) tb += 1;
                        }
                        ++e;
                    }
                    const long long t = ta + tb;
                    if (t > 0) {
                        r1x2 += ta * (2 * cum + t + 1);
                        tie += t * t * t - t;
                        cum += t;
                    }
                    i = e;
                }
                const double n = n0 + n1;
                const double u1 = 0.5 * (double)r1x2 - n0 * (n0 + 1.0) * 0.5;
                const double mu = n0 * n1 * 0.5;
                const double s = sqrt(n0 * n1 / 12.0 * ((n + 1.0) - (double)tie / (n * (n - 1.0))));
                double num = u1 - mu;
                num -= 0.5 * (double)((num > 0.0) - (num < 0.0));
                const double z = num / s;
                p_rank = dx_clip(erfc(fabs(z) * 0.70710678118654752440), 0.0, 1.0);
                if (isnan(z)) p_rank = NAN;
            }
        }
    }
    if (tid == 0) {
        double* o = out + g * DX_TG_NCOL;
        o[DX_TG_MEAN0] = mean0; o[DX_TG_MEAN1] = mean1; o[DX_TG_VAR0] = var0; o[DX_TG_VAR1] = var1;
        // (the Welch p-value -- a long serial continued fraction at a million degrees of freedom -- is filled in by
        // dx_two_group_welch_kernel, a thread per gene, instead of one thread of every CTA)
        o[DX_TG_PVAL_RANK] = p_rank;
        if (u1x2_out) u1x2_out[g] = r1x2 - (long long)(n0 * (n0 + 1.0));
        if (tie_out) tie_out[g] = tie;
        if (flags_out) flags_out[g] = flags;
    }
}

// Welch t-test p-values of the rows dx_two_group_kernel wrote (stats.py:116-178 arithmetic from the moments): a thread per gene
__global__ void dx_two_group_welch_kernel(int64_t G, double n0, double n1, double* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double* o = out + g * DX_TG_NCOL;
    o[DX_TG_PVAL_T] = dx_welch_p(o[DX_TG_MEAN0], o[DX_TG_MEAN1], o[DX_TG_VAR0], o[DX_TG_VAR1], n0, n1);
}

// ------------------------------------------------------------------------------------------------
// rank test of genes whose overflow list does not fit the on-chip sort of dx_two_group_kernel (normalised /
// log-transformed data: every non-zero is a non-count value; highly expressed genes).  One CTA of 1024 threads per
// flagged gene, bitonic sort of 64-bit elements (value key << 32 | block id) in a per-CTA slice of the caller's
// workspace: sub-sequences of TGB_TILE elements are sorted / merged in shared memory, the wider strides run on the
// workspace (L2-resident).  The tie blocks are then merged in parallel: every thread owns a contiguous piece of the
// sorted sequence and the runs of equal values that START in it; counts are integers, so rank sum and tie term are
// exact and independent of the order of the additions.
// ------------------------------------------------------------------------------------------------
constexpr int TGB_T = 1024;
constexpr int TGB_TILE = 8192;       // elements per shared-memory tile (64 KB)

__global__ void __launch_bounds__(TGB_T, 1)
dx_two_group_big_kernel(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* __restrict__ tail_all,
                        const int32_t* __restrict__ ovf_count, const int64_t* __restrict__ indptr,
                        const float* __restrict__ ovf_val, const uint8_t* __restrict__ ovf_group,
                        double* __restrict__ out, int64_t* __restrict__ u1x2_out, int64_t* __restrict__ tie_out,
                        int32_t* __restrict__ flags_out, unsigned long long* __restrict__ work, int64_t work_elems) {
    extern __shared__ __align__(16) unsigned long long tgb_tile[];       // [TGB_TILE]
    __shared__ unsigned s_cnt[2][DX_HIST_BINS + 1];
    __shared__ unsigned s_tail[2][DX_HIST_BINS + 1];
    __shared__ long long s_part[TGB_T / 32][2];
    __shared__ unsigned long long s_scan[TGB_T / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* E = work + (int64_t)blockIdx.x * work_elems;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        if (!(flags_out[g] & DX_FLAG_RANK_OVERFLOW)) continue;
        const int novf = ovf_count[g];
        const int m = novf + DX_HIST_BINS + 1;
        int64_t mp = TGB_TILE;
        while (mp < m) mp <<= 1;
        if (mp > work_elems) continue;            // workspace too small for this gene: the flag stays set
        const uint32_t* tail_g = tail_all + g * (int64_t)K * DX_HIST_BINS;
        const float* oval = ovf_val + indptr[g];
        const uint8_t* ogrp = ovf_group + indptr[g];
        __syncthreads();
        if (tid < 2 * DX_HIST_BINS) {
            const int k = tid / DX_HIST_BINS, j = tid % DX_HIST_BINS;
            unsigned c = 0;
            if (k == 0) c = tail_g[ga * DX_HIST_BINS + j];
            else if (gb >= 0) c = tail_g[gb * DX_HIST_BINS + j];
            else { for (int kk = 0; kk < K; ++kk) if (kk != ga) c += tail_g[kk * DX_HIST_BINS + j]; }
            s_tail[k][j] = c;
        }
        if (tid < 2) s_tail[tid][DX_HIST_BINS] = 0u;
        __syncthreads();
        if (tid < 2 * DX_HIST_BINS) {
            const int k = tid / DX_HIST_BINS, j = tid % DX_HIST_BINS;
            s_cnt[k][j + 1] = s_tail[k][j] - s_tail[k][j + 1];
        }
        auto member = [&](int t) -> int {
            const int og = ogrp[t];
            if (og == ga) return 0;
            if (gb >= 0) return (og == gb) ? 1 : -1;
            return (og < K) ? 1 : -1;
        };
        // overflow members per sample -> the zero block
        long long c0 = 0, c1 = 0;
        for (int t = tid; t < novf; t += TGB_T) {
            const int mb = member(t);
            c0 += mb == 0; c1 += mb == 1;
        }
        c0 = __reduce_add_sync(DX_FULL_MASK, (unsigned)c0); c1 = __reduce_add_sync(DX_FULL_MASK, (unsigned)c1);
        if (lane == 0) { s_part[warp][0] = c0; s_part[warp][1] = c1; }
        __syncthreads();
        if (tid == 0) {
            long long a = 0, b = 0;
            for (int w = 0; w < TGB_T / 32; ++w) { a += s_part[w][0]; b += s_part[w][1]; }
            s_cnt[0][0] = (unsigned)(n0 - (double)s_tail[0][0] - (double)a);
            s_cnt[1][0] = (unsigned)(n1 - (double)s_tail[1][0] - (double)b);
        }
        // ---- elements: id 0..64 = count block, 65 + t = overflow entry t; everything else sorts to the end ----
        for (int64_t i = tid; i < mp; i += TGB_T) {
            unsigned long long e = ~0ull;
            if (i <= DX_HIST_BINS) e = ((unsigned long long)dx_float_key((float)i) << 32) | (unsigned long long)i;
            else if (i < m && member((int)i - DX_HIST_BINS - 1) >= 0)
                e = ((unsigned long long)dx_float_key(oval[i - DX_HIST_BINS - 1]) << 32) | (unsigned long long)i;
            E[i] = e;
        }
        __syncthreads();
        // ---- bitonic sort, ascending --------------------------------------------------------------------------
        for (int64_t base = 0; base < mp; base += TGB_TILE) {               // every tile: all stages up to TGB_TILE
            for (int i = tid; i < TGB_TILE; i += TGB_T) tgb_tile[i] = E[base + i];
            __syncthreads();
            for (int k2 = 2; k2 <= TGB_TILE; k2 <<= 1) {
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < TGB_TILE; i += TGB_T) {
                        const int l = i ^ j;
                        if (l > i) {
                            unsigned long long a = tgb_tile[i], b = tgb_tile[l];
                            const bool up = (((base + i) & k2) == 0);
                            if ((a > b) == up) { tgb_tile[i] = b; tgb_tile[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int i = tid; i < TGB_TILE; i += TGB_T) E[base + i] = tgb_tile[i];
            __syncthreads();
        }
        for (int64_t k2 = 2 * (int64_t)TGB_TILE; k2 <= mp; k2 <<= 1) {
            for (int64_t j = k2 >> 1; j >= TGB_TILE; j >>= 1) {             // strides wider than a tile: on the workspace
                for (int64_t i = tid; i < mp; i += TGB_T) {
                    const int64_t l = i ^ j;
                    if (l > i) {
                        unsigned long long a = E[i], b = E[l];
                        const bool up = ((i & k2) == 0);
                        if ((a > b) == up) { E[i] = b; E[l] = a; }
                    }
                }
                __syncthreads();
            }
            for (int64_t base = 0; base < mp; base += TGB_TILE) {           // the rest of the merge inside the tiles
                for (int i = tid; i < TGB_TILE; i += TGB_T) tgb_tile[i] = E[base + i];
                __syncthreads();
                const bool up = ((base & k2) == 0);
                for (int j = TGB_TILE >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < TGB_TILE; i += TGB_T) {
                        const int l = i ^ j;
                        if (l > i) {
                            unsigned long long a = tgb_tile[i], b = tgb_tile[l];
                            if ((a > b) == up) { tgb_tile[i] = b; tgb_tile[l] = a; }
                        }
                    }
                    __syncthreads();
                }
                for (int i = tid; i < TGB_TILE; i += TGB_T) E[base + i] = tgb_tile[i];
                __syncthreads();
            }
        }
        // ---- tie blocks, in parallel -------------------------------------------------------------------------
        auto counts = [&](unsigned long long e, long long& ta, long long& tb) {
            ta = tb = 0;
            if (e == ~0ull) return;
            const unsigned id = (unsigned)(e & 0xffffffffull);
            if (id <= (unsigned)DX_HIST_BINS) { ta = s_cnt[0][id]; tb = s_cnt[1][id]; }
            else { const int mb = member((int)id - DX_HIST_BINS - 1); ta = mb == 0; tb = mb == 1; }
        };
        const int64_t per = (m + TGB_T - 1) / TGB_T;
        const int64_t lo = (int64_t)tid * per < m ? (int64_t)tid * per : m;
        const int64_t hi = lo + per < m ? lo + per : m;
        unsigned long long mine = 0;
        for (int64_t i = lo; i < hi; ++i) { long long ta, tb; counts(E[i], ta, tb); mine += (unsigned long long)(ta + tb); }
        // exclusive block scan of `mine` (thread order)
        unsigned long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(DX_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        unsigned long long before = 0;
        for (int w = 0; w < warp; ++w) before += s_scan[w];
        long long cum = (long long)(before + inc - mine);
        long long r1x2 = 0, tie = 0;
        {
            int64_t i = lo;
            // skip the tail of a run that started in an earlier piece
            if (i > 0 && i < hi) {
                const unsigned key_prev = (unsigned)(E[i - 1] >> 32);
                while (i < hi && E[i] != ~0ull && (unsigned)(E[i] >> 32) == key_prev) {
                    long long ta, tb; counts(E[i], ta, tb); cum += ta + tb; ++i;
                }
            }
            while (i < hi) {
                const unsigned long long e0 = E[i];
                if (e0 == ~0ull) break;
                const unsigned key = (unsigned)(e0 >> 32);
                long long ta = 0, tb = 0;
                int64_t e = i;
                while (e < m) {                                   // the run may extend beyond this thread's piece
                    const unsigned long long ee = E[e];
                    if (ee == ~0ull || (unsigned)(ee >> 32) != key) break;
                    long long a, b; counts(ee, a, b); ta += a; tb += b; ++e;
                }
                const long long t = ta + tb;
                if (t > 0) {
                    r1x2 += ta * (2 * cum + t + 1);
                    tie += t * t * t - t;
                    cum += t;
                }
                i = e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { r1x2 += __shfl_xor_sync(DX_FULL_MASK, r1x2, o); tie += __shfl_xor_sync(DX_FULL_MASK, tie, o); }
        __syncthreads();
        if (lane == 0) { s_part[warp][0] = r1x2; s_part[warp][1] = tie; }
        __syncthreads();
        if (tid == 0) {
            r1x2 = tie = 0;
            for (int w = 0; w < TGB_T / 32; ++w) { r1x2 += s_part[w][0]; tie += s_part[w][1]; }
            const double n = n0 + n1;
            const double u1 = 0.5 * (double)r1x2 - n0 * (n0 + 1.0) * 0.5;
            const double mu = n0 * n1 * 0.5;
            const double sdev = sqrt(n0 * n1 / 12.0 * ((n + 1.0) - (double)tie / (n * (n - 1.0))));
            double num = u1 - mu;
            num -= 0.5 * (double)((num > 0.0) - (num < 0.0));
            const double z = num / sdev;
            double p_rank = dx_clip(erfc(fabs(z) * 0.70710678118654752440), 0.0, 1.0);
            if (isnan(z)) p_rank = NAN;
            out[g * DX_TG_NCOL + DX_TG_PVAL_RANK] = p_rank;
            if (u1x2_out) u1x2_out[g] = r1x2 - (long long)(n0 * (n0 + 1.0));
            if (tie_out) tie_out[g] = tie;
            flags_out[g] &= ~DX_FLAG_RANK_OVERFLOW;
        }
        __syncthreads();
    }
}

inline unsigned grid1d(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }


// ---------------------------------------------------------------------------------------------------------
// Benjamini-Hochberg q-values (correction.py:5-24 -> statsmodels multipletests(method="fdr_bh") arithmetic:
// p sorted ascending, q_k = p_k / (k / n), running minimum from the right, clip at 1) WITHOUT a sort:
//   c_i = #{j : p_j <= p_i}  (the largest rank inside a tie block),  t_i = p_i / (c_i / n),
//   q_i = min(1, min_{j : p_j >= p_i} t_j)      -- identical values, two O(G^2 / SMs) all-pairs passes over
// shared-memory tiles (G = 2e4: ~4e8 compares per pass).  NaN p-values are excluded from n and stay NaN.
// ---------------------------------------------------------------------------------------------------------
constexpr int BH_T = 128, BH_R = 4, BH_TILE = 1024, BH_JS = 8;   // threads, p-values per thread, j-tile, j-slices (grid.y)
constexpr int BH_ST = 1024;                                       // threads of the single scan CTA

// order-preserving integer image of a p-value (p >= 0 or NaN): non-negative doubles sort like their bit patterns;
// NaN becomes the largest key, so it never satisfies "<= valid"
__device__ __forceinline__ long long dx_p_key(double p) {
    return (p == p) ? __double_as_longlong(p + 0.0) : 0x7fffffffffffffffLL;
}

// pass 1 (all pairs, register-tiled): cnt_i += #{j in slice : p_j <= p_i}; slice 0 also clears the sorted table
__global__ void __launch_bounds__(BH_T)
dx_bh_rank_kernel(int64_t G, const double* __restrict__ p, int* __restrict__ cnt, double* __restrict__ tsorted) {
    __shared__ long long sk[BH_TILE];
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * (BH_T * BH_R) + tid;
    long long ki[BH_R];
    int c[BH_R];
#pragma unroll
    for (int r = 0; r < BH_R; ++r) {
        const int64_t i = i0 + r * BH_T;
        ki[r] = (i < G) ? dx_p_key(p[i]) : -1;
        c[r] = 0;
        if (blockIdx.y == 0 && i < G) tsorted[i] = INFINITY;
    }
    const int64_t per = ((G + BH_JS - 1) / BH_JS + BH_TILE - 1) / BH_TILE * BH_TILE;
    const int64_t j_lo = (int64_t)blockIdx.y * per, j_hi = (j_lo + per < G) ? j_lo + per : G;
    for (int64_t t0 = j_lo; t0 < j_hi; t0 += BH_TILE) {
#pragma unroll
        for (int u = 0; u < BH_TILE / BH_T; ++u) {
            const int64_t j = t0 + tid + u * BH_T;
            sk[tid + u * BH_T] = (j < j_hi) ? dx_p_key(p[j]) : 0x7fffffffffffffffLL;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < BH_TILE; ++j) {
            const long long kj = sk[j];
#pragma unroll
            for (int r = 0; r < BH_R; ++r) c[r] += (kj <= ki[r]) ? 1 : 0;
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < BH_R; ++r) {
        const int64_t i = i0 + r * BH_T;
        if (i < G && c[r]) atomicAdd(cnt + i, c[r]);
    }
}

// n = #{non-NaN}; t_i = p_i / (cnt_i / n) goes to slot cnt_i - 1 of the table sorted by p (a tie block shares its
// last slot and one value; the other slots of the block stay +inf)
__global__ void __launch_bounds__(256)
dx_bh_scatter_kernel(int64_t G, const double* __restrict__ p, const int* __restrict__ cnt, double* __restrict__ tsorted) {
    __shared__ int s_cnt[8];
    const int tid = threadIdx.x;
    int nv = 0;
    for (int64_t j = tid; j < G; j += 256) { const double v = p[j]; nv += (v == v) ? 1 : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(DX_FULL_MASK, nv, o);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = nv;
    __syncthreads();
    int n = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) n += s_cnt[w];
    const int64_t i = (int64_t)blockIdx.x * 256 + tid;
    if (i < G) {
        const double pi = p[i];
        if (pi == pi) tsorted[cnt[i] - 1] = pi / ((double)cnt[i] / (double)n);      // statsmodels: p_sorted / (rank / n)
    }
}

// ---- ranks without a sort and without all pairs: monotone buckets ------------------------------------------------
// c_i = #{j : p_j <= p_i} = (elements in lower buckets) + (elements of the own bucket that are <= p_i).  The bucket map
// is monotone in p: p >= 2^-14 falls into one of 16384 linear buckets (uniform p-values: n / 16384 per bucket), smaller
// p into 4 buckets per binade (the exponent and two mantissa bits), so neither a null-dominated nor a
// signal-dominated vector piles up; p == 0 (underflow of strongly significant genes) and p == 1 (diffxpy's zero-variance
// rule) have the first and the last bucket to themselves, where the count is the bucket size.  Counting is
// exact, the result is the same integer the all-pairs kernel returns.  The 20k bucket counters fit in shared memory for
// the single-CTA scan.
constexpr int BH_LIN = 16384;
constexpr int BH_LB = 4037;                        // bucket 0: p == 0; log buckets 1 + ((bits of p) >> 50) for p < 2^-14  (1009 binades x 4)
constexpr int BH_NB = BH_LB + BH_LIN + 1;          // + linear buckets floor(p * 16384) in 1..16384

__device__ __forceinline__ int dx_bh_bucket(double p) {        // p >= 0, not NaN
    int b;
    if (p >= 0x1p-14) {
        const double sc = p * (double)BH_LIN;
        b = BH_LB + (sc >= (double)BH_LIN ? BH_LIN : (int)sc);
    } else {
        b = (p == 0.0) ? 0 : 1 + (int)(__double_as_longlong(p) >> 50);
    }
    return b < 0 ? 0 : (b >= BH_NB ? BH_NB - 1 : b);
}

__global__ void __launch_bounds__(256)
dx_bh_hist_kernel(int64_t G, const double* __restrict__ p, int* __restrict__ hist, double* __restrict__ tsorted) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= G) return;
    tsorted[i] = INFINITY;
    const double pi = p[i];
    if (pi == pi) atomicAdd(hist + dx_bh_bucket(pi), 1);
}

// exclusive scan of the bucket counts (single CTA); offs[BH_NB] = number of non-NaN p-values; clears the cursors
__global__ void __launch_bounds__(1024)
dx_bh_bucket_scan_kernel(int* __restrict__ hist_offs, int* __restrict__ cursor) {
    extern __shared__ int s_h[];             // [BH_NB + 1]
    __shared__ int s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = tid; b < BH_NB; b += 1024) { s_h[b] = hist_offs[b]; cursor[b] = 0; }
    __syncthreads();
    constexpr int PER = (BH_NB + 1023) / 1024;
    const int lo = tid * PER < BH_NB ? tid * PER : BH_NB, hi = (lo + PER < BH_NB) ? lo + PER : BH_NB;
    int sum = 0;
    for (int b = lo; b < hi; ++b) sum += s_h[b];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(DX_FULL_MASK, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    int run = before + inc - sum;
    for (int b = lo; b < hi; ++b) { const int c = s_h[b]; s_h[b] = run; run += c; }
    if (tid == 1023) s_h[BH_NB] = run;
    __syncthreads();
    for (int b = tid; b <= BH_NB; b += 1024) hist_offs[b] = s_h[b];
}

__global__ void __launch_bounds__(256)
dx_bh_place_kernel(int64_t G, const double* __restrict__ p, const int* __restrict__ offs, int* __restrict__ cursor,
                   long long* __restrict__ skey) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= G) return;
    const double pi = p[i];
    if (pi != pi) return;
    const int b = dx_bh_bucket(pi);
    skey[offs[b] + atomicAdd(cursor + b, 1)] = dx_p_key(pi);
}

// cnt_i, and t_i = p_i / (cnt_i / n) into slot cnt_i - 1 of the table sorted by p (a tie block shares its last slot and
// one value; the other slots of the block stay +inf)
__global__ void __launch_bounds__(256)
dx_bh_bucket_rank_kernel(int64_t G, const double* __restrict__ p, const int* __restrict__ offs,
                         const long long* __restrict__ skey, int* __restrict__ cnt, double* __restrict__ tsorted) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= G) return;
    const double pi = p[i];
    if (pi != pi) { cnt[i] = 0; return; }
    const int b = dx_bh_bucket(pi);
    const long long ki = dx_p_key(pi);
    const int lo = offs[b], hi = offs[b + 1];
    int c = lo;
    if (pi >= 1.0 || pi == 0.0) {      // (tested on p, not on the bucket index: ptxas 12.9 folds "b == 0 || b == BH_NB - 1"
                                       // into the predicate output of the clamping VIMNMX.RELU and takes this branch for every bucket)
        c = hi;                                              // p == 1 or p == 0: the whole bucket is one tie block
    } else {
        int c1 = 0, c2 = 0, c3 = 0, j = lo;
        for (; j + 4 <= hi; j += 4) {                        // independent loads: long tie blocks stay latency-tolerant
            const long long k0 = skey[j], k1 = skey[j + 1], k2 = skey[j + 2], k3 = skey[j + 3];
            c += (k0 <= ki) ? 1 : 0; c1 += (k1 <= ki) ? 1 : 0; c2 += (k2 <= ki) ? 1 : 0; c3 += (k3 <= ki) ? 1 : 0;
        }
        for (; j < hi; ++j) c += (skey[j] <= ki) ? 1 : 0;
        c += c1 + c2 + c3;
    }
    cnt[i] = c;
    const int n = offs[BH_NB];
    tsorted[c - 1] = pi / ((double)c / (double)n);          // statsmodels: p_sorted / (rank / n)
}

// ---- running minimum from the right, multi-CTA: chunk minima, then per-chunk suffix scan with the carry of the later
// chunks.  SORTED: the values are p_sorted[j] / ((j + 1) / n) (n = number of non-NaN = index of the first NaN) and the
// result is scattered through `order`; otherwise the values are the table of dx_bh_scatter_kernel, scanned in place.
__device__ __forceinline__ int64_t dx_first_nan(const double* __restrict__ ps, int64_t G) {
    int64_t lo = 0, hi = G;                 // NaNs sort last: first index with ps[idx] != ps[idx]
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const double v = ps[mid];
        if (v == v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <bool SORTED>
__device__ __forceinline__ double dx_bh_value(const double* __restrict__ src, int64_t j, int64_t G, double dn) {
    if (j >= G) return INFINITY;
    const double v = src[j];
    if (!SORTED) return v;
    return (v == v) ? v / ((double)(j + 1) / dn) : INFINITY;       // statsmodels: p_sorted / (rank / n)
}

template <bool SORTED>
__global__ void __launch_bounds__(BH_ST)
dx_bh_chunkmin_kernel(int64_t G, const double* __restrict__ src, double* __restrict__ chunk_min) {
    __shared__ double s_w[BH_ST / 32];
    __shared__ double s_n;
    const int tid = threadIdx.x;
    if (SORTED) { if (tid == 0) s_n = (double)dx_first_nan(src, G); __syncthreads(); }
    const double dn = SORTED ? s_n : 1.0;
    double m = dx_bh_value<SORTED>(src, (int64_t)blockIdx.x * BH_ST + tid, G, dn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(DX_FULL_MASK, m, o));
    if ((tid & 31) == 0) s_w[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        double r = INFINITY;
#pragma unroll
        for (int w = 0; w < BH_ST / 32; ++w) r = fmin(r, s_w[w]);
        chunk_min[blockIdx.x] = r;
    }
}

template <bool SORTED>
__global__ void __launch_bounds__(BH_ST)
dx_bh_suffix_kernel(int64_t G, double* __restrict__ src, const double* __restrict__ chunk_min, int n_chunks,
                    const int64_t* __restrict__ order, double* __restrict__ q) {
    __shared__ double s_w[BH_ST / 32];
    __shared__ double s_carry, s_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (SORTED) { if (tid == 0) s_n = (double)dx_first_nan(src, G); }
    // carry = minimum over all later chunks
    double c = INFINITY;
    for (int b = blockIdx.x + 1 + tid; b < n_chunks; b += BH_ST) c = fmin(c, chunk_min[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmin(c, __shfl_xor_sync(DX_FULL_MASK, c, o));
    if (lane == 0) s_w[warp] = c;
    __syncthreads();
    if (tid == 0) {
        double r = INFINITY;
#pragma unroll
        for (int w = 0; w < BH_ST / 32; ++w) r = fmin(r, s_w[w]);
        s_carry = r;
    }
    __syncthreads();
    const double carry = s_carry;
    const double dn = SORTED ? s_n : 1.0;
    const int64_t j = (int64_t)blockIdx.x * BH_ST + tid;
    const double v = dx_bh_value<SORTED>(src, j, G, dn);
    // suffix minimum inside the warp, then across the warps of the CTA
    double m = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_down_sync(DX_FULL_MASK, m, o);
        if (lane + o < 32) m = fmin(m, t);
    }
    __syncthreads();
    if (lane == 0) s_w[warp] = m;          // minimum of the whole warp
    __syncthreads();
    double later = carry;
    for (int w = warp + 1; w < BH_ST / 32; ++w) later = fmin(later, s_w[w]);
    m = fmin(m, later);
    if (j < G) {
        if (SORTED) {
            const double pj = src[j];
            q[order[j]] = (pj == pj) ? fmin(m, 1.0) : NAN;
        } else {
            src[j] = m;
        }
    }
}

// q_i = min(1, table[cnt_i - 1]) after the in-place suffix scan of the table
__global__ void __launch_bounds__(256)
dx_bh_gather_kernel(int64_t G, const double* __restrict__ p, const int* __restrict__ cnt, const double* __restrict__ table,
                    double* __restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= G) return;
    const double pi = p[i];
    q[i] = (pi == pi) ? fmin(table[cnt[i] - 1], 1.0) : NAN;
}

// ---- the whole correction in ONE CTA (n up to BH_FUSED_MAX): bucket histogram and offsets live in shared memory, the
// bucket-ordered keys / the table sorted by p / the ranks in the (L2-resident) workspace.  Same arithmetic and the same
// values as the multi-kernel path; replaces 7 launches (~50 us of launch latency for 20k p-values) by one.
// With an exchange buffer (several GPUs) the kernel first waits for the sequence flag of every rank -- the p-values
// were stored into the slot by the fit kernels of all ranks (dx_fit_args.cuh) -- and finally marks the exchange done.
constexpr int BH_FUSED_MAX = 32768;
constexpr int BH_FT = 1024;

struct BhExchange {
    unsigned char* local;        // null: p is an ordinary device vector
    int world;
    int64_t slot_bytes;
    double* pval_out;            // nullable: copy of the gathered p-values
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(BH_FT)
dx_bh_fused_kernel(int n, const double* __restrict__ p_in, double* __restrict__ q, double* __restrict__ tsorted,
                   long long* __restrict__ skey, int* __restrict__ cnt, BhExchange X) {
    extern __shared__ int s_h[];             // [BH_NB + 1]
    __shared__ int s_w[32];
    __shared__ double s_m[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* p = p_in;
    uint32_t seq = 0;
    if (X.local) {
        seq = *reinterpret_cast<volatile uint32_t*>(X.local + 264) + 1u;
        p = reinterpret_cast<const double*>(X.local + DX_EX_HEADER + (int64_t)(seq & 1u) * X.slot_bytes);
        if (tid < X.world) {
            const uint32_t* flag = reinterpret_cast<const uint32_t*>(X.local) + tid;
            unsigned long long t0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            while ((int32_t)(dx_ex_ld_acquire_sys(flag) - seq) < 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > X.timeout_ns) {          // a peer died: report instead of hanging the GPU
                    atomicExch(reinterpret_cast<uint32_t*>(X.local + 256), 1u + tid);
                    break;
                }
                __nanosleep(100);
            }
        }
    }
    for (int b = tid; b <= BH_NB; b += BH_FT) s_h[b] = 0;
    __syncthreads();
    // pass 1: bucket histogram; the table sorted by p starts at +inf
    for (int i = tid; i < n; i += BH_FT) {
        const double pi = __ldcg(p + i);
        tsorted[i] = INFINITY;
        if (X.pval_out) X.pval_out[i] = pi;
        if (pi == pi) atomicAdd(&s_h[dx_bh_bucket(pi)], 1);
    }
    __syncthreads();
    {   // exclusive scan of the bucket counts
        constexpr int PER = (BH_NB + BH_FT - 1) / BH_FT;
        const int lo = tid * PER < BH_NB ? tid * PER : BH_NB, hi = (lo + PER < BH_NB) ? lo + PER : BH_NB;
        int sum = 0;
        for (int b = lo; b < hi; ++b) sum += s_h[b];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(DX_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        int run = before + inc - sum;
        for (int b = lo; b < hi; ++b) { const int c = s_h[b]; s_h[b] = run; run += c; }
        if (tid == BH_FT - 1) s_total = run;
    }
    __syncthreads();
    const int n_valid = s_total;
    // pass 2: keys into their buckets; afterwards s_h[b] = end of bucket b (= start of bucket b + 1)
    for (int i = tid; i < n; i += BH_FT) {
        const double pi = __ldcg(p + i);
        if (pi == pi) skey[atomicAdd(&s_h[dx_bh_bucket(pi)], 1)] = dx_p_key(pi);
    }
    __syncthreads();
    // pass 3: c_i = #{p_j <= p_i}, t_i = p_i / (c_i / n) into slot c_i - 1 of the table sorted by p
    for (int i = tid; i < n; i += BH_FT) {
        const double pi = __ldcg(p + i);
        if (pi != pi) { cnt[i] = 0; continue; }
        const int b = dx_bh_bucket(pi);
        const long long ki = dx_p_key(pi);
        const int lo = b > 0 ? s_h[b - 1] : 0, hi = s_h[b];
        int c = lo;
        if (pi >= 1.0 || pi == 0.0) {
            c = hi;                                  // p == 1 or p == 0: the whole bucket is one tie block
        } else {
            for (int j = lo; j < hi; ++j) c += (skey[j] <= ki) ? 1 : 0;
        }
        cnt[i] = c;
        tsorted[c - 1] = pi / ((double)c / (double)n_valid);          // statsmodels: p_sorted / (rank / n)
    }
    __syncthreads();
    {   // pass 4: running minimum from the right over the table, in place
        const int per = (n + BH_FT - 1) / BH_FT;
        const int lo = tid * per < n ? tid * per : n, hi = (lo + per < n) ? lo + per : n;
        double m = INFINITY;
        for (int j = lo; j < hi; ++j) m = fmin(m, tsorted[j]);
        double sfx = m;                      // suffix minimum over the threads of the warp (this thread included)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_down_sync(DX_FULL_MASK, sfx, o);
            if (lane + o < 32) sfx = fmin(sfx, t);
        }
        if (lane == 0) s_m[warp] = sfx;      // minimum of the whole warp
        __syncthreads();
        double later = INFINITY;             // everything to the right of this thread's piece
        for (int w = warp + 1; w < BH_FT / 32; ++w) later = fmin(later, s_m[w]);
        const double nxt = __shfl_down_sync(DX_FULL_MASK, sfx, 1);
        if (lane < 31) later = fmin(later, nxt);
        double run = later;
        for (int j = hi - 1; j >= lo; --j) { run = fmin(run, tsorted[j]); tsorted[j] = run; }
    }
    __syncthreads();
    // pass 5: q_i = min(1, table[c_i - 1])
    for (int i = tid; i < n; i += BH_FT) {
        const double pi = __ldcg(p + i);
        q[i] = (pi == pi) ? fmin(tsorted[cnt[i] - 1], 1.0) : NAN;
    }
    if (X.local) {
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile uint32_t*>(X.local + 264) = seq;
    }
}

// ---- the whole correction in ONE CLUSTER of 8 CTAs (n up to BH_FUSED_MAX) ------------------------------------------------
// Same arithmetic as above, but 8192 threads instead of 1024 and every table in (distributed) shared memory: CTA c of the
// cluster owns elements [c EPC, (c + 1) EPC) of the input, the same range of positions of the table sorted by p, and
// buckets [c BPC, (c + 1) BPC) of the histogram; histogram increments, key placement, in-bucket counting and the final
// look-up go through distributed shared memory (mapa + remote atomics / loads), the phases are separated by the hardware
// cluster barrier.  20k p-values: ~12 us instead of ~100 us in one CTA or ~50 us of launch latency in 7 kernels.
constexpr int BH_CL = 8;
constexpr int BH_CT = 1024;
constexpr int BH_BPC = (BH_NB + 1 + BH_CL - 1) / BH_CL;          // buckets per CTA

__global__ void __cluster_dims__(BH_CL, 1, 1) __launch_bounds__(BH_CT, 1)
dx_bh_cluster_kernel(int n, const double* __restrict__ p_in, double* __restrict__ q, BhExchange X) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char bh_smem[];
    const int EPC = (n + BH_CL - 1) / BH_CL;                     // elements (and table positions) per CTA
    double* s_p = reinterpret_cast<double*>(bh_smem);             // [EPC] p-values of this CTA's elements
    long long* s_key = reinterpret_cast<long long*>(s_p + EPC);   // [EPC] bucket-ordered keys, positions of this CTA
    double* s_tab = reinterpret_cast<double*>(s_key + EPC);       // [EPC] table sorted by p, positions of this CTA
    int* s_cnt = reinterpret_cast<int*>(s_tab + EPC);             // [EPC] c_i of this CTA's elements
    int* s_off = s_cnt + ((EPC + 1) & ~1);                        // [BH_BPC] bucket counts -> offsets, buckets of this CTA
    __shared__ int s_tot[BH_CL];
    __shared__ double s_min[BH_CL];
    __shared__ int s_w[32];
    __shared__ double s_m[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i_lo = rank * EPC, n_loc = max(0, min(EPC, n - i_lo));

    const double* p = p_in;
    uint32_t seq = 0;
    if (X.local) {
        seq = *reinterpret_cast<volatile uint32_t*>(X.local + 264) + 1u;
        p = reinterpret_cast<const double*>(X.local + DX_EX_HEADER + (int64_t)(seq & 1u) * X.slot_bytes);
        if (tid < X.world) {
            const uint32_t* flag = reinterpret_cast<const uint32_t*>(X.local) + tid;
            unsigned long long t0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            while ((int32_t)(dx_ex_ld_acquire_sys(flag) - seq) < 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > X.timeout_ns) {          // a peer died: report instead of hanging the GPU
                    atomicExch(reinterpret_cast<uint32_t*>(X.local + 256), 1u + tid);
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < EPC; e += BH_CT) {
        double pi = NAN;
        if (e < n_loc) {
            pi = __ldcg(p + i_lo + e);
            if (X.pval_out) X.pval_out[i_lo + e] = pi;
        }
        s_p[e] = pi;
        s_tab[e] = INFINITY;
    }
    for (int b = tid; b < BH_BPC; b += BH_CT) s_off[b] = 0;
    cluster.sync();
    // phase 1: bucket histogram (remote shared-memory atomics)
    for (int e = tid; e < n_loc; e += BH_CT) {
        const double pi = s_p[e];
        if (pi == pi) {
            const int b = dx_bh_bucket(pi);
            atomicAdd(cluster.map_shared_rank(s_off, b / BH_BPC) + b % BH_BPC, 1);
        }
    }
    cluster.sync();
    // phase 2: exclusive scan of the bucket counts: inside the CTA, then across the cluster
    int excl_base;
    {
        constexpr int PER = (BH_BPC + BH_CT - 1) / BH_CT;
        const int lo = min(tid * PER, BH_BPC), hi = min(lo + PER, BH_BPC);
        int sum = 0;
        for (int b = lo; b < hi; ++b) sum += s_off[b];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(DX_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        excl_base = before + inc - sum;
        if (tid == BH_CT - 1) {
            const int total = before + inc;
            for (int r = 0; r < BH_CL; ++r) cluster.map_shared_rank(s_tot, r)[rank] = total;
        }
        cluster.sync();
        int base = 0;
        for (int r = 0; r < rank; ++r) base += s_tot[r];
        int run = base + excl_base;
        for (int b = lo; b < hi; ++b) { const int c = s_off[b]; s_off[b] = run; run += c; }
    }
    int n_valid = 0;
    for (int r = 0; r < BH_CL; ++r) n_valid += s_tot[r];
    cluster.sync();
    // phase 3: keys into their buckets; afterwards the offset of bucket b is its END (= start of bucket b + 1)
    for (int e = tid; e < n_loc; e += BH_CT) {
        const double pi = s_p[e];
        if (pi == pi) {
            const int b = dx_bh_bucket(pi);
            const int pos = atomicAdd(cluster.map_shared_rank(s_off, b / BH_BPC) + b % BH_BPC, 1);
            cluster.map_shared_rank(s_key, pos / EPC)[pos % EPC] = dx_p_key(pi);
        }
    }
    cluster.sync();
    // phase 4: c_i = #{p_j <= p_i}, t_i = p_i / (c_i / n) into slot c_i - 1 of the table sorted by p
    for (int e = tid; e < n_loc; e += BH_CT) {
        const double pi = s_p[e];
        if (pi != pi) { s_cnt[e] = 0; continue; }
        const int b = dx_bh_bucket(pi);
        const long long ki = dx_p_key(pi);
        const int hi = cluster.map_shared_rank(s_off, b / BH_BPC)[b % BH_BPC];
        const int lo = b > 0 ? cluster.map_shared_rank(s_off, (b - 1) / BH_BPC)[(b - 1) % BH_BPC] : 0;
        int c = lo;
        if (pi >= 1.0 || pi == 0.0) {
            c = hi;                                  // p == 1 or p == 0: the whole bucket is one tie block
        } else {
            for (int j = lo; j < hi; ++j) c += (cluster.map_shared_rank(s_key, j / EPC)[j % EPC] <= ki) ? 1 : 0;
        }
        s_cnt[e] = c;
        cluster.map_shared_rank(s_tab, (c - 1) / EPC)[(c - 1) % EPC] = pi / ((double)c / (double)n_valid);   // statsmodels: p_sorted / (rank / n)
    }
    cluster.sync();
    {   // phase 5: running minimum from the right over the table, in place: inside the CTA, then across the cluster
        const int per = (EPC + BH_CT - 1) / BH_CT;
        const int lo = min(tid * per, EPC), hi = min(lo + per, EPC);
        double m = INFINITY;
        for (int j = lo; j < hi; ++j) m = fmin(m, s_tab[j]);
        double sfx = m;                      // suffix minimum over the threads of the warp (this thread included)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_down_sync(DX_FULL_MASK, sfx, o);
            if (lane + o < 32) sfx = fmin(sfx, t);
        }
        if (lane == 0) s_m[warp] = sfx;      // minimum of the whole warp
        __syncthreads();
        double later = INFINITY;             // everything to the right of this thread's piece inside the CTA
        for (int w = warp + 1; w < BH_CT / 32; ++w) later = fmin(later, s_m[w]);
        const double nxt = __shfl_down_sync(DX_FULL_MASK, sfx, 1);
        if (lane < 31) later = fmin(later, nxt);
        if (tid == 0) {
            const double total = fmin(later, m);
            for (int r = 0; r < BH_CL; ++r) cluster.map_shared_rank(s_min, r)[rank] = total;
        }
        cluster.sync();
        for (int r = rank + 1; r < BH_CL; ++r) later = fmin(later, s_min[r]);
        double run = later;
        for (int j = hi - 1; j >= lo; --j) { run = fmin(run, s_tab[j]); s_tab[j] = run; }
    }
    cluster.sync();
    // phase 6: q_i = min(1, table[c_i - 1])
    for (int e = tid; e < n_loc; e += BH_CT) {
        const int c = s_cnt[e];
        q[i_lo + e] = (c > 0) ? fmin(cluster.map_shared_rank(s_tab, (c - 1) / EPC)[(c - 1) % EPC], 1.0) : NAN;
    }
    cluster.sync();          // nobody leaves while a peer may still read its tables
    if (X.local && rank == 0 && tid == 0) *reinterpret_cast<volatile uint32_t*>(X.local + 264) = seq;
}

static size_t dx_bh_cluster_smem(int n) {
    const int EPC = (n + BH_CL - 1) / BH_CL;
    return (size_t)EPC * 24 + (size_t)((EPC + 1) & ~1) * 4 + (size_t)BH_BPC * 4;
}
static cudaError_t dx_launch_bh_cluster(int n, const double* pval, double* qval, const BhExchange& X, cudaStream_t s) {
    const size_t smem = dx_bh_cluster_smem(n);
    cudaError_t e = cudaFuncSetAttribute(dx_bh_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dx_bh_cluster_kernel<<<BH_CL, BH_CT, smem, s>>>(n, pval, qval, X);
    return cudaGetLastError();
}

// several GPUs, more p-values than the fused kernel takes: wait for the flags, copy the slot out, mark the exchange done
__global__ void __launch_bounds__(1024)
dx_exchange_wait_kernel(int64_t n, BhExchange X) {
    const int tid = threadIdx.x;
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(X.local + 264) + 1u;
    const double* p = reinterpret_cast<const double*>(X.local + DX_EX_HEADER + (int64_t)(seq & 1u) * X.slot_bytes);
    if (tid < X.world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(X.local) + tid;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while ((int32_t)(dx_ex_ld_acquire_sys(flag) - seq) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > X.timeout_ns) { atomicExch(reinterpret_cast<uint32_t*>(X.local + 256), 1u + tid); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    for (int64_t i = tid; i < n; i += 1024) X.pval_out[i] = __ldcg(p + i);
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile uint32_t*>(X.local + 264) = seq;
}

}  // namespace

cudaError_t dx_launch_wald(int64_t n, const double* theta, const double* sd, double* pval, double* logp, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dx_wald_kernel<<<grid1d(n, 256), 256, 0, s>>>(n, theta, sd, pval, logp);
    return cudaGetLastError();
}
cudaError_t dx_launch_wald_from_fit(int64_t G, int p, int coef, const double* a_var, const double* fim_inv,
                                    double* theta, double* sd, double* pval, double* logp, int64_t coef_stride, cudaStream_t s) {
    if (G <= 0) return cudaSuccess;
    dx_wald_from_fit_kernel<<<grid1d(G, 256), 256, 0, s>>>(G, coef_stride > 0 ? coef_stride : G, p, coef, a_var, fim_inv, theta, sd,
                                                           pval, logp);
    return cudaGetLastError();
}
cudaError_t dx_launch_wald_chisq(int64_t G, int k, const double* theta, const double* cov,
                                 double* stat, double* pval, double* logp, cudaStream_t s) {
    if (G <= 0) return cudaSuccess;
    const int wpb = 4;
    const size_t smem = (size_t)wpb * (2 * k * k + k) * sizeof(double);
    dx_wald_chisq_kernel<<<grid1d(G, wpb), wpb * 32, smem, s>>>(G, k, theta, cov, stat, pval, logp);
    return cudaGetLastError();
}
cudaError_t dx_launch_lrt(int64_t n, const double* llf, const double* llr, int ddf, double* pval, double* logp, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dx_lrt_kernel<<<grid1d(n, 256), 256, 0, s>>>(n, llf, llr, ddf, pval, logp);
    return cudaGetLastError();
}
cudaError_t dx_launch_two_coef_z(int64_t n, const double* t0, const double* t1, const double* sd0, const double* sd1,
                                 double* pval, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dx_two_coef_z_kernel<<<grid1d(n, 256), 256, 0, s>>>(n, t0, t1, sd0, sd1, pval);
    return cudaGetLastError();
}
cudaError_t dx_launch_t_moments(int64_t n, const double* mu0, const double* mu1, const double* var0, const double* var1,
                                double n0, double n1, double* pval, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dx_t_moments_kernel<<<grid1d(n, 128), 128, 0, s>>>(n, mu0, mu1, var0, var1, n0, n1, pval);
    return cudaGetLastError();
}
cudaError_t dx_launch_two_group(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* tail,
                                const int32_t* ovf_count, const int64_t* indptr, const float* ovf_val, const uint8_t* ovf_group,
                                int do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, cudaStream_t s) {
    if (G <= 0) return cudaSuccess;
    dx_two_group_kernel<<<(unsigned)G, TG_T, 0, s>>>(G, K, ga, gb, n0, n1, tail, ovf_count, indptr, ovf_val, ovf_group, do_rank,
                                                     out, u1x2, tie, flags);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dx_two_group_welch_kernel<<<(unsigned)((G + 63) / 64), 64, 0, s>>>(G, n0, n1, out);
    return cudaGetLastError();
}

cudaError_t dx_launch_two_group_big(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* tail,
                                    const int32_t* ovf_count, const int64_t* indptr, const float* ovf_val,
                                    const uint8_t* ovf_group, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags,
                                    void* work, int64_t work_bytes, int64_t slice_bytes, int sm_count, cudaStream_t s) {
    if (G <= 0) return cudaSuccess;
    // one power-of-two slice of the workspace per CTA
    int64_t pow2 = TGB_TILE;
    while (pow2 * 2 * 8 <= slice_bytes) pow2 *= 2;
    int64_t grid = G < 2 * sm_count ? G : 2 * sm_count;
    if (grid > work_bytes / (pow2 * 8)) grid = work_bytes / (pow2 * 8);
    if (grid < 1 || slice_bytes < TGB_TILE * 8) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(dx_two_group_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TGB_TILE * 8);
    if (e != cudaSuccess) return e;
    dx_two_group_big_kernel<<<(unsigned)grid, TGB_T, TGB_TILE * 8, s>>>(G, K, ga, gb, n0, n1, tail, ovf_count, indptr, ovf_val,
                                                                       ovf_group, out, u1x2, tie, flags,
                                                                       (unsigned long long*)work, pow2);
    return cudaGetLastError();
}

size_t dx_bh_work_bytes_impl(int64_t n) {
    const int64_t n_chunks = (n + BH_ST - 1) / BH_ST;
    return (size_t)(8 * n + 8 * n_chunks + 8 * n + 4 * n + 4 * (2 * (int64_t)BH_NB + 2) + 16);
}

cudaError_t dx_launch_bh(int64_t n, const double* pval, double* qval, double* work, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    // work: table sorted by p [n] | chunk minima [ceil(n / 1024)] | bucket-ordered keys [n] | cnt [n] | bucket offsets
    // [BH_NB + 1] | bucket cursors [BH_NB]
    const int n_chunks = (int)((n + BH_ST - 1) / BH_ST);
    double* chunk_min = work + n;
    long long* skey = reinterpret_cast<long long*>(chunk_min + n_chunks);
    int* cnt = reinterpret_cast<int*>(skey + n);
    int* offs = cnt + n;
    int* cursor = offs + BH_NB + 1;
    static const bool all_pairs = getenv("DXB200_BH_ALLPAIRS") != nullptr;      // development hook: the previous formulation
    static const bool multi = getenv("DXB200_BH_MULTI") != nullptr;              // development hook: force the multi-kernel path
    const unsigned g256 = (unsigned)((n + 255) / 256);
    static const bool one_cta = getenv("DXB200_BH_ONECTA") != nullptr;           // development hook: the single-CTA kernel
    if (n <= BH_FUSED_MAX && !all_pairs && !multi && !one_cta) {
        BhExchange X;
        X.local = nullptr; X.world = 1; X.slot_bytes = 0; X.pval_out = nullptr; X.timeout_ns = 0;
        return dx_launch_bh_cluster((int)n, pval, qval, X, s);
    }
    if (n <= BH_FUSED_MAX && !all_pairs && !multi) {
        BhExchange X;
        X.local = nullptr; X.world = 1; X.slot_bytes = 0; X.pval_out = nullptr; X.timeout_ns = 0;
        cudaError_t e = cudaFuncSetAttribute(dx_bh_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (BH_NB + 1) * 4);
        if (e != cudaSuccess) return e;
        dx_bh_fused_kernel<<<1, BH_FT, (BH_NB + 1) * 4, s>>>((int)n, pval, qval, work, skey, cnt, X);
        return cudaGetLastError();
    }
    if (all_pairs) {
        cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(int), s);
        if (e != cudaSuccess) return e;
        const dim3 grid((unsigned)((n + BH_T * BH_R - 1) / (BH_T * BH_R)), BH_JS);
        dx_bh_rank_kernel<<<grid, BH_T, 0, s>>>(n, pval, cnt, work);
        dx_bh_scatter_kernel<<<g256, 256, 0, s>>>(n, pval, cnt, work);
    } else {
        cudaError_t e = cudaMemsetAsync(offs, 0, (size_t)(BH_NB + 1) * sizeof(int), s);
        if (e != cudaSuccess) return e;
        dx_bh_hist_kernel<<<g256, 256, 0, s>>>(n, pval, offs, work);
        e = cudaFuncSetAttribute(dx_bh_bucket_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (BH_NB + 1) * 4);
        if (e != cudaSuccess) return e;
        dx_bh_bucket_scan_kernel<<<1, 1024, (BH_NB + 1) * 4, s>>>(offs, cursor);
        dx_bh_place_kernel<<<g256, 256, 0, s>>>(n, pval, offs, cursor, skey);
        dx_bh_bucket_rank_kernel<<<g256, 256, 0, s>>>(n, pval, offs, skey, cnt, work);
    }
    dx_bh_chunkmin_kernel<false><<<n_chunks, BH_ST, 0, s>>>(n, work, chunk_min);
    dx_bh_suffix_kernel<false><<<n_chunks, BH_ST, 0, s>>>(n, work, chunk_min, n_chunks, nullptr, nullptr);
    dx_bh_gather_kernel<<<g256, 256, 0, s>>>(n, pval, cnt, work, qval);
    return cudaGetLastError();
}

cudaError_t dx_launch_bh_sorted(int64_t n, const double* p_sorted, const int64_t* order, double* qval, double* work,
                                cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int n_chunks = (int)((n + BH_ST - 1) / BH_ST);      // work: ceil(n / 1024) doubles
    dx_bh_chunkmin_kernel<true><<<n_chunks, BH_ST, 0, s>>>(n, p_sorted, work);
    dx_bh_suffix_kernel<true><<<n_chunks, BH_ST, 0, s>>>(n, const_cast<double*>(p_sorted), work, n_chunks, order, qval);
    return cudaGetLastError();
}

// q-values of the p-values that the fit kernels of all ranks stored into this rank's exchange buffer (dx_fit_args.cuh)
cudaError_t dx_launch_bh_exchange(int64_t n_total, void* local_buf, int world, int64_t slot_bytes, double* pval_all,
                                  double* qval, double* work, cudaStream_t s) {
    if (n_total <= 0 || !local_buf || world < 1 || world > DX_EX_MAX_WORLD || n_total * 8 > slot_bytes) return cudaErrorInvalidValue;
    BhExchange X;
    X.local = static_cast<unsigned char*>(local_buf); X.world = world; X.slot_bytes = slot_bytes; X.pval_out = pval_all;
    X.timeout_ns = 5000000000ull;
    static const bool one_cta = getenv("DXB200_BH_ONECTA") != nullptr;           // development hook: the single-CTA kernel
    if (n_total <= BH_FUSED_MAX && !one_cta) return dx_launch_bh_cluster((int)n_total, nullptr, qval, X, s);
    if (n_total <= BH_FUSED_MAX) {
        const int n_chunks = (int)((n_total + BH_ST - 1) / BH_ST);
        double* chunk_min = work + n_total;
        long long* skey = reinterpret_cast<long long*>(chunk_min + n_chunks);
        int* cnt = reinterpret_cast<int*>(skey + n_total);
        cudaError_t e = cudaFuncSetAttribute(dx_bh_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (BH_NB + 1) * 4);
        if (e != cudaSuccess) return e;
        dx_bh_fused_kernel<<<1, BH_FT, (BH_NB + 1) * 4, s>>>((int)n_total, nullptr, qval, work, skey, cnt, X);
        return cudaGetLastError();
    }
    if (!pval_all) return cudaErrorInvalidValue;
    dx_exchange_wait_kernel<<<1, 1024, 0, s>>>(n_total, X);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return dx_launch_bh(n_total, pval_all, qval, work, s);
}
