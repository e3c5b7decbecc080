// dx_fit_fast.cu -- K_fit, register-resident variant: the same batchglm training schedule per gene as
// dx_fit_grouped.cu (/root/reference/diffxpy/testing/tests.py:222-248 -> Estimator.initialize / train_sequence("DEFAULT") /
// finalize; restated in oracle/nb_glm.py, method_b="newton"), specialised for the designs the headline workloads use:
//   at most 32 design groups, at most 16 location coefficients, ONE dispersion coefficient shared by all groups
//   (constant dispersion model, DX_OPT_SCALE_CONST).
// One sub-warp of LANES (16 or 32) lanes per gene, lane <-> design group.  Everything that depends on the gene lives in
// registers: the group state (eta, mu, log1p(mu/r), 1/(r+mu), n, S), the pooled tail counts (lane <-> count bin), the
// coefficient a_i in lane i, and -- during the factorisation -- row i of X^T W X in lane i.  Shared memory holds only
// what lanes exchange: the two weight vectors, the sums X^T W X / X^T W ybar (a sparse weight -> sum product whose
// pattern depends on the design only and is built once per CTA), and the L factor for the back substitution.
// Reciprocals (pivots, 1/(r+mu), 1/r) use MUFU.RCP64H + two Newton steps instead of the IEEE division sequence.
// Per IWLS step this is ~3x fewer instructions and far fewer dependent shared-memory round trips (29 cycles each)
// than the general kernel, which matters twice: throughput on one GPU (20k genes, several waves of gene slots) and
// latency when the 20k genes are split over 8 GPUs (less than one wave per GPU).
#include <cstdlib>
#include "dx_common.cuh"
#include "dx_internal.h"
#include "dx_fit_args.cuh"

namespace {

extern __shared__ __align__(16) double smem_f[];

template <int L> __device__ __forceinline__ double ff_sum(double v, unsigned m) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
template <int L> __device__ __forceinline__ int ff_max_i(int v, unsigned m) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}

// 1 / d for |d| in the normal range well inside the exponent limits (pivots, r, r + mu): reciprocal seed + two Newton
// steps (relative error ~1e-16); anything else takes the IEEE division
__device__ __forceinline__ double ff_rcp(double d) {
    const double ad = fabs(d);
    if (ad > 1e-280 && ad < 1e280) {
        double x;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
        double e = fma(-d, x, 1.0);
        x = fma(x, e, x);
        e = fma(-d, x, 1.0);
        x = fma(x, e, x);
        e = fma(-d, x, 1.0);
        return fma(x, e, x);
    }
    return 1.0 / d;
}

__device__ __forceinline__ void ff_decode(int e, int& i, int& j) {      // lower-triangle entry e -> (row i, column j <= i)
    i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    while (i * (i + 1) / 2 > e) --i;
    j = e - i * (i + 1) / 2;
}

__host__ __device__ inline int ff_misc_bytes(int ntgt, int K) {
    return 2 * ((ntgt + 1 + 3) & ~3) + 2 * ((ntgt + 7) & ~7) + 8 + ((K + 7) & ~7);
}

// ---- overflow entries (counts above 64, non-integer values), constant dispersion: rarely executed, kept out of line ----
// unreduced lane partial of sum_ovf w [lgamma(r + y) - lgamma(r)]
template <int LANES> __device__ __noinline__ double ff_ovf_T(const float* __restrict__ rec0, int n, int ostep, double r, int sl) {
    const double lr = dx_log(r);
    double part = 0.0;
#pragma unroll 1
    for (int t = sl; t < n; t += LANES) {
        const float* rec = rec0 + t * ostep;
        const double y = (double)__ldg(rec);
        const double w = ostep == 2 ? (double)__float_as_uint(__ldg(rec + 1)) : 1.0;
        part += w * (dx_lgamma_ratio(r, y) + y * lr);
    }
    return part;
}
// unreduced lane partials of sum_ovf w [psi(r + y) - psi(r)] and sum_ovf w [psi1(r + y) - psi1(r)]
template <int LANES> __device__ __noinline__ void ff_ovf_DE(const float* __restrict__ rec0, int n, int ostep, double r, int sl,
                                                            double& D, double& E) {
    const double psi_r = dx_digamma(r), psi1_r = dx_trigamma(r);
#pragma unroll 1
    for (int t = sl; t < n; t += LANES) {
        const float* rec = rec0 + t * ostep;
        const double y = (double)__ldg(rec);
        const double w = ostep == 2 ? (double)__float_as_uint(__ldg(rec + 1)) : 1.0;
        D += w * (dx_digamma(r + y) - psi_r);
        E += w * (dx_trigamma(r + y) - psi1_r);
    }
}

#ifndef DX_FAST_MINB
#define DX_FAST_MINB 4          // resident CTAs (4 warps each) per SM the register allocation aims at: 128 registers per thread
#endif

template <int LANES, int PMAX>
__global__ void __launch_bounds__(128, DX_FAST_MINB)
dx_fit_fast_kernel(const FitArgs P) {
    constexpr int SPW = 32 / LANES;                 // gene slots per warp
    const int K = P.K, p = P.p, nent = P.nent, ntgt = nent + p, KL = P.KL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sl = lane % LANES, lane_base = (lane / LANES) * LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << lane_base);
    const bool closed_form = P.opts.init_a == DX_INIT_CLOSED_FORM;

    // ---- CTA tables: design rows, pseudo-inverse (closed-form init), the weight -> sum product list ---------------
    const int t_xl = 0, t_pinv = t_xl + K * p;
    const int t_li = (t_pinv + (closed_form ? p * KL : 0) + 1) & ~1;
    const int t_misc = t_li + 2 * P.nz_max;
    double* xl = smem_f + t_xl;
    double2* li = reinterpret_cast<double2*>(smem_f + t_li);                        // {value, index into the weight buffer}
    unsigned short* tg_ptr = reinterpret_cast<unsigned short*>(smem_f + t_misc);    // [ntgt + 1] item range of a target
    unsigned char* order = reinterpret_cast<unsigned char*>(tg_ptr + ((ntgt + 1 + 3) & ~3));   // targets, longest list first
    unsigned char* tlen = order + ((ntgt + 7) & ~7);                                // [ntgt] list length (bit 7: all-ones list)
    unsigned char* fulls = tlen + ((ntgt + 7) & ~7);                                // [8] targets that are plain sums of a weight vector
    unsigned char* lgrp = fulls + 8;                                                // [K] group -> distinct location row
    __shared__ int s_n_active, s_n_full, s_use_list;

    for (int i = tid; i < K * p; i += blockDim.x) xl[i] = P.xu_loc[i];
    if (closed_form) for (int i = tid; i < p * KL; i += blockDim.x) smem_f[t_pinv + i] = P.pinv_loc[i];
    for (int k = tid; k < K; k += blockDim.x) lgrp[k] = (unsigned char)(P.loc_group ? P.loc_group[k] : k);
    __syncthreads();
    // target e < nent: entry (i, j) of X^T W X = sum_k w1_k x_ki x_kj;  target nent + i: (X^T W ybar)_i = sum_k w2_k x_ki
    auto tgt_value = [&](int e, int k) -> double {
        if (e < nent) { int i, j; ff_decode(e, i, j); return xl[k * p + i] * xl[k * p + j]; }
        return xl[k * p + (e - nent)];
    };
    for (int e = tid; e < ntgt; e += blockDim.x) {
        int c = 0;
        bool ones = true;
        for (int k = 0; k < K; ++k) {
            const double v = tgt_value(e, k);
            c += (v != 0.0) ? 1 : 0;
            ones = ones && (v == 1.0);
        }
        tlen[e] = (unsigned char)(c | ((ones && c == K) ? 0x80 : 0));
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0, nfull = 0;
        for (int e = 0; e < ntgt; ++e) {
            const int c = tlen[e] & 0x7f;
            tg_ptr[e] = (unsigned short)run;
            if ((tlen[e] & 0x80) && nfull < 8) { fulls[nfull++] = (unsigned char)e; tlen[e] = 0; }
            else { tlen[e] = (unsigned char)c; run += c; }
        }
        tg_ptr[ntgt] = (unsigned short)run;
        s_n_full = nfull;
        s_use_list = run <= P.nz_max ? 1 : 0;
        s_n_active = 0;
    }
    __syncthreads();
    const bool use_list = s_use_list != 0;
    for (int e = tid; e < ntgt; e += blockDim.x) {
        const int c = tlen[e];
        if (c == 0) continue;
        int rank = 0;                     // position in the order "longest list first" (stable)
        for (int f = 0; f < ntgt; ++f) { const int cf = tlen[f]; rank += (cf > c || (cf == c && f < e)) ? 1 : 0; }
        order[rank] = (unsigned char)e;
        atomicAdd(&s_n_active, 1);
        if (use_list) {
            int t = tg_ptr[e];
            for (int k = 0; k < K; ++k) {
                const double v = tgt_value(e, k);
                if (v != 0.0) { li[t] = make_double2(v, __longlong_as_double((long long)(e < nent ? k : LANES + k))); ++t; }
            }
        }
    }
    __syncthreads();
    const int n_active = s_n_active, n_full = s_n_full;

    // ---- per-slot shared memory ------------------------------------------------------------------------------------
    const int slot = warp * SPW + lane / LANES;
    const int ntgt_pad = (ntgt + 1) & ~1;
    double* wbuf = smem_f + P.table_doubles + slot * P.slot_doubles;      // [2 LANES]: w1 | w2
    double* tg = wbuf + 2 * LANES;                                        // [ntgt] sums
    double* Lsc = tg + ntgt_pad;                                          // [PMAX][PMAX] L_ij (i > j) at [j * PMAX + i]
    double* dv = Lsc + PMAX * PMAX;                                       // [PMAX] 1 / d_j
    double* Cj = dv + ((PMAX + 1) & ~1);                                  // [64] pooled tail counts C(j) = sum_k c_k(j)
    for (int e = sl; e < ntgt; e += LANES) tg[e] = 0.0;                   // targets without products stay zero
    __syncwarp(mask);

    // group constants of this lane (lanes >= K carry an empty group: n = S = 0)
    const bool lane_has_group = sl < K;
    const int kk = lane_has_group ? sl : K - 1;
    const double nk = lane_has_group ? P.n_k[kk] : 0.0;
    const double offk = P.offset ? P.offset[kk] : 0.0;
    const double xs0 = P.xu_scale[0];
    const double* xrow = xl + kk * p;
    const dx_fit_opts& opts = P.opts;

    for (;;) {
    int64_t g;
    {
        unsigned q = 0;
        if (sl == 0) q = atomicAdd(P.queue, 1u);
        g = (int64_t)__shfl_sync(mask, q, 0, LANES);
    }
    if (g >= P.n_genes) break;

    // ---- sufficient statistics of the gene ---------------------------------------------------------------------------
    const uint32_t* __restrict__ tail = P.tail_all + g * (int64_t)K * DX_HIST_BINS;
    double Sk = 0.0, Qk = 0.0;
    int ymax_l = 0;
    if (lane_has_group) {
        // tail counts c_k(j) = #{y > j} never increase with j: stop at the first zero
        const uint4* row = reinterpret_cast<const uint4*>(tail + sl * DX_HIST_BINS);
        unsigned long long s_u = 0ull, q_u = 0ull;
#pragma unroll 1
        for (int c4 = 0; c4 < DX_HIST_BINS / 4; ++c4) {
            const uint4 v = __ldg(row + c4);
            if (v.x == 0u) break;
            const unsigned j0 = 8u * c4 + 1u;           // 2 j + 1 at j = 4 c4
            s_u += (unsigned long long)v.x + v.y + v.z + v.w;
            q_u += (unsigned long long)v.x * j0 + (unsigned long long)v.y * (j0 + 2u) + (unsigned long long)v.z * (j0 + 4u) +
                   (unsigned long long)v.w * (j0 + 6u);
            ymax_l = 4 * c4 + (v.w ? 4 : (v.z ? 3 : (v.y ? 2 : 1)));
            if (v.w == 0u) break;
        }
        Sk = (double)s_u;          // sum_j c(j) = sum of the counts; sum_j c(j) (2 j + 1) = sum of their squares
        Qk = (double)q_u;
    }
    const int ymax = ff_max_i<LANES>(ymax_l, mask);
    const int nrr = (ymax + LANES - 1) / LANES;     // count-bin rounds in use (1 for counts up to LANES)
    double lgc = 0.0;                               // sum_i lgamma(y_i + 1) = sum_j C(j) log(j + 1)
#pragma unroll 1
    for (int rr = 0; rr < nrr; ++rr) {              // lane <-> count bin j = sl + rr LANES (this lane reads only its own bins)
        unsigned long long acc = 0ull;
#pragma unroll 4
        for (int k = 0; k < K; ++k) acc += __ldg(tail + k * DX_HIST_BINS + rr * LANES + sl);
        const double c = (double)acc;
        Cj[rr * LANES + sl] = c;
        if (acc) lgc += c * dx_log((double)(sl + rr * LANES) + 1.0);
    }
    // overflow entries: plain list (value, group) or the records of dx_pack_overflow (layout in dx_overflow.cu)
    int novf = P.ovf_count[g];
    const float* oval = P.ovf_val + P.indptr[g];
    const uint8_t* ogrp = P.ovf_group + P.indptr[g];
    const float* oval_r = oval;
    int novf_r = novf, ostep = 1;
    if (novf > 0) {
        const float* ohead = nullptr;
        if (P.ovf_packed) {
            const int n_a = P.ovf_packed[2 * g], n_b = P.ovf_packed[2 * g + 1];
            if (n_a > 0) {
                ohead = oval;
                oval += 8 * K; ogrp += 8 * K;
                oval_r = oval + 2 * n_a;
                novf = n_a; novf_r = n_b; ostep = 2;
            }
        }
        if (ohead) {
            if (lane_has_group) {
                const float* h = ohead + 8 * sl;
                const double sy = __hiloint2double(__float_as_int(__ldg(h + 1)), __float_as_int(__ldg(h + 0)));
                const double sq = __hiloint2double(__float_as_int(__ldg(h + 3)), __float_as_int(__ldg(h + 2)));
                const double sg = __hiloint2double(__float_as_int(__ldg(h + 5)), __float_as_int(__ldg(h + 4)));
                const double cn = __hiloint2double(__float_as_int(__ldg(h + 7)), __float_as_int(__ldg(h + 6)));
                if (cn > 0.0) { Sk += sy; Qk += sq; lgc += sg; }
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                double sk = 0.0, qk = 0.0;
#pragma unroll 1
                for (int t = sl; t < novf; t += LANES) {
                    if (__ldg(ogrp + t) == k) {
                        const double y = (double)__ldg(oval + t);
                        sk += y; qk += y * y; lgc += dx_lgamma(y + 1.0);
                    }
                }
                sk = ff_sum<LANES>(sk, mask); qk = ff_sum<LANES>(qk, mask);
                if (sl == k) { Sk += sk; Qk += qk; }
            }
        }
    }
    lgc = ff_sum<LANES>(lgc, mask);
    const double isf = P.offset ? dx_exp(-offk) : 1.0;
    const double sum_all = ff_sum<LANES>(Sk, mask);
    int flags = 0;
    if (sum_all == 0.0 && novf == 0) flags |= DX_FLAG_ALL_ZERO;

    // ---- initialize (batchglm init_par); lane i holds a_i, b is the same in every lane ------------------------------
    double areg = 0.0, breg = 0.0;
    if (opts.init_a == DX_INIT_STANDARD) {
        const double n_total = ff_sum<LANES>(nk, mask);          // (every lane takes part in the shuffles)
        areg = (sl == 0) ? dx_clip(dx_log(sum_all / n_total), DX_ETA_MIN, DX_ETA_MAX) : 0.0;
    } else if (closed_form) {
        // batchglm closedform_glm_mean: log of the mean of x / size_factor over every distinct LOCATION design row,
        // pushed through the pseudo-inverse of those rows (np.linalg.lstsq in the reference)
        wbuf[sl] = Sk * isf;
        wbuf[LANES + sl] = nk;
        __syncwarp(mask);
        double lm = 0.0;
        if (sl < KL) {
            double num = 0.0, den = 0.0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                if (lgrp[k] == sl) { num += wbuf[k]; den += wbuf[LANES + k]; }
            }
            lm = dx_log(num / den + DX_TINY);
        }
        __syncwarp(mask);
        wbuf[sl] = lm;
        __syncwarp(mask);
        if (sl < p) {
            double acc = 0.0;
            const double* pinv = smem_f + t_pinv + sl * KL;
#pragma unroll 1
            for (int l = 0; l < KL; ++l) acc += pinv[l] * wbuf[l];
            areg = dx_clip(acc, DX_ETA_MIN, DX_ETA_MAX);
        }
        __syncwarp(mask);
    } else if (opts.init_a == DX_INIT_GIVEN) {
        if (sl < p) areg = dx_clip(P.a_var[(int64_t)sl * P.stride + g], DX_ETA_MIN, DX_ETA_MAX);
    }
    if (opts.init_b == DX_INIT_STANDARD) {
        const double n_total = ff_sum<LANES>(nk, mask);
        const double m = ff_sum<LANES>(Sk * isf, mask) / n_total;
        const double v = ff_sum<LANES>(Qk * isf * isf, mask) / n_total - m * m;
        const double denom = fmax(v - m, 2.2227587494850775e-162);
        breg = dx_clip(dx_log(m * m / denom + DX_TINY), DX_ETA_MIN, DX_ETA_MAX);
    } else if (opts.init_b == DX_INIT_GIVEN) {
        breg = dx_clip(P.b_var[g], DX_ETA_MIN, DX_ETA_MAX);
    }

    // ---- state evaluation ---------------------------------------------------------------------------------------------
    // T(r) = sum_j C(j) log(r + j) + sum_ovf [lgamma(r + y) - lgamma(r)] - sum_i lgamma(y_i + 1)
    auto eval_T = [&](double r) -> double {
        double part = 0.0;
#pragma unroll 1
        for (int rr = 0; rr < nrr; ++rr) {
            const double c = Cj[rr * LANES + sl];
            if (c != 0.0) part += c * dx_log(r + (double)(sl + rr * LANES));
        }
        if (novf > 0) part += ff_ovf_T<LANES>(oval_r, novf_r, ostep, r, sl);
        return ff_sum<LANES>(part, mask) - lgc;
    };
    // eta_k = clip(x_k . a + off_k) with a_i taken from lane i
    auto eval_eta = [&](double a_lane) -> double {
        double e = offk;
#pragma unroll
        for (int i = 0; i < PMAX; ++i)
            if (i < p) e += xrow[i] * __shfl_sync(mask, a_lane, i, LANES);
        return dx_clip(e, DX_ETA_MIN, DX_ETA_MAX);
    };
    auto eval_ll = [&](double eta_, double zeta_, double r_, double l_, double T_) -> double {
        const double part = Sk * (eta_ - zeta_ - l_) - nk * r_ * l_;
        return ff_sum<LANES>(part, mask) + T_;
    };

    double zeta = dx_clip(xs0 * breg, DX_ETA_MIN, DX_ETA_MAX);
    double r = dx_exp(zeta), inv_r = ff_rcp(r);
    double T_r = eval_T(r);
    double eta = eval_eta(areg);
    double mu = dx_exp(eta);
    double irm = ff_rcp(r + mu);
    double l1p = dx_log1p(mu * inv_r);
    double nll = -eval_ll(eta, zeta, r, l1p, T_r);

    // gradient G and curvature H of ll wrt the linear predictor of the dispersion at the current state
    // (scale_terms of the general kernel for a shared r): gradient wrt b = xs0 G, Hessian wrt b = xs0^2 H
    auto scale_terms = [&](double& G, double& H) {
        double D = 0.0, E = 0.0;
#pragma unroll 1
        for (int rr = 0; rr < nrr; ++rr) {
            const double c = Cj[rr * LANES + sl];
            if (c != 0.0) {
                const double inv = ff_rcp(r + (double)(sl + rr * LANES));
                D += c * inv;
                E -= c * inv * inv;
            }
        }
        if (novf > 0) ff_ovf_DE<LANES>(oval_r, novf_r, ostep, r, sl, D, E);
        const double Gk = -nk * l1p + (nk * mu - Sk) * irm;
        const double Hk = (nk * mu * mu + r * Sk) * (irm * irm);
        G = r * ff_sum<LANES>(Gk + D, mask);
        H = G + r * r * ff_sum<LANES>(E, mask) + r * ff_sum<LANES>(Hk, mask);
    };

    // X^T W X (row i of the lower triangle into Arow of lane i) and X^T W ybar (returned, lane i) at the current state
    double Arow[PMAX];
    auto loc_terms = [&]() -> double {
        const double q = r * irm;
        const double w1 = nk * mu * q;
        const double w2 = Sk * q - w1;
        wbuf[sl] = w1;
        wbuf[LANES + sl] = w2;
        __syncwarp(mask);
        if (use_list) {
#pragma unroll 1
            for (int a = sl; a < n_active; a += LANES) {
                const int e = order[a];
                double acc = 0.0;
#pragma unroll 2
                for (int t = tg_ptr[e], t1 = tg_ptr[e + 1]; t < t1; ++t) {
                    const double2 it = li[t];
                    acc += wbuf[(int)__double_as_longlong(it.y)] * it.x;
                }
                tg[e] = acc;
            }
        } else {
#pragma unroll 1
            for (int a = sl; a < n_active; a += LANES) {
                const int e = order[a];
                double acc = 0.0;
                if (e < nent) {
                    int i, j;
                    ff_decode(e, i, j);
#pragma unroll 4
                    for (int k = 0; k < K; ++k) acc += wbuf[k] * (xl[k * p + i] * xl[k * p + j]);
                } else {
#pragma unroll 4
                    for (int k = 0; k < K; ++k) acc += wbuf[LANES + k] * xl[k * p + (e - nent)];
                }
                tg[e] = acc;
            }
        }
#pragma unroll 1
        for (int f = 0; f < n_full; ++f) {            // all-ones lists (intercept): a plain sum of the weight vector
            const int e = fulls[f];
            const double v = ff_sum<LANES>(e < nent ? w1 : w2, mask);
            if (sl == 0) tg[e] = v;
        }
        __syncwarp(mask);
        const double* trow = tg + (sl < p ? sl * (sl + 1) / 2 : 0);
#pragma unroll
        for (int j = 0; j < PMAX; ++j) Arow[j] = (sl < p && j <= sl) ? trow[j] : 0.0;
        const double rhs = (sl < p) ? tg[nent + sl] : 0.0;
        __syncwarp(mask);
        return rhs;
    };

    // LDL^T of the matrix held row-wise in Arow (lane <-> row).  On exit Arow[j] (j < sl) = L_ij, also stored in Lsc;
    // dinv = 1 / d_sl.  Pivot guard of oracle/nb_glm.py:chol_guarded: d_j finite and above 2^-50 of the input diagonal.
    auto ldl_factor = [&](double& dinv) -> bool {
        double m_ii = 1.0;
#pragma unroll
        for (int j = 0; j < PMAX; ++j) if (j == sl) m_ii = Arow[j];
        dinv = 0.0;
        bool ok = true;
#pragma unroll
        for (int j = 0; j < PMAX; ++j) {
            if (j < p && ok) {
                const double u = Arow[j];                 // lane j: the pivot d_j; lanes i > j: A_ij before scaling
                const bool pivot_ok = isfinite(u) && u > DX_CHOL_PIVOT_RTOL * m_ii && m_ii > 0.0;
                const unsigned bal = __ballot_sync(mask, pivot_ok);
                if (!((bal >> (lane_base + j)) & 1u)) {
                    ok = false;
                } else {
                    const double inv = __shfl_sync(mask, ff_rcp(u), j, LANES);
                    if (sl == j) dinv = inv;
                    const double lij = u * inv;
                    if (sl > j && sl < p) Lsc[j * PMAX + sl] = lij;
                    __syncwarp(mask);
#pragma unroll
                    for (int c = j + 1; c < PMAX; ++c) {
                        if (c < p) {
                            const double lcj = Lsc[j * PMAX + c];
                            if (c <= sl) Arow[c] = fma(-u, lcj, Arow[c]);
                        }
                    }
                    if (sl > j) Arow[j] = lij;
                }
            }
        }
        __syncwarp(mask);
        return ok;
    };
    // solve (L D L^T) x = rhs; lane i holds rhs_i / returns x_i
    auto ldl_solve = [&](double dinv, double x) -> double {
#pragma unroll
        for (int j = 0; j < PMAX - 1; ++j) {
            if (j < p - 1) {
                const double xj = __shfl_sync(mask, x, j, LANES);
                if (sl > j && sl < p) x = fma(-Arow[j], xj, x);
            }
        }
        x *= dinv;
#pragma unroll
        for (int j = PMAX - 1; j > 0; --j) {
            if (j < p) {
                const double xj = __shfl_sync(mask, x, j, LANES);
                if (sl < j) x = fma(-Lsc[sl * PMAX + j], xj, x);
            }
        }
        return x;
    };

    // ---- train (batchglm numpy EstimatorGlm.train, per-gene form) ----------------------------------------------------
    const bool train_loc = opts.train_loc != 0, train_scale = opts.train_scale != 0;
    const int freq = train_scale ? (train_loc ? opts.update_b_freq : 1) : 0x7fffffff;
    int epochs_until_b = freq;
    bool loc_conv = false, fully = false;
    int step = 0;
#pragma unroll 1
    while (!fully && step < opts.max_steps) {
        double nll_new = nll;
        bool b_updated = false;
        if (epochs_until_b == 0) {
            if (train_scale && zeta < DX_ETA_SCALE_FROZEN) {
                // safeguarded Newton in b with a fixed (oracle/nb_glm.py:_b_step_newton); the state is committed step by
                // step and put back if the scale step as a whole made the likelihood worse
                const double b0 = breg;
                double f = -nll;
#pragma unroll 1
                for (int it = 0; it < opts.newton_max_iter; ++it) {
                    double G, H;
                    scale_terms(G, H);
                    const double gb = xs0 * G, hb = -(xs0 * xs0) * H;          // hb = -Hessian wrt b (positive near a maximum)
                    if (!(isfinite(hb) && isfinite(gb))) break;
                    const double sb = (hb > 0.0) ? gb / hb : (double)((gb > 0.0) - (gb < 0.0));
                    const double smax = fabs(sb);
                    const double sc = (smax > DX_NEWTON_MAX_STEP) ? DX_NEWTON_MAX_STEP / smax : 1.0;
                    const double pred = gb * (sc * sb);
                    if (sc * smax <= opts.newton_xtol || !(pred > DX_NEWTON_GAIN_RTOL * fabs(f))) break;
                    double t = 1.0, f_try = f, bt = breg, zeta_t = zeta, r_t = r, invr_t = inv_r, T_t = T_r, l_t = l1p, irm_t = irm;
                    bool accepted = false;
#pragma unroll 1
                    for (int h = 0; h < DX_NEWTON_MAX_HALVINGS; ++h) {
                        bt = dx_clip(breg + t * (sc * sb), DX_ETA_MIN, DX_ETA_MAX);
                        zeta_t = dx_clip(xs0 * bt, DX_ETA_MIN, DX_ETA_MAX);
                        r_t = dx_exp(zeta_t);
                        invr_t = ff_rcp(r_t);
                        T_t = eval_T(r_t);
                        irm_t = ff_rcp(r_t + mu);
                        l_t = dx_log1p(mu * invr_t);
                        f_try = eval_ll(eta, zeta_t, r_t, l_t, T_t);
                        if (isfinite(f_try) && f_try >= f) { accepted = true; break; }
                        t *= 0.5;
                    }
                    if (!accepted) break;
                    breg = bt; zeta = zeta_t; r = r_t; inv_r = invr_t; T_r = T_t; l1p = l_t; irm = irm_t;
                    f = f_try;
                    if (zeta >= DX_ETA_SCALE_FROZEN) break;
                }
                const double nll_prop = -f;
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    nll_new = nll_prop;
                } else {
                    breg = b0;
                    zeta = dx_clip(xs0 * breg, DX_ETA_MIN, DX_ETA_MAX);
                    r = dx_exp(zeta); inv_r = ff_rcp(r);
                    T_r = eval_T(r);
                    irm = ff_rcp(r + mu);
                    l1p = dx_log1p(mu * inv_r);
                }
            }
            epochs_until_b = freq;
            b_updated = true;
        } else {
            if (train_loc && !loc_conv) {
                const double rhs = loc_terms();
                double dinv;
                const bool ok = ldl_factor(dinv);
                const bool fin = __all_sync(mask, (sl < p) ? isfinite(rhs) : true);
                double atry = areg;
                if (ok && fin) {
                    const double delta = ldl_solve(dinv, rhs);
                    atry = dx_clip(areg + delta, DX_ETA_MIN, DX_ETA_MAX);
                } else {
                    flags |= DX_FLAG_SINGULAR_FIM;
                }
                const double eta_t = eval_eta(atry);
                const double mu_t = dx_exp(eta_t);
                const double irm_t = ff_rcp(r + mu_t);
                const double l_t = dx_log1p(mu_t * inv_r);
                const double nll_prop = -eval_ll(eta_t, zeta, r, l_t, T_r);
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    areg = atry; eta = eta_t; mu = mu_t; irm = irm_t; l1p = l_t;
                    nll_new = nll_prop;
                }
            }
            epochs_until_b -= 1;
        }
        const bool conv_f = ((nll - nll_new) / nll) < opts.lltol;
        nll = nll_new;
        ++step;
        if (b_updated) {
            if (loc_conv && conv_f) fully = true;
            loc_conv = fully;
        } else {
            loc_conv = loc_conv || conv_f;
            if (!train_scale) fully = fully || loc_conv;
        }
    }
    if (fully) flags |= DX_FLAG_CONVERGED;

    // ---- finalize (batchglm Estimator.finalize) ----------------------------------------------------------------------
    if (zeta >= DX_ETA_SCALE_FROZEN) flags |= DX_FLAG_SCALE_FROZEN;
    const double rhs_f = loc_terms();
    double G_f, H_f;
    scale_terms(G_f, H_f);
    double jac = ((sl < p) ? fabs(rhs_f) : 0.0) + ((sl == 0) ? fabs(xs0 * G_f) : 0.0);
    jac = ff_sum<LANES>(jac, mask) / ff_sum<LANES>(nk, mask);
    double dinv_f;
    const bool okf = ldl_factor(dinv_f);
    if (sl < p) dv[sl] = dinv_f;
    __syncwarp(mask);
    // inverse: lane c solves (L D L^T) x = e_c into registers (every lane reads the same L element: broadcast loads)
    double x[PMAX];
    if (okf) {
#pragma unroll
        for (int i = 0; i < PMAX; ++i) {
            double acc = (i == sl) ? 1.0 : 0.0;
            if (i < p) {
#pragma unroll
                for (int k = 0; k < i; ++k) acc = fma(-Lsc[k * PMAX + i], x[k], acc);
            }
            x[i] = acc;
        }
#pragma unroll
        for (int i = PMAX - 1; i >= 0; --i) {
            if (i < p) {
                double acc = x[i] * dv[i];
#pragma unroll
                for (int k = i + 1; k < PMAX; ++k) if (k < p) acc = fma(-Lsc[i * PMAX + k], x[k], acc);
                x[i] = acc;
            }
        }
    }
    __syncwarp(mask);
    // the factor is no longer needed: its storage takes the inverse (column c from lane c) for the symmetrisation
    if (okf && sl < p) {
#pragma unroll
        for (int i = 0; i < PMAX; ++i) if (i < p) Lsc[sl * PMAX + i] = x[i];
    }
    __syncwarp(mask);
    double var_cc = NAN;
    if (sl < p) {
        double* fo = P.fim_inv + g * (int64_t)p * p;
#pragma unroll
        for (int i = 0; i < PMAX; ++i) {
            if (i < p) {
                const double v = okf ? 0.5 * (x[i] + Lsc[i * PMAX + sl]) : NAN;
                fo[i * p + sl] = v;
                if (i == sl) var_cc = v;
            }
        }
        P.a_var[(int64_t)sl * P.stride + g] = areg;
    }
    if (sl == 0) {
        P.b_var[g] = breg;
        P.ll_out[g] = -nll;
        P.jac_out[g] = jac;
        P.niter_out[g] = step;
        P.flags_out[g] = flags;
    }
    if (P.epi.wald_coef >= 0 && sl == P.epi.wald_coef) dx_fit_epilogue_row(P.epi, g, areg, var_cc);
    __syncwarp(mask);
    }
    dx_fit_epilogue_publish(P.epi);
}

template <int LANES, int PMAX>
cudaError_t launch_fast(FitArgs& A, int sm_count, cudaStream_t stream) {
    constexpr int SPW = 32 / LANES;
    const int ntgt = A.nent + A.p;
    const bool closed_form = A.opts.init_a == DX_INIT_CLOSED_FORM;
    // capacity of the weight -> sum product list: every non-zero of the dense pattern up to 24 KB, else the dense loops
    int nz_cap = A.K * ntgt;
    if (nz_cap > 1536) nz_cap = 1536;
    A.nz_max = nz_cap;
    const int t_li = (A.K * A.p + (closed_form ? A.p * A.KL : 0) + 1) & ~1;
    A.table_doubles = (t_li + 2 * A.nz_max + (ff_misc_bytes(ntgt, A.K) + 7) / 8 + 1) & ~1;
    A.slot_doubles = 2 * LANES + ((ntgt + 1) & ~1) + PMAX * PMAX + ((PMAX + 1) & ~1) + DX_HIST_BINS;
    const int warps = 4;
    const size_t smem = ((size_t)A.table_doubles + (size_t)warps * SPW * A.slot_doubles) * 8;
    auto kern = dx_fit_fast_kernel<LANES, PMAX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (sm_count <= 0) {
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    // few genes (a shard of a multi-GPU run): one gene slot per warp-half is wasteful when whole SMs are idle, so spread
    // the slots over all SMs first -- CTAs with fewer warps, as many CTAs as the genes need
    const int slots_per_cta = warps * SPW;
    int64_t grid = (A.n_genes + slots_per_cta - 1) / slots_per_cta;
    if (grid > (int64_t)per_sm * sm_count) grid = (int64_t)per_sm * sm_count;
    if (grid < 1) grid = 1;
    e = dx_queue_acquire(stream, &A.queue);
    if (e != cudaSuccess) return e;
    if (A.epi.world > 1) {
        e = dx_queue_acquire(stream, &A.epi.done_ctas);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid, warps * 32, smem, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace

bool dx_fit_fast_eligible(int K, int p, int ps, const dx_fit_opts* opts) {
    return K <= 32 && p <= 16 && ps == 1 && (opts->reserved & DX_OPT_SCALE_CONST) != 0;
}

cudaError_t dx_launch_fit_fast(FitArgs& A, int sm_count, cudaStream_t stream) {
    const int K = A.K, p = A.p;
    if (K <= 16) {
        if (p <= 4) return launch_fast<16, 4>(A, sm_count, stream);
        if (p <= 9) return launch_fast<16, 9>(A, sm_count, stream);
        return launch_fast<16, 16>(A, sm_count, stream);
    }
    if (p <= 4) return launch_fast<32, 4>(A, sm_count, stream);
    if (p <= 9) return launch_fast<32, 9>(A, sm_count, stream);
    return launch_fast<32, 16>(A, sm_count, stream);
}
