// dx_fit_lane.cu -- K_fit, lane-resident variant: the batchglm training schedule per gene
// (/root/reference/diffxpy/testing/tests.py:222-248 -> Estimator.initialize / train_sequence("DEFAULT") / finalize;
// restated in oracle/nb_glm.py, method_b="newton") for the designs the headline workloads use:
//   at most 32 design groups, at most 9 location coefficients, ONE dispersion coefficient shared by all groups
//   (constant dispersion model, DX_OPT_SCALE_CONST), genes whose counts all sit in the tail-count table (integers up to 64;
//   a gene with overflow entries is marked and left to the general kernel, dx_fit_args.cuh).
//
// LPG lanes per gene (1, 2, 4, 8 or 16, chosen by the launcher from the gene count).  With LPG = 1 a THREAD owns a gene:
// no shuffles, no lane sits idle while another one factorises, and nearly every issued instruction is fp64 arithmetic on
// useful data -- the sub-warp-per-gene kernel (dx_fit_grouped.cu) spends 84 % of its issue slots on shuffles,
// shared-memory round trips and index arithmetic.  What a lane keeps:
//   registers      coefficients a[p], the packed lower triangle of X^T W X (45 doubles at p = 9), X^T W ybar, scalars;
//                  the LDL^T factorisation, both substitutions and the inverse run on compile-time indices only
//   shared memory  per gene and group: mu, 1/(r + mu), log1p(mu / r), S_k (sum of the counts), the group's likelihood
//                  term; per gene: the pooled tail counts C(j) = #{y > j}, j < 64 -- laid out [index][gene slot], so a
//                  warp's accesses are conflict free; per CTA: the distinct design rows, offsets, group sizes,
//                  pseudo-inverse, log(j + 1)
// Few genes (a gene range of a multi-GPU run) take a larger LPG, which shortens the latency of the launch.  The lanes of a
// gene share only the expensive, order-independent part of the work: lane l evaluates exp / log1p / reciprocals of the
// design groups l, l + LPG, ... and log / reciprocals of the count bins l, l + LPG, ... and stores the values.  EVERY sum
// over groups or bins is then formed by every lane on its own, from those stored values, in ascending index order with
// the same instructions (this file is compiled without implicit multiply-add contraction).  So all lanes of a gene hold
// identical bits without a single shuffle, and -- the point -- the results do not depend on LPG: a gene gets bit-identical
// coefficients, likelihood, Fisher inverse and p-value whether it is fitted as one of 20 000 genes on one GPU or as one of
// 2 500 on each of eight.  A trial state overwrites the stored state in place; the rare rejected step re-evaluates the
// state it came from.
#include <cstdlib>
#include "dx_common.cuh"
#include "dx_internal.h"
#include "dx_fit_args.cuh"
#include "dx_fastmath.cuh"

namespace {

extern __shared__ __align__(16) double smem_l[];

template <int L> __device__ __forceinline__ int fl_max_i(int v, unsigned m) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }       // lower triangle, j <= i

struct LaneLayout {          // shared-memory layout of one launch (doubles unless noted)
    int sp;                  // padded number of gene slots per CTA (stride between consecutive indices of a gene array)
    int spc;                 // the same for the uint32 tail-count array
    int t_xl, t_off, t_isf, t_nk, t_pinv, t_lgrp, t_logj;      // CTA tables (offsets in doubles)
    int t_state;             // [5 K][sp]: mu | 1/(r+mu) | log1p(mu/r) | S | likelihood term
    int t_bval;              // [64][sp] per-bin logarithms / reciprocals handed from the evaluating lane to all lanes (LPG > 1)
    int t_cj;                // uint32 [64][spc]
    int total_bytes;
};

__host__ __device__ inline LaneLayout lane_layout(int K, int p, int KL, bool closed_form, int threads, int lpg) {
    LaneLayout L;
    const int slots = threads / lpg;
    const int want = (32 / lpg) % 16;
    L.sp = slots + ((want - slots % 16) + 16) % 16;
    const int wantc = (32 / lpg) % 32;
    L.spc = slots + ((wantc - slots % 32) + 32) % 32;
    int o = 0;
    L.t_xl = o; o += K * p;
    L.t_off = o; o += K;
    L.t_isf = o; o += K;
    L.t_nk = o; o += K;
    L.t_pinv = o; o += closed_form ? p * KL : 0;
    L.t_lgrp = o; o += (K + 7) / 8;
    L.t_logj = o; o += DX_HIST_BINS;
    o = (o + 1) & ~1;
    L.t_state = o; o += 5 * K * L.sp;
    L.t_bval = o; o += (lpg > 1) ? DX_HIST_BINS * L.sp : 0;
    L.t_cj = o; o += (DX_HIST_BINS * L.spc + 1) / 2;
    L.total_bytes = o * 8;
    return L;
}

template <int LPG, int PMAX>
__global__ void __launch_bounds__(256, 1)
dx_fit_lane_kernel(const FitArgs P) {
    constexpr int NE = PMAX * (PMAX + 1) / 2;
    constexpr int GPW = 32 / LPG;                   // genes per warp
    const int K = P.K, p = P.p, KL = P.KL;
    const int tid = threadIdx.x, lane = tid & 31;
    const int sl = lane % LPG, lane_base = lane - sl;
    const unsigned mask = (LPG == 32) ? 0xffffffffu : (((1u << LPG) - 1u) << lane_base);
    const bool closed_form = P.opts.init_a == DX_INIT_CLOSED_FORM;
    const LaneLayout L = lane_layout(K, p, KL, closed_form, (int)blockDim.x, LPG);
    const int SP = L.sp, SPC = L.spc;
    const int GPL = (K + LPG - 1) / LPG;            // groups per lane

    // ---- CTA tables ---------------------------------------------------------------------------------------------------
    double* xl = smem_l + L.t_xl;
    double* offt = smem_l + L.t_off;
    double* isft = smem_l + L.t_isf;
    double* nkt = smem_l + L.t_nk;
    const double* pinv = smem_l + L.t_pinv;
    unsigned char* lgrp = reinterpret_cast<unsigned char*>(smem_l + L.t_lgrp);
    for (int i = tid; i < K * p; i += blockDim.x) xl[i] = P.xu_loc[i];
    for (int k = tid; k < K; k += blockDim.x) {
        const double o = P.offset ? P.offset[k] : 0.0;
        offt[k] = o;
        isft[k] = P.offset ? dx_exp(-o) : 1.0;
        nkt[k] = P.n_k[k];
        lgrp[k] = (unsigned char)(P.loc_group ? P.loc_group[k] : k);
    }
    if (closed_form) for (int i = tid; i < p * KL; i += blockDim.x) smem_l[L.t_pinv + i] = P.pinv_loc[i];
    for (int j = tid; j < DX_HIST_BINS; j += blockDim.x) smem_l[L.t_logj + j] = dx_log((double)j + 1.0);
    __syncthreads();
    const double* logj = smem_l + L.t_logj;
    double n_total = 0.0;
    for (int k = 0; k < K; ++k) n_total += nkt[k];
    const double xs0 = P.xu_scale[0];
    const dx_fit_opts& opts = P.opts;

    // ---- the gene slot of this lane -------------------------------------------------------------------------------------
    const int slot = tid / LPG;
    double* st_mu = smem_l + L.t_state + slot;                  // [k * SP]
    double* st_irm = st_mu + (size_t)K * SP;
    double* st_l1p = st_irm + (size_t)K * SP;
    double* st_S = st_l1p + (size_t)K * SP;
    double* st_t = st_S + (size_t)K * SP;
    double* bval = smem_l + L.t_bval + slot;                    // [j * SP]  (LPG > 1 only)
    uint32_t* cj = reinterpret_cast<uint32_t*>(smem_l + L.t_cj) + slot;      // [j * SPC]

    for (;;) {
    unsigned batch = 0;
    if (lane == 0) batch = atomicAdd(P.queue, 1u);
    batch = __shfl_sync(DX_FULL_MASK, batch, 0);
    if ((int64_t)batch * GPW >= P.n_genes) break;
    const int64_t g = (int64_t)batch * GPW + lane / LPG;
    bool take = g < P.n_genes;
    if (take && P.ovf_val && P.ovf_count[g] > DX_LANE_MAX_OVF) {          // overflow entries: this gene goes to the general kernel (dx_fit_args.cuh)
        if (sl == 0) P.niter_out[g] = -1;
        take = false;
    }
    if (take) {

    // ---- sufficient statistics of the gene ---------------------------------------------------------------------------
    const uint32_t* __restrict__ tail = P.tail_all + g * (int64_t)K * DX_HIST_BINS;
    int ymax_l = 0;
    // tail counts c_k(j) = #{y > j} never increase with j: a row is read 16 bytes at a time up to its first zero; the
    // first pieces of four rows are requested together (the loads are what this phase waits for)
#pragma unroll 1
    for (int m0 = 0; m0 < GPL; m0 += 4) {
        uint4 first[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = sl + LPG * (m0 + u);
            first[u] = (k < K) ? __ldg(reinterpret_cast<const uint4*>(tail + k * DX_HIST_BINS)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = sl + LPG * (m0 + u);
            if (k < K) {
                const uint4* row = reinterpret_cast<const uint4*>(tail + k * DX_HIST_BINS);
                unsigned long long s_u = 0ull, q_u = 0ull;
                int top = 0;
                uint4 v = first[u];
#pragma unroll 1
                for (int c4 = 0; c4 < DX_HIST_BINS / 4; ++c4) {
                    if (c4) v = __ldg(row + c4);
                    if (v.x == 0u) break;
                    const unsigned j0 = 8u * c4 + 1u;           // 2 j + 1 at j = 4 c4
                    s_u += (unsigned long long)v.x + v.y + v.z + v.w;
                    q_u += (unsigned long long)v.x * j0 + (unsigned long long)v.y * (j0 + 2u) +
                           (unsigned long long)v.z * (j0 + 4u) + (unsigned long long)v.w * (j0 + 6u);
                    top = 4 * c4 + (v.w ? 4 : (v.z ? 3 : (v.y ? 2 : 1)));
                    if (v.w == 0u) break;
                }
                ymax_l = max(ymax_l, top);
                st_S[k * SP] = (double)s_u;        // sum_j c(j) = sum of the counts
                st_mu[k * SP] = (double)q_u;       // sum_j c(j) (2 j + 1) = sum of their squares (scratch until the state exists)
            }
        }
    }
    const int ymax = fl_max_i<LPG>(ymax_l, mask);
#pragma unroll 1
    for (int j = sl; j < ymax; j += LPG) {                // pooled tail counts: lane <-> count bin
        unsigned acc = 0u;
#pragma unroll 4
        for (int k = 0; k < K; ++k) acc += __ldg(tail + k * DX_HIST_BINS + j);
        cj[j * SPC] = acc;
    }
    __syncwarp(mask);
    // from here on every lane forms every sum itself, in ascending index order
    double lgc = 0.0;                                     // sum_i lgamma(y_i + 1) = sum_j C(j) log(j + 1)
#pragma unroll 1
    for (int j = 1; j < ymax; ++j) lgc = fma((double)cj[j * SPC], logj[j], lgc);
    double sum_all = 0.0, Sisf = 0.0, Qisf = 0.0;       // sum S_k, sum S_k / sf_k, sum Q_k / sf_k^2
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const double sk = st_S[k * SP], q = st_mu[k * SP], f = isft[k];
        sum_all += sk; Sisf = fma(sk, f, Sisf); Qisf = fma(q * f, f, Qisf);
    }
    int flags = 0;
    if (sum_all == 0.0) flags |= DX_FLAG_ALL_ZERO;
    __syncwarp(mask);

    // ---- initialize (batchglm init_par): every lane holds all coefficients --------------------------------------------
    double a[PMAX];
#pragma unroll
    for (int i = 0; i < PMAX; ++i) a[i] = 0.0;
    double breg = 0.0;
    if (opts.init_a == DX_INIT_STANDARD) {
        a[0] = dx_clip(dx_log(sum_all / n_total), DX_ETA_MIN, DX_ETA_MAX);
    } else if (closed_form) {
        // batchglm closedform_glm_mean: log of the mean of x / size_factor over every distinct LOCATION design row,
        // pushed through the pseudo-inverse of those rows (np.linalg.lstsq in the reference).  Scratch: the state arrays.
        if (LPG == 1) {
#pragma unroll 1
            for (int lg = 0; lg < KL; ++lg) { st_irm[lg * SP] = 0.0; st_l1p[lg * SP] = 0.0; }
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const int lg = lgrp[k];
                st_irm[lg * SP] = fma(st_S[k * SP], isft[k], st_irm[lg * SP]);
                st_l1p[lg * SP] += nkt[k];
            }
#pragma unroll 1
            for (int lg = 0; lg < KL; ++lg) st_l1p[lg * SP] = dx_log(st_irm[lg * SP] / st_l1p[lg * SP] + DX_TINY);
        } else {          // (the same sums in the same order, location rows spread over the lanes)
#pragma unroll 1
            for (int lg = sl; lg < KL; lg += LPG) {
                double num = 0.0, den = 0.0;
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    if (lgrp[k] == lg) { num = fma(st_S[k * SP], isft[k], num); den += nkt[k]; }
                }
                st_l1p[lg * SP] = dx_log(num / den + DX_TINY);
            }
            __syncwarp(mask);
        }
#pragma unroll 1
        for (int l = 0; l < KL; ++l) {
            const double lm = st_l1p[l * SP];
#pragma unroll
            for (int i = 0; i < PMAX; ++i) if (i < p) a[i] = fma(pinv[i * KL + l], lm, a[i]);
        }
#pragma unroll
        for (int i = 0; i < PMAX; ++i) a[i] = (i < p) ? dx_clip(a[i], DX_ETA_MIN, DX_ETA_MAX) : 0.0;
        __syncwarp(mask);
    } else if (opts.init_a == DX_INIT_GIVEN) {
#pragma unroll
        for (int i = 0; i < PMAX; ++i) if (i < p) a[i] = dx_clip(P.a_var[(int64_t)i * P.stride + g], DX_ETA_MIN, DX_ETA_MAX);
    }
    if (opts.init_b == DX_INIT_STANDARD) {
        const double m = Sisf / n_total;
        const double v = Qisf / n_total - m * m;
        const double denom = fmax(v - m, 2.2227587494850775e-162);
        breg = dx_clip(dx_log(m * m / denom + DX_TINY), DX_ETA_MIN, DX_ETA_MAX);
    } else if (opts.init_b == DX_INIT_GIVEN) {
        breg = dx_clip(P.b_var[g], DX_ETA_MIN, DX_ETA_MAX);
    }

    // ---- state evaluation ---------------------------------------------------------------------------------------------
    // W design groups (or count bins) are evaluated side by side: independent dependency chains for the scheduler
    constexpr int W = (LPG <= 4) ? 4 : (LPG == 8 ? 2 : 1);
    // T(r) = sum_j C(j) log(r + j) - sum_i lgamma(y_i + 1)
    auto eval_T = [&](double r_) -> double {
        double part = 0.0;
        if (LPG == 1) {
#pragma unroll 1
            for (int j0 = 0; j0 < ymax; j0 += W) {
                double c[W], l[W];
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int j = j0 + u;
                    c[u] = (j < ymax) ? (double)cj[j * SPC] : 0.0;
                    l[u] = fl_log(r_ + (double)j);
                }
#pragma unroll
                for (int u = 0; u < W; ++u) part = fma(c[u], l[u], part);
            }
        } else {
            __syncwarp(mask);
#pragma unroll 1
            for (int j = sl; j < ymax; j += LPG) bval[j * SP] = fl_log(r_ + (double)j);
            __syncwarp(mask);
#pragma unroll 4
            for (int j = 0; j < ymax; ++j) part = fma((double)cj[j * SPC], bval[j * SP], part);
        }
        return part - lgc;
    };
    // the groups of this lane at coefficients a_ and dispersion r_: eta, mu, 1/(r + mu), log1p(mu / r) and the group's term
    // of the log-likelihood, S_k (eta_k - zeta - l_k) - n_k r l_k, are stored; the sum of the terms is returned
    auto eval_state = [&](const double (&a_)[PMAX], double r_, double invr_, double zeta_) -> double {
        __syncwarp(mask);
#pragma unroll 1
        for (int m0 = 0; m0 < GPL; m0 += W) {
            double e[W], mu_[W], ir[W], l[W];
            int kk[W];
#pragma unroll
            for (int u = 0; u < W; ++u) {
                const int k = sl + LPG * (m0 + u);
                kk[u] = (k < K) ? k : -1;
                const int kc = (k < K) ? k : 0;
                const double* xr = xl + kc * p;
                double t = offt[kc];
#pragma unroll
                for (int i = 0; i < PMAX; ++i) if (i < p) t = fma(xr[i], a_[i], t);
                e[u] = dx_clip(t, DX_ETA_MIN, DX_ETA_MAX);
            }
#pragma unroll
            for (int u = 0; u < W; ++u) mu_[u] = fl_exp(e[u]);
#pragma unroll
            for (int u = 0; u < W; ++u) ir[u] = fl_rcp_n(r_ + mu_[u]);
#pragma unroll
            for (int u = 0; u < W; ++u) l[u] = fl_log1p(mu_[u] * invr_);
#pragma unroll
            for (int u = 0; u < W; ++u) {
                if (kk[u] >= 0) {
                    const int k = kk[u];
                    st_mu[k * SP] = mu_[u]; st_irm[k * SP] = ir[u]; st_l1p[k * SP] = l[u];
                    st_t[k * SP] = fma(st_S[k * SP], (e[u] - zeta_) - l[u], -((nkt[k] * r_) * l[u]));
                }
            }
        }
        __syncwarp(mask);
        double part = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) part += st_t[k * SP];
        return part;
    };
    // the same at a new dispersion with mu kept (scale step): 1/(r + mu), log1p(mu / r) are overwritten.  eta_k enters the
    // likelihood only through sum_k S_k eta_k, which a scale step does not change: the r-dependent part is returned,
    // - sum_k [(S_k + n_k r) l_k + S_k zeta]
    auto eval_state_r = [&](double r_, double invr_, double zeta_) -> double {
        __syncwarp(mask);
#pragma unroll 1
        for (int m0 = 0; m0 < GPL; m0 += W) {
            double mu_[W], ir[W], l[W];
            int kk[W];
#pragma unroll
            for (int u = 0; u < W; ++u) {
                const int k = sl + LPG * (m0 + u);
                kk[u] = (k < K) ? k : -1;
                mu_[u] = st_mu[((k < K) ? k : 0) * SP];
            }
#pragma unroll
            for (int u = 0; u < W; ++u) ir[u] = fl_rcp_n(r_ + mu_[u]);
#pragma unroll
            for (int u = 0; u < W; ++u) l[u] = fl_log1p(mu_[u] * invr_);
#pragma unroll
            for (int u = 0; u < W; ++u) {
                if (kk[u] >= 0) {
                    const int k = kk[u];
                    const double sk = st_S[k * SP];
                    st_irm[k * SP] = ir[u]; st_l1p[k * SP] = l[u];
                    st_t[k * SP] = fma(fma(nkt[k], r_, sk), l[u], sk * zeta_);
                }
            }
        }
        __syncwarp(mask);
        double part = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) part -= st_t[k * SP];
        return part;
    };
    // sum_k S_k eta_k from the group part of the likelihood at the stored state
    auto sum_S_eta = [&](double grp_, double r_, double zeta_) -> double {
        double part = 0.0;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const double sk = st_S[k * SP];
            part += fma(fma(nkt[k], r_, sk), st_l1p[k * SP], sk * zeta_);
        }
        return grp_ + part;
    };

    double zeta = dx_clip(xs0 * breg, DX_ETA_MIN, DX_ETA_MAX);
    double r = fl_exp(zeta), inv_r = fl_rcp_n(r);
    double T_r = eval_T(r);
    double grp = eval_state(a, r, inv_r, zeta);
    double nll = -(grp + T_r);

    // gradient G and curvature H of ll wrt the linear predictor of the dispersion at the current state:
    // gradient wrt b = xs0 G, Hessian wrt b = xs0^2 H
    auto scale_terms = [&](double& G, double& H) {
        double D = 0.0, E = 0.0;
        if (LPG == 1) {
#pragma unroll 1
            for (int j0 = 0; j0 < ymax; j0 += W) {
                double ci[W], inv[W];
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int j = j0 + u;
                    ci[u] = (j < ymax) ? (double)cj[j * SPC] : 0.0;
                    inv[u] = fl_rcp_n(r + (double)j);
                }
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const double t = ci[u] * inv[u];
                    D += t;
                    E = fma(-t, inv[u], E);
                }
            }
        } else {
            __syncwarp(mask);
#pragma unroll 1
            for (int j = sl; j < ymax; j += LPG) bval[j * SP] = fl_rcp_n(r + (double)j);
            __syncwarp(mask);
#pragma unroll 4
            for (int j = 0; j < ymax; ++j) {
                const double inv = bval[j * SP];
                const double t = (double)cj[j * SPC] * inv;
                D += t;
                E = fma(-t, inv, E);
            }
        }
        double Gs = D, Hs = 0.0;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const double mu = st_mu[k * SP], irm = st_irm[k * SP], l1p = st_l1p[k * SP], S = st_S[k * SP], nk = nkt[k];
            Gs += fma(fma(nk, mu, -S), irm, -(nk * l1p));
            Hs = fma(fma(nk * mu, mu, r * S), irm * irm, Hs);
        }
        G = r * Gs;
        H = fma(r, Hs, fma(r * r, E, G));
    };

    // X^T W X (packed lower triangle) and X^T W ybar at the current state: every lane runs over all design groups.  The
    // design rows are the same in every thread, so zero entries are skipped by a uniform branch (categorical designs: 2-3
    // of 9 entries are non-zero)
    double A[NE], rhs[PMAX];
    auto loc_terms = [&]() {
#pragma unroll
        for (int e = 0; e < NE; ++e) A[e] = 0.0;
#pragma unroll
        for (int i = 0; i < PMAX; ++i) rhs[i] = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const double q = r * st_irm[k * SP];
            const double w1 = (nkt[k] * st_mu[k * SP]) * q;
            const double w2 = fma(st_S[k * SP], q, -w1);
            const double* xr = xl + k * p;
            double x[PMAX];
#pragma unroll
            for (int i = 0; i < PMAX; ++i) x[i] = (i < p) ? xr[i] : 0.0;
#pragma unroll
            for (int i = 0; i < PMAX; ++i) {
                if (x[i] != 0.0) {
                    rhs[i] = fma(w2, x[i], rhs[i]);
                    const double t = w1 * x[i];
#pragma unroll
                    for (int j = 0; j <= i; ++j) A[tri(i, j)] = fma(t, x[j], A[tri(i, j)]);
                }
            }
        }
    };

    // LDL^T in place: on exit A[tri(i, j)] = L_ij (j < i), dd[j] = 1 / d_j.  Pivot guard of oracle/nb_glm.py:chol_guarded:
    // d_j finite and above 2^-50 of the input diagonal.
    double dd[PMAX];
    auto ldl_factor = [&]() -> bool {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < PMAX; ++j) dd[j] = A[tri(j, j)];
#pragma unroll
        for (int j = 0; j < PMAX; ++j) {
            if (j < p) {
                const double d = A[tri(j, j)], m_jj = dd[j];
                const bool pivot_ok = isfinite(d) && d > DX_CHOL_PIVOT_RTOL * m_jj && m_jj > 0.0;
                ok = ok && pivot_ok;
                const double inv = pivot_ok ? fl_rcp(d) : 0.0;
                dd[j] = inv;
                double lc[PMAX];
#pragma unroll
                for (int c = j + 1; c < PMAX; ++c) lc[c] = A[tri(c, j)] * inv;
#pragma unroll
                for (int i = j + 1; i < PMAX; ++i) {
                    if (i < p) {
                        const double u = A[tri(i, j)];
#pragma unroll
                        for (int c = j + 1; c <= i; ++c) A[tri(i, c)] = fma(-u, lc[c], A[tri(i, c)]);
                    }
                }
#pragma unroll
                for (int c = j + 1; c < PMAX; ++c) A[tri(c, j)] = lc[c];
            }
        }
        return ok;
    };
    // solve (L D L^T) x = rhs in place
    auto ldl_solve = [&](double (&x)[PMAX]) {
#pragma unroll
        for (int i = 1; i < PMAX; ++i) {
            if (i < p) {
#pragma unroll
                for (int j = 0; j < i; ++j) x[i] = fma(-A[tri(i, j)], x[j], x[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < PMAX; ++i) x[i] *= dd[i];
#pragma unroll
        for (int i = PMAX - 2; i >= 0; --i) {
            if (i < p - 1) {
#pragma unroll
                for (int j = i + 1; j < PMAX; ++j) if (j < p) x[i] = fma(-A[tri(j, i)], x[j], x[i]);
            }
        }
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
                // safeguarded Newton in b with a fixed (oracle/nb_glm.py:_b_step_newton).  Trial dispersions overwrite the
                // r-dependent state in place; it is re-evaluated if the line search, or the step as a whole, is given up
                const double b0 = breg, zeta0 = zeta, r0 = r, invr0 = inv_r, T0 = T_r;
                const double s_eta = sum_S_eta(grp, r, zeta);
                double f = -nll;
                bool dirty = false;          // the stored 1/(r+mu), log1p(mu/r) belong to a rejected trial
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
                    double t = 1.0, f_try = f, bt = breg, zeta_t = zeta, r_t = r, invr_t = inv_r, T_t = T_r;
                    bool accepted = false;
#pragma unroll 1
                    for (int h = 0; h < DX_NEWTON_MAX_HALVINGS; ++h) {
                        bt = dx_clip(breg + t * (sc * sb), DX_ETA_MIN, DX_ETA_MAX);
                        zeta_t = dx_clip(xs0 * bt, DX_ETA_MIN, DX_ETA_MAX);
                        r_t = fl_exp(zeta_t);
                        invr_t = fl_rcp_n(r_t);
                        T_t = eval_T(r_t);
                        f_try = s_eta + eval_state_r(r_t, invr_t, zeta_t) + T_t;
                        if (isfinite(f_try) && f_try >= f) { accepted = true; break; }
                        t *= 0.5;
                    }
                    if (!accepted) { dirty = true; break; }
                    breg = bt; zeta = zeta_t; r = r_t; inv_r = invr_t; T_r = T_t;
                    f = f_try;
                    if (zeta >= DX_ETA_SCALE_FROZEN) break;
                }
                const double nll_prop = -f;
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    nll_new = nll_prop;
                } else {
                    dirty = dirty || (breg != b0);
                    breg = b0; zeta = zeta0; r = r0; inv_r = invr0; T_r = T0;
                }
                if (dirty) (void)eval_state_r(r, inv_r, zeta);
                grp = -nll_new - T_r;          // group part of the likelihood at the committed state
            }
            epochs_until_b = freq;
            b_updated = true;
        } else {
            if (train_loc && !loc_conv) {
                loc_terms();
                const bool ok = ldl_factor();
                bool fin = true;
#pragma unroll
                for (int i = 0; i < PMAX; ++i) fin = fin && ((i < p) ? isfinite(rhs[i]) : true);
                double atry[PMAX];
                if (ok && fin) {
                    ldl_solve(rhs);
#pragma unroll
                    for (int i = 0; i < PMAX; ++i) atry[i] = (i < p) ? dx_clip(a[i] + rhs[i], DX_ETA_MIN, DX_ETA_MAX) : 0.0;
                } else {
                    flags |= DX_FLAG_SINGULAR_FIM;
#pragma unroll
                    for (int i = 0; i < PMAX; ++i) atry[i] = a[i];
                }
                // pass 0: the trial coefficients; pass 1 (only after a rejection): the kept coefficients again, which
                // puts their state back -- one copy of the evaluation code serves both
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const double grp_t = eval_state(atry, r, inv_r, zeta);
                    if (pass == 1) break;
                    const double nll_prop = -(grp_t + T_r);
                    if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                    if (isfinite(nll_prop) && !(nll_prop > nll)) {
#pragma unroll
                        for (int i = 0; i < PMAX; ++i) a[i] = atry[i];
                        grp = grp_t;
                        nll_new = nll_prop;
                        break;
                    }
#pragma unroll
                    for (int i = 0; i < PMAX; ++i) atry[i] = a[i];
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
    loc_terms();
    double G_f, H_f;
    scale_terms(G_f, H_f);
    double jac = fabs(xs0 * G_f);
#pragma unroll
    for (int i = 0; i < PMAX; ++i) jac += (i < p) ? fabs(rhs[i]) : 0.0;
    jac /= n_total;
    const bool okf = ldl_factor();
    // inverse in place: L <- L^-1 (unit lower triangular), then A^-1 = L^-T D^-1 L^-1 (lower triangle; mirrored on store)
#pragma unroll
    for (int j = 0; j < PMAX; ++j) {
#pragma unroll
        for (int i = j + 1; i < PMAX; ++i) {
            if (i < p) {
                double s = A[tri(i, j)];
#pragma unroll
                for (int k = j + 1; k < i; ++k) s = fma(A[tri(i, k)], A[tri(k, j)], s);
                A[tri(i, j)] = -s;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PMAX; ++i) {
        if (i < p) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double s = (j == i) ? dd[i] : dd[i] * A[tri(i, j)];
#pragma unroll
                for (int k = i + 1; k < PMAX; ++k) if (k < p) s = fma(A[tri(k, i)] * dd[k], A[tri(k, j)], s);
                A[tri(i, j)] = s;
            }
        }
    }
    double* fo = P.fim_inv + g * (int64_t)p * p;
    double var_cc = NAN, theta_c = NAN;
#pragma unroll
    for (int i = 0; i < PMAX; ++i) {
        if (i < p) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double v = okf ? A[tri(i, j)] : NAN;
                if ((tri(i, j) % LPG) == sl) { fo[i * p + j] = v; fo[j * p + i] = v; }
            }
            if ((i % LPG) == sl) P.a_var[(int64_t)i * P.stride + g] = a[i];
            if (i == P.epi.wald_coef) { var_cc = okf ? A[tri(i, i)] : NAN; theta_c = a[i]; }
        }
    }
    if (sl == 0) {
        P.b_var[g] = breg;
        P.ll_out[g] = -nll;
        P.jac_out[g] = jac;
        P.niter_out[g] = step;
        P.flags_out[g] = flags;
        if (P.epi.wald_coef >= 0) dx_fit_epilogue_row(P.epi, g, theta_c, var_cc);
    }
    }   // gene of this lane
    __syncwarp();
    }   // batches
    dx_fit_epilogue_publish(P.epi);
}

template <int LPG, int PMAX>
cudaError_t launch_lane(FitArgs& A, int threads, int grid, size_t smem, cudaStream_t stream) {
    auto kern = dx_fit_lane_kernel<LPG, PMAX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = dx_queue_acquire(stream, &A.queue);
    if (e != cudaSuccess) return e;
    if (A.epi.world > 1 && A.epi.publish) {
        e = dx_queue_acquire(stream, &A.epi.done_ctas);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid, threads, smem, stream>>>(A);
    return cudaGetLastError();
}

template <int PMAX>
cudaError_t launch_lane_p(FitArgs& A, int lpg, int threads, int grid, size_t smem, cudaStream_t stream) {
    switch (lpg) {
        case 1: return launch_lane<1, PMAX>(A, threads, grid, smem, stream);
        case 2: return launch_lane<2, PMAX>(A, threads, grid, smem, stream);
        case 4: return launch_lane<4, PMAX>(A, threads, grid, smem, stream);
        case 8: return launch_lane<8, PMAX>(A, threads, grid, smem, stream);
        default: return launch_lane<16, PMAX>(A, threads, grid, smem, stream);
    }
}

}  // namespace

bool dx_fit_lane_eligible(int K, int p, int ps, const dx_fit_opts* opts) {
    return K <= 32 && p <= 9 && ps == 1 && (opts->reserved & DX_OPT_SCALE_CONST) != 0;
}

cudaError_t dx_launch_fit_lane(FitArgs& A, int sm_count, cudaStream_t stream) {
    if (sm_count <= 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const bool closed_form = A.opts.init_a == DX_INIT_CLOSED_FORM;
    const int smem_cap = 220 * 1024;
    const char* lpg_env = getenv("DXB200_FIT_LPG");        // tuning / test hook: lanes per gene (read at every launch)
    const char* thr_env = getenv("DXB200_FIT_THREADS");    // tuning / test hook: threads per CTA
    // lanes per gene (any choice gives the same bits): a thread per gene is the cheapest in issue slots and wins as soon as
    // the genes fill the SMs (measured on B200, 16 groups x 9 coefficients: 20k genes 0.26 ms at 1 lane, 0.30 at 2, 0.39 at 4;
    // 10k genes 0.23 / 0.22 / 0.25; 5k genes 0.22 at 1, 0.175 at 4, 0.195 at 8; 2.5k genes 0.21 at 1, 0.14 at 4 and 8, 0.17 at 16)
    int lpg = 1;
    if (A.n_genes * 4 <= (int64_t)sm_count * 256) lpg = 4;
    else if (A.n_genes * 2 <= (int64_t)sm_count * 256) lpg = 2;
    while (lpg > 1 && lpg > A.K) lpg >>= 1;
    if (lpg_env) { const int v = atoi(lpg_env); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) lpg = v; }
    // threads per CTA: what the genes need when spread over all SMs, rounded up to whole warps, within the shared memory
    int64_t lanes = A.n_genes * lpg;
    int threads = (int)(((lanes + sm_count - 1) / sm_count + 31) / 32 * 32);
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    if (thr_env) { const int v = atoi(thr_env); if (v >= 32 && v <= 256 && v % 32 == 0) threads = v; }
    while (threads > 32 && lane_layout(A.K, A.p, A.KL, closed_form, threads, lpg).total_bytes > smem_cap) threads -= 32;
    const LaneLayout L = lane_layout(A.K, A.p, A.KL, closed_form, threads, lpg);
    if (L.total_bytes > smem_cap) return cudaErrorInvalidConfiguration;
    int64_t grid = (lanes + threads - 1) / threads;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    if (A.p <= 4) return launch_lane_p<4>(A, lpg, threads, (int)grid, (size_t)L.total_bytes, stream);
    return launch_lane_p<9>(A, lpg, threads, (int)grid, (size_t)L.total_bytes, stream);
}
