// dx_fit_grouped.cu -- K_fit (grouped design): the whole batchglm training schedule for one gene runs
// inside ONE WARP, from initialisation to finalize, on the per-(gene, group) sufficient statistics that
// K_ingest produced.  No pass over the count matrix, no host loop, no global synchronisation: genes are
// independent, so the reference's lock-step loop over all genes
//   (/root/reference/diffxpy/testing/tests.py:222-248 -> batchglm numpy Estimator.initialize /
//    train_sequence("DEFAULT") / finalize; restated in oracle/nb_glm.py, method_b="newton")
// is exactly equivalent to one independent loop per gene.
//
// For design group k (distinct row of [design_loc | design_scale | log size factor], n_k cells):
//   eta_k = x_k.a + off_k, mu_k = exp(eta_k), zeta_k = z_k.b, r_k = exp(zeta_k)
//   ll   = sum_k [ sum_j c_k(j) log1p((j - mu_k)/(r_k + mu_k)) + S_k eta_k - n_k r_k log1p(mu_k / r_k) ]
//          + sum_ovf [ G(r_k, y) - y log1p(mu_k / r_k) ] - sum_i lgamma(y_i + 1)
//   X^T W X   = sum_k n_k W_k x_k x_k^T,   W_k = mu_k r_k / (r_k + mu_k)
//   X^T W ybar = sum_k x_k ( S_k r_k/(r_k + mu_k) - n_k W_k )
// with c_k(j) the tail counts #{i in k: y_i > j} (integer counts <= 64) and S_k = sum_{i in k} y_i.
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int WARPS_PER_CTA = 1;     // one gene per CTA: the block scheduler balances the uneven per-gene cost

struct Ctx {
    int K, p, ps, lane, ymax, novf;
    bool shared_r;                         // every group has the same scale design row (constant dispersion)
    unsigned long long ovf_groups_lo[4];   // bitmask of groups that own overflow entries (K <= 255)
    const double* __restrict__ xl;   // [K][p]
    const double* __restrict__ xs;   // [K][ps]
    const double* __restrict__ off;  // [K] or null
    const double* __restrict__ nk;   // [K]
    const uint32_t* __restrict__ tail;  // [K][64]
    const float* __restrict__ oval;
    const uint8_t* __restrict__ ogrp;
    double *mu, *eta, *rr, *zeta, *S, *w1, *w2; // [K]; w1, w2 are scratch
    double *A, *L, *rhs, *a, *at;              // p*p, p*p, p, p, p
    double *b, *bt, *gb, *sb, *Hb, *Lb;        // ps, ps, ps, ps, ps*ps, ps*ps
    double *Cj;                                // [64] tail counts summed over groups (shared_r only)
    double lgam_const, n_total;
    double T_r;                                // part of the likelihood that depends on r only (cached)
};

__device__ __forceinline__ bool ctx_group_has_ovf(const Ctx& c, int k) {
    return (c.ovf_groups_lo[k >> 6] >> (k & 63)) & 1ull;
}

// T(r) = sum_k sum_j c_k(j) log(r_k + j) + sum_ovf [lgamma(r_k + y) - lgamma(r_k)]
__device__ __noinline__ double compute_T(Ctx& c) {
    double part = 0.0;
    if (c.shared_r) {
        const double r = c.rr[0];
#pragma unroll 1
        for (int j = c.lane; j < c.ymax; j += 32) {
            const double cnt = c.Cj[j];
            if (cnt != 0.0) part += cnt * log(r + (double)j);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) {
            const double r = c.rr[k];
#pragma unroll 1
            for (int j = c.lane; j < c.ymax; j += 32) {
                const unsigned cnt = c.tail[k * DX_HIST_BINS + j];
                if (cnt) part += (double)cnt * log(r + (double)j);
            }
        }
    }
#pragma unroll 1
    for (int t = c.lane; t < c.novf; t += 32) {
        const double y = (double)c.oval[t];
        const double r = c.rr[c.ogrp[t]];
        part += dx_lgamma_ratio(r, y) + y * log(r);
    }
    return dx_warp_sum(part);
}

// r_k = exp(clip(z_k . b)) and the cached T(r); returns max_k zeta_k
__device__ __noinline__ double set_b(Ctx& c, const double* bvec) {
    double zmax = -INFINITY;
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        double z = 0.0;
#pragma unroll 1
        for (int i = 0; i < c.ps; ++i) z += c.xs[k * c.ps + i] * bvec[i];
        z = dx_clip(z, DX_ETA_MIN, DX_ETA_MAX);
        c.zeta[k] = z;
        c.rr[k] = exp(z);
        zmax = fmax(zmax, z);
    }
    __syncwarp();
    c.T_r = compute_T(c);
    return dx_warp_max(zmax);
}

// mu_k = exp(clip(x_k . a + off_k))
__device__ __noinline__ void compute_mu(Ctx& c, const double* avec) {
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        double e = c.off ? c.off[k] : 0.0;
#pragma unroll 1
        for (int i = 0; i < c.p; ++i) e += c.xl[k * c.p + i] * avec[i];
        e = dx_clip(e, DX_ETA_MIN, DX_ETA_MAX);
        c.eta[k] = e;
        c.mu[k] = exp(e);
    }
    __syncwarp();
}

// log-likelihood from the current mu / eta / rr / T_r:
//   ll = T(r) - sum lgamma(y+1) + sum_k [ S_k (eta_k - log(r_k + mu_k)) - n_k r_k log1p(mu_k / r_k) ]
__device__ __noinline__ double ll_state(Ctx& c) {
    double part = 0.0;
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        const double m = c.mu[k], r = c.rr[k];
        part += c.S[k] * (c.eta[k] - log(r + m)) - c.nk[k] * r * log1p(m / r);
    }
    return dx_warp_sum(part) + (c.T_r - c.lgam_const);
}

__device__ double ll_at(Ctx& c, const double* avec) { compute_mu(c, avec); return ll_state(c); }

// A = X^T W X (full symmetric p x p in c.A), rhs = X^T W ybar; needs current mu, rr
__device__ __noinline__ void loc_terms(Ctx& c) {
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        const double m = c.mu[k], r = c.rr[k];
        const double q = r / (r + m);
        const double nw = c.nk[k] * m * q;
        c.w1[k] = nw;
        c.w2[k] = c.S[k] * q - nw;
    }
    __syncwarp();
    const int p = c.p, nent = p * (p + 1) / 2;
#pragma unroll 1
    for (int e = c.lane; e < nent; e += 32) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
#pragma unroll 1
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
#pragma unroll 1
        while (i * (i + 1) / 2 > e) --i;
        const int j = e - i * (i + 1) / 2;
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) acc += c.w1[k] * c.xl[k * p + i] * c.xl[k * p + j];
        c.A[i * p + j] = acc;
        c.A[j * p + i] = acc;
    }
#pragma unroll 1
    for (int i = c.lane; i < p; i += 32) {
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) acc += c.w2[k] * c.xl[k * p + i];
        c.rhs[i] = acc;
    }
    __syncwarp();
}

// gradient (c.gb) and NEGATIVE Hessian (c.Hb) of ll wrt b at the current mu / rr / zeta
__device__ __noinline__ void scale_terms(Ctx& c) {
    const int ps = c.ps;
    if (c.shared_r) {
        // constant dispersion: sum the digamma / trigamma differences over all groups at once (lanes over j)
        const double r = c.rr[0];
        double D = 0.0, E = 0.0;
#pragma unroll 1
        for (int j = c.lane; j < c.ymax; j += 32) {
            const double cnt = c.Cj[j];
            if (cnt != 0.0) {
                const double inv = 1.0 / (r + (double)j);
                D += cnt * inv;
                E -= cnt * inv * inv;
            }
        }
        if (c.novf > 0) {
            const double psi_r = dx_digamma(r), psi1_r = dx_trigamma(r);
#pragma unroll 1
            for (int t = c.lane; t < c.novf; t += 32) {
                const double y = (double)c.oval[t];
                D += dx_digamma(r + y) - psi_r;
                E += dx_trigamma(r + y) - psi1_r;
            }
        }
        double G = 0.0, H = 0.0;
#pragma unroll 1
        for (int k = c.lane; k < c.K; k += 32) {
            const double m = c.mu[k], n = c.nk[k], S = c.S[k];
            const double rm = r + m;
            G += -n * log1p(m / r) + (n * m - S) / rm;
            H += (n * m * m + r * S) / (rm * rm);
        }
        G = r * (dx_warp_sum(G + D));
        H = G + r * r * dx_warp_sum(E) + r * dx_warp_sum(H);
        if (c.lane == 0) {
#pragma unroll 1
            for (int i = 0; i < ps; ++i) {
                c.gb[i] = c.xs[i] * G;
#pragma unroll 1
                for (int j = 0; j < ps; ++j) c.Hb[i * ps + j] = -c.xs[i] * c.xs[j] * H;
            }
        }
        __syncwarp();
        return;
    }
    // general scale model: lane <-> group, serial over j (results stay with the lane)
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        const double r = c.rr[k];
        double D = 0.0, E = 0.0;
#pragma unroll 1
        for (int j = 0; j < c.ymax; ++j) {
            const unsigned cnt = c.tail[k * DX_HIST_BINS + j];
            if (cnt) {
                const double inv = 1.0 / (r + (double)j);
                D += (double)cnt * inv;
                E -= (double)cnt * inv * inv;
            }
        }
        c.w1[k] = D;
        c.w2[k] = E;
    }
    __syncwarp();
    if (c.novf > 0) {
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) {
            if (!ctx_group_has_ovf(c, k)) continue;
            const double r = c.rr[k];
            const double psi_r = dx_digamma(r), psi1_r = dx_trigamma(r);
            double D = 0.0, E = 0.0;
#pragma unroll 1
            for (int t = c.lane; t < c.novf; t += 32) {
                if (c.ogrp[t] == k) {
                    const double y = (double)c.oval[t];
                    D += dx_digamma(r + y) - psi_r;
                    E += dx_trigamma(r + y) - psi1_r;
                }
            }
            D = dx_warp_sum(D);
            E = dx_warp_sum(E);
            if (c.lane == 0) { c.w1[k] += D; c.w2[k] += E; }
        }
        __syncwarp();
    }
#pragma unroll 1
    for (int k = c.lane; k < c.K; k += 32) {
        const double r = c.rr[k], m = c.mu[k], n = c.nk[k], S = c.S[k];
        const double rm = r + m;
        const double g = r * (c.w1[k] - n * log1p(m / r) + (n * m - S) / rm);
        const double h = g + r * r * c.w2[k] + r * (n * m * m + r * S) / (rm * rm);
        c.w1[k] = g;
        c.w2[k] = h;
    }
    __syncwarp();
#pragma unroll 1
    for (int e = c.lane; e < ps * ps; e += 32) {
        const int i = e / ps, j = e - i * ps;
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) acc += c.w2[k] * c.xs[k * ps + i] * c.xs[k * ps + j];
        c.Hb[e] = -acc;
    }
#pragma unroll 1
    for (int i = c.lane; i < ps; i += 32) {
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < c.K; ++k) acc += c.w1[k] * c.xs[k * ps + i];
        c.gb[i] = acc;
    }
    __syncwarp();
}

// Newton direction s (c.sb) from c.gb and c.Hb (= -Hessian): solve (Hb + lam I) s = gb, lam grows until the
// guarded Cholesky succeeds (oracle/nb_glm.py:_newton_direction); returns false on non-finite input
__device__ __noinline__ bool newton_direction(Ctx& c) {
    const int ps = c.ps;
    if (ps == 1) {
        const double h = c.Hb[0], g = c.gb[0];
        if (!(isfinite(h) && isfinite(g))) return false;
        if (c.lane == 0) c.sb[0] = (h > 0.0) ? g / h : (double)((g > 0.0) - (g < 0.0));
        __syncwarp();
        return true;
    }
    bool fin = true;
#pragma unroll 1
    for (int e = c.lane; e < ps * ps; e += 32) fin = fin && isfinite(c.Hb[e]);
#pragma unroll 1
    for (int i = c.lane; i < ps; i += 32) fin = fin && isfinite(c.gb[i]);
    if (!__all_sync(DX_FULL_MASK, fin)) return false;
    const double diag = (c.lane < ps) ? c.Hb[c.lane * ps + c.lane] : 0.0;   // ps <= 32: one diagonal entry per lane
    const double base = fmax(dx_warp_max(fabs(diag)), 1e-300);
    double lam = 0.0;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        if (c.lane < ps) c.Hb[c.lane * ps + c.lane] = diag + lam;
        __syncwarp();
        const bool ok = chol_warp(c.Hb, c.Lb, ps, c.lane);
        __syncwarp();
        if (ok) {
            if (c.lane == 0) chol_solve_serial(c.Lb, ps, c.gb, c.sb);
            __syncwarp();
            return true;
        }
        lam = (lam == 0.0) ? base * 1e-6 : lam * 10.0;
    }
#pragma unroll 1
    for (int i = c.lane; i < ps; i += 32) { const double g = c.gb[i]; c.sb[i] = (double)((g > 0.0) - (g < 0.0)); }
    __syncwarp();
    return true;
}

// safeguarded Newton in b with a fixed (oracle/nb_glm.py:_b_step_newton); c.b is updated in place; returns ll at
// the final b.  Requires c.mu / c.eta to hold mu(a) and f_cur = ll(a, c.b).
__device__ __noinline__ double scale_newton(Ctx& c, double f_cur, int max_iter, double xtol) {
    double zmax = set_b(c, c.b);
    if (zmax >= DX_ETA_SCALE_FROZEN) return f_cur;
    double f = f_cur;
#pragma unroll 1
    for (int it = 0; it < max_iter; ++it) {
        // (rr, zeta, T_r correspond to c.b here)
        scale_terms(c);
        if (!newton_direction(c)) break;
        double smax = 0.0;
#pragma unroll 1
        for (int i = c.lane; i < c.ps; i += 32) smax = fmax(smax, fabs(c.sb[i]));
        smax = dx_warp_max(smax);
        const double sc = (smax > DX_NEWTON_MAX_STEP) ? DX_NEWTON_MAX_STEP / smax : 1.0;
        double pred = 0.0;
#pragma unroll 1
        for (int i = c.lane; i < c.ps; i += 32) pred += c.gb[i] * (sc * c.sb[i]);
        pred = dx_warp_sum(pred);
        // converged: the step is below xtol or its predicted gain is below the resolution of the likelihood
        if (sc * smax <= xtol || !(pred > DX_NEWTON_GAIN_RTOL * fabs(f))) break;
        double t = 1.0, f_try = f;
        bool accepted = false;
#pragma unroll 1
        for (int h = 0; h < DX_NEWTON_MAX_HALVINGS; ++h) {
#pragma unroll 1
            for (int i = c.lane; i < c.ps; i += 32)
                c.bt[i] = dx_clip(c.b[i] + t * (sc * c.sb[i]), DX_ETA_MIN, DX_ETA_MAX);
            __syncwarp();
            zmax = set_b(c, c.bt);
            f_try = ll_state(c);
            if (isfinite(f_try) && f_try >= f) { accepted = true; break; }
            t *= 0.5;
        }
        if (!accepted) break;
#pragma unroll 1
        for (int i = c.lane; i < c.ps; i += 32) c.b[i] = c.bt[i];
        __syncwarp();
        f = f_try;
        if (zmax >= DX_ETA_SCALE_FROZEN) break;
    }
    set_b(c, c.b);
    return f;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 16)
dx_fit_grouped_kernel(int64_t n_genes, int K, int p, int ps,
                      const double* __restrict__ xu_loc, const double* __restrict__ xu_scale,
                      const double* __restrict__ offset, const double* __restrict__ n_k,
                      const double* __restrict__ pinv_loc, const int32_t* __restrict__ loc_group, int KL,
                      const uint32_t* __restrict__ tail_all, const int32_t* __restrict__ ovf_count,
                      const int64_t* __restrict__ indptr, const float* __restrict__ ovf_val,
                      const uint8_t* __restrict__ ovf_group, dx_fit_opts opts,
                      double* __restrict__ a_var, double* __restrict__ b_var, double* __restrict__ fim_inv,
                      double* __restrict__ ll_out, double* __restrict__ jac_out,
                      int32_t* __restrict__ niter_out, int32_t* __restrict__ flags_out, int per_warp_doubles) {
    extern __shared__ __align__(16) double smem_d[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
    if (g >= n_genes) return;
    double* base = smem_d + (size_t)warp * per_warp_doubles;
    Ctx c;
    c.K = K; c.p = p; c.ps = ps; c.lane = lane;
    c.xl = xu_loc; c.xs = xu_scale; c.off = offset; c.nk = n_k;
    c.tail = tail_all + g * (int64_t)K * DX_HIST_BINS;
    c.novf = ovf_count[g];
    c.oval = ovf_val + indptr[g];
    c.ogrp = ovf_group + indptr[g];
    c.mu = base; c.eta = c.mu + K; c.rr = c.eta + K; c.zeta = c.rr + K; c.S = c.zeta + K; c.w1 = c.S + K; c.w2 = c.w1 + K;
    c.A = c.w2 + K; c.L = c.A + p * p; c.rhs = c.L + p * p; c.a = c.rhs + p; c.at = c.a + p;
    c.b = c.at + p; c.bt = c.b + ps; c.gb = c.bt + ps; c.sb = c.gb + ps; c.Hb = c.sb + ps; c.Lb = c.Hb + ps * ps;
    c.Cj = c.Lb + ps * ps;
    {   // constant dispersion model <=> every group carries the same scale design row
        bool same = true;
#pragma unroll 1
        for (int k = lane; k < K; k += 32)
#pragma unroll 1
            for (int i = 0; i < ps; ++i) same = same && (xu_scale[k * ps + i] == xu_scale[i]);
        c.shared_r = __all_sync(DX_FULL_MASK, same);
    }
    c.T_r = 0.0;

    // ---- sufficient statistics of the gene ---------------------------------------------------------
    int ymax = 0;
    double lgc = 0.0;
    double cj0 = 0.0, cj1 = 0.0;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        // lanes over j (two bins each): S_k = sum_j c_k(j), sum y^2 = sum_j c(j)(2j+1), sum lgamma(y+1) = sum_j c(j) log(j+1)
        const unsigned c0 = c.tail[k * DX_HIST_BINS + lane], c1 = c.tail[k * DX_HIST_BINS + 32 + lane];
        double s = (double)c0 + (double)c1;
        double q = (double)c0 * (2.0 * lane + 1.0) + (double)c1 * (2.0 * (lane + 32) + 1.0);
        cj0 += (double)c0;
        cj1 += (double)c1;
        if (c0) ymax = max(ymax, lane + 1);
        if (c1) ymax = max(ymax, lane + 33);
        s = dx_warp_sum(s);
        q = dx_warp_sum(q);
        if (lane == 0) { c.S[k] = s; c.w2[k] = q; }
    }
    c.Cj[lane] = cj0;
    c.Cj[lane + 32] = cj1;
    if (cj0 != 0.0) lgc += cj0 * log((double)lane + 1.0);
    if (cj1 != 0.0) lgc += cj1 * log((double)lane + 33.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ymax = max(ymax, __shfl_xor_sync(DX_FULL_MASK, ymax, o));
    c.ymax = ymax;
    c.ovf_groups_lo[0] = c.ovf_groups_lo[1] = c.ovf_groups_lo[2] = c.ovf_groups_lo[3] = 0ull;
    __syncwarp();
    if (c.novf > 0) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            double s = 0.0, q = 0.0;
            bool any = false;
#pragma unroll 1
            for (int t = lane; t < c.novf; t += 32) {
                if (c.ogrp[t] == k) {
                    const double y = (double)c.oval[t];
                    s += y; q += y * y; lgc += lgamma(y + 1.0); any = true;
                }
            }
            s = dx_warp_sum(s); q = dx_warp_sum(q);
            if (__any_sync(DX_FULL_MASK, any)) {
                c.ovf_groups_lo[k >> 6] |= (1ull << (k & 63));
                if (lane == 0) { c.S[k] += s; c.w2[k] += q; }
            }
        }
        __syncwarp();
    }
    c.lgam_const = dx_warp_sum(lgc);
    double n_total = 0.0, sum_all = 0.0, m_sf = 0.0, e2_sf = 0.0;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const double isf = offset ? exp(-offset[k]) : 1.0;
        n_total += n_k[k];
        sum_all += c.S[k];
        m_sf += c.S[k] * isf;
        e2_sf += c.w2[k] * isf * isf;
    }
    c.n_total = n_total;
    int flags = 0;
    if (sum_all == 0.0 && c.novf == 0) flags |= DX_FLAG_ALL_ZERO;

    // ---- initialize (batchglm init_par) -----------------------------------------------------------
    if (opts.init_a == DX_INIT_STANDARD) {
#pragma unroll 1
        for (int i = lane; i < p; i += 32) c.a[i] = (i == 0) ? dx_clip(log(sum_all / n_total), DX_ETA_MIN, DX_ETA_MAX) : 0.0;
    } else if (opts.init_a == DX_INIT_CLOSED_FORM) {
        // batchglm closedform_glm_mean: log of the mean of x / size_factor over every distinct LOCATION design
        // row, pushed through the pseudo-inverse of those rows (np.linalg.lstsq in the reference)
#pragma unroll 1
        for (int l = lane; l < KL; l += 32) {
            double num = 0.0, den = 0.0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                if ((loc_group ? loc_group[k] : k) == l) {
                    num += c.S[k] * (offset ? exp(-offset[k]) : 1.0);
                    den += n_k[k];
                }
            }
            c.w1[l] = log(num / den + DX_TINY);
        }
        __syncwarp();
#pragma unroll 1
        for (int i = lane; i < p; i += 32) {
            double acc = 0.0;
#pragma unroll 1
            for (int l = 0; l < KL; ++l) acc += pinv_loc[i * KL + l] * c.w1[l];
            c.a[i] = dx_clip(acc, DX_ETA_MIN, DX_ETA_MAX);
        }
        __syncwarp();
    } else if (opts.init_a == DX_INIT_ALL_ZERO) {
#pragma unroll 1
        for (int i = lane; i < p; i += 32) c.a[i] = 0.0;
    } else {
#pragma unroll 1
        for (int i = lane; i < p; i += 32) c.a[i] = dx_clip(a_var[(int64_t)i * n_genes + g], DX_ETA_MIN, DX_ETA_MAX);
    }
    if (opts.init_b == DX_INIT_STANDARD) {
        const double m = m_sf / n_total;
        const double v = e2_sf / n_total - m * m;
        const double denom = fmax(v - m, 2.2227587494850775e-162);
#pragma unroll 1
        for (int i = lane; i < ps; i += 32) c.b[i] = (i == 0) ? dx_clip(log(m * m / denom + DX_TINY), DX_ETA_MIN, DX_ETA_MAX) : 0.0;
    } else if (opts.init_b == DX_INIT_ALL_ZERO) {
#pragma unroll 1
        for (int i = lane; i < ps; i += 32) c.b[i] = 0.0;
    } else {
#pragma unroll 1
        for (int i = lane; i < ps; i += 32) c.b[i] = dx_clip(b_var[(int64_t)i * n_genes + g], DX_ETA_MIN, DX_ETA_MAX);
    }
    __syncwarp();

    // ---- train (batchglm numpy EstimatorGlm.train, per-gene form) -----------------------------------
    const bool train_loc = opts.train_loc != 0, train_scale = opts.train_scale != 0;
    const int freq = train_scale ? (train_loc ? opts.update_b_freq : 1) : 0x7fffffff;
    int epochs_until_b = freq;
    bool loc_conv = false, fully = false;
    set_b(c, c.b);
    double nll = -ll_at(c, c.a);
    int step = 0;
#pragma unroll 1
    while (!fully && step < opts.max_steps) {
        double nll_new = nll;
        bool b_updated = false;
        if (epochs_until_b == 0) {
            if (train_scale) {
                const double bold = (lane < ps) ? c.b[lane] : 0.0;    // ps <= 32: one coefficient per lane
                compute_mu(c, c.a);
                const double f = scale_newton(c, -nll, opts.newton_max_iter, opts.newton_xtol);
                const double nll_prop = -f;
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    nll_new = nll_prop;
                } else {
                    if (lane < ps) c.b[lane] = bold;
                    __syncwarp();
                    set_b(c, c.b);
                }
            }
            epochs_until_b = freq;
            b_updated = true;
        } else {
            if (train_loc && !loc_conv) {
                compute_mu(c, c.a);
                loc_terms(c);
                const bool ok = chol_warp(c.A, c.L, p, lane);
                bool fin = true;
#pragma unroll 1
                for (int i = lane; i < p; i += 32) fin = fin && isfinite(c.rhs[i]);
                fin = __all_sync(DX_FULL_MASK, fin);
                if (ok && fin) {
                    if (lane == 0) chol_solve_serial(c.L, p, c.rhs, c.rhs);
                    __syncwarp();
#pragma unroll 1
                    for (int i = lane; i < p; i += 32) c.at[i] = dx_clip(c.a[i] + c.rhs[i], DX_ETA_MIN, DX_ETA_MAX);
                } else {
                    flags |= DX_FLAG_SINGULAR_FIM;
#pragma unroll 1
                    for (int i = lane; i < p; i += 32) c.at[i] = c.a[i];
                }
                __syncwarp();
                const double nll_prop = -ll_at(c, c.at);
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
#pragma unroll 1
                    for (int i = lane; i < p; i += 32) c.a[i] = c.at[i];
                    nll_new = nll_prop;
                }
                __syncwarp();
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

    // ---- finalize (batchglm Estimator.finalize) -----------------------------------------------------
    const double zmax = set_b(c, c.b);
    if (zmax >= DX_ETA_SCALE_FROZEN) flags |= DX_FLAG_SCALE_FROZEN;
    compute_mu(c, c.a);
    loc_terms(c);
    double jac = 0.0;
#pragma unroll 1
    for (int i = lane; i < p; i += 32) jac += fabs(c.rhs[i]);
    const bool okf = chol_warp(c.A, c.L, p, lane);
    __syncwarp();
    // inverse: lane col solves L L^T x = e_col serially into its own column of c.A (A is no longer needed)
#pragma unroll 1
    for (int col = lane; col < p; col += 32) {
        double* x = c.A + col * p;     // column stored as a row; the inverse is symmetric
        if (okf) {
#pragma unroll 1
            for (int i = 0; i < p; ++i) {
                double s = (i == col) ? 1.0 : 0.0;
#pragma unroll 1
                for (int k = 0; k < i; ++k) s -= c.L[i * p + k] * x[k];
                x[i] = s / c.L[i * p + i];
            }
#pragma unroll 1
            for (int i = p - 1; i >= 0; --i) {
                double s = x[i];
#pragma unroll 1
                for (int k = i + 1; k < p; ++k) s -= c.L[k * p + i] * x[k];
                x[i] = s / c.L[i * p + i];
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < p; ++i) x[i] = NAN;
        }
    }
    __syncwarp();
    double* fo = fim_inv + g * (int64_t)p * p;
#pragma unroll 1
    for (int e = lane; e < p * p; e += 32) {
        const int i = e / p, j = e - i * p;
        fo[e] = okf ? 0.5 * (c.A[i * p + j] + c.A[j * p + i]) : NAN;
    }
    scale_terms(c);
#pragma unroll 1
    for (int i = lane; i < ps; i += 32) jac += fabs(c.gb[i]);
    jac = dx_warp_sum(jac) / n_total;
#pragma unroll 1
    for (int i = lane; i < p; i += 32) a_var[(int64_t)i * n_genes + g] = c.a[i];
#pragma unroll 1
    for (int i = lane; i < ps; i += 32) b_var[(int64_t)i * n_genes + g] = c.b[i];
    if (lane == 0) {
        ll_out[g] = -nll;
        jac_out[g] = jac;
        niter_out[g] = step;
        flags_out[g] = flags;
    }
}

}  // namespace

cudaError_t dx_launch_fit_grouped(int64_t n_genes, int K, int p, int ps,
                                  const double* xu_loc, const double* xu_scale, const double* offset,
                                  const double* n_k, const double* pinv_loc, const int32_t* loc_group, int n_loc_groups,
                                  const uint32_t* tail, const int32_t* ovf_count, const int64_t* indptr,
                                  const float* ovf_val, const uint8_t* ovf_group, const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, cudaStream_t stream) {
    if (n_genes <= 0) return cudaSuccess;
    const int per_warp = 7 * K + 2 * p * p + 3 * p + 4 * ps + 2 * ps * ps + DX_HIST_BINS + 2;
    const size_t smem = (size_t)per_warp * WARPS_PER_CTA * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(dx_fit_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((n_genes + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    dx_fit_grouped_kernel<<<grid, WARPS_PER_CTA * 32, smem, stream>>>(
        n_genes, K, p, ps, xu_loc, xu_scale, offset, n_k, pinv_loc, loc_group, n_loc_groups, tail, ovf_count, indptr, ovf_val, ovf_group,
        *opts, a_var, b_var, fim_inv, ll, jac, niter, flags, per_warp);
    return cudaGetLastError();
}
