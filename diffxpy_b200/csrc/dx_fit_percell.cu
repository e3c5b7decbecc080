// dx_fit_percell.cu -- K_fit (per-cell design): the batchglm training schedule for designs whose rows do NOT
// collapse into a few groups (continuous covariates / spline bases, per-cell size factors).  One CTA per gene
// (dynamic gene queue); every training step streams all cells of the gene once:
//   eta_i = x_i . a + log s_i, mu_i = exp(eta_i), r_i = exp(z_i . b)
//   W_i = mu_i r_i / (r_i + mu_i),  X^T W X and X^T W ybar accumulated in fp64 registers (upper triangle),
//   ll_i = G(r_i, y_i) + y_i (eta_i - log1p(mu_i / r_i)) - r_i log1p(mu_i / r_i) - lgamma(y_i + 1)
// The design matrix (coefficient-major, L2 resident) is read with coalesced loads; the gene's sparse column is
// densified tile by tile in shared memory so zero and non-zero cells run the same code.
// Restates /root/reference/diffxpy/testing/tests.py:222-248 (batchglm numpy Estimator) exactly like
// dx_fit_grouped.cu; CPU twin: oracle/nb_glm.py (method_b="newton").
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int PT = 256;            // threads per CTA
constexpr int TILE = 2048;         // cells densified per tile
constexpr int NW = PT / 32;


enum PassMode { PASS_LL = 0, PASS_LOC = 1, PASS_SCALE = 2 };

struct Smem {
    float* ytile;        // [TILE]
    double *a, *at, *b, *bt;            // p, p, ps, ps
    double *A, *A2, *L, *rhs, *rhs2;    // p*p, p*p, p*p, p, p
    double *gb, *sb, *Hb, *Lb;          // ps, ps, ps*ps, ps*ps
    double* red;         // [NW][nacc]
    double* scal;        // small scalar mailbox
    int* ictl;           // integer control words
};

// small-count fast paths (exact same values as the lgamma / digamma differences up to rounding)
__device__ __forceinline__ double g_ratio(double r, double y) {
    const int yi = (int)y;
    if ((double)yi == y && yi >= 1 && yi <= 12) {
        // sum_{j<y} log(1 + j/r) = log(prod_{j<y} (1 + j/r)): one log; the product stays far below overflow for
        // y <= 12 unless r < 1e-24, where the lgamma form takes over
        if (yi == 1) return 0.0;
        if (r > 1e-24) {
            const double ir = 1.0 / r;
            double prod = 1.0 + ir;
#pragma unroll 1
            for (int j = 2; j < yi; ++j) prod *= fma((double)j, ir, 1.0);
            return log(prod);
        }
    }
    return dx_lgamma_ratio(r, y);
}
__device__ __forceinline__ void psi_diffs(double r, double y, double& d0, double& d1) {
    const int yi = (int)y;
    if ((double)yi == y && yi >= 1 && yi <= 12) {
        d0 = 0.0; d1 = 0.0;
#pragma unroll 1
        for (int j = 0; j < yi; ++j) { const double inv = 1.0 / (r + (double)j); d0 += inv; d1 -= inv * inv; }
        return;
    }
    d0 = dx_digamma(r + y) - dx_digamma(r);
    d1 = dx_trigamma(r + y) - dx_trigamma(r);
}

// deterministic block reduction of `n` per-thread partials (acc[0..n)) into out[0..n): warp shuffles, then a
// fixed-order sum over the warps
template <int NMAX>
__device__ __forceinline__ void block_reduce(double (&acc)[NMAX], int n, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < NMAX; ++e) {
        if (e < n) {
            const double v = dx_warp_sum(acc[e]);
            if (lane == 0) red[warp * NMAX + e] = v;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += PT) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * NMAX + e];
        out[e] = s;
    }
    __syncthreads();
}

template <int PMAX, int PSMAX>
struct Pass {
    static constexpr int NLOC = PMAX * (PMAX + 1) / 2 + PMAX + 1;
    static constexpr int NSC = PSMAX * (PSMAX + 1) / 2 + PSMAX + 1;
    static constexpr int NRED = (NLOC > NSC) ? NLOC : NSC;

    // One pass over all cells of the gene at (avec, bvec).  Results: ll (returned to every thread);
    // PASS_LOC: A_out (p x p, symmetric), rhs_out (p); PASS_SCALE: gb (ps), Hb (= -Hessian, ps x ps).
    static __device__ double run(int mode, const Smem& sm, const double* avec, const double* bvec,
                                 int64_t N, int p, int ps, const double* __restrict__ dl, const double* __restrict__ ds,
                                 const double* __restrict__ log_sf, const int32_t* __restrict__ rows,
                                 const float* __restrict__ vals, int64_t e_start, int64_t e_end,
                                 double* A_out, double* rhs_out, double lgam_const, bool scale_const) {
        const int tid = threadIdx.x;
        double acc[NRED];
#pragma unroll
        for (int e = 0; e < NRED; ++e) acc[e] = 0.0;
        double av[PMAX], bv[PSMAX];
#pragma unroll
        for (int j = 0; j < PMAX; ++j) av[j] = (j < p) ? avec[j] : 0.0;
#pragma unroll
        for (int j = 0; j < PSMAX; ++j) bv[j] = (j < ps) ? bvec[j] : 0.0;
        // constant dispersion model (one scale coefficient, identical design value in every cell): r is a per-pass scalar
        const double z_const = scale_const ? __ldg(ds) : 0.0;
        const double zeta_const = scale_const ? dx_clip(z_const * bv[0], DX_ETA_MIN, DX_ETA_MAX) : 0.0;
        const double r_const = scale_const ? exp(zeta_const) : 0.0;
        if (tid == 0) sm.ictl[0] = 0;       // entry cursor (relative to e_start)
        __syncthreads();
        for (int64_t t0 = 0; t0 < N; t0 += TILE) {
            const int tn = (int)((N - t0 < TILE) ? (N - t0) : TILE);
            for (int c = tid; c < tn; c += PT) sm.ytile[c] = 0.0f;
            const int cur = sm.ictl[0];
            __syncthreads();
            if (tid == 0) sm.ictl[1] = 0x7fffffff;
            __syncthreads();
            {   // scatter the entries whose cell falls into this tile (rows ascend inside a gene)
                int64_t e = e_start + cur + tid;
                const int64_t t1 = t0 + tn;
#pragma unroll 1
                for (; e < e_end; e += PT) {
                    const int64_t row = rows[e];
                    if (row >= t1) break;
                    sm.ytile[(int)(row - t0)] = vals[e];
                }
                const int64_t stop = (e < e_end) ? e : e_end;     // first entry this thread did not consume
                atomicMin(&sm.ictl[1], (int)(stop - e_start));
            }
            __syncthreads();
            if (tid == 0) sm.ictl[0] = sm.ictl[1];
#pragma unroll 1
            for (int c = tid; c < tn; c += PT) {
                const int64_t i = t0 + c;
                const double y = (double)sm.ytile[c];
                double x[PMAX];
                double eta = log_sf ? log_sf[i] : 0.0;
                {
                    const double* px = dl + i;           // coefficient-major design: column j of cell i at px + j * N
#pragma unroll
                    for (int j = 0; j < PMAX; ++j) {
                        x[j] = (j < p) ? *px : 0.0;
                        px += N;
                        eta += x[j] * av[j];
                    }
                }
                eta = dx_clip(eta, DX_ETA_MIN, DX_ETA_MAX);
                double z[PSMAX];
                double zeta = zeta_const;
                if (!scale_const) {
                    zeta = 0.0;
                    const double* pz = ds + i;
#pragma unroll
                    for (int j = 0; j < PSMAX; ++j) {
                        z[j] = (j < ps) ? __ldg(pz) : 0.0;
                        pz += N;
                        zeta += z[j] * bv[j];
                    }
                    zeta = dx_clip(zeta, DX_ETA_MIN, DX_ETA_MAX);
                } else {
#pragma unroll
                    for (int j = 0; j < PSMAX; ++j) z[j] = (j == 0) ? z_const : 0.0;
                }
                const double mu = exp(eta), r = scale_const ? r_const : exp(zeta);
                const double l1p = log1p(mu / r);
                double ll = -r * l1p;
                if (y != 0.0) ll += g_ratio(r, y) + y * (eta - l1p);
                if (mode == PASS_LOC) {
                    const double q = r / (r + mu);
                    const double w = mu * q, sc = (y - mu) * q;
                    int e = 0;
#pragma unroll
                    for (int a = 0; a < PMAX; ++a) {
                        const double wx = w * x[a];
#pragma unroll
                        for (int b2 = 0; b2 <= a; ++b2) { acc[e] += wx * x[b2]; ++e; }
                    }
#pragma unroll
                    for (int a = 0; a < PMAX; ++a) acc[PMAX * (PMAX + 1) / 2 + a] += sc * x[a];
                    acc[NLOC - 1] += ll;
                } else if (mode == PASS_SCALE) {
                    double d0 = 0.0, d1 = 0.0;
                    if (y != 0.0) psi_diffs(r, y, d0, d1);
                    const double rm = r + mu;
                    const double g = r * (d0 - l1p + (mu - y) / rm);
                    const double h = g + r * r * d1 + r * (mu * mu + r * y) / (rm * rm);
                    int e = 0;
#pragma unroll
                    for (int a = 0; a < PSMAX; ++a) {
                        const double hz = h * z[a];
#pragma unroll
                        for (int b2 = 0; b2 <= a; ++b2) { acc[e] += hz * z[b2]; ++e; }
                    }
#pragma unroll
                    for (int a = 0; a < PSMAX; ++a) acc[PSMAX * (PSMAX + 1) / 2 + a] += g * z[a];
                    acc[NSC - 1] += ll;
                } else {
                    acc[0] += ll;
                }
            }
            __syncthreads();
        }
        // reduce
        const int n = (mode == PASS_LOC) ? NLOC : ((mode == PASS_SCALE) ? NSC : 1);
        double* out = sm.scal + 8;           // scratch for the reduced vector (NRED doubles)
        block_reduce<NRED>(acc, n, sm.red, out);
        double ll_tot;
        if (mode == PASS_LOC) {
            for (int e = tid; e < PMAX * (PMAX + 1) / 2; e += PT) {
                int a = 0;
                while ((a + 1) * (a + 2) / 2 <= e) ++a;
                const int b2 = e - a * (a + 1) / 2;
                if (a < p && b2 < p) { A_out[a * p + b2] = out[e]; A_out[b2 * p + a] = out[e]; }
            }
            for (int a = tid; a < p; a += PT) rhs_out[a] = out[PMAX * (PMAX + 1) / 2 + a];
            ll_tot = out[NLOC - 1];
        } else if (mode == PASS_SCALE) {
            for (int e = tid; e < PSMAX * (PSMAX + 1) / 2; e += PT) {
                int a = 0;
                while ((a + 1) * (a + 2) / 2 <= e) ++a;
                const int b2 = e - a * (a + 1) / 2;
                if (a < ps && b2 < ps) { sm.Hb[a * ps + b2] = -out[e]; sm.Hb[b2 * ps + a] = -out[e]; }
            }
            for (int a = tid; a < ps; a += PT) sm.gb[a] = out[PSMAX * (PSMAX + 1) / 2 + a];
            ll_tot = out[NSC - 1];
        } else {
            ll_tot = out[0];
        }
        __syncthreads();
        return ll_tot - lgam_const;
    }
};

// Newton direction by warp 0 (same rule as dx_fit_grouped.cu / oracle _newton_direction); result in sm.sb;
// returns false on non-finite input.  Must be called by all threads of warp 0 only.
__device__ bool percell_direction(const Smem& sm, int ps, int lane) {
    if (ps == 1) {
        const double h = sm.Hb[0], g = sm.gb[0];
        if (!(isfinite(h) && isfinite(g))) return false;
        if (lane == 0) sm.sb[0] = (h > 0.0) ? g / h : (double)((g > 0.0) - (g < 0.0));
        __syncwarp();
        return true;
    }
    bool fin = true;
    for (int e = lane; e < ps * ps; e += 32) fin = fin && isfinite(sm.Hb[e]);
    for (int i = lane; i < ps; i += 32) fin = fin && isfinite(sm.gb[i]);
    if (!__all_sync(DX_FULL_MASK, fin)) return false;
    const double diag = (lane < ps) ? sm.Hb[lane * ps + lane] : 0.0;
    const double base = fmax(dx_warp_max(fabs(diag)), 1e-300);
    double lam = 0.0;
    for (int it = 0; it < 12; ++it) {
        if (lane < ps) sm.Hb[lane * ps + lane] = diag + lam;
        __syncwarp();
        const bool ok = chol_warp(sm.Hb, sm.Lb, ps, lane);
        __syncwarp();
        if (ok) {
            if (lane == 0) chol_solve_serial(sm.Lb, ps, sm.gb, sm.sb);
            __syncwarp();
            return true;
        }
        lam = (lam == 0.0) ? base * 1e-6 : lam * 10.0;
    }
    for (int i = lane; i < ps; i += 32) { const double g = sm.gb[i]; sm.sb[i] = (double)((g > 0.0) - (g < 0.0)); }
    __syncwarp();
    return true;
}

// max_i clip(z_i . b): decides whether the scale predictor sits on the numerical plateau (all threads get it)
__device__ double scale_eta_max(const Smem& sm, const double* __restrict__ ds, const double* bvec, int64_t N, int ps,
                                bool scale_const) {
    if (scale_const) return dx_clip(__ldg(ds) * bvec[0], DX_ETA_MIN, DX_ETA_MAX);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double m = -INFINITY;
#pragma unroll 1
    for (int64_t i = tid; i < N; i += PT) {
        double z = 0.0;
        for (int j = 0; j < ps; ++j) z += ds[(int64_t)j * N + i] * bvec[j];
        m = fmax(m, dx_clip(z, DX_ETA_MIN, DX_ETA_MAX));
    }
    m = dx_warp_max(m);
    if (lane == 0) sm.red[warp] = m;
    __syncthreads();
    double zmax = -INFINITY;
    for (int w = 0; w < NW; ++w) zmax = fmax(zmax, sm.red[w]);
    __syncthreads();
    return zmax;
}

template <int PMAX, int PSMAX>
__global__ void __launch_bounds__(PT)
dx_fit_percell_kernel(int64_t n_genes, int64_t N, int p, int ps,
                      const double* __restrict__ dl, const double* __restrict__ ds, const double* __restrict__ log_sf,
                      const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows, const float* __restrict__ vals,
                      dx_fit_opts opts, double* __restrict__ a_var, double* __restrict__ b_var,
                      double* __restrict__ fim_inv, double* __restrict__ ll_out, double* __restrict__ jac_out,
                      int32_t* __restrict__ niter_out, int32_t* __restrict__ flags_out, unsigned int* __restrict__ gene_queue) {
    using P = Pass<PMAX, PSMAX>;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem sm;
    {
        double* d = reinterpret_cast<double*>(smraw);
        sm.a = d; d += p; sm.at = d; d += p; sm.b = d; d += ps; sm.bt = d; d += ps;
        sm.A = d; d += p * p; sm.A2 = d; d += p * p; sm.L = d; d += p * p; sm.rhs = d; d += p; sm.rhs2 = d; d += p;
        sm.gb = d; d += ps; sm.sb = d; d += ps; sm.Hb = d; d += ps * ps; sm.Lb = d; d += ps * ps;
        sm.red = d; d += NW * P::NRED;
        sm.scal = d; d += 8 + P::NRED;
        sm.ictl = reinterpret_cast<int*>(d); d += 4;
        sm.ytile = reinterpret_cast<float*>(d);
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool scale_const = (ps == 1) && (opts.reserved & DX_OPT_SCALE_CONST);
    __shared__ int s_gene;
    for (;;) {
        if (tid == 0) s_gene = (int)atomicAdd(gene_queue, 1u);
        __syncthreads();
        const int64_t g = s_gene;
        __syncthreads();
        if (g >= n_genes) return;
        const int64_t e0 = indptr[g], e1 = indptr[g + 1];
        // constant of the likelihood and start values
        double lg = 0.0;
        for (int64_t e = e0 + tid; e < e1; e += PT) lg += lgamma((double)vals[e] + 1.0);
        {
            double one[1] = {lg};
            block_reduce<1>(one, 1, sm.red, sm.scal);
        }
        const double lgam_const = sm.scal[0];
        for (int i = tid; i < p; i += PT) sm.a[i] = dx_clip(a_var[(int64_t)i * n_genes + g], DX_ETA_MIN, DX_ETA_MAX);
        for (int i = tid; i < ps; i += PT) sm.b[i] = dx_clip(b_var[(int64_t)i * n_genes + g], DX_ETA_MIN, DX_ETA_MAX);
        __syncthreads();
        int flags = (e1 == e0) ? DX_FLAG_ALL_ZERO : 0;
        const bool train_loc = opts.train_loc != 0, train_scale = opts.train_scale != 0;
        const int freq = train_scale ? (train_loc ? opts.update_b_freq : 1) : 0x7fffffff;
        int epochs_until_b = freq;
        bool loc_conv = false, fully = false, have_terms = true;
        double nll = -P::run(PASS_LOC, sm, sm.a, sm.b, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, sm.A, sm.rhs, lgam_const, scale_const);
        int step = 0;
        while (!fully && step < opts.max_steps) {
            double nll_new = nll;
            bool b_updated = false;
            if (epochs_until_b == 0) {
                if (train_scale) {
                    // ---- safeguarded Newton in b (oracle _b_step_newton) -------------------------------------
                    for (int i = tid; i < ps; i += PT) sm.bt[i] = sm.b[i];       // bt = working copy, b = old
                    __syncthreads();
                    double zmax = scale_eta_max(sm, ds, sm.bt, N, ps, scale_const);
                    bool moved = false;
                    double f = -nll;
                    if (zmax < DX_ETA_SCALE_FROZEN) {
                        for (int it = 0; it < opts.newton_max_iter; ++it) {
                            P::run(PASS_SCALE, sm, sm.a, sm.bt, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, nullptr, nullptr, lgam_const, scale_const);
                            if (warp == 0) {
                                bool ok = percell_direction(sm, ps, lane);
                                double smax = 0.0, pred = 0.0;
                                if (ok) {
                                    for (int i = lane; i < ps; i += 32) smax = fmax(smax, fabs(sm.sb[i]));
                                    smax = dx_warp_max(smax);
                                    const double sc = (smax > DX_NEWTON_MAX_STEP) ? DX_NEWTON_MAX_STEP / smax : 1.0;
                                    for (int i = lane; i < ps; i += 32) { sm.sb[i] *= sc; pred += sm.gb[i] * sm.sb[i]; }
                                    pred = dx_warp_sum(pred);
                                    smax *= sc;
                                    if (smax <= opts.newton_xtol || !(pred > DX_NEWTON_GAIN_RTOL * fabs(f))) ok = false;
                                }
                                if (lane == 0) sm.ictl[2] = ok ? 1 : 0;
                            }
                            __syncthreads();
                            if (!sm.ictl[2]) break;
                            // line search from bt along sb; candidate kept in gb (free after the direction)
                            double t = 1.0, f_try = f;
                            bool accepted = false;
                            for (int h = 0; h < DX_NEWTON_MAX_HALVINGS; ++h) {
                                for (int i = tid; i < ps; i += PT)
                                    sm.gb[i] = dx_clip(sm.bt[i] + t * sm.sb[i], DX_ETA_MIN, DX_ETA_MAX);
                                __syncthreads();
                                f_try = P::run(PASS_LL, sm, sm.a, sm.gb, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, nullptr, nullptr, lgam_const, scale_const);
                                if (isfinite(f_try) && f_try >= f) { accepted = true; break; }
                                t *= 0.5;
                            }
                            if (!accepted) break;
                            for (int i = tid; i < ps; i += PT) sm.bt[i] = sm.gb[i];
                            __syncthreads();
                            f = f_try;
                            moved = true;
                            zmax = scale_eta_max(sm, ds, sm.bt, N, ps, scale_const);
                            if (zmax >= DX_ETA_SCALE_FROZEN) break;
                        }
                    }
                    const double nll_prop = -f;
                    if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                    if (isfinite(nll_prop) && !(nll_prop > nll)) {
                        for (int i = tid; i < ps; i += PT) sm.b[i] = sm.bt[i];
                        nll_new = nll_prop;
                        if (moved) have_terms = false;
                    }
                    __syncthreads();
                }
                epochs_until_b = freq;
                b_updated = true;
            } else {
                if (train_loc && !loc_conv) {
                    if (!have_terms) {
                        P::run(PASS_LOC, sm, sm.a, sm.b, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, sm.A, sm.rhs, lgam_const, scale_const);
                        have_terms = true;
                    }
                    if (warp == 0) {
                        const bool ok = chol_warp(sm.A, sm.L, p, lane);
                        bool fin = true;
                        for (int i = lane; i < p; i += 32) fin = fin && isfinite(sm.rhs[i]);
                        fin = __all_sync(DX_FULL_MASK, fin);
                        __syncwarp();
                        if (ok && fin) {
                            if (lane == 0) chol_solve_serial(sm.L, p, sm.rhs, sm.rhs2);
                            __syncwarp();
                            for (int i = lane; i < p; i += 32) sm.at[i] = dx_clip(sm.a[i] + sm.rhs2[i], DX_ETA_MIN, DX_ETA_MAX);
                        } else {
                            for (int i = lane; i < p; i += 32) sm.at[i] = sm.a[i];
                        }
                        if (lane == 0) sm.ictl[2] = (ok && fin) ? 1 : 0;
                    }
                    __syncthreads();
                    if (!sm.ictl[2]) flags |= DX_FLAG_SINGULAR_FIM;
                    const double nll_prop = -P::run(PASS_LOC, sm, sm.at, sm.b, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, sm.A2, sm.rhs2, lgam_const, scale_const);
                    if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                    if (isfinite(nll_prop) && !(nll_prop > nll)) {
                        for (int i = tid; i < p; i += PT) { sm.a[i] = sm.at[i]; sm.rhs[i] = sm.rhs2[i]; }
                        for (int e = tid; e < p * p; e += PT) sm.A[e] = sm.A2[e];
                        nll_new = nll_prop;
                    }
                    __syncthreads();
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
        // ---- finalize --------------------------------------------------------------------------------------
        if (!have_terms)
            P::run(PASS_LOC, sm, sm.a, sm.b, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, sm.A, sm.rhs, lgam_const, scale_const);
        P::run(PASS_SCALE, sm, sm.a, sm.b, N, p, ps, dl, ds, log_sf, rows, vals, e0, e1, nullptr, nullptr, lgam_const, scale_const);
        if (scale_eta_max(sm, ds, sm.b, N, ps, scale_const) >= DX_ETA_SCALE_FROZEN) flags |= DX_FLAG_SCALE_FROZEN;
        if (warp == 0) {
            double jac = 0.0;
            for (int i = lane; i < p; i += 32) jac += fabs(sm.rhs[i]);
            for (int i = lane; i < ps; i += 32) jac += fabs(sm.gb[i]);
            jac = dx_warp_sum(jac) / (double)N;
            const bool okf = chol_warp(sm.A, sm.L, p, lane);
            __syncwarp();
            for (int col = lane; col < p; col += 32) {
                double* x = sm.A2 + col * p;
                if (okf) {
                    for (int i = 0; i < p; ++i) {
                        double s = (i == col) ? 1.0 : 0.0;
                        for (int k = 0; k < i; ++k) s -= sm.L[i * p + k] * x[k];
                        x[i] = s / sm.L[i * p + i];
                    }
                    for (int i = p - 1; i >= 0; --i) {
                        double s = x[i];
                        for (int k = i + 1; k < p; ++k) s -= sm.L[k * p + i] * x[k];
                        x[i] = s / sm.L[i * p + i];
                    }
                } else {
                    for (int i = 0; i < p; ++i) x[i] = NAN;
                }
            }
            __syncwarp();
            double* fo = fim_inv + g * (int64_t)p * p;
            for (int e = lane; e < p * p; e += 32) {
                const int i = e / p, j = e - i * p;
                fo[e] = okf ? 0.5 * (sm.A2[i * p + j] + sm.A2[j * p + i]) : NAN;
            }
            for (int i = lane; i < p; i += 32) a_var[(int64_t)i * n_genes + g] = sm.a[i];
            for (int i = lane; i < ps; i += 32) b_var[(int64_t)i * n_genes + g] = sm.b[i];
            if (lane == 0) {
                ll_out[g] = -nll;
                jac_out[g] = jac;
                niter_out[g] = step;
                flags_out[g] = flags;
            }
        }
        __syncthreads();
    }
}

template <int PMAX, int PSMAX>
cudaError_t launch_t(int64_t n_genes, int64_t n_cells, int p, int ps, const double* dl, const double* ds,
                     const double* log_sf, const int64_t* indptr, const int32_t* rows, const float* vals,
                     const dx_fit_opts* opts, double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                     int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream) {
    using P = Pass<PMAX, PSMAX>;
    const size_t n_d = (size_t)2 * p + 2 * ps + 3 * p * p + 2 * p + 2 * ps + 2 * ps * ps + NW * P::NRED + 8 + P::NRED + 4;
    const size_t smem = n_d * sizeof(double) + TILE * sizeof(float) + 16;
    auto kern = dx_fit_percell_kernel<PMAX, PSMAX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned int* queue = nullptr;          // per-launch work queue slot (safe with concurrent launches on other streams)
    e = dx_queue_acquire(stream, &queue);
    if (e != cudaSuccess) return e;
    int64_t grid = (int64_t)sm_count * per_sm;
    if (grid > n_genes) grid = n_genes;
    kern<<<(unsigned)grid, PT, smem, stream>>>(n_genes, n_cells, p, ps, dl, ds, log_sf, indptr, rows, vals, *opts,
                                               a_var, b_var, fim_inv, ll, jac, niter, flags, queue);
    return cudaGetLastError();
}

}  // namespace

cudaError_t dx_launch_fit_percell(int64_t n_genes, int64_t n_cells, int p, int ps,
                                  const double* design_loc, const double* design_scale, const double* log_sf,
                                  const int64_t* indptr, const int32_t* rows, const float* vals,
                                  const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream) {
    if (n_genes <= 0) return cudaSuccess;
    if (p > 16 || ps > 8) return cudaErrorNotSupported;
#define DX_PC(PM, PSM) return launch_t<PM, PSM>(n_genes, n_cells, p, ps, design_loc, design_scale, log_sf, indptr, rows, \
                                                 vals, opts, a_var, b_var, fim_inv, ll, jac, niter, flags, sm_count, stream)
    if (ps == 1) {
        if (p <= 4) DX_PC(4, 1);
        if (p <= 8) DX_PC(8, 1);
        if (p <= 12) DX_PC(12, 1);
        DX_PC(16, 1);
    }
    if (p <= 4) DX_PC(4, 8);
    if (p <= 8) DX_PC(8, 8);
    if (p <= 12) DX_PC(12, 8);
    DX_PC(16, 8);
#undef DX_PC
}
