// dx_fit_grouped.cu -- K_fit (grouped design): the whole batchglm training schedule for one gene runs
// inside ONE SUB-WARP (8, 16 or 32 lanes, picked from the number of design groups / coefficients), from
// initialisation to finalize, on the per-(gene, group) sufficient statistics that K_ingest produced.
// No pass over the count matrix, no host loop, no global synchronisation: genes are independent, so the
// reference's lock-step loop over all genes
//   (/root/reference/diffxpy/testing/tests.py:222-248 -> batchglm numpy Estimator.initialize /
//    train_sequence("DEFAULT") / finalize; restated in oracle/nb_glm.py, method_b="newton")
// is exactly equivalent to one independent loop per gene.
//
// For design group k (distinct row of [design_loc | design_scale | log size factor], n_k cells):
//   eta_k = x_k.a + off_k, mu_k = dx_exp(eta_k), zeta_k = z_k.b, r_k = dx_exp(zeta_k), l_k = dx_log1p(mu_k / r_k)
//   ll   = T(r) + sum_k [ S_k (eta_k - zeta_k - l_k) - n_k r_k l_k ] - sum_i dx_lgamma(y_i + 1)
//   T(r) = sum_k sum_j c_k(j) dx_log(r_k + j) + sum_ovf [ dx_lgamma(r_k + y) - dx_lgamma(r_k) ]
//   X^T W X    = sum_k n_k W_k x_k x_k^T,   W_k = mu_k r_k / (r_k + mu_k)
//   X^T W ybar = sum_k x_k ( S_k r_k/(r_k + mu_k) - n_k W_k )
// with c_k(j) the tail counts #{i in k: y_i > j} (integer counts <= 64) and S_k = sum_{i in k} y_i.
//
// Work layout: lane <-> design group for everything per group (exp / log1p / reciprocal are evaluated once
// per (a, b) state and cached in shared memory, in a current and a trial copy that swap on acceptance);
// lane <-> matrix row for the p x p LDL^T factorisation and the triangular solves; lane <-> entry of the
// lower triangle for X^T W X, which is a product-table x weight contraction (the products x_ki x_kj do not
// depend on the gene and are staged once per CTA).  Coefficient vectors live in registers (lane i holds
// coefficient i).
#include <cstdlib>
#include "dx_common.cuh"
#include "dx_internal.h"
#include "dx_fit_args.cuh"

namespace {

extern __shared__ __align__(16) double smem_d[];


// helpers on the IWLS path are inlined (their Slot fields then live in registers); the rarely executed scale-model
// helpers stay out of line to keep the hot loop inside the instruction cache
#ifndef DX_FIT_INLINE_HOT
#define DX_FIT_INLINE_HOT 1
#endif
#if DX_FIT_INLINE_HOT
#define DX_HOT __forceinline__
#else
#define DX_HOT __noinline__
#endif

template <int SW> __device__ __forceinline__ double sw_sum(double v, unsigned m) {
#pragma unroll
    for (int o = SW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
template <int SW> __device__ __forceinline__ double sw_max(double v, unsigned m) {
#pragma unroll
    for (int o = SW / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int SW> __device__ __forceinline__ int sw_max_i(int v, unsigned m) {
#pragma unroll
    for (int o = SW / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int SW> __device__ __forceinline__ double sw_bcast(double v, int src, unsigned m) {
    return __shfl_sync(m, v, src, SW);
}
template <int SW> __device__ __forceinline__ bool sw_all(bool v, unsigned m) {
    return (__ballot_sync(m, v) & m) == m;
}

// All arrays of a slot are addressed as offsets (in doubles) into smem_d, so every access is a shared-memory
// instruction and "swap current and trial state" is an integer swap.
struct Slot {
    unsigned mask;
    int sl;                                  // lane inside the sub-warp
    int K, p, ps, ymax, novf, nent;
    bool shared_r, has_nz;
    // CTA-wide tables
    int t_xl, t_xs, t_off, t_nk, t_nzv;      // [K][p], [K][ps], [K], [K], non-zero design products [nz]
    const unsigned short* nz_ptr;            // [nent + 1] entry -> range of its non-zero products (shared memory)
    const unsigned char* nz_k;               // [nz] design group of every non-zero product
    const unsigned char *ei, *ej;            // entry -> (row, column) of the lower triangle (shared memory)
    // per-slot state: a-dependent, b-dependent, (a, b)-dependent; current and trial copies
    int eta, mu, eta_t, mu_t;
    int zeta, rr, zeta_t, rr_t;
    int l1p, irm, l1p_t, irm_t;
    int S, w1, w2, Cj, A, vec, Hb, Hw, Ainv;
    double T_r, lgam_const, n_total;
    unsigned long long ovf_groups[4];        // bitmask of groups that own overflow entries (K <= 255)
    const uint32_t* __restrict__ tail;       // [K][64]
    // overflow entries.  Plain list: (oval[t], ogrp[t]), t < novf, ostep = 1.  Packed by dx_pack_overflow: records
    // (value, multiplicity) every ostep = 2 floats; `oval/ogrp/novf` view the per-(group, value) records and
    // `oval_r/novf_r` the per-value records the shared-dispersion loops use (same list when not packed).
    const float* __restrict__ oval;
    const uint8_t* __restrict__ ogrp;
    const float* __restrict__ oval_r;
    const float* __restrict__ ohead;         // packed: per group {sum y, sum y^2, sum lgamma(y + 1), count} as 4 doubles
    int novf_r, ostep;
};

__device__ __forceinline__ double slot_ovf_weight(const float* rec, int ostep) {
    return ostep == 2 ? (double)__float_as_uint(__ldg(rec + 1)) : 1.0;
}

__device__ __forceinline__ bool slot_group_has_ovf(const Slot& s, int k) {
    return (s.ovf_groups[k >> 6] >> (k & 63)) & 1ull;
}

// T(r) = sum_k sum_j c_k(j) dx_log(r_k + j) + sum_ovf [dx_lgamma(r_k + y) - dx_lgamma(r_k)]
template <int SW> __device__ __noinline__ double compute_T(const Slot& s, int o_rr) {
    const double* rr = smem_d + o_rr;
    double part = 0.0;
    if (s.shared_r) {
        const double r = rr[0];
        const double* Cj = smem_d + s.Cj;
#pragma unroll 1
        for (int j = s.sl; j < s.ymax; j += SW) {
            const double cnt = Cj[j];
            if (cnt != 0.0) part += cnt * dx_log(r + (double)j);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < s.K; ++k) {
            const double r = rr[k];
#pragma unroll 1
            for (int j = s.sl; j < s.ymax; j += SW) {
                const unsigned cnt = __ldg(s.tail + k * DX_HIST_BINS + j);
                if (cnt) part += (double)cnt * dx_log(r + (double)j);
            }
        }
    }
    if (s.shared_r) {
        const double r = rr[0], lr = dx_log(r);
#pragma unroll 1
        for (int t = s.sl; t < s.novf_r; t += SW) {
            const float* rec = s.oval_r + t * s.ostep;
            const double y = (double)__ldg(rec);
            part += slot_ovf_weight(rec, s.ostep) * (dx_lgamma_ratio(r, y) + y * lr);
        }
    } else {
#pragma unroll 1
        for (int t = s.sl; t < s.novf; t += SW) {
            const float* rec = s.oval + t * s.ostep;
            const double y = (double)__ldg(rec);
            const double r = rr[__ldg(s.ogrp + t * s.ostep)];
            part += slot_ovf_weight(rec, s.ostep) * (dx_lgamma_ratio(r, y) + y * dx_log(r));
        }
    }
    return sw_sum<SW>(part, s.mask);
}

// zeta_k = clip(z_k . b), r_k = dx_exp(zeta_k) into the given buffers, T(r) into T_out; returns max_k zeta_k.
// breg: lane i holds b_i.
template <int SW> __device__ __noinline__ double set_b(const Slot& s, double breg, int o_zeta, int o_rr, double& T_out) {
    double* vec = smem_d + s.vec;
    if (s.sl < s.ps) vec[s.sl] = breg;
    __syncwarp(s.mask);
    const double* xs = smem_d + s.t_xs;
    double zmax = -INFINITY;
#pragma unroll 1
    for (int k = s.sl; k < s.K; k += SW) {
        double z = 0.0;
#pragma unroll 1
        for (int i = 0; i < s.ps; ++i) z += xs[k * s.ps + i] * vec[i];
        z = dx_clip(z, DX_ETA_MIN, DX_ETA_MAX);
        smem_d[o_zeta + k] = z;
        smem_d[o_rr + k] = dx_exp(z);
        zmax = fmax(zmax, z);
    }
    __syncwarp(s.mask);
    T_out = compute_T<SW>(s, o_rr);
    return sw_max<SW>(zmax, s.mask);
}

// eta_k = clip(x_k . a + off_k), mu_k = dx_exp(eta_k); areg: lane i holds a_i
template <int SW> __device__ DX_HOT void set_a(const Slot& s, double areg, int o_eta, int o_mu) {
    double* vec = smem_d + s.vec;
    if (s.sl < s.p) vec[s.sl] = areg;
    __syncwarp(s.mask);
    const double* xl = smem_d + s.t_xl;
    const int p = s.p;
#pragma unroll 1
    for (int k = s.sl; k < s.K; k += SW) {
        double e = smem_d[s.t_off + k];
#pragma unroll 3
        for (int i = 0; i < p; ++i) e += xl[k * p + i] * vec[i];
        e = dx_clip(e, DX_ETA_MIN, DX_ETA_MAX);
        smem_d[o_eta + k] = e;
        smem_d[o_mu + k] = dx_exp(e);
    }
    __syncwarp(s.mask);
}

// l_k = dx_log1p(mu_k / r_k), irm_k = 1 / (r_k + mu_k)
template <int SW> __device__ DX_HOT void refresh(const Slot& s, int o_mu, int o_rr, int o_l1p, int o_irm) {
#pragma unroll 1
    for (int k = s.sl; k < s.K; k += SW) {
        const double m = smem_d[o_mu + k], r = smem_d[o_rr + k];
        smem_d[o_irm + k] = 1.0 / (r + m);
        smem_d[o_l1p + k] = dx_log1p(m / r);
    }
    __syncwarp(s.mask);
}

template <int SW> __device__ __forceinline__ double ll_state(const Slot& s, int o_eta, int o_zeta, int o_rr, int o_l1p, double T) {
    double part = 0.0;
#pragma unroll 1
    for (int k = s.sl; k < s.K; k += SW) {
        const double l = smem_d[o_l1p + k];
        part += smem_d[s.S + k] * (smem_d[o_eta + k] - smem_d[o_zeta + k] - l) - smem_d[s.t_nk + k] * smem_d[o_rr + k] * l;
    }
    return sw_sum<SW>(part, s.mask) + (T - s.lgam_const);
}

// A = X^T W X (lower triangle, column-major in s.A) from the current state; returns (X^T W ybar)_i in lane i
template <int SW> __device__ DX_HOT double loc_terms(const Slot& s) {
    const int K = s.K, p = s.p, nent = s.nent;
    double* w1 = smem_d + s.w1;
    double* w2 = smem_d + s.w2;
#pragma unroll 1
    for (int k = s.sl; k < K; k += SW) {
        const double q = smem_d[s.rr + k] * smem_d[s.irm + k];
        const double nw = smem_d[s.t_nk + k] * smem_d[s.mu + k] * q;
        w1[k] = nw;
        w2[k] = smem_d[s.S + k] * q - nw;
    }
    __syncwarp(s.mask);
    const double* xl = smem_d + s.t_xl;
    double* A = smem_d + s.A;
#pragma unroll 1
    for (int e = s.sl; e < nent; e += SW) {
        const int i = s.ei[e], j = s.ej[e];
        double acc = 0.0;
        if (s.has_nz) {
            // categorical designs: most products x_ki x_kj vanish; the CTA-wide list holds the others (same k order,
            // so the sum is bit-identical to the dense loop)
            const double* nzv = smem_d + s.t_nzv;
#pragma unroll 1
            for (int t = s.nz_ptr[e], t1 = s.nz_ptr[e + 1]; t < t1; ++t) acc += w1[s.nz_k[t]] * nzv[t];
        } else {
#pragma unroll 4
            for (int k = 0; k < K; ++k) acc += w1[k] * (xl[k * p + i] * xl[k * p + j]);
        }
        A[j * p + i] = acc;
    }
    double rhs = 0.0;
    if (s.sl < p) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) rhs += w2[k] * xl[k * p + s.sl];
    }
    __syncwarp(s.mask);
    return rhs;
}

// In-place LDL^T of the n x n SPD matrix whose lower triangle is stored column-major in A (lane <-> row).
// On exit A[j*n + i] (i > j) holds L_ij (unit lower), the diagonal holds d_j and lane j returns 1/d_j in dinv.
// Pivot guard of oracle/nb_glm.py:chol_guarded: d_j must be finite and exceed 2^-50 of the input diagonal.
template <int SW> __device__ __noinline__ bool ldl_factor(double* A, int n, int sl, unsigned mask, double& dinv) {
    const double m_ii = (sl < n) ? A[sl * n + sl] : 1.0;
    dinv = 0.0;
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        const double d = A[j * n + j];
        const double mjj = sw_bcast<SW>(m_ii, j, mask);
        if (!(isfinite(d) && d > DX_CHOL_PIVOT_RTOL * mjj && mjj > 0.0)) { ok = false; break; }
        const double inv = 1.0 / d;
        if (sl == j) dinv = inv;
        const bool act = sl > j && sl < n;
        double lij = 0.0;
        if (act) {
            lij = A[j * n + sl] * inv;
#pragma unroll 1
            for (int c = j + 1; c <= sl; ++c) A[c * n + sl] -= lij * A[j * n + c];
        }
        __syncwarp(mask);
        if (act) A[j * n + sl] = lij;     // column j is not read again before the solves
    }
    __syncwarp(mask);
    return ok;
}

// solve (L D L^T) x = rhs; lane i holds rhs_i / returns x_i
template <int SW> __device__ __forceinline__ double ldl_solve(const double* A, int n, int sl, unsigned mask, double dinv, double x) {
#pragma unroll 1
    for (int j = 0; j < n - 1; ++j) {
        const double xj = sw_bcast<SW>(x, j, mask);
        if (sl > j && sl < n) x -= A[j * n + sl] * xj;
    }
    x *= dinv;
#pragma unroll 1
    for (int j = n - 1; j > 0; --j) {
        const double xj = sw_bcast<SW>(x, j, mask);
        if (sl < j) x -= A[sl * n + j] * xj;
    }
    return x;
}

// gradient (returned: lane i holds g_i) and NEGATIVE Hessian (s.Hb, full symmetric ps x ps) of ll wrt b at the
// current state
template <int SW> __device__ __noinline__ double scale_terms(const Slot& s) {
    const int ps = s.ps, K = s.K;
    const double* xs = smem_d + s.t_xs;
    double* Hb = smem_d + s.Hb;
    if (s.shared_r) {
        // constant dispersion: sum the digamma / trigamma differences over all groups at once (lanes over j)
        const double r = smem_d[s.rr];
        const double* Cj = smem_d + s.Cj;
        double D = 0.0, E = 0.0;
#pragma unroll 1
        for (int j = s.sl; j < s.ymax; j += SW) {
            const double cnt = Cj[j];
            if (cnt != 0.0) {
                const double inv = 1.0 / (r + (double)j);
                D += cnt * inv;
                E -= cnt * inv * inv;
            }
        }
        if (s.novf > 0) {
            const double psi_r = dx_digamma(r), psi1_r = dx_trigamma(r);
#pragma unroll 1
            for (int t = s.sl; t < s.novf_r; t += SW) {
                const float* rec = s.oval_r + t * s.ostep;
                const double y = (double)__ldg(rec), w = slot_ovf_weight(rec, s.ostep);
                D += w * (dx_digamma(r + y) - psi_r);
                E += w * (dx_trigamma(r + y) - psi1_r);
            }
        }
        double G = 0.0, H = 0.0;
#pragma unroll 1
        for (int k = s.sl; k < K; k += SW) {
            const double m = smem_d[s.mu + k], n = smem_d[s.t_nk + k], S = smem_d[s.S + k];
            const double irm = smem_d[s.irm + k];
            G += -n * smem_d[s.l1p + k] + (n * m - S) * irm;
            H += (n * m * m + r * S) * (irm * irm);
        }
        G = r * sw_sum<SW>(G + D, s.mask);
        H = G + r * r * sw_sum<SW>(E, s.mask) + r * sw_sum<SW>(H, s.mask);
        for (int e = s.sl; e < ps * ps; e += SW) {
            const int i = e / ps, j = e - i * ps;
            Hb[e] = -xs[i] * xs[j] * H;
        }
        __syncwarp(s.mask);
        return (s.sl < ps) ? xs[s.sl] * G : 0.0;
    }
    // general scale model: lane <-> group, serial over j (results stay with the lane)
    double* w1 = smem_d + s.w1;
    double* w2 = smem_d + s.w2;
#pragma unroll 1
    for (int k = s.sl; k < K; k += SW) {
        const double r = smem_d[s.rr + k];
        double D = 0.0, E = 0.0;
#pragma unroll 1
        for (int j = 0; j < s.ymax; ++j) {
            const unsigned cnt = __ldg(s.tail + k * DX_HIST_BINS + j);
            if (cnt) {
                const double inv = 1.0 / (r + (double)j);
                D += (double)cnt * inv;
                E -= (double)cnt * inv * inv;
            }
        }
        w1[k] = D;
        w2[k] = E;
    }
    __syncwarp(s.mask);
    if (s.novf > 0) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            if (!slot_group_has_ovf(s, k)) continue;
            const double r = smem_d[s.rr + k];
            const double psi_r = dx_digamma(r), psi1_r = dx_trigamma(r);
            double D = 0.0, E = 0.0;
#pragma unroll 1
            for (int t = s.sl; t < s.novf; t += SW) {
                if (__ldg(s.ogrp + t * s.ostep) == k) {
                    const float* rec = s.oval + t * s.ostep;
                    const double y = (double)__ldg(rec), w = slot_ovf_weight(rec, s.ostep);
                    D += w * (dx_digamma(r + y) - psi_r);
                    E += w * (dx_trigamma(r + y) - psi1_r);
                }
            }
            D = sw_sum<SW>(D, s.mask);
            E = sw_sum<SW>(E, s.mask);
            if (s.sl == 0) { w1[k] += D; w2[k] += E; }
        }
        __syncwarp(s.mask);
    }
#pragma unroll 1
    for (int k = s.sl; k < K; k += SW) {
        const double r = smem_d[s.rr + k], m = smem_d[s.mu + k], n = smem_d[s.t_nk + k], S = smem_d[s.S + k];
        const double irm = smem_d[s.irm + k];
        const double g = r * (w1[k] - n * smem_d[s.l1p + k] + (n * m - S) * irm);
        const double h = g + r * r * w2[k] + r * (n * m * m + r * S) * (irm * irm);
        w1[k] = g;
        w2[k] = h;
    }
    __syncwarp(s.mask);
#pragma unroll 1
    for (int e = s.sl; e < ps * ps; e += SW) {
        const int i = e / ps, j = e - i * ps;
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) acc += w2[k] * xs[k * ps + i] * xs[k * ps + j];
        Hb[e] = -acc;
    }
    double g = 0.0;
    if (s.sl < ps) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) g += w1[k] * xs[k * ps + s.sl];
    }
    __syncwarp(s.mask);
    return g;
}

// Newton direction (lane i returns s_i) from the gradient (lane registers) and s.Hb (= -Hessian): solve
// (Hb + lam I) s = g, lam grows until the guarded factorisation succeeds (oracle/nb_glm.py:_newton_direction);
// returns false on non-finite input
template <int SW> __device__ __noinline__ bool newton_direction(const Slot& s, double g, double& sb) {
    const int ps = s.ps;
    const double* Hb = smem_d + s.Hb;
    if (ps == 1) {
        const double h = Hb[0];
        const double g0 = sw_bcast<SW>(g, 0, s.mask);
        sb = 0.0;
        if (!(isfinite(h) && isfinite(g0))) return false;
        sb = (h > 0.0) ? g0 / h : (double)((g0 > 0.0) - (g0 < 0.0));
        return true;
    }
    bool fin = (s.sl < ps) ? isfinite(g) : true;
#pragma unroll 1
    for (int e = s.sl; e < ps * ps; e += SW) fin = fin && isfinite(Hb[e]);
    sb = 0.0;
    if (!sw_all<SW>(fin, s.mask)) return false;
    const double diag = (s.sl < ps) ? Hb[s.sl * ps + s.sl] : 0.0;
    const double base = fmax(sw_max<SW>(fabs(diag), s.mask), 1e-300);
    double* Hw = smem_d + s.Hw;
    double lam = 0.0;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
#pragma unroll 1
        for (int e = s.sl; e < ps * ps; e += SW) Hw[e] = Hb[e];
        __syncwarp(s.mask);
        if (s.sl < ps) Hw[s.sl * ps + s.sl] = diag + lam;
        __syncwarp(s.mask);
        double dinv;
        const bool ok = ldl_factor<SW>(Hw, ps, s.sl, s.mask, dinv);
        if (ok) {
            sb = ldl_solve<SW>(Hw, ps, s.sl, s.mask, dinv, (s.sl < ps) ? g : 0.0);
            return true;
        }
        lam = (lam == 0.0) ? base * 1e-6 : lam * 10.0;
    }
    sb = (s.sl < ps) ? (double)((g > 0.0) - (g < 0.0)) : 0.0;
    return true;
}

__device__ __forceinline__ void swap_i(int& a, int& b) { const int t = a; a = b; b = t; }

// safeguarded Newton in b with a fixed (oracle/nb_glm.py:_b_step_newton).  breg (lane i holds b_i) and the current
// b / (a, b) state are updated in place; returns ll at the final b.  f_cur = ll at entry.
template <int SW> __device__ __noinline__ double scale_newton(Slot& s, double& breg, double& zmax, double f_cur, int max_iter, double xtol) {
    if (zmax >= DX_ETA_SCALE_FROZEN) return f_cur;
    double f = f_cur;
#pragma unroll 1
    for (int it = 0; it < max_iter; ++it) {
        const double g = scale_terms<SW>(s);
        double sb;
        if (!newton_direction<SW>(s, g, sb)) break;
        const double smax = sw_max<SW>((s.sl < s.ps) ? fabs(sb) : 0.0, s.mask);
        const double sc = (smax > DX_NEWTON_MAX_STEP) ? DX_NEWTON_MAX_STEP / smax : 1.0;
        const double pred = sw_sum<SW>((s.sl < s.ps) ? g * (sc * sb) : 0.0, s.mask);
        // converged: the step is below xtol or its predicted gain is below the resolution of the likelihood
        if (sc * smax <= xtol || !(pred > DX_NEWTON_GAIN_RTOL * fabs(f))) break;
        double t = 1.0, f_try = f, bt = breg, T_t = 0.0, zmax_t = zmax;
        bool accepted = false;
#pragma unroll 1
        for (int h = 0; h < DX_NEWTON_MAX_HALVINGS; ++h) {
            bt = dx_clip(breg + t * (sc * sb), DX_ETA_MIN, DX_ETA_MAX);
            zmax_t = set_b<SW>(s, bt, s.zeta_t, s.rr_t, T_t);
            refresh<SW>(s, s.mu, s.rr_t, s.l1p_t, s.irm_t);
            f_try = ll_state<SW>(s, s.eta, s.zeta_t, s.rr_t, s.l1p_t, T_t);
            if (isfinite(f_try) && f_try >= f) { accepted = true; break; }
            t *= 0.5;
        }
        if (!accepted) break;
        breg = bt;
        swap_i(s.zeta, s.zeta_t); swap_i(s.rr, s.rr_t); swap_i(s.l1p, s.l1p_t); swap_i(s.irm, s.irm_t);
        s.T_r = T_t;
        zmax = zmax_t;
        f = f_try;
        if (zmax >= DX_ETA_SCALE_FROZEN) break;
    }
    return f;
}


template <int SW>
__global__ void __launch_bounds__(256)
dx_fit_grouped_kernel(const FitArgs P) {
    constexpr int SLOTS_PER_WARP = 32 / SW;
    const int K = P.K, p = P.p, ps = P.ps, nent = P.nent;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slots_per_cta = (blockDim.x >> 5) * SLOTS_PER_WARP;

    // ---- CTA tables: design rows, offsets, group sizes, product table, entry decode --------------------------
    const int t_xl = 0, t_xs = t_xl + K * p, t_off = t_xs + K * ps, t_nk = t_off + K, t_nzv = t_nk + K;
    const int t_dec = t_nzv + P.nz_max;                              // byte tables live here
    unsigned char* dec_i = reinterpret_cast<unsigned char*>(smem_d + t_dec);
    unsigned char* dec_j = dec_i + ((nent + 7) & ~7);
    unsigned short* nz_ptr = reinterpret_cast<unsigned short*>(dec_j + ((nent + 7) & ~7));
    unsigned char* nz_k = reinterpret_cast<unsigned char*>(nz_ptr + ((nent + 1 + 3) & ~3));
    __shared__ int s_nz_total;
    for (int i = tid; i < K * p; i += blockDim.x) smem_d[t_xl + i] = P.xu_loc[i];
    for (int i = tid; i < K * ps; i += blockDim.x) smem_d[t_xs + i] = P.xu_scale[i];
    for (int i = tid; i < K; i += blockDim.x) {
        smem_d[t_off + i] = P.offset ? P.offset[i] : 0.0;
        smem_d[t_nk + i] = P.n_k[i];
    }
    for (int e = tid; e < nent; e += blockDim.x) {
        int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        dec_i[e] = (unsigned char)i;
        dec_j[e] = (unsigned char)(e - i * (i + 1) / 2);
    }
    __syncthreads();
    // sparse table of the non-zero design products x_ki x_kj (count, exclusive scan, fill); too many -> dense loop
    for (int e = tid; e < nent; e += blockDim.x) {
        int c = 0;
        for (int k = 0; k < K; ++k) c += (smem_d[t_xl + k * p + dec_i[e]] * smem_d[t_xl + k * p + dec_j[e]] != 0.0) ? 1 : 0;
        nz_ptr[e] = (unsigned short)c;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int e = 0; e < nent; ++e) { const int c = nz_ptr[e]; nz_ptr[e] = (unsigned short)run; run += c; }
        nz_ptr[nent] = (unsigned short)run;
        s_nz_total = run;
    }
    __syncthreads();
    const bool has_nz = s_nz_total <= P.nz_max;
    if (has_nz) {
        for (int e = tid; e < nent; e += blockDim.x) {
            int t = nz_ptr[e];
            for (int k = 0; k < K; ++k) {
                const double v = smem_d[t_xl + k * p + dec_i[e]] * smem_d[t_xl + k * p + dec_j[e]];
                if (v != 0.0) { nz_k[t] = (unsigned char)k; smem_d[t_nzv + t] = v; ++t; }
            }
        }
    }
    __syncthreads();

    // persistent gene slots: every sub-warp pulls its next gene from the launch's queue, so a slot whose gene converged
    // early (6 .. 18 steps on the bench workload) does not idle until the slowest gene of its CTA is done
    const int slot = warp * SLOTS_PER_WARP + lane / SW;
    for (;;) {
    int64_t g;
    {
        unsigned q = 0;
        if (lane % SW == 0) q = atomicAdd(P.queue, 1u);
        g = (int64_t)__shfl_sync((SW == 32) ? 0xffffffffu : (((1u << SW) - 1u) << ((lane / SW) * SW)), q, (lane / SW) * SW);
    }
    if (g >= P.n_genes) break;
    // second launch after the lane-resident kernel: only the genes it left marked (niter = -1) are fitted here
    if (P.only_deferred && P.niter_out[g] >= 0) continue;

    Slot s;
    s.sl = lane % SW;
    s.mask = (SW == 32) ? 0xffffffffu : (((1u << SW) - 1u) << ((lane / SW) * SW));
    s.K = K; s.p = p; s.ps = ps; s.nent = nent;
    s.has_nz = has_nz;
    s.nz_ptr = nz_ptr; s.nz_k = nz_k;
    s.t_xl = t_xl; s.t_xs = t_xs; s.t_off = t_off; s.t_nk = t_nk; s.t_nzv = t_nzv;
    s.ei = dec_i; s.ej = dec_j;
    {
        int o = P.table_doubles + slot * P.slot_doubles;
        s.eta = o; o += K; s.mu = o; o += K; s.eta_t = o; o += K; s.mu_t = o; o += K;
        s.zeta = o; o += K; s.rr = o; o += K; s.zeta_t = o; o += K; s.rr_t = o; o += K;
        s.l1p = o; o += K; s.irm = o; o += K; s.l1p_t = o; o += K; s.irm_t = o; o += K;
        s.S = o; o += K;
        s.w1 = o; s.Ainv = o; o += K; s.w2 = o; o += K; s.Cj = o; o += DX_HIST_BINS + P.ainv_extra;   // Ainv aliases w1 | w2 | Cj
        s.A = o; o += p * p;
        s.vec = o; o += (p > ps ? p : ps);
        s.Hb = o; o += ps * ps; s.Hw = o; o += ps * ps;
    }
    const int sl = s.sl;
    const unsigned mask = s.mask;
    s.tail = P.tail_all + g * (int64_t)K * DX_HIST_BINS;
    s.novf = P.ovf_val ? P.ovf_count[g] : 0;          // (no list: the caller vouches that no gene has overflow entries)
    s.oval = P.ovf_val + P.indptr[g];
    s.ogrp = P.ovf_group + P.indptr[g];
    s.oval_r = s.oval; s.novf_r = s.novf; s.ostep = 1; s.ohead = nullptr;
    if (P.ovf_packed && s.novf > 0) {
        const int n_a = P.ovf_packed[2 * g], n_b = P.ovf_packed[2 * g + 1];
        if (n_a > 0) {       // layout written by dx_pack_overflow_kernel (dx_overflow.cu)
            s.ohead = s.oval;
            s.oval += 8 * K; s.ogrp += 8 * K;
            s.oval_r = s.oval + 2 * n_a;
            s.novf = n_a; s.novf_r = n_b; s.ostep = 2;
        }
    }
    {   // constant dispersion model <=> every group carries the same scale design row
        bool same = true;
#pragma unroll 1
        for (int k = sl; k < K; k += SW)
#pragma unroll 1
            for (int i = 0; i < ps; ++i) same = same && (smem_d[t_xs + k * ps + i] == smem_d[t_xs + i]);
        s.shared_r = sw_all<SW>(same, mask);
    }
    s.T_r = 0.0;

    // ---- sufficient statistics of the gene ---------------------------------------------------------
    double* Sk = smem_d + s.S;
    double* w1 = smem_d + s.w1;
    double* w2 = smem_d + s.w2;
    int ymax = 0;
    double lgc = 0.0;
    {
        constexpr int R = DX_HIST_BINS / SW;
        double cj[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cj[r] = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            // lanes over j: S_k = sum_j c_k(j), sum y^2 = sum_j c(j)(2j+1), sum dx_lgamma(y+1) = sum_j c(j) dx_log(j+1)
            double sk = 0.0, qk = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = sl + r * SW;
                const unsigned c = __ldg(s.tail + k * DX_HIST_BINS + j);
                const double cd = (double)c;
                sk += cd;
                qk += cd * (2.0 * j + 1.0);
                cj[r] += cd;
                if (c) ymax = max(ymax, j + 1);
            }
            sk = sw_sum<SW>(sk, mask);
            qk = sw_sum<SW>(qk, mask);
            if (sl == 0) { Sk[k] = sk; w2[k] = qk; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = sl + r * SW;
            smem_d[s.Cj + j] = cj[r];
            if (cj[r] != 0.0) lgc += cj[r] * dx_log((double)j + 1.0);
        }
    }
    s.ymax = sw_max_i<SW>(ymax, mask);
    s.ovf_groups[0] = s.ovf_groups[1] = s.ovf_groups[2] = s.ovf_groups[3] = 0ull;
    __syncwarp(mask);
    if (s.ohead) {
#pragma unroll 1
        for (int k = sl; k < K; k += SW) {
            const float* h = s.ohead + 8 * k;
            const double sy = __hiloint2double(__float_as_int(__ldg(h + 1)), __float_as_int(__ldg(h + 0)));
            const double sq = __hiloint2double(__float_as_int(__ldg(h + 3)), __float_as_int(__ldg(h + 2)));
            const double sg = __hiloint2double(__float_as_int(__ldg(h + 5)), __float_as_int(__ldg(h + 4)));
            const double cn = __hiloint2double(__float_as_int(__ldg(h + 7)), __float_as_int(__ldg(h + 6)));
            if (cn > 0.0) { Sk[k] += sy; w2[k] += sq; lgc += sg; }
        }
#pragma unroll 1
        for (int k = 0; k < K; ++k) {      // every lane needs the whole mask
            const float* h = s.ohead + 8 * k;
            const double cn = __hiloint2double(__float_as_int(__ldg(h + 7)), __float_as_int(__ldg(h + 6)));
            if (cn > 0.0) s.ovf_groups[k >> 6] |= (1ull << (k & 63));
        }
        __syncwarp(mask);
    } else if (s.novf > 0) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            double sk = 0.0, qk = 0.0;
            bool any = false;
#pragma unroll 1
            for (int t = sl; t < s.novf; t += SW) {
                if (__ldg(s.ogrp + t) == k) {
                    const double y = (double)__ldg(s.oval + t);
                    sk += y; qk += y * y; lgc += dx_lgamma(y + 1.0); any = true;
                }
            }
            sk = sw_sum<SW>(sk, mask); qk = sw_sum<SW>(qk, mask);
            if (__ballot_sync(mask, any) & mask) {
                s.ovf_groups[k >> 6] |= (1ull << (k & 63));
                if (sl == 0) { Sk[k] += sk; w2[k] += qk; }
            }
        }
        __syncwarp(mask);
    }
    s.lgam_const = sw_sum<SW>(lgc, mask);
    double n_total = 0.0, sum_all = 0.0, m_sf = 0.0, e2_sf = 0.0;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const double isf = P.offset ? dx_exp(-smem_d[t_off + k]) : 1.0;
        n_total += smem_d[t_nk + k];
        sum_all += Sk[k];
        m_sf += Sk[k] * isf;
        e2_sf += w2[k] * isf * isf;
    }
    s.n_total = n_total;
    int flags = 0;
    if (sum_all == 0.0 && s.novf == 0) flags |= DX_FLAG_ALL_ZERO;
    __syncwarp(mask);

    // ---- initialize (batchglm init_par); lane i holds a_i / b_i ------------------------------------
    const dx_fit_opts& opts = P.opts;
    double areg = 0.0, breg = 0.0;
    if (opts.init_a == DX_INIT_STANDARD) {
        areg = (sl == 0) ? dx_clip(dx_log(sum_all / n_total), DX_ETA_MIN, DX_ETA_MAX) : 0.0;
    } else if (opts.init_a == DX_INIT_CLOSED_FORM) {
        // batchglm closedform_glm_mean: log of the mean of x / size_factor over every distinct LOCATION design
        // row, pushed through the pseudo-inverse of those rows (np.linalg.lstsq in the reference)
        const int KL = P.KL;
#pragma unroll 1
        for (int l = sl; l < KL; l += SW) {
            double num = 0.0, den = 0.0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                if ((P.loc_group ? P.loc_group[k] : k) == l) {
                    num += Sk[k] * (P.offset ? dx_exp(-smem_d[t_off + k]) : 1.0);
                    den += smem_d[t_nk + k];
                }
            }
            w1[l] = dx_log(num / den + DX_TINY);
        }
        __syncwarp(mask);
        if (sl < p) {
            double acc = 0.0;
#pragma unroll 1
            for (int l = 0; l < KL; ++l) acc += P.pinv_loc[sl * KL + l] * w1[l];
            areg = dx_clip(acc, DX_ETA_MIN, DX_ETA_MAX);
        }
        __syncwarp(mask);
    } else if (opts.init_a == DX_INIT_GIVEN) {
        if (sl < p) areg = dx_clip(P.a_var[(int64_t)sl * P.stride + g], DX_ETA_MIN, DX_ETA_MAX);
    }
    if (opts.init_b == DX_INIT_STANDARD) {
        const double m = m_sf / n_total;
        const double v = e2_sf / n_total - m * m;
        const double denom = fmax(v - m, 2.2227587494850775e-162);
        breg = (sl == 0) ? dx_clip(dx_log(m * m / denom + DX_TINY), DX_ETA_MIN, DX_ETA_MAX) : 0.0;
    } else if (opts.init_b == DX_INIT_GIVEN) {
        if (sl < ps) breg = dx_clip(P.b_var[(int64_t)sl * P.stride + g], DX_ETA_MIN, DX_ETA_MAX);
    }

    // ---- train (batchglm numpy EstimatorGlm.train, per-gene form) -----------------------------------
    const bool train_loc = opts.train_loc != 0, train_scale = opts.train_scale != 0;
    const int freq = train_scale ? (train_loc ? opts.update_b_freq : 1) : 0x7fffffff;
    int epochs_until_b = freq;
    bool loc_conv = false, fully = false;
    double zmax = set_b<SW>(s, breg, s.zeta, s.rr, s.T_r);
    set_a<SW>(s, areg, s.eta, s.mu);
    refresh<SW>(s, s.mu, s.rr, s.l1p, s.irm);
    double nll = -ll_state<SW>(s, s.eta, s.zeta, s.rr, s.l1p, s.T_r);
    int step = 0;
#pragma unroll 1
    while (!fully && step < opts.max_steps) {
        double nll_new = nll;
        bool b_updated = false;
        if (epochs_until_b == 0) {
            if (train_scale) {
                const double bold = breg;
                const double f = scale_newton<SW>(s, breg, zmax, -nll, opts.newton_max_iter, opts.newton_xtol);
                const double nll_prop = -f;
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    nll_new = nll_prop;
                } else {
                    breg = bold;
                    zmax = set_b<SW>(s, breg, s.zeta, s.rr, s.T_r);
                    refresh<SW>(s, s.mu, s.rr, s.l1p, s.irm);
                }
            }
            epochs_until_b = freq;
            b_updated = true;
        } else {
            if (train_loc && !loc_conv) {
                const double rhs = loc_terms<SW>(s);
                double dinv;
                const bool ok = ldl_factor<SW>(smem_d + s.A, p, sl, mask, dinv);
                const bool fin = sw_all<SW>((sl < p) ? isfinite(rhs) : true, mask);
                double atry = areg;
                if (ok && fin) {
                    const double delta = ldl_solve<SW>(smem_d + s.A, p, sl, mask, dinv, (sl < p) ? rhs : 0.0);
                    atry = dx_clip(areg + delta, DX_ETA_MIN, DX_ETA_MAX);
                } else {
                    flags |= DX_FLAG_SINGULAR_FIM;
                }
                set_a<SW>(s, atry, s.eta_t, s.mu_t);
                refresh<SW>(s, s.mu_t, s.rr, s.l1p_t, s.irm_t);
                const double nll_prop = -ll_state<SW>(s, s.eta_t, s.zeta, s.rr, s.l1p_t, s.T_r);
                if (!isfinite(nll_prop)) flags |= DX_FLAG_NONFINITE;
                if (isfinite(nll_prop) && !(nll_prop > nll)) {
                    areg = atry;
                    swap_i(s.eta, s.eta_t); swap_i(s.mu, s.mu_t); swap_i(s.l1p, s.l1p_t); swap_i(s.irm, s.irm_t);
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

    // ---- finalize (batchglm Estimator.finalize) -----------------------------------------------------
    if (zmax >= DX_ETA_SCALE_FROZEN) flags |= DX_FLAG_SCALE_FROZEN;
    const double rhs_f = loc_terms<SW>(s);
    const double gb_f = scale_terms<SW>(s);                 // (may use w1 / w2: X^T W X is already in s.A)
    double jac = ((sl < p) ? fabs(rhs_f) : 0.0) + ((sl < ps) ? fabs(gb_f) : 0.0);
    jac = sw_sum<SW>(jac, mask) / n_total;
    double dinv_f;
    const bool okf = ldl_factor<SW>(smem_d + s.A, p, sl, mask, dinv_f);
    double* Ainv = smem_d + s.Ainv;                         // aliases w1 | w2 | Cj: none of them is needed any more
    {   // inverse: lane c solves (L D L^T) x = e_c serially into its own column (all columns in parallel)
        double* dv = smem_d + s.vec;
        if (sl < p) dv[sl] = dinv_f;
        __syncwarp(mask);
        if (okf && sl < p) {
            const double* L = smem_d + s.A;                 // L_ik (i > k) at L[k * p + i]
            double* x = Ainv + sl * p;
#pragma unroll 1
            for (int i = 0; i < p; ++i) {
                double acc = (i == sl) ? 1.0 : 0.0;
#pragma unroll 1
                for (int k = 0; k < i; ++k) acc -= L[k * p + i] * x[k];
                x[i] = acc;
            }
#pragma unroll 1
            for (int i = p - 1; i >= 0; --i) {
                double acc = x[i] * dv[i];
#pragma unroll 1
                for (int k = i + 1; k < p; ++k) acc -= L[i * p + k] * x[k];
                x[i] = acc;
            }
        }
    }
    __syncwarp(mask);
    double* fo = P.fim_inv + g * (int64_t)p * p;
#pragma unroll 1
    for (int e = sl; e < p * p; e += SW) {
        const int i = e / p, j = e - i * p;
        fo[e] = okf ? 0.5 * (Ainv[i * p + j] + Ainv[j * p + i]) : NAN;
    }
    if (sl < p) P.a_var[(int64_t)sl * P.stride + g] = areg;
    if (sl < ps) P.b_var[(int64_t)sl * P.stride + g] = breg;
    if (sl == 0) {
        P.ll_out[g] = -nll;
        P.jac_out[g] = jac;
        P.niter_out[g] = step;
        P.flags_out[g] = flags;
    }
    if (P.epi.wald_coef >= 0 && sl == P.epi.wald_coef)
        dx_fit_epilogue_row(P.epi, g, areg, okf ? Ainv[sl * p + sl] : NAN);
    __syncwarp(mask);
    }
    dx_fit_epilogue_publish(P.epi);
}

template <int SW>
cudaError_t launch_sw(FitArgs& A, int sm_count, cudaStream_t stream) {
    const int slots_per_warp = 32 / SW;
    const size_t table_bytes = (size_t)A.table_doubles * 8, slot_bytes = (size_t)A.slot_doubles * 8;
    // warps per CTA: maximise the number of resident gene slots per SM under the shared-memory budget
    const size_t budget = 228 * 1024;      // per SM; every resident CTA also pays 1 KB of system-reserved shared memory
    int best_w = 1;
    long best_res = -1;
    const int cand[] = {1, 2, 3, 4, 6, 8};
    for (int w : cand) {
        const size_t need = table_bytes + (size_t)w * slots_per_warp * slot_bytes + 1024 + 64;
        if (need > 227 * 1024) continue;
        long ctas = (long)(budget / need);
        if (ctas > 32) ctas = 32;
        long warps = ctas * w;
        if (warps > 40) warps = 40;          // beyond ~40 warps the register file is the limit anyway
        const long res = warps * 1000 - w;   // ties -> fewer warps per CTA (finer-grained scheduling)
        if (res > best_res) { best_res = res; best_w = w; }
    }
    if (best_res < 0) return cudaErrorInvalidConfiguration;
    static const char* w_env = getenv("DXB200_FIT_WARPS");       // tuning hook (development only)
    if (w_env && atoi(w_env) >= 1 && atoi(w_env) <= 8 &&
        table_bytes + (size_t)atoi(w_env) * slots_per_warp * slot_bytes <= budget) best_w = atoi(w_env);
    const int slots_per_cta = best_w * slots_per_warp;
    const size_t smem = table_bytes + (size_t)slots_per_cta * slot_bytes;
    cudaError_t e = cudaFuncSetAttribute(dx_fit_grouped_kernel<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dx_fit_grouped_kernel<SW>, best_w * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (sm_count <= 0) {
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    int64_t grid = (A.n_genes + slots_per_cta - 1) / slots_per_cta;        // never more CTAs than gene slots needed
    if (grid > (int64_t)per_sm * sm_count) grid = (int64_t)per_sm * sm_count;
    if (grid < 1) grid = 1;                // an empty shard still publishes its exchange flag
    e = dx_queue_acquire(stream, &A.queue);
    if (e != cudaSuccess) return e;
    if (A.epi.world > 1) {
        e = dx_queue_acquire(stream, &A.epi.done_ctas);
        if (e != cudaSuccess) return e;
    }
    dx_fit_grouped_kernel<SW><<<(unsigned)grid, best_w * 32, smem, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace

cudaError_t dx_launch_fit_grouped(int64_t n_genes, int K, int p, int ps,
                                  const double* xu_loc, const double* xu_scale, const double* offset,
                                  const double* n_k, const double* pinv_loc, const int32_t* loc_group, int n_loc_groups,
                                  const uint32_t* tail, const int32_t* ovf_count, const int64_t* indptr,
                                  const float* ovf_val, const uint8_t* ovf_group, const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, int64_t coef_stride, const int32_t* ovf_packed,
                                  const DxFitEpilogueHost* epi, cudaStream_t stream) {
    FitArgs A;
    A.epi.wald_coef = -1;
    A.epi.theta = A.epi.sd = A.epi.pval = A.epi.logp = nullptr;
    A.epi.world = 1; A.epi.rank = 0; A.epi.publish = 1; A.epi.gene_offset = 0; A.epi.slot_bytes = 0; A.epi.done_ctas = nullptr;
    for (int r = 0; r < DX_EX_MAX_WORLD; ++r) A.epi.peer[r] = nullptr;
    if (epi) {
        A.epi.wald_coef = epi->wald_coef;
        A.epi.theta = epi->theta; A.epi.sd = epi->sd; A.epi.pval = epi->pval; A.epi.logp = epi->logp;
        if (epi->world > 1) {
            if (epi->world > DX_EX_MAX_WORLD || epi->rank < 0 || epi->rank >= epi->world || !epi->peer_bufs) return cudaErrorInvalidValue;
            A.epi.world = epi->world; A.epi.rank = epi->rank;
            A.epi.gene_offset = epi->gene_offset; A.epi.slot_bytes = epi->slot_bytes;
            for (int r = 0; r < epi->world; ++r) A.epi.peer[r] = static_cast<unsigned char*>(epi->peer_bufs[r]);
        }
    }
    if (n_genes <= 0 && A.epi.world <= 1) return cudaSuccess;
    A.ovf_packed = ovf_packed;
    A.stride = coef_stride > 0 ? coef_stride : n_genes;
    A.n_genes = n_genes; A.K = K; A.p = p; A.ps = ps; A.KL = n_loc_groups;
    A.xu_loc = xu_loc; A.xu_scale = xu_scale; A.offset = offset; A.n_k = n_k; A.pinv_loc = pinv_loc; A.loc_group = loc_group;
    A.tail_all = tail; A.ovf_count = ovf_count; A.indptr = indptr; A.ovf_val = ovf_val; A.ovf_group = ovf_group;
    A.opts = *opts;
    A.a_var = a_var; A.b_var = b_var; A.fim_inv = fim_inv; A.ll_out = ll; A.jac_out = jac; A.niter_out = niter; A.flags_out = flags;
    A.nent = p * (p + 1) / 2;
    const char* impl_env = getenv("DXB200_FIT_IMPL");      // test hook: "g" forces the general kernel (read at every launch)
    A.only_deferred = 0;
    if (!(impl_env && impl_env[0] == 'g') && dx_fit_lane_eligible(K, p, ps, opts) && (reinterpret_cast<uintptr_t>(tail) & 15u) == 0) {
        // lane-resident kernel (it stores its p-values but leaves the exchange flag alone), then the general kernel for the
        // genes it left marked (long overflow lists), which publishes for the fit as a whole
        if (!ovf_val) return dx_launch_fit_lane(A, 0, stream);      // no overflow lists at all: nothing can be left over
        FitArgs B = A;
        B.epi.publish = 0;
        cudaError_t e = dx_launch_fit_lane(B, 0, stream);
        if (e != cudaSuccess) return e;
        A.only_deferred = 1;
    }
    static const char* nz_env = getenv("DXB200_FIT_NZ");         // tuning hook (development only)
    A.nz_max = nz_env ? atoi(nz_env) : 96;   // capacity of the non-zero design-product list (0.9 KB); denser designs use the dense loop
    A.table_doubles = K * p + K * ps + 2 * K + A.nz_max + (2 * ((A.nent + 7) & ~7) + 2 * ((A.nent + 1 + 3) & ~3) + A.nz_max + 7) / 8 + 2;
    A.ainv_extra = (p * p > 2 * K + DX_HIST_BINS) ? p * p - (2 * K + DX_HIST_BINS) : 0;
    A.slot_doubles = 13 * K + 2 * K + DX_HIST_BINS + A.ainv_extra + p * p + (p > ps ? p : ps) + 2 * ps * ps;
    A.slot_doubles = (A.slot_doubles + 1) & ~1;
    A.table_doubles = (A.table_doubles + 1) & ~1;
    const char* sw_env = getenv("DXB200_FIT_SW");          // test hook: sub-warp width of the general kernel (16 or 32)
    const int sw_min = sw_env ? atoi(sw_env) : 0;
    // the sub-warp width follows the coefficient counts (lane <-> matrix row in the factorisations); design groups beyond
    // the width are strided over by the lanes
    if (K <= 8 && p <= 8 && ps <= 8 && sw_min <= 8) return launch_sw<8>(A, 0, stream);
    if (p <= 16 && ps <= 16 && sw_min <= 16) return launch_sw<16>(A, 0, stream);
    return launch_sw<32>(A, 0, stream);
}
