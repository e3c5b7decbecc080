// dx_fit_tile.cu -- K_fit (per-cell designs) as a gene-tiled fp64 contraction.
//
// The batchglm training schedule (/root/reference/diffxpy/testing/tests.py:222-248 -> Estimator.initialize /
// train_sequence("DEFAULT") / finalize; restated in oracle/nb_glm.py, method_b="newton") for designs whose rows do NOT
// collapse into a few groups: continuous covariates, spline bases, per-cell size factors.  Every evaluation of the model
// touches every cell of every gene, so the path is bound by fp64 arithmetic, not by HBM:
//   eta_ig = x_i . a_g + log s_i,  mu = exp(eta),  r = exp(z_i . b_g),  W = mu r / (r + mu),  S = (y - mu) r / (r + mu)
//   X^T W X [g][(j,k)] = sum_i W_ig  x_ij x_ik        X^T S [g][j] = sum_i S_ig x_ij
// The products Z_i,(j,k) = x_ij x_ik do not depend on the gene.  A CTA therefore owns a TILE of GT genes and walks the
// cells in tiles of CT:
//   phase 0  the design rows of the cell tile are staged in shared memory (prefetched one tile ahead into registers), the
//            product table Z[CT][32 NE] is built ONCE for all genes of the tile, the genes' sparse columns are densified
//   phase 1  a thread evaluates a 2-cell x 4-gene block of (eta, exp, log1p, likelihood term, W, S): eight independent
//            dependency chains per thread, written to shared memory as W[CT][GT], S[CT][GT]
//   phase 2  C[g][e] += W[c][g] Z[c][e] as a register-blocked contraction: a warp owns MG genes and a slice of the cells,
//            lane l owns the entries l, l + 32, ...: MG x NE accumulators per lane, weights are warp-uniform broadcast
//            loads, Z loads are conflict free.  The score splits as X^T S = sum_i y_i q_i x_i - sum_i W_i x_i: the second sum
//            rides in the unused tail of the last entry slot (entries x_j with the same weights W), the first one runs
//            over the NON-ZERO cells only (a short loop of the warp that densified the gene's column)
// The genes of a tile run the schedule in lockstep (batchglm's epochs are global: every update_b_freq-th step is the
// dispersion step for all genes); each gene keeps its own convergence state and query point, and a pass evaluates every
// active gene at its own point.  The Newton iterations of the dispersion step are a per-gene state machine over passes that
// return likelihood AND derivatives at the query point, so an accepted line-search point needs no second pass.
// Results are deterministic: no atomics in any sum, fixed reduction orders.
// CPU twin: oracle/nb_glm.py (fit_nb, method_b="newton").
#include <cstdlib>
#include <type_traits>
#include "dx_common.cuh"
#include "dx_internal.h"
#include "dx_fastmath.cuh"

namespace {

constexpr int NT = 256;            // threads per CTA
constexpr int NWP = NT / 32;
constexpr int TB = 64;             // per-gene tables of the count-dependent terms cover y = 0 .. TB - 1

enum { SC_NLL = 0, SC_NLLNEW, SC_F, SC_T, SC_LGAM, SC_R, SC_INVR, SC_LL, SC_G, SC_H, SC_ZMAX, SC_COUNT };
enum { IC_FLAGS = 0, IC_LOCCONV, IC_FULLY, IC_HAVE, IC_DOK, IC_NSTATE, IC_NIT, IC_NHALV, IC_MOVED, IC_NITER, IC_COUNT };
enum { NS_DONE = 0, NS_DERIV = 1, NS_LINE = 2 };
enum { PASS_LOC = 0, PASS_SCALE = 1 };

// diagnostics: passes over the cells and gene tiles since the last reset (dx_fit_percell_counters)
__device__ unsigned long long g_tile_counters[4];          // location passes, dispersion passes, dispersion epochs, gene tiles

struct TileArgs {
    int64_t n_genes, N;
    int p, ps;
    const double *dl, *ds, *log_sf;
    const int64_t* indptr;
    const int32_t* rows;
    const float* vals;
    dx_fit_opts opts;
    double *a_var, *b_var, *fim_inv, *ll, *jac;
    int32_t *niter, *flags;
    unsigned int* queue;
};

struct Lay {          // shared-memory layout (offsets in doubles)
    int a, qa, delta, rhs;          // [p][GT]
    int b, bt, qb, sb, gb;          // [ps][GT]
    int hg;                         // [GT][ps * ps]  -(Hessian) of the dispersion step (general scale model)
    int sc, ic, ent;                // per-gene scalars, ints, entry ranges
    int gt, lf;                     // [GT][TB] table G(r, y) (constant dispersion model); log(y!) for y < TB
    int t0, t1;                     // [GT][TB] psi differences: inside the region (the dispersion pass has no product table)
    int rn;                         // [p][GT] sum over the non-zero cells of y q x_j
    int pair;                       // uint8 entry -> (j, k) tables
    int tile;                       // start of the region shared by the passes and the small dense algebra between them
    int Z, W, S, X, D, Ls, Y, YI, NZ;   // inside the region; X / D / Ls / Y / YI / NZ exist twice (tile parity): second copy at
    int stg, xstg;                  // + stg (Ls, Y, YI, NZ) resp. + xstg (X, D)
    int scratch;                    // inside the region, behind the reduced pass result
    int total_bytes;
};

template <int NE, int MG, int GT, int CT, bool SCONST>
__host__ __device__ inline Lay make_layout(int p, int ps) {
    constexpr int ZW = 32 * NE, ZWS = SCONST ? 0 : 96, ZWM = (ZW > ZWS) ? ZW : ZWS;
    constexpr int WST = GT + 2, XST = CT + 1, KS = NWP / (GT / MG);
    Lay L;
    // everything whose size does not depend on (p, ps) comes first: those offsets are compile-time constants in the kernel
    int o = 0;
    L.sc = o; o += SC_COUNT * GT;
    L.ic = o; o += (IC_COUNT * GT + 1) / 2;
    L.ent = o; o += 2 * GT;
    L.gt = o; o += SCONST ? GT * TB : 0;
    L.lf = o; o += TB;
    L.pair = o; o += (2 * ZW + 2 * 96 + 7) / 8;
    o = (o + 1) & ~1;
    L.tile = o;
    int t = 0;
    L.Z = t; L.t0 = t; L.t1 = t + GT * TB;
    t += (SCONST && CT * ZWM < 2 * GT * TB) ? 2 * GT * TB : CT * ZWM;
    L.W = t; t += CT * WST + 8;
    L.S = t; t += CT * WST;
    L.Ls = t; t += CT;
    L.Y = t; t += GT * CT / 2;
    L.YI = t; t += GT * CT / 8;
    L.NZ = t; t += GT * (CT + 8) / 8;
    t = (t + 1) & ~1;
    L.stg = t - L.Ls;
    t += L.stg;
    const int pm = (p > ps) ? p : ps;
    L.scratch = GT * ZWM;
    int need = L.scratch + NWP * (pm * pm + 2 * pm);          // extraction / solves next to the reduced pass result
    if (need < NWP * 2 * pm * pm) need = NWP * 2 * pm * pm;   // finalize: factor and inverse per warp
    if (need < KS * GT * ZWM) need = KS * GT * ZWM;           // partial sums of the cell slices
    if (t < need) t = need;
    o = L.tile + t;
    // sizes that depend on the design width (X, D relative to the region like the other pass arrays)
    L.X = o - L.tile; o += p * XST; o = (o + 1) & ~1;
    L.D = o - L.tile; o += SCONST ? 0 : ps * XST; o = (o + 1) & ~1;
    L.xstg = o - L.tile - L.X;
    o += L.xstg;
    L.a = o; o += p * GT; L.qa = o; o += p * GT; L.delta = o; o += p * GT; L.rhs = o; o += p * GT; L.rn = o; o += p * GT;
    L.b = o; o += ps * GT; L.bt = o; o += ps * GT; L.qb = o; o += ps * GT; L.sb = o; o += ps * GT; L.gb = o; o += ps * GT;
    L.hg = o; o += SCONST ? 0 : GT * ps * ps;
    L.total_bytes = o * 8;
    return L;
}

// ---- count-dependent terms outside the tables (y >= TB, non-integer values, per-cell dispersion) ----------------------
__device__ __noinline__ double g_ratio_slow(double r, double y) {
    const int yi = (int)y;
    if ((double)yi == y && yi >= 1 && yi <= 12) {
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
__device__ __noinline__ void psi_diffs_slow(double r, double y, double& d0, double& d1) {
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

// guarded Cholesky in place (Lm holds the full symmetric matrix, n <= 32): the rule of chol_warp (dx_common.cuh)
__device__ __noinline__ bool chol_inplace(double* Lm, int n, int lane) {
    const double dorig = (lane < n) ? Lm[lane * n + lane] : 0.0;
    __syncwarp();
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        const double v = Lm[j * n + j];
        const double mjj = __shfl_sync(DX_FULL_MASK, dorig, j);
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

// Newton direction of the dispersion step for one gene (rule of oracle _newton_direction): H (= -Hessian, ps x ps,
// overwritten), g -> s; one warp.  false: non-finite input.
__device__ bool newton_direction(double* H, double* Lb, const double* g, double* s, int ps, int lane) {
    if (ps == 1) {
        const double h = H[0], gg = g[0];
        if (!(isfinite(h) && isfinite(gg))) return false;
        if (lane == 0) s[0] = (h > 0.0) ? gg / h : (double)((gg > 0.0) - (gg < 0.0));
        __syncwarp();
        return true;
    }
    bool fin = true;
    for (int e = lane; e < ps * ps; e += 32) fin = fin && isfinite(H[e]);
    for (int i = lane; i < ps; i += 32) fin = fin && isfinite(g[i]);
    if (!__all_sync(DX_FULL_MASK, fin)) return false;
    const double diag = (lane < ps) ? H[lane * ps + lane] : 0.0;
    const double base = fmax(dx_warp_max(fabs(diag)), 1e-300);
    double lam = 0.0;
    for (int it = 0; it < 12; ++it) {
        for (int e = lane; e < ps * ps; e += 32) Lb[e] = H[e];
        __syncwarp();
        if (lane < ps) Lb[lane * ps + lane] = diag + lam;
        __syncwarp();
        const bool ok = chol_inplace(Lb, ps, lane);
        __syncwarp();
        if (ok) {
            if (lane == 0) chol_solve_serial(Lb, ps, g, s);
            __syncwarp();
            return true;
        }
        lam = (lam == 0.0) ? base * 1e-6 : lam * 10.0;
    }
    for (int i = lane; i < ps; i += 32) { const double gg = g[i]; s[i] = (double)((gg > 0.0) - (gg < 0.0)); }
    __syncwarp();
    return true;
}

template <int NE, int MG, int GT, int CT, int PMAXC, int MINB, bool SCONST>
__global__ void __launch_bounds__(NT, MINB) dx_fit_tile_kernel(const TileArgs A) {
    constexpr int ZW = 32 * NE, WST = GT + 2, XST = CT + 1;
    constexpr int NGH = GT / MG, KS = NWP / NGH, GP1 = GT / NWP, CP1 = CT / 32;
    constexpr int XPT = (PMAXC * CT + NT - 1) / NT, DPT = SCONST ? 1 : (8 * CT + NT - 1) / NT;
    static_assert(NGH * KS == NWP && GP1 * NWP == GT && CP1 * 32 == CT && (GP1 % 2) == 0 && (MG % 2) == 0, "tile shape");
    extern __shared__ __align__(16) double sm[];
    const int p = A.p, ps = A.ps;
    const int64_t N = A.N;
    const Lay L = make_layout<NE, MG, GT, CT, SCONST>(p, ps);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = p * (p + 1) / 2, Es = ps * (ps + 1) / 2;
    double* const a_s = sm + L.a; double* const qa_s = sm + L.qa; double* const delta_s = sm + L.delta; double* const rhs_s = sm + L.rhs;
    double* const b_s = sm + L.b; double* const bt_s = sm + L.bt; double* const qb_s = sm + L.qb; double* const sb_s = sm + L.sb;
    double* const gb_s = sm + L.gb;
    double* const sc = sm + L.sc;
    int* const ic = reinterpret_cast<int*>(sm + L.ic);
    int64_t* const ent = reinterpret_cast<int64_t*>(sm + L.ent);
    double* const Gt = sm + L.gt; double* const T0 = sm + L.tile + L.t0; double* const T1 = sm + L.tile + L.t1;
    double* const lfact = sm + L.lf;
    double* const rn_s = sm + L.rn;
    uint8_t* const pa = reinterpret_cast<uint8_t*>(sm + L.pair);
    uint8_t* const pb = pa + ZW; uint8_t* const pas = pb + ZW; uint8_t* const pbs = pas + 96;
    double* const tile = sm + L.tile;
    double* const Zs = tile + L.Z; double* const Ws = tile + L.W; double* const Ss = tile + L.S;
    // staging buffers of a cell tile, twice (tile parity): design rows, offsets, dense values of the genes' columns, the
    // table index of every value (255: outside the tables), the list of the non-zero cells per gene
    double* const Xs0 = tile + L.X; double* const Ds0 = tile + L.D; double* const Lsf0 = tile + L.Ls;
    float* const Ys0 = reinterpret_cast<float*>(tile + L.Y);
    uint8_t* const Yi0 = reinterpret_cast<uint8_t*>(tile + L.YI);
    uint8_t* const NZl0 = reinterpret_cast<uint8_t*>(tile + L.NZ);
    double* const Cred = tile;
    double* const scratch = tile + L.scratch;
    __shared__ int s_tile;
    __shared__ unsigned s_mask, s_acc;
#define SCG(k, g) sc[(k) * GT + (g)]
#define ICG(k, g) ic[(k) * GT + (g)]

    // entry -> (j, k) tables.  Location pass: products x_j x_k first, the plain columns x_j in the tail of the last slot, zeros
    // between.  Dispersion pass (general scale model): products z_j z_k in slots 0-1, plain columns z_j in slot 2.
    for (int e = tid; e < ZW; e += NT) {
        int j = 255, k = 255;
        if (e < E) { int i = 0; while ((i + 1) * (i + 2) / 2 <= e) ++i; j = i; k = e - i * (i + 1) / 2; }
        else if (e >= ZW - p) { j = e - (ZW - p); k = 254; }
        pa[e] = (uint8_t)j; pb[e] = (uint8_t)k;
    }
    for (int e = tid; e < 96; e += NT) {
        int j = 255, k = 255;
        if (e < Es) { int i = 0; while ((i + 1) * (i + 2) / 2 <= e) ++i; j = i; k = e - i * (i + 1) / 2; }
        else if (e >= 64 && e < 64 + ps) { j = e - 64; k = 254; }
        pas[e] = (uint8_t)j; pbs[e] = (uint8_t)k;
    }
    for (int j = tid; j < TB; j += NT) lfact[j] = lgamma((double)j + 1.0);
    const bool train_loc = A.opts.train_loc != 0, train_scale = A.opts.train_scale != 0;
    const int freq = train_scale ? (train_loc ? A.opts.update_b_freq : 1) : 0x7fffffff;
    const double z_const = SCONST ? __ldg(A.ds) : 0.0;

    // ------------------------------------------------------------------------------------------------------------------
    // One pass over all cells: every gene of `mask` is evaluated at (qa, qb).
    //   PASS_LOC    SC_LL = sum of the likelihood terms; Cred[g][.] = X^T W X (packed) and sum_i W_i x_i (tail of the last
    //               slot); rn[j][g] = sum over the non-zero cells of y q x_j
    //   PASS_SCALE  SC_LL, and the derivatives of the likelihood in b: constant dispersion -> SC_G, SC_H (second derivative);
    //               general -> Cred[g][.] = sum h z z^T (packed, slots 0-1) and sum g z (slot 2); SC_ZMAX = max_i clip(z_i . b)
    // ------------------------------------------------------------------------------------------------------------------
    auto run_pass = [&](auto kind_tag, const unsigned mask) {
        constexpr int KIND = decltype(kind_tag)::value;
        constexpr bool GEMM = (KIND == PASS_LOC) || !SCONST;
        constexpr int NEK = (KIND == PASS_LOC) ? NE : 3;          // entry slots of the contraction
        constexpr int NEW = (KIND == PASS_LOC) ? NE : 2;          // slots [0, NEW) are weighted by Ws, the rest by Ss
        constexpr int ZWK = 32 * NEK;
        if (tid == 0) atomicAdd(&g_tile_counters[KIND], 1ull);
        __syncthreads();
        // per-gene constants of the pass
        if (SCONST) {
            if (tid < GT) {
                const double zeta = dx_clip(z_const * qb_s[tid], DX_ETA_MIN, DX_ETA_MAX);
                const double r = exp(zeta);
                const double invr = 1.0 / r;
                SCG(SC_R, tid) = r; SCG(SC_INVR, tid) = invr; SCG(SC_ZMAX, tid) = zeta;
                double* gt = Gt + tid * TB; double* t0 = T0 + tid * TB; double* t1 = T1 + tid * TB;
                if ((mask >> tid) & 1u) {
                    double sg = 0.0, s0 = 0.0, s1 = 0.0;
                    gt[0] = 0.0; t0[0] = 0.0; t1[0] = 0.0;
#pragma unroll 1
                    for (int j = 0; j < TB - 1; ++j) {        // G(r, y) = sum_{j<y} log1p(j / r); psi differences likewise
                        sg += fl_log1p((double)j * invr);
                        gt[j + 1] = sg;
                        if (KIND == PASS_SCALE) {
                            const double iv = 1.0 / (r + (double)j);
                            s0 += iv; s1 -= iv * iv;
                            t0[j + 1] = s0; t1[j + 1] = s1;
                        }
                    }
                }
            }
        }
        const int gh = warp % NGH, ks = warp / NGH;
        const unsigned grp_mask = (MG >= 32) ? mask : ((mask >> (gh * MG)) & ((1u << MG) - 1u));
        // Which genes this warp evaluates in phase 1 (slot u <-> gene gu[u]).  Location passes: its own GP1 genes.  Dispersion
        // passes of the constant dispersion model (no contraction): the genes of the pass are dealt out evenly, ceil(n / 8) to a
        // warp -- on average half the genes of a tile sit out a Newton pass of the lockstep schedule (they are done with their
        // line search), and a pass costs as much as its busiest warp.
        constexpr bool COMPACT = (KIND == PASS_SCALE) && SCONST;
        int gu[GP1];
        int ng_w = 0;
        if (COMPACT) {
            const int n_pass = __popc(mask);
            const int gpw = (n_pass + NWP - 1) / NWP;
#pragma unroll
            for (int u = 0; u < GP1; ++u) {
                const int idx = warp * gpw + u;          // position in the ascending list of the genes of the pass
                int g = 0;
                if (u < gpw && idx < n_pass) {
                    unsigned m = mask;
                    for (int k = 0; k < idx; ++k) m &= m - 1u;          // drop the idx lowest set bits
                    g = __ffs(m) - 1;
                    ng_w = u + 1;
                }
                gu[u] = g;
            }
        } else {
#pragma unroll
            for (int u = 0; u < GP1; ++u) gu[u] = warp * GP1 + u;
            ng_w = ((mask >> (warp * GP1)) & ((1u << GP1) - 1u)) ? GP1 : 0;
        }
        const bool p1_mask = ng_w > 0;
        const double* const wbase = Ws + gh * MG;
        const double* const vbase = Ss + gh * MG;
        int zjk[GEMM ? NEK : 1];          // this lane's entries: j | k << 8
        if (GEMM) {
            const uint8_t* const pja = (KIND == PASS_LOC) ? pa : pas;
            const uint8_t* const pjb = (KIND == PASS_LOC) ? pb : pbs;
#pragma unroll
            for (int i = 0; i < NEK; ++i) zjk[i] = (int)pja[lane + 32 * i] | ((int)pjb[lane + 32 * i] << 8);
        }
        double acc[GEMM ? MG : 1][GEMM ? NEK : 1];
#pragma unroll
        for (int m = 0; m < (GEMM ? MG : 1); ++m)
#pragma unroll
            for (int i = 0; i < (GEMM ? NEK : 1); ++i) acc[m][i] = 0.0;
        double llacc[GP1], gacc[GP1], hacc[GP1], zmx[GP1], rnz[GP1];
        // score over the non-zero cells: lane <-> coefficient; up to 16 coefficients the two half-warps take alternate cells
        const int nz_s = (p <= 16) ? 2 : 1, nz_j = (p <= 16) ? (lane & 15) : lane, nz_h = (p <= 16) ? (lane >> 4) : 0;
        int cur[GP1];          // entries consumed so far (offset from the gene's first entry)
#pragma unroll
        for (int u = 0; u < GP1; ++u) {
            llacc[u] = 0.0; gacc[u] = 0.0; hacc[u] = 0.0; zmx[u] = -INFINITY; rnz[u] = 0.0;
            cur[u] = 0;
        }
        // register prefetch of the next cell tile: design rows, offsets, candidate entries of the sparse columns
        double xr[XPT], dr[DPT], lr = 0.0;
        int32_t yrow[GP1];
        float yval[GP1];
        auto prefetch = [&](int64_t c0) {
#pragma unroll
            for (int i = 0; i < XPT; ++i) {
                const int idx = tid + NT * i, j = idx / CT, c = idx % CT;
                xr[i] = (j < p && c0 + c < N) ? A.dl[(int64_t)j * N + c0 + c] : 0.0;
            }
            if (!SCONST) {
#pragma unroll
                for (int i = 0; i < DPT; ++i) {
                    const int idx = tid + NT * i, j = idx / CT, c = idx % CT;
                    dr[i] = (j < ps && c0 + c < N) ? A.ds[(int64_t)j * N + c0 + c] : 0.0;
                }
            }
            if (tid < CT) lr = (c0 + tid < N) ? (A.log_sf ? A.log_sf[c0 + tid] : 0.0) : DX_ETA_MIN;          // (padding: mu ~ 1e-130)
#pragma unroll
            for (int u = 0; u < GP1; ++u) {
                const int64_t e = ent[gu[u]] + cur[u] + lane;
                const bool in = u < ng_w && e < ent[GT + gu[u]];
                yrow[u] = in ? A.rows[e] : 0x7fffffff;
                yval[u] = in ? A.vals[e] : 0.0f;
            }
        };
        // stage the prefetched tile (registers -> buffers of parity pb), densify the sparse columns, fetch the tile after it
        int nzn_next[GP1];
        auto stage = [&](const int pb, const int64_t c0) {
            double* const Xs = Xs0 + pb * L.xstg; double* const Ds = Ds0 + pb * L.xstg; double* const Lsf = Lsf0 + pb * L.stg;
            float* const Ys = Ys0 + pb * L.stg * 2; uint8_t* const Yi = Yi0 + pb * L.stg * 8; uint8_t* const NZl = NZl0 + pb * L.stg * 8;
#pragma unroll
            for (int i = 0; i < XPT; ++i) {
                const int idx = tid + NT * i, j = idx / CT, c = idx % CT;
                if (j < p) Xs[j * XST + c] = xr[i];
            }
            if (!SCONST) {
#pragma unroll
                for (int i = 0; i < DPT; ++i) {
                    const int idx = tid + NT * i, j = idx / CT, c = idx % CT;
                    if (j < ps) Ds[j * XST + c] = dr[i];
                }
            }
            if (tid < CT) Lsf[tid] = lr;
#pragma unroll
            for (int u = 0; u < GP1; ++u) {
                float* yrow_s = Ys + (warp * GP1 + u) * CT;
#pragma unroll
                for (int v = 0; v < CP1; ++v) yrow_s[lane + 32 * v] = 0.0f;
            }
            {
                uint32_t* yi_w = reinterpret_cast<uint32_t*>(Yi + warp * GP1 * CT);
                for (int i = lane; i < GP1 * CT / 4; i += 32) yi_w[i] = 0u;
            }
            __syncwarp();
            const int64_t c1 = (N - c0 < CT) ? N : c0 + CT;
#pragma unroll
            for (int u = 0; u < GP1; ++u) {
                float* yrow_s = Ys + (warp * GP1 + u) * CT;
                uint8_t* yi_s = Yi + (warp * GP1 + u) * CT;
                uint8_t* nzl = NZl + (warp * GP1 + u) * (CT + 8);
                auto put = [&](const int c, const float v, const int at) {
                    // (a canonical shard -- cells ascending, a cell at most once per gene -- never trips these guards; a
                    // malformed one must not write outside the tile)
                    if (c < 0 || at >= CT + 8) return;
                    const int vi = (int)v;
                    yrow_s[c] = v;
                    yi_s[c] = (vi >= 0 && vi < TB && (float)vi == v) ? (uint8_t)vi : (uint8_t)255;
                    nzl[at] = (uint8_t)c;
                };
                bool hit = (int64_t)yrow[u] < c1;
                if (hit) put((int)(yrow[u] - c0), yval[u], lane);
                int n = __popc(__ballot_sync(DX_FULL_MASK, hit));
                int tot = n;
                cur[u] += n;
                while (n == 32) {                       // more than 32 entries of this gene inside the cell tile
                    const int64_t e = ent[gu[u]] + cur[u] + lane;
                    hit = false;
                    if (e < ent[GT + gu[u]]) {
                        const int64_t row = A.rows[e];
                        if (row < c1) { put((int)(row - c0), A.vals[e], tot + lane); hit = true; }
                    }
                    n = __popc(__ballot_sync(DX_FULL_MASK, hit));
                    tot += n;
                    cur[u] += n;
                }
                nzn_next[u] = (tot < CT + 8) ? tot : CT + 8;
            }
            if (c1 < N) prefetch(c1);
        };
        prefetch(0);
        __syncthreads();          // tables and constants visible
        stage(0, 0);
        __syncthreads();

        int pbuf = 0;
#pragma unroll 1
        for (int64_t c0 = 0; c0 < N; c0 += CT, pbuf ^= 1) {
            const int tn = (int)((N - c0 < CT) ? (N - c0) : CT);
            const double* const Xs = Xs0 + pbuf * L.xstg; const double* const Ds = Ds0 + pbuf * L.xstg;
            const double* const Lsf = Lsf0 + pbuf * L.stg;
            const float* const Ys = Ys0 + pbuf * L.stg * 2; const uint8_t* const Yi = Yi0 + pbuf * L.stg * 8;
            const uint8_t* const NZl = NZl0 + pbuf * L.stg * 8;
            const double* const dsrc_tile = (KIND == PASS_LOC) ? Xs : Ds;
            int nzn[GP1];
#pragma unroll
            for (int u = 0; u < GP1; ++u) nzn[u] = nzn_next[u];
            if (GEMM) {
                // product table of the cell tile, shared by all genes of the CTA: a warp per cell, lane <-> entry
#pragma unroll 4
                for (int c = warp; c < tn; c += NWP) {
#pragma unroll
                    for (int i = 0; i < NEK; ++i) {
                        const int j = zjk[i] & 255, k = zjk[i] >> 8;
                        double z = 0.0;
                        if (j != 255) {
                            z = dsrc_tile[j * XST + c];
                            if (k != 254) z *= dsrc_tile[k * XST + c];
                        }
                        Zs[c * ZWK + lane + 32 * i] = z;
                    }
                }
            }
            // ---- phase 1: per (cell, gene) model evaluation -------------------------------------------------------------
            if (p1_mask) {
                double eta[CP1][GP1], zeta[CP1][GP1];
#pragma unroll
                for (int v = 0; v < CP1; ++v) {
                    const double base = Lsf[lane + 32 * v];
#pragma unroll
                    for (int u = 0; u < GP1; ++u) { eta[v][u] = base; zeta[v][u] = 0.0; }
                }
#pragma unroll 2
                for (int j = 0; j < p; ++j) {
                    double xv[CP1], av[GP1];
#pragma unroll
                    for (int v = 0; v < CP1; ++v) xv[v] = Xs[j * XST + lane + 32 * v];
                    if (COMPACT) {
#pragma unroll
                        for (int u = 0; u < GP1; ++u) av[u] = qa_s[j * GT + gu[u]];
                    } else {
                        const double2* ap = reinterpret_cast<const double2*>(qa_s + j * GT + warp * GP1);
#pragma unroll
                        for (int u = 0; u < GP1; u += 2) { const double2 t = ap[u / 2]; av[u] = t.x; av[u + 1] = t.y; }
                    }
#pragma unroll
                    for (int v = 0; v < CP1; ++v)
#pragma unroll
                        for (int u = 0; u < GP1; ++u) eta[v][u] = fma(xv[v], av[u], eta[v][u]);
                }
                if (!SCONST) {
#pragma unroll 1
                    for (int j = 0; j < ps; ++j) {
                        double xv[CP1], bv[GP1];
#pragma unroll
                        for (int v = 0; v < CP1; ++v) xv[v] = Ds[j * XST + lane + 32 * v];
                        const double2* bp = reinterpret_cast<const double2*>(qb_s + j * GT + warp * GP1);
#pragma unroll
                        for (int u = 0; u < GP1; u += 2) { const double2 t = bp[u / 2]; bv[u] = t.x; bv[u + 1] = t.y; }
#pragma unroll
                        for (int v = 0; v < CP1; ++v)
#pragma unroll
                            for (int u = 0; u < GP1; ++u) zeta[v][u] = fma(xv[v], bv[u], zeta[v][u]);
                    }
                }
#pragma unroll
                for (int v = 0; v < CP1; ++v) {
                    const int c = lane + 32 * v;
                    const bool valid = c < tn;
                    double wv[GP1], sv[GP1];
#pragma unroll
                    for (int u = 0; u < GP1; ++u) {
                        if (COMPACT && u >= ng_w) break;          // (warp-uniform)
                        const int g = gu[u], slot = warp * GP1 + u;
                        const double y = (double)Ys[slot * CT + c];
                        const int yi = Yi[slot * CT + c];          // table index of the value (255: not in the tables)
                        double e = eta[v][u];
                        if ((__double2hiint(e) & 0x7fffffff) > 0x4071b000) e = dx_clip(e, DX_ETA_MIN, DX_ETA_MAX);          // |e| > 283 (or NaN)
                        const double mu = fl_exp(e);
                        double r, invr;
                        if (SCONST) { r = SCG(SC_R, g); invr = SCG(SC_INVR, g); }
                        else {
                            const double zc = dx_clip(zeta[v][u], DX_ETA_MIN, DX_ETA_MAX);
                            if (valid) zmx[u] = fmax(zmx[u], zc);
                            r = fl_exp(zc);
                            invr = fl_rcp_n(r);
                        }
                        const double inv = fl_rcp_n(r + mu);
                        const double q = r * inv;
                        const double l1p = fl_log1p(mu * invr);
                        double G = 0.0, d0 = 0.0, d1 = 0.0;
                        if (SCONST) {
                            G = Gt[g * TB + (yi & (TB - 1))];
                            if (KIND == PASS_SCALE) { d0 = T0[g * TB + (yi & (TB - 1))]; d1 = T1[g * TB + (yi & (TB - 1))]; }
                        }
                        if (SCONST ? (yi == 255) : (y != 0.0)) {          // counts beyond the tables, non-integer values, per-cell r
                            G = g_ratio_slow(r, y);
                            if (KIND == PASS_SCALE) psi_diffs_slow(r, y, d0, d1);
                        }
                        // (cells that pad a ragged last tile carry the offset -297.8 and a zero design row: mu ~ 1e-130, and every
                        // term below vanishes against the sums it joins)
                        llacc[u] += fma(y, e - l1p, fma(-r, l1p, G));
                        if (KIND == PASS_LOC) {
                            wv[u] = mu * q;
                            sv[u] = y * q;
                        } else {
                            const double gi = r * (d0 - l1p + (mu - y) * inv);
                            const double hi = gi + r * r * d1 + r * (mu * mu + r * y) * (inv * inv);
                            if (SCONST) { gacc[u] += gi; hacc[u] += hi; }
                            else { wv[u] = hi; sv[u] = gi; }
                        }
                    }
                    if (GEMM) {
                        double2* wp = reinterpret_cast<double2*>(Ws + c * WST + warp * GP1);
                        double2* sp = reinterpret_cast<double2*>(Ss + c * WST + warp * GP1);
#pragma unroll
                        for (int u = 0; u < GP1; u += 2) {
                            wp[u / 2] = make_double2(wv[u], wv[u + 1]);
                            sp[u / 2] = make_double2(sv[u], sv[u + 1]);
                        }
                    }
                }
                if (KIND == PASS_LOC) {
                    // score, first part: sum over the non-zero cells of (y q) x_j; lane <-> coefficient j
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < GP1; ++u) {
                        const uint8_t* nzl = NZl + (warp * GP1 + u) * (CT + 8);
                        const double* scol = Ss + warp * GP1 + u;
                        const double* xrow = Xs + ((nz_j < p) ? nz_j : 0) * XST;
                        const int n = nzn[u];
                        int i = nz_h;
#pragma unroll 1
                        for (; i + nz_s < n; i += 2 * nz_s) {
                            const int ca = nzl[i], cb = nzl[i + nz_s];
                            const double sa = scol[ca * WST], xa = xrow[ca], sb2 = scol[cb * WST], xb = xrow[cb];
                            rnz[u] = fma(sa, xa, rnz[u]);
                            rnz[u] = fma(sb2, xb, rnz[u]);
                        }
                        if (i < n) { const int ca = nzl[i]; rnz[u] = fma(scol[ca * WST], xrow[ca], rnz[u]); }
                    }
                }
            }
            if (GEMM) {
                __syncthreads();
                // ---- phase 2: C[g][e] += W[c][g] Z[c][e] ----------------------------------------------------------------
                if (grp_mask) {
#pragma unroll 2
                    for (int k = ks; k < tn; k += KS) {
                        double z[NEK];
                        const double* zrow = Zs + k * ZWK + lane;
#pragma unroll
                        for (int i = 0; i < NEK; ++i) z[i] = zrow[32 * i];
                        const double2* wrow = reinterpret_cast<const double2*>(wbase + k * WST);
                        const double2* vrow = reinterpret_cast<const double2*>(vbase + k * WST);
#pragma unroll
                        for (int m = 0; m < MG; m += 2) {
                            const double2 w = wrow[m / 2];
#pragma unroll
                            for (int i = 0; i < NEW; ++i) {
                                acc[m][i] = fma(w.x, z[i], acc[m][i]);
                                acc[m + 1][i] = fma(w.y, z[i], acc[m + 1][i]);
                            }
                            if (NEW < NEK) {
                                const double2 v = vrow[m / 2];
#pragma unroll
                                for (int i = NEW; i < NEK; ++i) {
                                    acc[m][i] = fma(v.x, z[i], acc[m][i]);
                                    acc[m + 1][i] = fma(v.y, z[i], acc[m + 1][i]);
                                }
                            }
                        }
                    }
                }
            }
            // the next tile is staged into the buffers of the other parity (last read in phase 1 of the previous tile, which
            // every warp has left): phase 0 shares the interval of the contraction, two barriers per tile instead of three
            if (c0 + CT < N) stage(pbuf ^ 1, c0 + CT);
            __syncthreads();
        }
        // ---- end of pass: the per-gene sums -------------------------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < GP1; ++u) {
            if (u >= ng_w) break;
            const int g = gu[u];
            const double l = dx_warp_sum(llacc[u]);
            const double gsum = (KIND == PASS_SCALE && SCONST) ? dx_warp_sum(gacc[u]) : 0.0;
            const double hsum = (KIND == PASS_SCALE && SCONST) ? dx_warp_sum(hacc[u]) : 0.0;
            const double zm = SCONST ? 0.0 : dx_warp_max(zmx[u]);
            if (lane == 0) {
                SCG(SC_LL, g) = l;
                if (KIND == PASS_SCALE && SCONST) { SCG(SC_G, g) = gsum; SCG(SC_H, g) = hsum; }
                if (!SCONST) SCG(SC_ZMAX, g) = zm;
            }
            if (KIND == PASS_LOC) {
                double rsum = rnz[u];
                if (nz_s == 2) rsum += __shfl_xor_sync(DX_FULL_MASK, rsum, 16);
                if (lane < p) rn_s[lane * GT + g] = rsum;
            }
        }
        if (GEMM) {
#pragma unroll
            for (int m = 0; m < MG; ++m)
#pragma unroll
                for (int i = 0; i < NEK; ++i)
                    Cred[((size_t)ks * GT + gh * MG + m) * ZWK + lane + 32 * i] = acc[m][i];
            __syncthreads();
            if (KS > 1) {
                for (int idx = tid; idx < GT * ZWK; idx += NT) {
                    double s = Cred[idx];
#pragma unroll
                    for (int q = 1; q < KS; ++q) s += Cred[(size_t)q * GT * ZWK + idx];
                    Cred[idx] = s;
                }
            }
        }
        __syncthreads();
    };
    auto LOC = std::integral_constant<int, PASS_LOC>();
    auto SCALE = std::integral_constant<int, PASS_SCALE>();

    // X^T W X and X^T S of the genes in `which` out of the reduced pass result: the information matrix goes to global memory
    // (fim_inv holds it until finalize inverts it), the score to rhs, the IWLS step (X^T W X)^-1 X^T S to delta
    auto extract_loc = [&](const unsigned which, const int64_t g0) {
        double* const Lw = scratch + warp * (p * p + 2 * p);
        double* const rv = Lw + p * p;
        double* const xv = rv + p;
        for (int g = warp; g < GT; g += NWP) {
            if (!((which >> g) & 1u)) continue;
            const double* cg = Cred + (size_t)g * ZW;
            for (int e = lane; e < p * p; e += 32) {
                const int i = e / p, j = e - i * p;
                Lw[e] = (i >= j) ? cg[i * (i + 1) / 2 + j] : cg[j * (j + 1) / 2 + i];
            }
            for (int i = lane; i < p; i += 32) { const double v = rn_s[i * GT + g] - cg[ZW - p + i]; rv[i] = v; rhs_s[i * GT + g] = v; }
            __syncwarp();
            double* fo = A.fim_inv + (g0 + g) * (int64_t)p * p;
            for (int e = lane; e < p * p; e += 32) fo[e] = Lw[e];
            bool fin = true;
            for (int i = lane; i < p; i += 32) fin = fin && isfinite(rv[i]);
            fin = __all_sync(DX_FULL_MASK, fin);
            const bool ok = chol_inplace(Lw, p, lane);
            __syncwarp();
            if (ok && fin) {
                if (lane == 0) chol_solve_serial(Lw, p, rv, xv);
                __syncwarp();
                for (int i = lane; i < p; i += 32) delta_s[i * GT + g] = xv[i];
            }
            if (lane == 0) ICG(IC_DOK, g) = (ok && fin) ? 1 : 0;
            __syncwarp();
        }
    };

    for (;;) {
        if (tid == 0) s_tile = (int)atomicAdd(A.queue, 1u);
        __syncthreads();
        const int64_t g0 = (int64_t)s_tile * GT;
        __syncthreads();
        if (g0 >= A.n_genes) return;
        const int ng = (int)((A.n_genes - g0 < GT) ? (A.n_genes - g0) : GT);
        const unsigned valid_mask = (ng >= 32) ? 0xffffffffu : ((1u << ng) - 1u);
        if (tid == 0) atomicAdd(&g_tile_counters[3], 1ull);
        // ---- start values, per-gene constants ----------------------------------------------------------------------------
        if (tid < GT) {
            const int g = tid;
            const bool v = g < ng;
            const int64_t e0 = v ? A.indptr[g0 + g] : 0, e1 = v ? A.indptr[g0 + g + 1] : 0;
            ent[g] = e0; ent[GT + g] = e1;
            for (int i = 0; i < p; ++i) {
                const double av = v ? dx_clip(A.a_var[(int64_t)i * A.n_genes + g0 + g], DX_ETA_MIN, DX_ETA_MAX) : 0.0;
                a_s[i * GT + g] = av; qa_s[i * GT + g] = av; delta_s[i * GT + g] = 0.0; rhs_s[i * GT + g] = 0.0;
            }
            for (int i = 0; i < ps; ++i) {
                const double bv = v ? dx_clip(A.b_var[(int64_t)i * A.n_genes + g0 + g], DX_ETA_MIN, DX_ETA_MAX) : 0.0;
                b_s[i * GT + g] = bv; qb_s[i * GT + g] = bv; bt_s[i * GT + g] = bv; sb_s[i * GT + g] = 0.0; gb_s[i * GT + g] = 0.0;
            }
            ICG(IC_FLAGS, g) = (e1 == e0) ? DX_FLAG_ALL_ZERO : 0;
            ICG(IC_LOCCONV, g) = 0; ICG(IC_FULLY, g) = v ? 0 : 1; ICG(IC_HAVE, g) = 1; ICG(IC_DOK, g) = 0;
            ICG(IC_NSTATE, g) = NS_DONE; ICG(IC_NITER, g) = 0;
        }
        __syncthreads();
        for (int g = warp; g < GT; g += NWP) {
            double lg = 0.0;
#pragma unroll 4
            for (int64_t e = ent[g] + lane; e < ent[GT + g]; e += 32) {
                const double y = (double)A.vals[e];
                const int yi = (int)y;
                lg += (yi >= 0 && yi < TB && (double)yi == y) ? lfact[yi] : lgamma(y + 1.0);
            }
            lg = dx_warp_sum(lg);
            if (lane == 0) SCG(SC_LGAM, g) = lg;
        }
        run_pass(LOC, valid_mask);
        if (tid < GT) SCG(SC_NLL, tid) = -(SCG(SC_LL, tid) - SCG(SC_LGAM, tid));
        extract_loc(valid_mask, g0);
        __syncthreads();

        int epochs_until_b = freq;
        int step = 0;
#pragma unroll 1
        while (step < A.opts.max_steps) {
            // genes still training
            if (tid == 0) s_mask = 0u;
            __syncthreads();
            if (tid < GT && !ICG(IC_FULLY, tid)) atomicOr(&s_mask, 1u << tid);
            __syncthreads();
            const unsigned active = s_mask;
            if (!active) break;
            if (tid < GT) SCG(SC_NLLNEW, tid) = SCG(SC_NLL, tid);
            const bool b_epoch = (epochs_until_b == 0);
            if (b_epoch) {
                if (train_scale) {
                    // ---- safeguarded Newton in b (oracle _b_step_newton), all active genes in lockstep passes -------------
                    if (tid == 0) atomicAdd(&g_tile_counters[2], 1ull);
                    if (tid < GT) {
                        const int g = tid;
                        const bool on = (active >> g) & 1u;
                        for (int i = 0; i < ps; ++i) { bt_s[i * GT + g] = b_s[i * GT + g]; qb_s[i * GT + g] = b_s[i * GT + g]; }
                        ICG(IC_NSTATE, g) = on ? NS_DERIV : NS_DONE;
                        ICG(IC_NIT, g) = 0; ICG(IC_NHALV, g) = 0; ICG(IC_MOVED, g) = 0;
                        SCG(SC_F, g) = -SCG(SC_NLL, g);
                    }
                    for (;;) {
                        if (tid == 0) s_mask = 0u;
                        __syncthreads();
                        if (tid < GT && ICG(IC_NSTATE, tid) != NS_DONE) atomicOr(&s_mask, 1u << tid);
                        __syncthreads();
                        const unsigned nmask = s_mask;
                        if (!nmask) break;
                        run_pass(SCALE, nmask);
                        // accept / reject the evaluated points (a thread per gene), then directions (a warp per gene)
                        if (tid == 0) s_acc = 0u;
                        __syncthreads();
                        if (tid < GT && ((nmask >> tid) & 1u)) {
                            const int g = tid;
                            const int st = ICG(IC_NSTATE, g);
                            bool want_dir = false;
                            if (st == NS_DERIV) {
                                if (SCG(SC_ZMAX, g) >= DX_ETA_SCALE_FROZEN) ICG(IC_NSTATE, g) = NS_DONE;
                                else want_dir = true;
                            } else {
                                const double f_try = SCG(SC_LL, g) - SCG(SC_LGAM, g);
                                if (isfinite(f_try) && f_try >= SCG(SC_F, g)) {
                                    for (int i = 0; i < ps; ++i) bt_s[i * GT + g] = qb_s[i * GT + g];
                                    SCG(SC_F, g) = f_try;
                                    ICG(IC_MOVED, g) = 1;
                                    const int it = ICG(IC_NIT, g) + 1;
                                    ICG(IC_NIT, g) = it;
                                    if (SCG(SC_ZMAX, g) >= DX_ETA_SCALE_FROZEN || it >= A.opts.newton_max_iter) ICG(IC_NSTATE, g) = NS_DONE;
                                    else want_dir = true;
                                } else {
                                    const int h = ICG(IC_NHALV, g) + 1;
                                    ICG(IC_NHALV, g) = h;
                                    if (h >= DX_NEWTON_MAX_HALVINGS) ICG(IC_NSTATE, g) = NS_DONE;
                                    else {
                                        const double t = SCG(SC_T, g) * 0.5;
                                        SCG(SC_T, g) = t;
                                        for (int i = 0; i < ps; ++i)
                                            qb_s[i * GT + g] = dx_clip(bt_s[i * GT + g] + t * sb_s[i * GT + g], DX_ETA_MIN, DX_ETA_MAX);
                                    }
                                }
                            }
                            if (want_dir) atomicOr(&s_acc, 1u << g);
                        }
                        __syncthreads();
                        const unsigned dirs = s_acc;
                        {
                            double* const Hw = scratch + warp * (2 * ps * ps + 2 * ps);
                            double* const Lb = Hw + ps * ps;
                            double* const gv = Lb + ps * ps;
                            double* const sv = gv + ps;
                            for (int g = warp; g < GT; g += NWP) {
                                if (!((dirs >> g) & 1u)) continue;
                                if (SCONST) {
                                    if (lane == 0) { Hw[0] = -SCG(SC_H, g) * z_const * z_const; gv[0] = SCG(SC_G, g) * z_const; }
                                } else {
                                    const double* cg = Cred + (size_t)g * 96;
                                    for (int e = lane; e < ps * ps; e += 32) {
                                        const int i = e / ps, j = e - i * ps;
                                        Hw[e] = -((i >= j) ? cg[i * (i + 1) / 2 + j] : cg[j * (j + 1) / 2 + i]);
                                    }
                                    for (int i = lane; i < ps; i += 32) gv[i] = cg[64 + i];
                                }
                                __syncwarp();
                                bool ok = newton_direction(Hw, Lb, gv, sv, ps, lane);
                                double smax = 0.0, pred = 0.0;
                                if (ok) {
                                    for (int i = lane; i < ps; i += 32) smax = fmax(smax, fabs(sv[i]));
                                    smax = dx_warp_max(smax);
                                    const double scl = (smax > DX_NEWTON_MAX_STEP) ? DX_NEWTON_MAX_STEP / smax : 1.0;
                                    for (int i = lane; i < ps; i += 32) { sv[i] *= scl; pred += gv[i] * sv[i]; }
                                    pred = dx_warp_sum(pred);
                                    smax *= scl;
                                    if (smax <= A.opts.newton_xtol || !(pred > DX_NEWTON_GAIN_RTOL * fabs(SCG(SC_F, g)))) ok = false;
                                }
                                __syncwarp();
                                if (ok) {
                                    for (int i = lane; i < ps; i += 32) {
                                        sb_s[i * GT + g] = sv[i];
                                        qb_s[i * GT + g] = dx_clip(bt_s[i * GT + g] + sv[i], DX_ETA_MIN, DX_ETA_MAX);
                                    }
                                    if (lane == 0) { SCG(SC_T, g) = 1.0; ICG(IC_NHALV, g) = 0; ICG(IC_NSTATE, g) = NS_LINE; }
                                } else if (lane == 0) {
                                    ICG(IC_NSTATE, g) = NS_DONE;
                                }
                                __syncwarp();
                            }
                        }
                        __syncthreads();
                    }
                    if (tid < GT && ((active >> tid) & 1u)) {
                        const int g = tid;
                        const double nll_prop = -SCG(SC_F, g);
                        if (!isfinite(nll_prop)) ICG(IC_FLAGS, g) |= DX_FLAG_NONFINITE;
                        if (isfinite(nll_prop) && !(nll_prop > SCG(SC_NLL, g))) {
                            for (int i = 0; i < ps; ++i) b_s[i * GT + g] = bt_s[i * GT + g];
                            SCG(SC_NLLNEW, g) = nll_prop;
                            if (ICG(IC_MOVED, g)) ICG(IC_HAVE, g) = 0;
                        }
                    }
                    __syncthreads();
                }
                epochs_until_b = freq;
            } else {
                if (train_loc) {
                    // genes that take an IWLS step in this epoch
                    if (tid == 0) { s_mask = 0u; s_acc = 0u; }
                    __syncthreads();
                    if (tid < GT && ((active >> tid) & 1u) && !ICG(IC_LOCCONV, tid)) {
                        atomicOr(&s_mask, 1u << tid);
                        if (!ICG(IC_HAVE, tid)) atomicOr(&s_acc, 1u << tid);
                    }
                    __syncthreads();
                    const unsigned amask = s_mask, refresh = s_acc;
                    if (amask) {
                        if (refresh) {          // the dispersion moved since X^T W X was formed: re-evaluate at (a, b)
                            if (tid < GT) {
                                for (int i = 0; i < p; ++i) qa_s[i * GT + tid] = a_s[i * GT + tid];
                                for (int i = 0; i < ps; ++i) qb_s[i * GT + tid] = b_s[i * GT + tid];
                            }
                            run_pass(LOC, refresh);
                            extract_loc(refresh, g0);
                            if (tid < GT && ((refresh >> tid) & 1u)) ICG(IC_HAVE, tid) = 1;
                            __syncthreads();
                        }
                        if (tid < GT) {
                            const int g = tid;
                            const bool ok = ICG(IC_DOK, g) != 0;
                            if (((amask >> g) & 1u) && !ok) ICG(IC_FLAGS, g) |= DX_FLAG_SINGULAR_FIM;
                            for (int i = 0; i < p; ++i) {
                                const double av = a_s[i * GT + g];
                                qa_s[i * GT + g] = ok ? dx_clip(av + delta_s[i * GT + g], DX_ETA_MIN, DX_ETA_MAX) : av;
                            }
                            for (int i = 0; i < ps; ++i) qb_s[i * GT + g] = b_s[i * GT + g];
                        }
                        run_pass(LOC, amask);
                        if (tid == 0) s_acc = 0u;
                        __syncthreads();
                        if (tid < GT && ((amask >> tid) & 1u)) {
                            const int g = tid;
                            const double nll_prop = -(SCG(SC_LL, g) - SCG(SC_LGAM, g));
                            if (!isfinite(nll_prop)) ICG(IC_FLAGS, g) |= DX_FLAG_NONFINITE;
                            if (isfinite(nll_prop) && !(nll_prop > SCG(SC_NLL, g))) {
                                for (int i = 0; i < p; ++i) a_s[i * GT + g] = qa_s[i * GT + g];
                                SCG(SC_NLLNEW, g) = nll_prop;
                                atomicOr(&s_acc, 1u << g);
                            }
                        }
                        __syncthreads();
                        extract_loc(s_acc, g0);
                        __syncthreads();
                    }
                }
                epochs_until_b -= 1;
            }
            ++step;
            // per-gene convergence bookkeeping (batchglm's two-stage rule)
            if (tid < GT && ((active >> tid) & 1u)) {
                const int g = tid;
                const double nll = SCG(SC_NLL, g), nll_new = SCG(SC_NLLNEW, g);
                const bool conv_f = ((nll - nll_new) / nll) < A.opts.lltol;
                SCG(SC_NLL, g) = nll_new;
                ICG(IC_NITER, g) = step;
                int loc_conv = ICG(IC_LOCCONV, g), fully = 0;
                if (b_epoch) {
                    if (loc_conv && conv_f) fully = 1;
                    loc_conv = fully;
                } else {
                    loc_conv = loc_conv || conv_f;
                    if (!train_scale) fully = loc_conv;
                }
                ICG(IC_LOCCONV, g) = loc_conv;
                if (fully) { ICG(IC_FULLY, g) = 1; ICG(IC_FLAGS, g) |= DX_FLAG_CONVERGED; }
            }
            __syncthreads();
        }
        // ---- finalize -------------------------------------------------------------------------------------------------------
        if (tid == 0) s_mask = 0u;
        __syncthreads();
        if (tid < GT) {
            if (tid < ng && !ICG(IC_HAVE, tid)) atomicOr(&s_mask, 1u << tid);
            for (int i = 0; i < p; ++i) qa_s[i * GT + tid] = a_s[i * GT + tid];
            for (int i = 0; i < ps; ++i) qb_s[i * GT + tid] = b_s[i * GT + tid];
        }
        __syncthreads();
        const unsigned refresh = s_mask;
        if (refresh) {
            run_pass(LOC, refresh);
            extract_loc(refresh, g0);
            __syncthreads();
        }
        run_pass(SCALE, valid_mask);          // the gradient in b for the jacobian norm, and the plateau flag
        if (tid < GT) {
            const int g = tid;
            if (SCONST) gb_s[g] = SCG(SC_G, g) * z_const;
            else for (int i = 0; i < ps; ++i) gb_s[i * GT + g] = Cred[(size_t)g * 96 + 64 + i];
            if (g < ng && SCG(SC_ZMAX, g) >= DX_ETA_SCALE_FROZEN) ICG(IC_FLAGS, g) |= DX_FLAG_SCALE_FROZEN;
        }
        __syncthreads();
        {
            double* const Lw = tile + warp * (2 * p * p);
            double* const Iv = Lw + p * p;
            for (int g = warp; g < ng; g += NWP) {
                double jac = 0.0;
                for (int i = lane; i < p; i += 32) jac += fabs(rhs_s[i * GT + g]);
                for (int i = lane; i < ps; i += 32) jac += fabs(gb_s[i * GT + g]);
                jac = dx_warp_sum(jac) / (double)N;
                double* fo = A.fim_inv + (g0 + g) * (int64_t)p * p;
                for (int e = lane; e < p * p; e += 32) Lw[e] = fo[e];
                __syncwarp();
                const bool okf = chol_inplace(Lw, p, lane);
                __syncwarp();
                for (int col = lane; col < p; col += 32) {
                    double* x = Iv + col * p;
                    if (okf) {
                        for (int i = 0; i < p; ++i) {
                            double s = (i == col) ? 1.0 : 0.0;
                            for (int k = 0; k < i; ++k) s -= Lw[i * p + k] * x[k];
                            x[i] = s / Lw[i * p + i];
                        }
                        for (int i = p - 1; i >= 0; --i) {
                            double s = x[i];
                            for (int k = i + 1; k < p; ++k) s -= Lw[k * p + i] * x[k];
                            x[i] = s / Lw[i * p + i];
                        }
                    }
                }
                __syncwarp();
                for (int e = lane; e < p * p; e += 32) {
                    const int i = e / p, j = e - i * p;
                    fo[e] = okf ? 0.5 * (Iv[i * p + j] + Iv[j * p + i]) : NAN;
                }
                for (int i = lane; i < p; i += 32) A.a_var[(int64_t)i * A.n_genes + g0 + g] = a_s[i * GT + g];
                for (int i = lane; i < ps; i += 32) A.b_var[(int64_t)i * A.n_genes + g0 + g] = b_s[i * GT + g];
                if (lane == 0) {
                    A.ll[g0 + g] = -SCG(SC_NLL, g);
                    A.jac[g0 + g] = jac;
                    A.niter[g0 + g] = ICG(IC_NITER, g);
                    A.flags[g0 + g] = ICG(IC_FLAGS, g);
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
#undef SCG
#undef ICG
}

template <int NE, int MG, int GT, int CT, int PMAXC, int MINB, bool SCONST>
cudaError_t launch_cfg(const TileArgs& A0, int sm_count, cudaStream_t stream) {
    TileArgs A = A0;
    const Lay L = make_layout<NE, MG, GT, CT, SCONST>(A.p, A.ps);
    auto kern = dx_fit_tile_kernel<NE, MG, GT, CT, PMAXC, MINB, SCONST>;
    if (L.total_bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes);
    if (e != cudaSuccess) return e;
    e = dx_queue_acquire(stream, &A.queue);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, (size_t)L.total_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (A.n_genes + GT - 1) / GT;
    if (grid > (int64_t)sm_count * per_sm) grid = (int64_t)sm_count * per_sm;
    kern<<<(unsigned)grid, NT, (size_t)L.total_bytes, stream>>>(A);
    return cudaGetLastError();
}

template <bool SCONST>
cudaError_t launch_p(const TileArgs& A, int sm_count, cudaStream_t stream) {
    const int need = A.p * (A.p + 1) / 2 + A.p;          // entries of X^T W X (packed) + X^T S
    // up to 12 coefficients: two CTAs per SM (128 registers, < 113 KB of shared memory each), so that one tile's barriers and
    // dependency chains are covered by the other's arithmetic; wider designs need the registers for the accumulators
    //                     NE  MG  GT  CT  PMAXC MINB
    if (need <= 32)  return launch_cfg<1, 8, 32, 32, 7, 2, SCONST>(A, sm_count, stream);
    if (need <= 64)  return launch_cfg<2, 8, 32, 32, 9, 2, SCONST>(A, sm_count, stream);
    if (need <= 96) {
        const char* env = getenv("DXB200_TILE_MINB");          // tuning hook
        if (env && atoi(env) == 1) return launch_cfg<3, 16, 32, 64, 12, 1, SCONST>(A, sm_count, stream);
        return launch_cfg<3, 8, 32, 32, 12, 2, SCONST>(A, sm_count, stream);
    }
    if (need <= 160) return launch_cfg<5, 8, 32, 64, 16, 1, SCONST>(A, sm_count, stream);
    if (need <= 352) return launch_cfg<11, 4, 32, 32, 24, 1, SCONST>(A, sm_count, stream);
    return launch_cfg<18, 2, 16, 32, 32, 1, SCONST>(A, sm_count, stream);
}

}  // namespace

// The file is compiled twice (Makefile): DX_TILE_PART 1 holds the instantiations of the constant dispersion model and the
// entry points, DX_TILE_PART 2 those of the general scale model -- two translation units build side by side instead of one
// that takes three minutes.  Each part has its own copy of the diagnostics counters.
#ifndef DX_TILE_PART
#define DX_TILE_PART 0          // both in one object
#endif

#define DX_TILE_FILL_ARGS()                                                                                              \
    TileArgs A;                                                                                                          \
    A.n_genes = n_genes; A.N = n_cells; A.p = p; A.ps = ps;                                                              \
    A.dl = design_loc; A.ds = design_scale; A.log_sf = log_sf;                                                           \
    A.indptr = indptr; A.rows = rows; A.vals = vals; A.opts = *opts;                                                     \
    A.a_var = a_var; A.b_var = b_var; A.fim_inv = fim_inv; A.ll = ll; A.jac = jac; A.niter = niter; A.flags = flags;     \
    A.queue = nullptr

static cudaError_t tile_counters(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_tile_counters, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    if (reset) {
        const unsigned long long zero[4] = {0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_tile_counters, zero, sizeof(zero));
    }
    return e;
}

#if DX_TILE_PART == 2 || DX_TILE_PART == 0
cudaError_t dx_fit_percell_counters_general(unsigned long long* out, int reset) { return tile_counters(out, reset); }

cudaError_t dx_launch_fit_percell_general(int64_t n_genes, int64_t n_cells, int p, int ps,
                                          const double* design_loc, const double* design_scale, const double* log_sf,
                                          const int64_t* indptr, const int32_t* rows, const float* vals,
                                          const dx_fit_opts* opts,
                                          double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                          int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream) {
    DX_TILE_FILL_ARGS();
    return launch_p<false>(A, sm_count, stream);
}
#else
cudaError_t dx_fit_percell_counters_general(unsigned long long* out, int reset);
cudaError_t dx_launch_fit_percell_general(int64_t n_genes, int64_t n_cells, int p, int ps,
                                          const double* design_loc, const double* design_scale, const double* log_sf,
                                          const int64_t* indptr, const int32_t* rows, const float* vals,
                                          const dx_fit_opts* opts,
                                          double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                          int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream);
#endif

#if DX_TILE_PART == 1 || DX_TILE_PART == 0
cudaError_t dx_fit_percell_counters_impl(unsigned long long* out, int reset) {
    unsigned long long a[4], b[4] = {0, 0, 0, 0};
    cudaError_t e = tile_counters(a, reset);
    if (e != cudaSuccess) return e;
#if DX_TILE_PART == 1
    e = dx_fit_percell_counters_general(b, reset);
    if (e != cudaSuccess) return e;
#endif
    for (int i = 0; i < 4; ++i) out[i] = a[i] + b[i];
    return cudaSuccess;
}

cudaError_t dx_launch_fit_percell(int64_t n_genes, int64_t n_cells, int p, int ps,
                                  const double* design_loc, const double* design_scale, const double* log_sf,
                                  const int64_t* indptr, const int32_t* rows, const float* vals,
                                  const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream) {
    if (n_genes <= 0) return cudaSuccess;
    if (p < 1 || p > 32 || ps < 1 || ps > 8) return cudaErrorNotSupported;
    const bool scale_const = (ps == 1) && (opts->reserved & DX_OPT_SCALE_CONST);
    if (!scale_const)
        return dx_launch_fit_percell_general(n_genes, n_cells, p, ps, design_loc, design_scale, log_sf, indptr, rows, vals, opts,
                                             a_var, b_var, fim_inv, ll, jac, niter, flags, sm_count, stream);
    DX_TILE_FILL_ARGS();
    return launch_p<true>(A, sm_count, stream);
}
#endif
