// dx_api.cu -- the extern "C" surface of libdxb200.so (include/dxb200.h): argument checks, error
// strings, launch plumbing.  No torch types, no allocation except inside the *_host entry point.
#include <cstdio>
#include <cstring>
#include <string>
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {
thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }
int fail_cuda(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return DX_ERR_CUDA;
}
int sm_count_cached() {
    static int sm = 0;
    if (sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm <= 0) sm = 148;
    }
    return sm;
}
bool bad_opts(const dx_fit_opts* o) {
    return !o || o->max_steps < 0 || o->update_b_freq < 1 || o->newton_max_iter < 0 || !(o->lltol >= 0.0);
}
}  // namespace

#define DX_CHECK(cond, msg) do { if (!(cond)) return fail(DX_ERR_INVALID, msg); } while (0)
#define DX_CUDA(call, where) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda(e__, where); } while (0)

extern "C" {

int dx_version(void) { return 100; }
const char* dx_last_error(void) { return g_err.c_str(); }
int dx_device_sm_count(int* out) {
    DX_CHECK(out, "dx_device_sm_count: null out");
    *out = sm_count_cached();
    return DX_OK;
}

int dx_group_suffstats(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                       const uint8_t* cell_group, int32_t n_groups,
                       uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_group_suffstats: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_group_suffstats: n_groups must be in 1..255");
    DX_CHECK(indptr && cell_group && hist && ovf_count, "dx_group_suffstats: null pointer");
    DX_CUDA(dx_launch_ingest(n_genes, indptr, rows, vals, cell_group, n_groups, hist, ovf_count, ovf_val, ovf_group,
                             sm_count_cached(), (cudaStream_t)stream), "dx_group_suffstats");
    return DX_OK;
}

int dx_gene_moments(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                    const double* inv_sf, double* sum1, double* sum2, void* stream) {
    DX_CHECK(n_genes >= 0 && indptr && sum1 && sum2, "dx_gene_moments: bad arguments");
    DX_CUDA(dx_launch_gene_moments(n_genes, indptr, rows, vals, inv_sf, sum1, sum2, (cudaStream_t)stream), "dx_gene_moments");
    return DX_OK;
}

int dx_nb_fit_grouped(int64_t n_genes, int32_t n_groups, int32_t p_loc, int32_t p_scale,
                      const double* xu_loc, const double* xu_scale, const double* offset, const double* n_k,
                      const double* pinv_loc, const int32_t* loc_group, int32_t n_loc_groups,
                      const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                      const float* ovf_val, const uint8_t* ovf_group, const dx_fit_opts* opts,
                      double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                      int32_t* niter, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_nb_fit_grouped: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_nb_fit_grouped: n_groups must be in 1..255");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF && p_scale >= 1 && p_scale <= DX_MAX_COEF,
             "dx_nb_fit_grouped: coefficient counts must be in 1..32");
    DX_CHECK(!bad_opts(opts), "dx_nb_fit_grouped: bad options");
    DX_CHECK(xu_loc && xu_scale && n_k && hist && ovf_count && indptr, "dx_nb_fit_grouped: null input pointer");
    DX_CHECK(a_var && b_var && fim_inv && ll && jac && niter && flags, "dx_nb_fit_grouped: null output pointer");
    DX_CHECK(opts->init_a != DX_INIT_CLOSED_FORM || (pinv_loc && n_loc_groups >= 1 && n_loc_groups <= n_groups),
             "dx_nb_fit_grouped: closed-form init needs pinv_loc and 1 <= n_loc_groups <= n_groups");
    DX_CHECK(opts->init_b != DX_INIT_CLOSED_FORM, "dx_nb_fit_grouped: closed-form init_b must be resolved by the host");
    DX_CUDA(dx_launch_fit_grouped(n_genes, n_groups, p_loc, p_scale, xu_loc, xu_scale, offset, n_k, pinv_loc, loc_group,
                                  n_loc_groups, hist, ovf_count, indptr, ovf_val, ovf_group, opts, a_var, b_var, fim_inv, ll, jac, niter,
                                  flags, (cudaStream_t)stream), "dx_nb_fit_grouped");
    return DX_OK;
}

int dx_nb_fit_percell(int64_t n_genes, int64_t n_cells, int32_t p_loc, int32_t p_scale,
                      const double* design_loc, const double* design_scale, const double* log_sf,
                      const int64_t* indptr, const int32_t* rows, const float* vals, const dx_fit_opts* opts,
                      double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                      int32_t* niter, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0 && n_cells >= 1, "dx_nb_fit_percell: bad sizes");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF && p_scale >= 1 && p_scale <= DX_MAX_COEF,
             "dx_nb_fit_percell: coefficient counts must be in 1..32");
    DX_CHECK(!bad_opts(opts), "dx_nb_fit_percell: bad options");
    DX_CHECK(opts->init_a == DX_INIT_GIVEN && opts->init_b == DX_INIT_GIVEN,
             "dx_nb_fit_percell: start values must be given (DX_INIT_GIVEN)");
    DX_CHECK(design_loc && design_scale && indptr, "dx_nb_fit_percell: null input pointer");
    DX_CHECK(a_var && b_var && fim_inv && ll && jac && niter && flags, "dx_nb_fit_percell: null output pointer");
    DX_CUDA(dx_launch_fit_percell(n_genes, n_cells, p_loc, p_scale, design_loc, design_scale, log_sf, indptr, rows, vals,
                                  opts, a_var, b_var, fim_inv, ll, jac, niter, flags, sm_count_cached(),
                                  (cudaStream_t)stream), "dx_nb_fit_percell");
    return DX_OK;
}

int dx_wald_test(int64_t n, const double* theta, const double* sd, double* pval, double* log_pval, void* stream) {
    DX_CHECK(n >= 0 && theta && sd && pval, "dx_wald_test: bad arguments");
    DX_CUDA(dx_launch_wald(n, theta, sd, pval, log_pval, (cudaStream_t)stream), "dx_wald_test");
    return DX_OK;
}

int dx_wald_from_fit(int64_t n_genes, int32_t p_loc, int32_t coef, const double* a_var, const double* fim_inv,
                     double* theta, double* sd, double* pval, double* log_pval, void* stream) {
    DX_CHECK(n_genes >= 0 && a_var && fim_inv && pval, "dx_wald_from_fit: bad arguments");
    DX_CHECK(coef >= 0 && coef < p_loc, "dx_wald_from_fit: coefficient index out of range");
    DX_CUDA(dx_launch_wald_from_fit(n_genes, p_loc, coef, a_var, fim_inv, theta, sd, pval, log_pval, (cudaStream_t)stream),
            "dx_wald_from_fit");
    return DX_OK;
}

int dx_wald_chisq(int64_t n_genes, int32_t k, const double* theta, const double* cov,
                  double* stat, double* pval, double* log_pval, void* stream) {
    DX_CHECK(n_genes >= 0 && theta && cov && pval, "dx_wald_chisq: bad arguments");
    DX_CHECK(k >= 1 && k <= DX_MAX_COEF, "dx_wald_chisq: k must be in 1..32");
    DX_CUDA(dx_launch_wald_chisq(n_genes, k, theta, cov, stat, pval, log_pval, (cudaStream_t)stream), "dx_wald_chisq");
    return DX_OK;
}

int dx_lrt(int64_t n, const double* ll_full, const double* ll_reduced, int32_t delta_df,
           double* pval, double* log_pval, void* stream) {
    DX_CHECK(n >= 0 && ll_full && ll_reduced && pval, "dx_lrt: bad arguments");
    DX_CHECK(delta_df >= 1, "dx_lrt: delta_df must be >= 1");
    DX_CUDA(dx_launch_lrt(n, ll_full, ll_reduced, delta_df, pval, log_pval, (cudaStream_t)stream), "dx_lrt");
    return DX_OK;
}

int dx_two_coef_z(int64_t n, const double* t0, const double* t1, const double* sd0, const double* sd1,
                  double* pval, void* stream) {
    DX_CHECK(n >= 0 && t0 && t1 && sd0 && sd1 && pval, "dx_two_coef_z: bad arguments");
    DX_CUDA(dx_launch_two_coef_z(n, t0, t1, sd0, sd1, pval, (cudaStream_t)stream), "dx_two_coef_z");
    return DX_OK;
}

int dx_t_test_moments(int64_t n, const double* mu0, const double* mu1, const double* var0, const double* var1,
                      double n0, double n1, double* pval, void* stream) {
    DX_CHECK(n >= 0 && mu0 && mu1 && var0 && var1 && pval, "dx_t_test_moments: bad arguments");
    DX_CUDA(dx_launch_t_moments(n, mu0, mu1, var0, var1, n0, n1, pval, (cudaStream_t)stream), "dx_t_test_moments");
    return DX_OK;
}

int dx_bh_qvalues(int64_t n, const double* pval, double* qval, double* work, void* stream) {
    DX_CHECK(n >= 0 && pval && qval && work, "dx_bh_qvalues: bad arguments");
    DX_CUDA(dx_launch_bh(n, pval, qval, work, (cudaStream_t)stream), "dx_bh_qvalues");
    return DX_OK;
}

int dx_two_group_tests(int64_t n_genes, double n0, double n1,
                       const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                       const float* ovf_val, const uint8_t* ovf_group,
                       int32_t do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0 && hist && ovf_count && indptr && out, "dx_two_group_tests: bad arguments");
    DX_CHECK(n0 >= 1 && n1 >= 1, "dx_two_group_tests: both groups need at least one cell");
    DX_CHECK(n0 + n1 <= 2097151.0 || !do_rank, "dx_two_group_tests: rank sums are exact up to 2^21-1 cells");
    DX_CUDA(dx_launch_two_group(n_genes, n0, n1, hist, ovf_count, indptr, ovf_val, ovf_group, do_rank, out, u1x2, tie,
                                flags, (cudaStream_t)stream), "dx_two_group_tests");
    return DX_OK;
}

// ---- host-buffer entry point ------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

int dx_wald_grouped_host(int64_t n_cells, int64_t n_genes, const int64_t* indptr, const int32_t* rows,
                         const float* vals, const uint8_t* cell_group, int32_t n_groups,
                         int32_t p_loc, int32_t p_scale, const double* xu_loc, const double* xu_scale,
                         const double* offset, const double* n_k, const double* pinv_loc,
                         const int32_t* loc_group, int32_t n_loc_groups,
                         const dx_fit_opts* opts, int32_t coef,
                         double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                         int32_t* niter, int32_t* flags, double* pval, double* log_pval, double* sd) {
    DX_CHECK(n_cells >= 1 && n_genes >= 0 && indptr && cell_group, "dx_wald_grouped_host: bad arguments");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_wald_grouped_host: n_groups must be in 1..255");
    DX_CHECK(coef >= 0 && coef < p_loc, "dx_wald_grouped_host: coefficient index out of range");
    DX_CHECK(!bad_opts(opts), "dx_wald_grouped_host: bad options");
    const int64_t nnz = indptr[n_genes];
    const int64_t G = n_genes;
    DevBuf d_indptr, d_rows, d_vals, d_grp, d_hist, d_oc, d_ov, d_og, d_xl, d_xs, d_off, d_nk, d_pinv, d_lg;
    DevBuf d_a, d_b, d_fi, d_ll, d_jac, d_ni, d_fl, d_p, d_lp, d_sd;
    cudaStream_t s;
    DX_CUDA(cudaStreamCreate(&s), "stream");
    int rc = DX_OK;
#define HX(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail_cuda(e__, "dx_wald_grouped_host"); goto done; } } while (0)
    {
        HX(d_indptr.alloc((G + 1) * 8)); HX(d_rows.alloc(nnz * 4)); HX(d_vals.alloc(nnz * 4)); HX(d_grp.alloc(n_cells));
        HX(d_hist.alloc((size_t)G * n_groups * DX_HIST_BINS * 4)); HX(d_oc.alloc(G * 4)); HX(d_ov.alloc(nnz * 4)); HX(d_og.alloc(nnz));
        HX(d_xl.alloc((size_t)n_groups * p_loc * 8)); HX(d_xs.alloc((size_t)n_groups * p_scale * 8));
        HX(d_off.alloc((size_t)n_groups * 8)); HX(d_nk.alloc((size_t)n_groups * 8)); HX(d_pinv.alloc((size_t)n_groups * p_loc * 8));
        HX(d_a.alloc((size_t)G * p_loc * 8)); HX(d_b.alloc((size_t)G * p_scale * 8)); HX(d_fi.alloc((size_t)G * p_loc * p_loc * 8));
        HX(d_ll.alloc(G * 8)); HX(d_jac.alloc(G * 8)); HX(d_ni.alloc(G * 4)); HX(d_fl.alloc(G * 4));
        HX(d_p.alloc(G * 8)); HX(d_lp.alloc(G * 8)); HX(d_sd.alloc(G * 8));
        HX(cudaMemcpyAsync(d_indptr.p, indptr, (G + 1) * 8, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_rows.p, rows, nnz * 4, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_vals.p, vals, nnz * 4, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_grp.p, cell_group, n_cells, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_xl.p, xu_loc, (size_t)n_groups * p_loc * 8, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_xs.p, xu_scale, (size_t)n_groups * p_scale * 8, cudaMemcpyHostToDevice, s));
        if (offset) HX(cudaMemcpyAsync(d_off.p, offset, (size_t)n_groups * 8, cudaMemcpyHostToDevice, s));
        HX(cudaMemcpyAsync(d_nk.p, n_k, (size_t)n_groups * 8, cudaMemcpyHostToDevice, s));
        HX(d_lg.alloc((size_t)n_groups * 4));
        if (pinv_loc) HX(cudaMemcpyAsync(d_pinv.p, pinv_loc, (size_t)(n_loc_groups > 0 ? n_loc_groups : n_groups) * p_loc * 8, cudaMemcpyHostToDevice, s));
        if (loc_group) HX(cudaMemcpyAsync(d_lg.p, loc_group, (size_t)n_groups * 4, cudaMemcpyHostToDevice, s));
        if (opts->init_a == DX_INIT_GIVEN) HX(cudaMemcpyAsync(d_a.p, a_var, (size_t)G * p_loc * 8, cudaMemcpyHostToDevice, s));
        if (opts->init_b == DX_INIT_GIVEN) HX(cudaMemcpyAsync(d_b.p, b_var, (size_t)G * p_scale * 8, cudaMemcpyHostToDevice, s));
        rc = dx_group_suffstats(G, (const int64_t*)d_indptr.p, (const int32_t*)d_rows.p, (const float*)d_vals.p,
                                (const uint8_t*)d_grp.p, n_groups, (uint32_t*)d_hist.p, (int32_t*)d_oc.p,
                                (float*)d_ov.p, (uint8_t*)d_og.p, s);
        if (rc) goto done;
        rc = dx_nb_fit_grouped(G, n_groups, p_loc, p_scale, (const double*)d_xl.p, (const double*)d_xs.p,
                               offset ? (const double*)d_off.p : nullptr, (const double*)d_nk.p,
                               pinv_loc ? (const double*)d_pinv.p : nullptr,
                               loc_group ? (const int32_t*)d_lg.p : nullptr, n_loc_groups, (const uint32_t*)d_hist.p,
                               (const int32_t*)d_oc.p, (const int64_t*)d_indptr.p, (const float*)d_ov.p,
                               (const uint8_t*)d_og.p, opts, (double*)d_a.p, (double*)d_b.p, (double*)d_fi.p,
                               (double*)d_ll.p, (double*)d_jac.p, (int32_t*)d_ni.p, (int32_t*)d_fl.p, s);
        if (rc) goto done;
        rc = dx_wald_from_fit(G, p_loc, coef, (const double*)d_a.p, (const double*)d_fi.p, nullptr, (double*)d_sd.p,
                              (double*)d_p.p, (double*)d_lp.p, s);
        if (rc) goto done;
        HX(cudaMemcpyAsync(a_var, d_a.p, (size_t)G * p_loc * 8, cudaMemcpyDeviceToHost, s));
        HX(cudaMemcpyAsync(b_var, d_b.p, (size_t)G * p_scale * 8, cudaMemcpyDeviceToHost, s));
        if (fim_inv) HX(cudaMemcpyAsync(fim_inv, d_fi.p, (size_t)G * p_loc * p_loc * 8, cudaMemcpyDeviceToHost, s));
        if (ll) HX(cudaMemcpyAsync(ll, d_ll.p, G * 8, cudaMemcpyDeviceToHost, s));
        if (jac) HX(cudaMemcpyAsync(jac, d_jac.p, G * 8, cudaMemcpyDeviceToHost, s));
        if (niter) HX(cudaMemcpyAsync(niter, d_ni.p, G * 4, cudaMemcpyDeviceToHost, s));
        if (flags) HX(cudaMemcpyAsync(flags, d_fl.p, G * 4, cudaMemcpyDeviceToHost, s));
        if (pval) HX(cudaMemcpyAsync(pval, d_p.p, G * 8, cudaMemcpyDeviceToHost, s));
        if (log_pval) HX(cudaMemcpyAsync(log_pval, d_lp.p, G * 8, cudaMemcpyDeviceToHost, s));
        if (sd) HX(cudaMemcpyAsync(sd, d_sd.p, G * 8, cudaMemcpyDeviceToHost, s));
        HX(cudaStreamSynchronize(s));
    }
done:
#undef HX
    cudaStreamDestroy(s);
    return rc;
}

}  // extern "C"
