// dx_internal.h -- launcher prototypes shared by the translation units of libdxb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dxb200.h"

#define DX_INGEST_THREADS 128

int dx_ingest_private_bins(int K);
// zero-initialised work-queue counter for one launch (ring of 4096 slots, safe across concurrent streams)
cudaError_t dx_queue_acquire(cudaStream_t stream, unsigned int** out);
cudaError_t dx_launch_ingest(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                             const uint8_t* cell_group, int K, uint32_t* hist, int32_t* ovf_count,
                             float* ovf_val, uint8_t* ovf_group, const int32_t* gene_order, int flags, int sm_count,
                             cudaStream_t stream, const int32_t* pieces = nullptr, int64_t n_pieces = 0,
                             const int32_t* multi_genes = nullptr, int64_t n_multi = 0);
cudaError_t dx_launch_gene_moments(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                                   const double* inv_sf, double* sum1, double* sum2, cudaStream_t stream);
// what K_fit does for a gene after finalize (both nullable parts): the single-coefficient Wald row, and on several GPUs
// the store of the p-value into every rank's exchange buffer (peer_bufs: host array of `world` device pointers)
struct DxFitEpilogueHost {
    int wald_coef;                        // < 0: none
    double *theta, *sd, *pval, *logp;     // device, [n_genes], nullable
    int world, rank;                      // world <= 1: no exchange
    int64_t gene_offset, slot_bytes;
    void* const* peer_bufs;
};
cudaError_t dx_launch_fit_grouped(int64_t n_genes, int K, int p, int ps,
                                  const double* xu_loc, const double* xu_scale, const double* offset,
                                  const double* n_k, const double* pinv_loc, const int32_t* loc_group, int n_loc_groups,
                                  const uint32_t* tail, const int32_t* ovf_count, const int64_t* indptr,
                                  const float* ovf_val, const uint8_t* ovf_group, const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, int64_t coef_stride, const int32_t* ovf_packed,
                                  const DxFitEpilogueHost* epi, cudaStream_t stream);
cudaError_t dx_launch_pack_overflow(int64_t n_genes, int K, int min_entries, const int32_t* ovf_count,
                                    const int64_t* indptr, float* ovf_val, uint8_t* ovf_group, int32_t* ovf_packed,
                                    int sm_count, cudaStream_t stream);
cudaError_t dx_launch_fit_percell(int64_t n_genes, int64_t n_cells, int p, int ps,
                                  const double* design_loc, const double* design_scale, const double* log_sf,
                                  const int64_t* indptr, const int32_t* rows, const float* vals,
                                  const dx_fit_opts* opts,
                                  double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                                  int32_t* niter, int32_t* flags, int sm_count, cudaStream_t stream);
cudaError_t dx_fit_percell_counters_impl(unsigned long long* out, int reset);
cudaError_t dx_launch_wald(int64_t n, const double* theta, const double* sd, double* pval, double* logp, cudaStream_t s);
cudaError_t dx_launch_wald_from_fit(int64_t G, int p, int coef, const double* a_var, const double* fim_inv,
                                    double* theta, double* sd, double* pval, double* logp, int64_t coef_stride, cudaStream_t s);
cudaError_t dx_launch_wald_chisq(int64_t G, int k, const double* theta, const double* cov,
                                 double* stat, double* pval, double* logp, cudaStream_t s);
cudaError_t dx_launch_lrt(int64_t n, const double* llf, const double* llr, int ddf, double* pval, double* logp, cudaStream_t s);
cudaError_t dx_launch_two_coef_z(int64_t n, const double* t0, const double* t1, const double* sd0, const double* sd1,
                                 double* pval, cudaStream_t s);
cudaError_t dx_launch_t_moments(int64_t n, const double* mu0, const double* mu1, const double* var0, const double* var1,
                                double n0, double n1, double* pval, cudaStream_t s);
cudaError_t dx_launch_two_group(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* tail,
                                const int32_t* ovf_count, const int64_t* indptr, const float* ovf_val, const uint8_t* ovf_group,
                                int do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, cudaStream_t s);
cudaError_t dx_launch_two_group_big(int64_t G, int K, int ga, int gb, double n0, double n1, const uint32_t* tail,
                                    const int32_t* ovf_count, const int64_t* indptr, const float* ovf_val,
                                    const uint8_t* ovf_group, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags,
                                    void* work, int64_t work_bytes, int64_t slice_bytes, int sm_count, cudaStream_t s);
size_t dx_bh_work_bytes_impl(int64_t n);
cudaError_t dx_launch_bh(int64_t n, const double* pval, double* qval, double* work, cudaStream_t s);
cudaError_t dx_launch_bh_exchange(int64_t n_total, void* local_buf, int world, int64_t slot_bytes, double* pval_all,
                                  double* qval, double* work, cudaStream_t s);
cudaError_t dx_launch_bh_sorted(int64_t n, const double* p_sorted, const int64_t* order, double* qval, double* work,
                                cudaStream_t s);
cudaError_t dx_launch_peer_allgather(int64_t n, const double* src, void* const* peer_bufs, int world, int rank,
                                     int64_t slot_bytes, double* dst, int sm_count, cudaStream_t stream);

// K8 (dx_transpose.cu): cell-major CSR -> gene-major shard of the gene range [lo, lo + n_out)
size_t dx_transpose_work_bytes_impl(int64_t n_cells, int n_out, int sm_count);
cudaError_t dx_launch_csr_to_csc(int64_t n_cells, const int64_t* indptr, const int32_t* cols, const float* vals, int lo, int n_out,
                                 int64_t* out_ptr, int32_t* out_rows, float* out_vals, void* work, int sm_count,
                                 cudaStream_t stream);
// staged variant (append-only streams: no write amplification); n_cells < 2^27, column indices sorted inside a cell
size_t dx_transpose_staged_work_bytes_impl(int64_t n_cells, int n_out, int64_t nnz, int sm_count);
cudaError_t dx_launch_csr_to_csc_staged(int64_t n_cells, int64_t nnz, const int64_t* indptr, const int32_t* cols, const float* vals,
                                        int lo, int n_out, int64_t* out_ptr, int32_t* out_rows, float* out_vals, void* work,
                                        int64_t work_bytes, int sm_count, cudaStream_t stream);
