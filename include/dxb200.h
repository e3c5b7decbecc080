/*
 * dxb200.h -- C ABI of libdxb200.so: the B200 (sm_100a) implementation of diffxpy's per-gene
 * negative-binomial GLM fit + Wald / LRT statistics and the model-free Welch-t / Wilcoxon tests.
 *
 * Boundary replaced: the batchglm Estimator protocol that diffxpy drives from
 *   /root/reference/diffxpy/testing/tests.py:26-249   (_fit: InputDataGLM -> Estimator -> initialize ->
 *                                                       train_sequence -> finalize)
 * and the statistics that follow it
 *   /root/reference/diffxpy/stats/stats.py:9-39, 42-75, 116-178, 181-218, 221-282
 *   /root/reference/diffxpy/testing/det.py:783-811, 1545-1620, 1672-1743
 *   /root/reference/diffxpy/testing/correction.py:5-24
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is DEVICE memory owned by the caller unless the
 *     function name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - all calls are stream-ordered and never synchronise, except *_host convenience entry points.
 *   - return 0 on success, a DX_ERR_* code otherwise; dx_last_error() gives the message (thread local).
 *   - the count matrix shard is gene-major compressed (CSC of the cells x genes matrix):
 *       indptr[G+1] int64, rows[nnz] int32 (cell index, ascending inside a gene), vals[nnz] float32.
 *     When rows and vals are 16-byte aligned (any cudaMalloc / pool allocation) K_ingest streams them with TMA bulk
 *     copies in 16-byte units: both arrays must then be readable up to the next 16-byte boundary past their last
 *     element (allocation granularity guarantees it; the extra bytes are never used).  Unaligned arrays take the
 *     plain-load kernel.
 *   - work queues of the dynamic-scheduling kernels live in a 1 KB device array owned by the library (one slot per
 *     launch, ring of 256): launches on different streams do not interfere.
 */
#ifndef DXB200_H
#define DXB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DX_OK 0
#define DX_ERR_INVALID 1
#define DX_ERR_CUDA 2
#define DX_ERR_UNSUPPORTED 3

/* per-gene status bits (column `err` of the Wald summary, det.py:839-840) */
#define DX_FLAG_CONVERGED 1
#define DX_FLAG_SINGULAR_FIM 2
#define DX_FLAG_SCALE_FROZEN 4
#define DX_FLAG_NONFINITE 8
#define DX_FLAG_ALL_ZERO 16
#define DX_FLAG_RANK_OVERFLOW 32 /* rank test: more distinct non-count values than the on-chip sort holds */

#define DX_HIST_BINS 64      /* direct-addressed count bins y = 1..64 per (gene, design group) */
#define DX_MAX_GROUPS 255    /* distinct design rows handled by the grouped path; group id 255 = skip cell */
#define DX_SKIP_GROUP 255
#define DX_MAX_COEF 32       /* max coefficients of the location / scale model */
#define DX_MAX_SCALE_COEF_PERCELL 8   /* per-cell path: max coefficients of a scale model with numeric covariates */

/* init modes: batchglm init_par (tests.py:36-37, 222-229) */
#define DX_INIT_GIVEN 0        /* a_var / b_var hold the start values */
#define DX_INIT_STANDARD 1     /* a: intercept = log(mean); b: intercept = log(method-of-moments size) */
#define DX_INIT_CLOSED_FORM 2  /* a: pinv(unique design) @ log(group means)  (location model only) */
#define DX_INIT_ALL_ZERO 3

/* dx_fit_opts.reserved bits */
#define DX_OPT_SCALE_CONST 1   /* p_scale == 1 and design_scale is the same value in every cell / design group (the caller
                                  verified it).  Per-cell path: the dispersion is evaluated once per pass instead of once per
                                  cell.  Grouped path: selects the register-resident kernel when n_groups <= 32, p_loc <= 16 */

typedef struct dx_fit_opts {
    int32_t max_steps;        /* train(): max_steps, DEFAULT strategy 1000 */
    int32_t update_b_freq;    /* train(): update_b_freq, DEFAULT strategy 5 */
    int32_t train_loc;        /* 0 when the closed-form location init is exact (saturated design) */
    int32_t train_scale;      /* 0 == quick_scale */
    int32_t init_a;           /* DX_INIT_* */
    int32_t init_b;           /* DX_INIT_* (closed form for b is resolved by the host into GIVEN) */
    int32_t newton_max_iter;  /* scale Newton iterations per scale step (100) */
    int32_t reserved;         /* option bits: DX_OPT_SCALE_CONST */
    double lltol;             /* batchglm LLTOL_BY_FEATURE 1e-10 */
    double newton_xtol;       /* 1e-8 */
} dx_fit_opts;

int dx_version(void);
const char* dx_last_error(void);
int dx_device_sm_count(int* out);

/* ---- K_ingest: one pass over the CSC shard -> per (gene, design group) count histograms ---------
 * Replaces the per-iteration dense passes of batchglm (tests.py:242-245) and x.mean / x.power(2).mean
 * (det.py:1567-1584, 1693-1709) by sufficient statistics.
 *   cell_group[n_cells] : design-group id of every cell (0..n_groups-1, 255 = ignore the cell)
 *   hist[G][n_groups][64] uint32 : TAIL COUNTS c(j) = number of cells of the group whose count y is an
 *                         integer with j < y <= 64 (j = 0..63).  The NB likelihood needs sum_j c(j) f(r + j);
 *                         the bins are recovered as #{y == j+1} = c(j) - c(j+1).
 *   ovf_*               : every other non-zero entry (non-integer, negative or > 64) of gene g is
 *                         copied to ovf_val/ovf_group[indptr[g] .. indptr[g]+ovf_count[g]), in a
 *                         deterministic order
 */
int dx_group_suffstats(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                       const uint8_t* cell_group, int32_t n_groups,
                       uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group,
                       void* stream);
/* Same, with what a resident shard can know about itself:
 *   gene_order[n_genes] (nullable): permutation of 0..n_genes-1, the order in which genes are handed to the SMs
 *                       (largest first shortens the tail of the launch); results do not depend on it
 *   flags: DX_SHARD_INTEGER_COUNTS = every stored value is a non-negative integer below 2^24 (verified by the
 *          caller); the kernel then skips the per-entry integrality test
 *          DX_GROUPS_FIT_U16 = every group of cell_group has at most 65535 cells; the on-chip histograms then use
 *          16-bit bins, which lets more warps stay resident */
#define DX_SHARD_INTEGER_COUNTS 1
#define DX_GROUPS_FIT_U16 2      /* no design group has more than 65535 cells (verified by the caller): 16-bit on-chip bins */
#define DX_SHARD_NO_OVERFLOW 4   /* every stored value is an integer in 1..64 (verified by the caller): no gene has overflow
                                    entries, so the pass that rewrites the overflow lists of genes counted in parts is skipped */
int dx_group_suffstats_ex(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                          const uint8_t* cell_group, int32_t n_groups,
                          uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group,
                          const int32_t* gene_order, int32_t flags, void* stream);

/* Same, driven by a work list that is prepared once with the resident shard (like gene_order): the unit of work is a
 * PIECE, a whole gene or one of several equal parts of a gene that is too large to be one unit -- with few genes per GPU
 * (a gene range of a multi-GPU run) the largest gene would otherwise set the duration of the pass.
 *   pieces[n_pieces][4] int32, 16-byte aligned: {gene, first element relative to the gene's start (a multiple of 256),
 *                       number of elements, 1 if the gene has several pieces else 0}; every gene appears at least once
 *                       (an empty gene as one piece of 0 elements); the queue hands them out in list order (largest
 *                       first shortens the tail)
 *   multi_genes[n_multi] the genes with several pieces: their rows are zeroed first, their parts add tail counts with
 *                       atomics, and those that turn out to hold overflow entries are counted once more as whole genes
 *                       (so the overflow lists are the same as in the gene-wise pass).  Results do not depend on the list. */
int dx_group_suffstats_pieces(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                              const uint8_t* cell_group, int32_t n_groups,
                              uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group,
                              const int32_t* pieces, int64_t n_pieces, const int32_t* multi_genes, int64_t n_multi,
                              int32_t flags, void* stream);

/* ---- pageable host array -> device array, staged through pinned buffers by eight host threads (each with its own stream;
 * the caller's stream is ordered after them).  src_kind 0: n_elems bytes as they are; 1: float64 -> float32; 2: int64 ->
 * int32 (what scipy matrices may hold).  Blocks until the host array has been read; the device copy completes in stream
 * order.  This is the upload of the count matrix diffxpy hands over as host memory (utils.py:20-33). */
int dx_upload(void* dst_device, const void* src_host, int64_t n_elems, int32_t src_kind, void* stream);

/* ---- K8: cell-major CSR -> gene-major shard (the layout every kernel above streams) ---------------
 * Replaces the per-chunk densification of column blocks from the row-major matrix inside batchglm's InputDataGLM
 * (constructed at tests.py:177-191; data arrives as scipy CSR / AnnData.X, utils.py:20-33).  Device pointers.
 *   indptr[n_cells + 1] int64, cols[nnz] int32 (a cell holds a gene at most once: canonical CSR), vals[nnz] float32
 *   gene range [gene_lo, gene_lo + n_genes_out): the entries of other genes are skipped (gene-sharded runs)
 *   out_indptr[n_genes_out + 1] int64, out_rows / out_vals sized for the entries of the range (<= nnz); cells ascending
 *   inside every gene (stable, deterministic, no atomics on the output)
 *   work: dx_csr_to_csc_work_bytes(n_cells, n_genes_out) bytes.  n_genes_out <= 56320. */
int dx_csr_to_csc_work_bytes(int64_t n_cells, int32_t n_genes_out, int64_t* bytes);
int dx_csr_to_csc(int64_t n_cells, const int64_t* indptr, const int32_t* cols, const float* vals,
                  int32_t gene_lo, int32_t n_genes_out, int64_t* out_indptr, int32_t* out_rows, float* out_vals,
                  void* work, void* stream);
/* Staged variant, same result: the entries first go to append-only segments per (block of 32 genes, chunk of cells), then
 * from there to their genes, so every write stream stays sector-friendly (dx_csr_to_csc writes 4-byte elements to one open
 * position per gene and CTA: 7x the algorithmic DRAM traffic on a 100k x 20k matrix).  Needs the column indices sorted
 * inside every cell (canonical CSR), n_cells < 2^27 and a workspace of dx_csr_to_csc_staged_work_bytes(n_cells,
 * n_genes_out, nnz) bytes (about 8 nnz). */
int dx_csr_to_csc_staged_work_bytes(int64_t n_cells, int32_t n_genes_out, int64_t nnz, int64_t* bytes);
int dx_csr_to_csc_staged(int64_t n_cells, int64_t nnz, const int64_t* indptr, const int32_t* cols, const float* vals,
                         int32_t gene_lo, int32_t n_genes_out, int64_t* out_indptr, int32_t* out_rows, float* out_vals,
                         void* work, int64_t work_bytes, void* stream);

/* ---- K_fit (grouped): the whole batchglm training schedule + finalize, one warp per gene ----------
 * Replaces Estimator.initialize / train_sequence / finalize (tests.py:222-248) when the rows of
 * [design_loc | design_scale | log size factor] take n_groups <= 255 distinct values.
 *   xu_loc[n_groups][p_loc], xu_scale[n_groups][p_scale] : distinct design rows times constraints
 *   offset[n_groups] (nullable) : log size factor of the group;  n_k[n_groups] : cells per group
 *   closed-form location init (batchglm closedform_glm_mean): loc_group[n_groups] maps a group to its
 *   distinct LOCATION design row (nullable = identity), pinv_loc[p_loc][n_loc_groups] is the pseudo-inverse
 *   of those distinct rows; both nullable unless init_a == CLOSED_FORM
 *   a_var[p_loc][G], b_var[p_scale][G] in/out;  fim_inv[G][p_loc][p_loc];  ll, jac [G];  niter, flags [G]
 */
int dx_nb_fit_grouped(int64_t n_genes, int32_t n_groups, int32_t p_loc, int32_t p_scale,
                      const double* xu_loc, const double* xu_scale, const double* offset, const double* n_k,
                      const double* pinv_loc, const int32_t* loc_group, int32_t n_loc_groups,
                      const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                      const float* ovf_val, const uint8_t* ovf_group,
                      const dx_fit_opts* opts,
                      double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                      int32_t* niter, int32_t* flags, void* stream);

/* ---- K_pack: multiplicity compression of long overflow lists (optional, between K_ingest and K_fit) ---------------
 * Genes whose counts exceed 64 in many cells (mean expression in the tens or hundreds) carry overflow lists as long
 * as the cell count, which K_fit walks at every likelihood evaluation.  dx_pack_overflow rewrites, IN PLACE, the list
 * of every gene with at least min_entries entries (<= 0: DX_PACK_MIN_ENTRIES) as records (value, multiplicity) plus
 * per-group sums -- layout in diffxpy_b200/csrc/dx_overflow.cu -- and writes ovf_packed[n_genes][2], the two record
 * counts of a packed gene or {0, 0} for a gene it left alone.  Packed arrays are only understood by
 * dx_nb_fit_grouped_packed (same arguments as dx_nb_fit_grouped plus ovf_packed; a null ovf_packed = plain lists):
 * run dx_group_pair_tests / dx_two_group_tests on the statistics BEFORE packing them.  Results agree with the plain
 * path to summation order (<= 1e-12 relative on the log-likelihood); nothing changes for genes without long lists.
 * Replaces nothing in the reference: it bounds the cost of tests.py:242-245 on highly expressed genes. */
#define DX_PACK_MIN_ENTRIES 256
int dx_pack_overflow(int64_t n_genes, int32_t n_groups, int32_t min_entries, const int32_t* ovf_count,
                     const int64_t* indptr, float* ovf_val, uint8_t* ovf_group, int32_t* ovf_packed, void* stream);
int dx_nb_fit_grouped_packed(int64_t n_genes, int32_t n_groups, int32_t p_loc, int32_t p_scale,
                             const double* xu_loc, const double* xu_scale, const double* offset, const double* n_k,
                             const double* pinv_loc, const int32_t* loc_group, int32_t n_loc_groups,
                             const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                             const float* ovf_val, const uint8_t* ovf_group, const int32_t* ovf_packed,
                             const dx_fit_opts* opts,
                             double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                             int32_t* niter, int32_t* flags, void* stream);

/* ---- K_fit + Wald in one launch (the product path of de.test.wald with one tested coefficient) ------------------------
 * dx_nb_fit_grouped_packed followed, per gene and inside the same kernel, by the single-coefficient Wald row
 * (det.py:783-802, stats.py:181-218): theta = a_var[coef], sd = sqrt(max(fim_inv[g][coef][coef], tiny)), pval, log_pval
 * ([n_genes] each, nullable).  A NaN variance (singular Fisher information) gives NaN rows, as in the reference.
 * Several GPUs (world > 1): the p-value of gene g is ALSO stored into the current slot of every rank's exchange buffer
 * (dx_exchange_create below) at index gene_offset + g, and the last CTA publishes this rank's sequence flag;
 * dx_bh_qvalues_exchange on every rank then waits for all flags and corrects the complete vector -- no separate gather
 * launch.  peer_bufs: HOST array of `world` device pointers (own buffer at [rank]); world == 1: no exchange, peer_bufs
 * may be null.
 * opts->reserved & DX_OPT_SCALE_CONST (p_scale == 1 and the same dispersion design value in every group, verified by
 * the caller) with n_groups <= 32 and p_loc <= 16 selects the register-resident kernel (dx_fit_fast.cu).  * ovf_val / ovf_group may be NULL when the caller knows that no gene has overflow entries (every stored value an integer in
 * 1..64, DX_SHARD_NO_OVERFLOW): the launch that fits genes with overflow entries is then skipped. */
int dx_nb_fit_wald_grouped(int64_t n_genes, int32_t n_groups, int32_t p_loc, int32_t p_scale,
                           const double* xu_loc, const double* xu_scale, const double* offset, const double* n_k,
                           const double* pinv_loc, const int32_t* loc_group, int32_t n_loc_groups,
                           const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                           const float* ovf_val, const uint8_t* ovf_group, const int32_t* ovf_packed,
                           const dx_fit_opts* opts,
                           double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                           int32_t* niter, int32_t* flags,
                           int32_t coef, double* theta, double* sd, double* pval, double* log_pval,
                           void* const* peer_bufs, int32_t world, int32_t rank, int64_t slot_bytes, int64_t gene_offset,
                           void* stream);
/* Benjamini-Hochberg q-values (correction.py:5-24) of the n_total p-values that dx_nb_fit_wald_grouped of ALL ranks
 * stored into this rank's exchange buffer: waits (on the stream) for every rank's flag, corrects, marks the exchange
 * done.  pval_all (nullable for n_total <= 32768): copy of the gathered p-values.  work: dx_bh_work_bytes(n_total). */
int dx_bh_qvalues_exchange(int64_t n_total, void* local_buf, int32_t world, int64_t slot_bytes, double* pval_all,
                           double* qval, double* work, void* stream);

/* ---- multi-GPU: all-gather of per-gene result rows over NVLink peer memory (one node, one process per GPU) ----------
 * Replaces nothing in the reference (it is single-process); it is the one exchange of the sharded path:
 * the p-values of every rank's genes before the Benjamini-Hochberg step (det.py:108-117 needs all of them).
 *   dx_exchange_create  allocates this rank's exchange buffer (512 + 2 * slot_bytes) and returns a CUDA IPC handle
 *                       (DX_IPC_HANDLE_BYTES bytes) that the host side sends to the other ranks with its own plumbing;
 *   dx_exchange_open    maps a peer's buffer;  dx_exchange_close / dx_exchange_destroy undo the two;
 *   dx_peer_allgather_f64  stores src[0..n) into the current slot of EVERY rank's buffer at row `rank`, publishes the
 *                       exchange's sequence number in the peers' flag rows, waits (on the stream) for the flags of all
 *                       ranks and copies the world * n gathered doubles to dst.  peer_bufs: HOST array of `world`
 *                       device pointers, own buffer at [rank].  Every rank makes the same sequence of calls with the
 *                       same n; the sequence number is kept in the buffer, so the call can sit in a replayed CUDA graph.
 *   dx_exchange_status  (synchronous) 0, or 1 + the rank whose flag did not arrive within 5 s. */
#define DX_IPC_HANDLE_BYTES 64
int dx_exchange_create(int64_t slot_bytes, void** local_buf, void* ipc_handle_out);
int dx_exchange_open(const void* ipc_handle, void** peer_buf);
int dx_exchange_close(void* peer_buf);
int dx_exchange_destroy(void* local_buf);
int dx_exchange_status(const void* local_buf, int32_t* failed_rank_plus_one);
int dx_peer_allgather_f64(int64_t n, const double* src, void* const* peer_bufs, int32_t world, int32_t rank,
                          int64_t slot_bytes, double* dst, void* stream);

/* ---- K_fit (per cell): same schedule for designs that do not collapse into groups ----------------
 * General path: arbitrary numeric design, per-cell size factors (tests.py:177-191 arguments).  A CTA owns a tile of
 * genes and walks the cells once per model evaluation; X^T W X is a register-blocked fp64 contraction of per-(cell, gene)
 * weights with the design-column products, which are shared by all genes of the tile (csrc/dx_fit_tile.cu).
 * p_loc <= DX_MAX_COEF (32), p_scale <= DX_MAX_SCALE_COEF_PERCELL (8).
 *   design_loc[p_loc][n_cells], design_scale[p_scale][n_cells] (coefficient-major, constraints applied)
 *   log_sf[n_cells] nullable.  a_var / b_var must hold the start values (DX_INIT_GIVEN).
 */
int dx_nb_fit_percell(int64_t n_genes, int64_t n_cells, int32_t p_loc, int32_t p_scale,
                      const double* design_loc, const double* design_scale, const double* log_sf,
                      const int64_t* indptr, const int32_t* rows, const float* vals,
                      const dx_fit_opts* opts,
                      double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                      int32_t* niter, int32_t* flags, void* stream);

/* diagnostics of the per-cell path (synchronous): out4 = {location passes, dispersion passes, dispersion epochs, gene
 * tiles} summed over the CTAs since the last reset; a pass = one sweep of a gene tile over all cells. */
int dx_fit_percell_counters(uint64_t* out4, int32_t reset);

/* ---- column moments of the CSC shard (initialisation of the per-cell path, det.py:775-781 mean) ---
 *   inv_sf[n_cells] nullable: moments of x / size_factor.  sum1, sum2, nnz: [G] */
int dx_gene_moments(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                    const double* inv_sf, double* sum1, double* sum2, void* stream);

/* ---- K_stats: stats.py restated on device -------------------------------------------------------
 * wald_test (stats.py:181-218): pval = 2 (1 - Phi(|theta/sd|)), log_pval = accurate log of the tail. */
int dx_wald_test(int64_t n, const double* theta, const double* sd, double* pval, double* log_pval, void* stream);
/* Wald rows straight from the fit: theta = a_var[coef], sd = sqrt(max(fim_inv[g][coef][coef], tiny))
 * (det.py:792-802) */
int dx_wald_from_fit(int64_t n_genes, int32_t p_loc, int32_t coef, const double* a_var, const double* fim_inv,
                     double* theta, double* sd, double* pval, double* log_pval, void* stream);
/* wald_test_chisq (stats.py:221-282): theta[k][G], cov[G][k][k]; NaN where cov is numerically singular */
int dx_wald_chisq(int64_t n_genes, int32_t k, const double* theta, const double* cov,
                  double* stat, double* pval, double* log_pval, void* stream);
/* likelihood_ratio_test (stats.py:9-39) */
int dx_lrt(int64_t n, const double* ll_full, const double* ll_reduced, int32_t delta_df,
           double* pval, double* log_pval, void* stream);
/* two_coef_z_test (stats.py:285-326) */
int dx_two_coef_z(int64_t n, const double* t0, const double* t1, const double* sd0, const double* sd1,
                  double* pval, void* stream);
/* t_test_moments (stats.py:116-178) */
int dx_t_test_moments(int64_t n, const double* mu0, const double* mu1, const double* var0, const double* var1,
                      double n0, double n1, double* pval, void* stream);

/* Benjamini-Hochberg q-values of the non-NaN p-values (correction.py:5-24 -> statsmodels fdr_bh arithmetic,
 * bit-identical values), without a sort: ranks by counting over monotone buckets of p, then a running minimum from
 * the right.  work: dx_bh_work_bytes(n) bytes of device memory (20 n + 8 ceil(n / 1024) + 0.17 MB), 8-byte aligned. */
int dx_bh_work_bytes(int64_t n, int64_t* bytes);
int dx_bh_qvalues(int64_t n, const double* pval, double* qval, double* work, void* stream);
/* Same q-values for large n, after a device sort done by the caller: p_sorted ascending with NaNs last, order[j] =
 * position of p_sorted[j] in the unsorted vector (what torch.sort / cub return); qval is written unsorted.
 * work: ceil(n / 1024) doubles. */
int dx_bh_qvalues_sorted(int64_t n, const double* p_sorted, const int64_t* order, double* qval, double* work, void* stream);

/* ---- K_twogroup: det.py:1545-1620 (Welch t) and det.py:1672-1743 (Mann-Whitney U) from the
 * 2-group sufficient statistics of dx_group_suffstats (group 0 = first sample, group 1 = second).
 * out[G][DX_TG_NCOL] doubles; u1x2 / tie are the exact integers 2*U1 and sum(t^3 - t). */
#define DX_TG_MEAN0 0
#define DX_TG_MEAN1 1
#define DX_TG_VAR0 2
#define DX_TG_VAR1 3
#define DX_TG_PVAL_T 4
#define DX_TG_PVAL_RANK 5
#define DX_TG_NCOL 6
int dx_two_group_tests(int64_t n_genes, double n0, double n1,
                       const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                       const float* ovf_val, const uint8_t* ovf_group,
                       int32_t do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* stream);

/* The same two tests between any two of the n_groups samples of ONE dx_group_suffstats pass (pairwise / versus-rest
 * comparisons, tests.py:1097-1330, 1333-1506, without re-reading the matrix): sample 0 = group_a; sample 1 =
 * group_b, or the union of all other groups when group_b < 0.  n_a, n_b: cells per sample. */
int dx_group_pair_tests(int64_t n_genes, int32_t n_groups, int32_t group_a, int32_t group_b, double n_a, double n_b,
                        const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                        const float* ovf_val, const uint8_t* ovf_group,
                        int32_t do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* stream);

/* Rank test of the genes dx_group_pair_tests / dx_two_group_tests flagged DX_FLAG_RANK_OVERFLOW (more than 4031 values
 * outside the count bins: normalised or log-transformed data, highly expressed genes).  Same arguments (the arrays of the
 * first call, its out / u1x2 / tie / flags are updated in place and the flag cleared) plus a device workspace: every
 * CTA sorts one gene in a slice of slice_bytes (rounded down to a power of two) of it, so a gene with n_ovf such values
 * needs slice_bytes >= 8 * max(8192, next_pow2(n_ovf + 65)); the launch uses min(n_genes, 2 per SM,
 * work_bytes / slice_bytes) CTAs.  A gene whose list does not fit a slice keeps its flag.
 * Replaces the same reference lines as dx_group_pair_tests (stats.py:42-75 via det.py:1672-1743). */
#define DX_RANK_WORK_MIN_BYTES 65536
int dx_group_pair_tests_big(int64_t n_genes, int32_t n_groups, int32_t group_a, int32_t group_b, double n_a, double n_b,
                            const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                            const float* ovf_val, const uint8_t* ovf_group,
                            double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* work, int64_t work_bytes,
                            int64_t slice_bytes, void* stream);

/* ---- host-buffer entry point (what bench.py's e2e leg and a foreign-language binding call) ---------
 * Full Wald test of one coefficient on a HOST CSC matrix with a grouped design: copies the inputs to
 * the device (chunk c + 1 crosses PCIe while chunk c is ingested and fitted), runs ingest + fit + Wald + the
 * Benjamini-Hochberg correction over all genes (det.py:108-117, correction.py:5-24), copies the per-gene rows back.
 * Pointers are host memory (pinned for full PCIe rate).  out_* are [G] (a_var [p_loc][G], fim_inv [G][p_loc][p_loc]);
 * fim_inv .. qval are nullable.  The device workspace is cached between calls; dx_host_release() frees it. */
int dx_wald_grouped_host(int64_t n_cells, int64_t n_genes, const int64_t* indptr, const int32_t* rows,
                         const float* vals, const uint8_t* cell_group, int32_t n_groups,
                         int32_t p_loc, int32_t p_scale, const double* xu_loc, const double* xu_scale,
                         const double* offset, const double* n_k, const double* pinv_loc,
                         const int32_t* loc_group, int32_t n_loc_groups,
                         const dx_fit_opts* opts, int32_t coef,
                         double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                         int32_t* niter, int32_t* flags, double* pval, double* log_pval, double* sd, double* qval);
int dx_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* DXB200_H */
