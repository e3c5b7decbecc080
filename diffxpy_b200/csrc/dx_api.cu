// dx_api.cu -- the extern "C" surface of libdxb200.so (include/dxb200.h): argument checks, error
// strings, launch plumbing.  No torch types, no allocation except inside the *_host entry point.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
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

// ---- pageable host memory -> device, staged through pinned buffers by several host threads ---------------------------
// Every thread owns a slice of the array, two pinned 4 MB buffers and a stream: it copies (and converts) a piece into a
// buffer while the DMA of its previous piece is in flight.  One thread copies ~9 GB/s, eight keep PCIe 5 x16 (~52 GB/s)
// busy.  The caller's stream is ordered after all of them.  Called through ctypes the whole transfer runs without the
// Python GIL, so an interpreter thread can build design matrices meanwhile (diffxpy_b200/device.py: prefetch).
namespace {
constexpr int UP_THREADS = 16;          // lanes allocated; dx_upload uses up_lanes() of them
constexpr size_t UP_BUF = 4u << 20;
struct UpLane { void* pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; cudaStream_t st = nullptr; cudaEvent_t done = nullptr; };
std::mutex g_up_mutex;
UpLane g_up[UP_THREADS];
bool g_up_ready = false;
cudaEvent_t g_up_start = nullptr;

// lanes in use: DXB200_UPLOAD_THREADS (1..16), default 8 (16-core host of a B200 box: 8, 12 and 16 lanes all reach 35-44 GB/s from pageable memory: bound by the host copy)
int up_lanes() {
    static const int n = [] {
        const char* env = getenv("DXB200_UPLOAD_THREADS");
        const int v = env ? atoi(env) : 8;
        return v < 1 ? 1 : (v > UP_THREADS ? UP_THREADS : v);
    }();
    return n;
}

template <typename S, typename D> void up_convert(D* dst, const S* src, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = (D)src[i]; }
}  // namespace

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
                             nullptr, 0, sm_count_cached(), (cudaStream_t)stream), "dx_group_suffstats");
    return DX_OK;
}

int dx_group_suffstats_ex(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                          const uint8_t* cell_group, int32_t n_groups,
                          uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group,
                          const int32_t* gene_order, int32_t flags, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_group_suffstats_ex: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_group_suffstats_ex: n_groups must be in 1..255");
    DX_CHECK(indptr && cell_group && hist && ovf_count, "dx_group_suffstats_ex: null pointer");
    DX_CHECK((flags & ~(DX_SHARD_INTEGER_COUNTS | DX_GROUPS_FIT_U16 | DX_SHARD_NO_OVERFLOW)) == 0, "dx_group_suffstats_ex: unknown flag");
    DX_CUDA(dx_launch_ingest(n_genes, indptr, rows, vals, cell_group, n_groups, hist, ovf_count, ovf_val, ovf_group,
                             gene_order, flags, sm_count_cached(), (cudaStream_t)stream), "dx_group_suffstats_ex");
    return DX_OK;
}

int dx_group_suffstats_pieces(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                              const uint8_t* cell_group, int32_t n_groups,
                              uint32_t* hist, int32_t* ovf_count, float* ovf_val, uint8_t* ovf_group,
                              const int32_t* pieces, int64_t n_pieces, const int32_t* multi_genes, int64_t n_multi,
                              int32_t flags, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_group_suffstats_pieces: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_group_suffstats_pieces: n_groups must be in 1..255");
    DX_CHECK(indptr && cell_group && hist && ovf_count, "dx_group_suffstats_pieces: null pointer");
    DX_CHECK(pieces && n_pieces >= n_genes && n_multi >= 0 && (n_multi == 0 || multi_genes),
             "dx_group_suffstats_pieces: the work list must hold at least one piece per gene");
    DX_CHECK((reinterpret_cast<uintptr_t>(pieces) & 15u) == 0, "dx_group_suffstats_pieces: the work list must be 16-byte aligned");
    DX_CHECK((flags & ~(DX_SHARD_INTEGER_COUNTS | DX_GROUPS_FIT_U16 | DX_SHARD_NO_OVERFLOW)) == 0,
             "dx_group_suffstats_pieces: unknown flag");
    DX_CUDA(dx_launch_ingest(n_genes, indptr, rows, vals, cell_group, n_groups, hist, ovf_count, ovf_val, ovf_group,
                             nullptr, flags, sm_count_cached(), (cudaStream_t)stream, pieces, n_pieces, multi_genes, n_multi),
            "dx_group_suffstats_pieces");
    return DX_OK;
}

int dx_upload(void* dst_device, const void* src_host, int64_t n_elems, int32_t src_kind, void* stream) {
    // src_kind: 0 = bytes copied as they are (n_elems bytes); 1 = float64 -> float32; 2 = int64 -> int32
    DX_CHECK(dst_device && src_host && n_elems >= 0 && src_kind >= 0 && src_kind <= 2, "dx_upload: bad arguments");
    if (n_elems == 0) return DX_OK;
    std::lock_guard<std::mutex> lock(g_up_mutex);
    int dev = 0;
    DX_CUDA(cudaGetDevice(&dev), "dx_upload");
    if (!g_up_ready) {
        for (int k = 0; k < UP_THREADS; ++k) {
            for (int b = 0; b < 2; ++b) {
                DX_CUDA(cudaHostAlloc(&g_up[k].pin[b], UP_BUF, cudaHostAllocDefault), "dx_upload: pinned staging buffer");
                DX_CUDA(cudaEventCreateWithFlags(&g_up[k].ev[b], cudaEventDisableTiming | cudaEventBlockingSync), "dx_upload");
            }
            DX_CUDA(cudaStreamCreateWithFlags(&g_up[k].st, cudaStreamNonBlocking), "dx_upload");
            DX_CUDA(cudaEventCreateWithFlags(&g_up[k].done, cudaEventDisableTiming), "dx_upload");
        }
        DX_CUDA(cudaEventCreateWithFlags(&g_up_start, cudaEventDisableTiming), "dx_upload");
        g_up_ready = true;
    }
    // the destination may still be in use by work queued on the caller's stream (a recycled allocation): the copy lanes
    // start behind it
    DX_CUDA(cudaEventRecord(g_up_start, (cudaStream_t)stream), "dx_upload");
    for (int k = 0; k < UP_THREADS; ++k) DX_CUDA(cudaStreamWaitEvent(g_up[k].st, g_up_start, 0), "dx_upload");
    const size_t dst_size = src_kind == 0 ? 1 : 4, src_size = src_kind == 0 ? 1 : 8;
    const size_t per_buf = UP_BUF / dst_size;                       // elements per staging buffer
    const size_t n = (size_t)n_elems;
    const int lanes = up_lanes();
    size_t slice = (n + lanes - 1) / lanes;
    slice = (slice + 63) / 64 * 64;
    cudaError_t errs[UP_THREADS];
    std::thread workers[UP_THREADS];
    int used = 0;
    for (int k = 0; k < lanes; ++k) {
        errs[k] = cudaSuccess;
        const size_t lo = (size_t)k * slice, hi = std::min(n, lo + slice);
        if (lo >= hi) break;
        ++used;
        workers[k] = std::thread([=, &errs]() {
            cudaSetDevice(dev);
            UpLane& L = g_up[k];
            int b = 0;
            // the staging buffers may still feed the last copies of the PREVIOUS call (rows, then vals, are uploaded back to
            // back): every buffer waits for its own event before it is overwritten (a never-recorded event is complete)
            bool busy[2] = {true, true};
            for (size_t pos = lo; pos < hi; pos += per_buf, b ^= 1) {
                const size_t m = std::min(per_buf, hi - pos);
                if (busy[b]) { cudaError_t e = cudaEventSynchronize(L.ev[b]); if (e != cudaSuccess) { errs[k] = e; return; } }
                const char* src = static_cast<const char*>(src_host) + pos * src_size;
                if (src_kind == 0) memcpy(L.pin[b], src, m);
                else if (src_kind == 1) up_convert(static_cast<float*>(L.pin[b]), reinterpret_cast<const double*>(src), m);
                else up_convert(static_cast<int32_t*>(L.pin[b]), reinterpret_cast<const int64_t*>(src), m);
                cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst_device) + pos * dst_size, L.pin[b], m * dst_size,
                                                cudaMemcpyHostToDevice, L.st);
                if (e == cudaSuccess) e = cudaEventRecord(L.ev[b], L.st);
                if (e != cudaSuccess) { errs[k] = e; return; }
                busy[b] = true;
            }
            errs[k] = cudaEventRecord(L.done, L.st);
        });
    }
    for (int k = 0; k < used; ++k) workers[k].join();
    for (int k = 0; k < used; ++k) DX_CUDA(errs[k], "dx_upload");
    for (int k = 0; k < used; ++k) DX_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_up[k].done, 0), "dx_upload");
    return DX_OK;
}

int dx_csr_to_csc_work_bytes(int64_t n_cells, int32_t n_genes_out, int64_t* bytes) {
    DX_CHECK(bytes && n_cells >= 0 && n_genes_out >= 0, "dx_csr_to_csc_work_bytes: bad arguments");
    *bytes = (int64_t)dx_transpose_work_bytes_impl(n_cells, n_genes_out, sm_count_cached());
    return DX_OK;
}

int dx_csr_to_csc(int64_t n_cells, const int64_t* indptr, const int32_t* cols, const float* vals,
                  int32_t gene_lo, int32_t n_genes_out, int64_t* out_indptr, int32_t* out_rows, float* out_vals,
                  void* work, void* stream) {
    DX_CHECK(n_cells >= 0 && gene_lo >= 0 && n_genes_out >= 0, "dx_csr_to_csc: bad sizes");
    DX_CHECK(indptr && out_indptr && work, "dx_csr_to_csc: null pointer");
    DX_CHECK((int64_t)n_genes_out * 4 <= 220 * 1024, "dx_csr_to_csc: more than 56320 genes in the range (one shared-memory counter per gene)");
    DX_CUDA(dx_launch_csr_to_csc(n_cells, indptr, cols, vals, gene_lo, n_genes_out, out_indptr, out_rows, out_vals, work,
                                 sm_count_cached(), (cudaStream_t)stream), "dx_csr_to_csc");
    return DX_OK;
}

int dx_csr_to_csc_staged_work_bytes(int64_t n_cells, int32_t n_genes_out, int64_t nnz, int64_t* bytes) {
    DX_CHECK(bytes && n_cells >= 0 && n_genes_out >= 0 && nnz >= 0, "dx_csr_to_csc_staged_work_bytes: bad arguments");
    *bytes = (int64_t)dx_transpose_staged_work_bytes_impl(n_cells, n_genes_out, nnz, sm_count_cached());
    return DX_OK;
}

int dx_csr_to_csc_staged(int64_t n_cells, int64_t nnz, const int64_t* indptr, const int32_t* cols, const float* vals,
                         int32_t gene_lo, int32_t n_genes_out, int64_t* out_indptr, int32_t* out_rows, float* out_vals,
                         void* work, int64_t work_bytes, void* stream) {
    DX_CHECK(n_cells >= 0 && nnz >= 0 && gene_lo >= 0 && n_genes_out >= 0, "dx_csr_to_csc_staged: bad sizes");
    DX_CHECK(indptr && out_indptr && work, "dx_csr_to_csc_staged: null pointer");
    DX_CHECK((int64_t)n_genes_out * 4 <= 220 * 1024, "dx_csr_to_csc_staged: more than 56320 genes in the range");
    DX_CHECK(n_cells < (1ll << 27), "dx_csr_to_csc_staged: at most 2^27 - 1 cells (the record key holds the cell in 27 bits)");
    DX_CHECK(work_bytes >= (int64_t)dx_transpose_staged_work_bytes_impl(n_cells, n_genes_out, nnz, sm_count_cached()),
             "dx_csr_to_csc_staged: workspace smaller than dx_csr_to_csc_staged_work_bytes");
    DX_CUDA(dx_launch_csr_to_csc_staged(n_cells, nnz, indptr, cols, vals, gene_lo, n_genes_out, out_indptr, out_rows, out_vals,
                                        work, work_bytes, sm_count_cached(), (cudaStream_t)stream), "dx_csr_to_csc_staged");
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
                                  flags, 0, nullptr, nullptr, (cudaStream_t)stream), "dx_nb_fit_grouped");
    return DX_OK;
}

int dx_pack_overflow(int64_t n_genes, int32_t n_groups, int32_t min_entries, const int32_t* ovf_count,
                     const int64_t* indptr, float* ovf_val, uint8_t* ovf_group, int32_t* ovf_packed, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_pack_overflow: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_pack_overflow: n_groups must be in 1..255");
    DX_CHECK(ovf_count && indptr && ovf_val && ovf_group && ovf_packed, "dx_pack_overflow: null pointer");
    DX_CUDA(dx_launch_pack_overflow(n_genes, n_groups, min_entries > 0 ? min_entries : DX_PACK_MIN_ENTRIES, ovf_count, indptr,
                                    ovf_val, ovf_group, ovf_packed, sm_count_cached(), (cudaStream_t)stream),
            "dx_pack_overflow");
    return DX_OK;
}

int dx_nb_fit_grouped_packed(int64_t n_genes, int32_t n_groups, int32_t p_loc, int32_t p_scale,
                             const double* xu_loc, const double* xu_scale, const double* offset, const double* n_k,
                             const double* pinv_loc, const int32_t* loc_group, int32_t n_loc_groups,
                             const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                             const float* ovf_val, const uint8_t* ovf_group, const int32_t* ovf_packed,
                             const dx_fit_opts* opts,
                             double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                             int32_t* niter, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0, "dx_nb_fit_grouped_packed: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_nb_fit_grouped_packed: n_groups must be in 1..255");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF && p_scale >= 1 && p_scale <= DX_MAX_COEF,
             "dx_nb_fit_grouped_packed: coefficient counts must be in 1..32");
    DX_CHECK(!bad_opts(opts), "dx_nb_fit_grouped_packed: bad options");
    DX_CHECK(xu_loc && xu_scale && n_k && hist && ovf_count && indptr, "dx_nb_fit_grouped_packed: null input pointer");
    DX_CHECK(a_var && b_var && fim_inv && ll && jac && niter && flags, "dx_nb_fit_grouped_packed: null output pointer");
    DX_CHECK(opts->init_a != DX_INIT_CLOSED_FORM || (pinv_loc && n_loc_groups >= 1 && n_loc_groups <= n_groups),
             "dx_nb_fit_grouped_packed: closed-form init needs pinv_loc and 1 <= n_loc_groups <= n_groups");
    DX_CHECK(opts->init_b != DX_INIT_CLOSED_FORM, "dx_nb_fit_grouped_packed: closed-form init_b must be resolved by the host");
    DX_CUDA(dx_launch_fit_grouped(n_genes, n_groups, p_loc, p_scale, xu_loc, xu_scale, offset, n_k, pinv_loc, loc_group,
                                  n_loc_groups, hist, ovf_count, indptr, ovf_val, ovf_group, opts, a_var, b_var, fim_inv, ll, jac, niter,
                                  flags, 0, ovf_packed, nullptr, (cudaStream_t)stream), "dx_nb_fit_grouped_packed");
    return DX_OK;
}

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
                           void* stream) {
    DX_CHECK(n_genes >= 0, "dx_nb_fit_wald_grouped: n_genes < 0");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_nb_fit_wald_grouped: n_groups must be in 1..255");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF && p_scale >= 1 && p_scale <= DX_MAX_COEF,
             "dx_nb_fit_wald_grouped: coefficient counts must be in 1..32");
    DX_CHECK(!bad_opts(opts), "dx_nb_fit_wald_grouped: bad options");
    DX_CHECK(xu_loc && xu_scale && n_k && hist && ovf_count && indptr, "dx_nb_fit_wald_grouped: null input pointer");
    DX_CHECK(a_var && b_var && fim_inv && ll && jac && niter && flags, "dx_nb_fit_wald_grouped: null output pointer");
    DX_CHECK(opts->init_a != DX_INIT_CLOSED_FORM || (pinv_loc && n_loc_groups >= 1 && n_loc_groups <= n_groups),
             "dx_nb_fit_wald_grouped: closed-form init needs pinv_loc and 1 <= n_loc_groups <= n_groups");
    DX_CHECK(opts->init_b != DX_INIT_CLOSED_FORM, "dx_nb_fit_wald_grouped: closed-form init_b must be resolved by the host");
    DX_CHECK(coef >= 0 && coef < p_loc, "dx_nb_fit_wald_grouped: coefficient index out of range");
    DX_CHECK(world >= 1 && world <= 16 && rank >= 0 && rank < world, "dx_nb_fit_wald_grouped: world must be 1..16, rank inside it");
    DX_CHECK(world == 1 || (peer_bufs && gene_offset >= 0 && (gene_offset + n_genes) * 8 <= slot_bytes),
             "dx_nb_fit_wald_grouped: exchange needs peer_bufs and a slot that holds gene_offset + n_genes doubles");
    DxFitEpilogueHost epi;
    epi.wald_coef = coef; epi.theta = theta; epi.sd = sd; epi.pval = pval; epi.logp = log_pval;
    epi.world = world; epi.rank = rank; epi.gene_offset = gene_offset; epi.slot_bytes = slot_bytes; epi.peer_bufs = peer_bufs;
    DX_CUDA(dx_launch_fit_grouped(n_genes, n_groups, p_loc, p_scale, xu_loc, xu_scale, offset, n_k, pinv_loc, loc_group,
                                  n_loc_groups, hist, ovf_count, indptr, ovf_val, ovf_group, opts, a_var, b_var, fim_inv, ll, jac, niter,
                                  flags, 0, ovf_packed, &epi, (cudaStream_t)stream), "dx_nb_fit_wald_grouped");
    return DX_OK;
}

int dx_bh_qvalues_exchange(int64_t n_total, void* local_buf, int32_t world, int64_t slot_bytes, double* pval_all,
                           double* qval, double* work, void* stream) {
    DX_CHECK(n_total >= 1 && local_buf && qval && work, "dx_bh_qvalues_exchange: bad arguments");
    DX_CHECK(world >= 1 && world <= 16, "dx_bh_qvalues_exchange: world must be 1..16");
    DX_CHECK(n_total * 8 <= slot_bytes, "dx_bh_qvalues_exchange: slot too small for n_total doubles");
    DX_CUDA(dx_launch_bh_exchange(n_total, local_buf, world, slot_bytes, pval_all, qval, work, (cudaStream_t)stream),
            "dx_bh_qvalues_exchange");
    return DX_OK;
}

int dx_nb_fit_percell(int64_t n_genes, int64_t n_cells, int32_t p_loc, int32_t p_scale,
                      const double* design_loc, const double* design_scale, const double* log_sf,
                      const int64_t* indptr, const int32_t* rows, const float* vals, const dx_fit_opts* opts,
                      double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                      int32_t* niter, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0 && n_cells >= 1, "dx_nb_fit_percell: bad sizes");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF, "dx_nb_fit_percell: location coefficients must be in 1..32");
    DX_CHECK(p_scale >= 1 && p_scale <= DX_MAX_SCALE_COEF_PERCELL,
             "dx_nb_fit_percell: a scale model with numeric covariates takes 1..8 coefficients (categorical scale models of "
             "any width collapse into design groups and use dx_nb_fit_grouped)");
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

int dx_fit_percell_counters(uint64_t* out4, int32_t reset) {
    DX_CHECK(out4, "dx_fit_percell_counters: null out");
    DX_CUDA(cudaDeviceSynchronize(), "dx_fit_percell_counters");
    DX_CUDA(dx_fit_percell_counters_impl(reinterpret_cast<unsigned long long*>(out4), reset), "dx_fit_percell_counters");
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
    DX_CUDA(dx_launch_wald_from_fit(n_genes, p_loc, coef, a_var, fim_inv, theta, sd, pval, log_pval, 0, (cudaStream_t)stream),
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

int dx_bh_work_bytes(int64_t n, int64_t* bytes) {
    DX_CHECK(n >= 0 && bytes, "dx_bh_work_bytes: bad arguments");
    *bytes = (int64_t)dx_bh_work_bytes_impl(n);
    return DX_OK;
}

int dx_bh_qvalues(int64_t n, const double* pval, double* qval, double* work, void* stream) {
    DX_CHECK(n >= 0 && pval && qval && work, "dx_bh_qvalues: bad arguments");
    DX_CUDA(dx_launch_bh(n, pval, qval, work, (cudaStream_t)stream), "dx_bh_qvalues");
    return DX_OK;
}

int dx_bh_qvalues_sorted(int64_t n, const double* p_sorted, const int64_t* order, double* qval, double* work, void* stream) {
    DX_CHECK(n >= 0 && p_sorted && order && qval && work, "dx_bh_qvalues_sorted: bad arguments");
    DX_CUDA(dx_launch_bh_sorted(n, p_sorted, order, qval, work, (cudaStream_t)stream), "dx_bh_qvalues_sorted");
    return DX_OK;
}

int dx_two_group_tests(int64_t n_genes, double n0, double n1,
                       const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                       const float* ovf_val, const uint8_t* ovf_group,
                       int32_t do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0 && hist && ovf_count && indptr && out, "dx_two_group_tests: bad arguments");
    DX_CHECK(n0 >= 1 && n1 >= 1, "dx_two_group_tests: both groups need at least one cell");
    DX_CHECK(n0 + n1 <= 2097151.0 || !do_rank, "dx_two_group_tests: rank sums are exact up to 2^21-1 cells");
    DX_CUDA(dx_launch_two_group(n_genes, 2, 0, 1, n0, n1, hist, ovf_count, indptr, ovf_val, ovf_group, do_rank, out, u1x2, tie,
                                flags, (cudaStream_t)stream), "dx_two_group_tests");
    return DX_OK;
}

int dx_group_pair_tests(int64_t n_genes, int32_t n_groups, int32_t group_a, int32_t group_b, double n_a, double n_b,
                        const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                        const float* ovf_val, const uint8_t* ovf_group,
                        int32_t do_rank, double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* stream) {
    DX_CHECK(n_genes >= 0 && hist && ovf_count && indptr && out, "dx_group_pair_tests: bad arguments");
    DX_CHECK(n_groups >= 2 && n_groups <= DX_MAX_GROUPS, "dx_group_pair_tests: n_groups must be in 2..255");
    DX_CHECK(group_a >= 0 && group_a < n_groups && group_b < n_groups && group_b != group_a,
             "dx_group_pair_tests: group ids out of range");
    DX_CHECK(n_a >= 1 && n_b >= 1, "dx_group_pair_tests: both samples need at least one cell");
    DX_CHECK(n_a + n_b <= 2097151.0 || !do_rank, "dx_group_pair_tests: rank sums are exact up to 2^21-1 cells");
    DX_CUDA(dx_launch_two_group(n_genes, n_groups, group_a, group_b, n_a, n_b, hist, ovf_count, indptr, ovf_val, ovf_group,
                                do_rank, out, u1x2, tie, flags, (cudaStream_t)stream), "dx_group_pair_tests");
    return DX_OK;
}

int dx_group_pair_tests_big(int64_t n_genes, int32_t n_groups, int32_t group_a, int32_t group_b, double n_a, double n_b,
                            const uint32_t* hist, const int32_t* ovf_count, const int64_t* indptr,
                            const float* ovf_val, const uint8_t* ovf_group,
                            double* out, int64_t* u1x2, int64_t* tie, int32_t* flags, void* work, int64_t work_bytes,
                            int64_t slice_bytes, void* stream) {
    DX_CHECK(n_genes >= 0 && hist && ovf_count && indptr && out && flags, "dx_group_pair_tests_big: bad arguments");
    DX_CHECK(n_groups >= 2 && n_groups <= DX_MAX_GROUPS, "dx_group_pair_tests_big: n_groups must be in 2..255");
    DX_CHECK(group_a >= 0 && group_a < n_groups && group_b < n_groups && group_b != group_a,
             "dx_group_pair_tests_big: group ids out of range");
    DX_CHECK(n_a >= 1 && n_b >= 1, "dx_group_pair_tests_big: both samples need at least one cell");
    DX_CHECK(n_a + n_b <= 2097151.0, "dx_group_pair_tests_big: rank sums are exact up to 2^21-1 cells");
    DX_CHECK(work && slice_bytes >= DX_RANK_WORK_MIN_BYTES && work_bytes >= slice_bytes,
             "dx_group_pair_tests_big: workspace smaller than one slice of DX_RANK_WORK_MIN_BYTES");
    DX_CUDA(dx_launch_two_group_big(n_genes, n_groups, group_a, group_b, n_a, n_b, hist, ovf_count, indptr, ovf_val, ovf_group,
                                    out, u1x2, tie, flags, work, work_bytes, slice_bytes, sm_count_cached(), (cudaStream_t)stream),
            "dx_group_pair_tests_big");
    return DX_OK;
}

// ---- exchange buffers in peer memory (multi-GPU gather of result rows) ---------------------------------------------
int dx_exchange_create(int64_t slot_bytes, void** local_buf, void* ipc_handle_out) {
    DX_CHECK(slot_bytes > 0 && local_buf && ipc_handle_out, "dx_exchange_create: bad arguments");
    void* p = nullptr;
    const size_t total = 512 + 2 * (size_t)slot_bytes;
    DX_CUDA(cudaMalloc(&p, total), "dx_exchange_create: cudaMalloc");
    cudaError_t e = cudaMemset(p, 0, total);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); DX_CUDA(e, "dx_exchange_create: IPC handle"); }
    static_assert(sizeof(cudaIpcMemHandle_t) == DX_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    *local_buf = p;
    return DX_OK;
}

int dx_exchange_open(const void* ipc_handle, void** peer_buf) {
    DX_CHECK(ipc_handle && peer_buf, "dx_exchange_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    DX_CUDA(cudaIpcOpenMemHandle(peer_buf, h, cudaIpcMemLazyEnablePeerAccess), "dx_exchange_open");
    return DX_OK;
}

int dx_exchange_close(void* peer_buf) {
    DX_CHECK(peer_buf, "dx_exchange_close: null pointer");
    DX_CUDA(cudaIpcCloseMemHandle(peer_buf), "dx_exchange_close");
    return DX_OK;
}

int dx_exchange_destroy(void* local_buf) {
    DX_CHECK(local_buf, "dx_exchange_destroy: null pointer");
    DX_CUDA(cudaFree(local_buf), "dx_exchange_destroy");
    return DX_OK;
}

int dx_exchange_status(const void* local_buf, int32_t* failed_rank_plus_one) {
    DX_CHECK(local_buf && failed_rank_plus_one, "dx_exchange_status: bad arguments");
    uint32_t w = 0;
    DX_CUDA(cudaMemcpy(&w, static_cast<const unsigned char*>(local_buf) + 256, 4, cudaMemcpyDeviceToHost), "dx_exchange_status");
    *failed_rank_plus_one = (int32_t)w;
    return DX_OK;
}

int dx_peer_allgather_f64(int64_t n, const double* src, void* const* peer_bufs, int32_t world, int32_t rank,
                          int64_t slot_bytes, double* dst, void* stream) {
    DX_CHECK(n >= 0 && (src || n == 0) && peer_bufs, "dx_peer_allgather_f64: bad arguments");
    DX_CHECK(world >= 1 && world <= 16 && rank >= 0 && rank < world, "dx_peer_allgather_f64: world must be 1..16, rank inside it");
    DX_CHECK((int64_t)world * n * 8 <= slot_bytes, "dx_peer_allgather_f64: slot too small for world * n doubles");
    DX_CUDA(dx_launch_peer_allgather(n, src, peer_bufs, world, rank, slot_bytes, dst, sm_count_cached(), (cudaStream_t)stream),
            "dx_peer_allgather_f64");
    return DX_OK;
}

// ---- host-buffer entry point ------------------------------------------------------------------------
// The host shard is cut into chunks of genes with about equal non-zero counts; chunk c+1 crosses PCIe on the copy
// stream while chunk c is ingested, fitted and tested on the compute stream, so the call costs about one
// host-to-device transfer of the matrix.  Device workspace is cached per process (grown on demand).
namespace {
struct HostPipe {
    std::mutex mu;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    std::vector<cudaEvent_t> ev;
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf bufs[32];
    int device = -1;
    cudaError_t ensure(int slot, size_t bytes, void** out) {
        Buf& b = bufs[slot];
        if (b.cap < bytes) {
            if (b.p) cudaFree(b.p);
            b.p = nullptr; b.cap = 0;
            const size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&b.p, want);
            if (e != cudaSuccess) return e;
            b.cap = want;
        }
        *out = b.p;
        return cudaSuccess;
    }
    cudaError_t init(int n_events) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) {           // workspace belongs to one device: drop it when the caller switched devices
            for (Buf& b : bufs) { b.p = nullptr; b.cap = 0; }
            s_copy = s_comp = nullptr; ev.clear();
            device = dev;
        }
        if (!s_copy) { e = cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
        if (!s_comp) { e = cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
        while ((int)ev.size() < n_events) {
            cudaEvent_t x;
            e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            ev.push_back(x);
        }
        return cudaSuccess;
    }
};
HostPipe g_pipe;
}  // namespace

// frees the device workspace, streams and events the *_host entry points cache between calls
int dx_host_release(void) {
    std::lock_guard<std::mutex> lock(g_pipe.mu);
    if (g_pipe.device >= 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(g_pipe.device);
        for (HostPipe::Buf& b : g_pipe.bufs) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        for (cudaEvent_t e : g_pipe.ev) cudaEventDestroy(e);
        g_pipe.ev.clear();
        if (g_pipe.s_copy) cudaStreamDestroy(g_pipe.s_copy);
        if (g_pipe.s_comp) cudaStreamDestroy(g_pipe.s_comp);
        g_pipe.s_copy = g_pipe.s_comp = nullptr;
        g_pipe.device = -1;
        cudaSetDevice(cur);
    }
    return DX_OK;
}

int dx_wald_grouped_host(int64_t n_cells, int64_t n_genes, const int64_t* indptr, const int32_t* rows,
                         const float* vals, const uint8_t* cell_group, int32_t n_groups,
                         int32_t p_loc, int32_t p_scale, const double* xu_loc, const double* xu_scale,
                         const double* offset, const double* n_k, const double* pinv_loc,
                         const int32_t* loc_group, int32_t n_loc_groups,
                         const dx_fit_opts* opts, int32_t coef,
                         double* a_var, double* b_var, double* fim_inv, double* ll, double* jac,
                         int32_t* niter, int32_t* flags, double* pval, double* log_pval, double* sd, double* qval) {
    DX_CHECK(n_cells >= 1 && n_genes >= 0 && indptr && cell_group, "dx_wald_grouped_host: bad arguments");
    DX_CHECK(n_groups >= 1 && n_groups <= DX_MAX_GROUPS, "dx_wald_grouped_host: n_groups must be in 1..255");
    DX_CHECK(p_loc >= 1 && p_loc <= DX_MAX_COEF && p_scale >= 1 && p_scale <= DX_MAX_COEF,
             "dx_wald_grouped_host: coefficient counts must be in 1..32");
    DX_CHECK(coef >= 0 && coef < p_loc, "dx_wald_grouped_host: coefficient index out of range");
    DX_CHECK(!bad_opts(opts), "dx_wald_grouped_host: bad options");
    DX_CHECK(xu_loc && xu_scale && n_k && a_var && b_var, "dx_wald_grouped_host: null pointer");
    DX_CHECK(opts->init_a != DX_INIT_CLOSED_FORM || (pinv_loc && n_loc_groups >= 1 && n_loc_groups <= n_groups),
             "dx_wald_grouped_host: closed-form init needs pinv_loc and 1 <= n_loc_groups <= n_groups");
    if (n_genes == 0) return DX_OK;
    const int64_t nnz = indptr[n_genes];
    const int64_t G = n_genes;
    const int K = n_groups, p = p_loc, ps = p_scale;
    std::lock_guard<std::mutex> lock(g_pipe.mu);
    // chunk boundaries: about equal non-zero counts, at least 256 genes, at most 32 chunks
    int n_chunks = (int)((nnz >> 24) + 1);                       // ~16M non-zeros (128 MB) per chunk
    if (n_chunks > 32) n_chunks = 32;
    if (const char* env = getenv("DXB200_HOST_CHUNKS")) { const int v = atoi(env); if (v >= 1 && v <= 30) n_chunks = v; }
    else if ((int64_t)n_chunks * 256 > G) n_chunks = (int)((G + 255) / 256);
    if ((int64_t)n_chunks > G) n_chunks = (int)G;
    if (n_chunks < 1) n_chunks = 1;
    std::vector<int64_t> cut(n_chunks + 1);
    cut[0] = 0; cut[n_chunks] = G;
    for (int c = 1; c < n_chunks; ++c) {
        const int64_t target = nnz / n_chunks * c;
        const int64_t* it = std::lower_bound(indptr + cut[c - 1], indptr + G, target);
        int64_t g0 = it - indptr;
        if (g0 < cut[c - 1]) g0 = cut[c - 1];
        cut[c] = g0;
    }
    DX_CUDA(g_pipe.init(n_chunks + 2), "dx_wald_grouped_host: streams");
    cudaStream_t sc = g_pipe.s_copy, sx = g_pipe.s_comp;
    int rc = DX_OK;
    void *d_indptr, *d_rows, *d_vals, *d_grp, *d_hist, *d_oc, *d_ov, *d_og, *d_xl, *d_xs, *d_off, *d_nk, *d_pinv, *d_lg;
    void *d_a, *d_b, *d_fi, *d_ll, *d_jac, *d_ni, *d_fl, *d_p, *d_lp, *d_sd, *d_pk, *d_q, *d_bhw;
#define HX(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail_cuda(e__, "dx_wald_grouped_host"); goto done; } } while (0)
    {
        int sl = 0;
        HX(g_pipe.ensure(sl++, (G + 1) * 8, &d_indptr)); HX(g_pipe.ensure(sl++, nnz * 4 + 16, &d_rows));
        HX(g_pipe.ensure(sl++, nnz * 4 + 16, &d_vals)); HX(g_pipe.ensure(sl++, n_cells, &d_grp));
        HX(g_pipe.ensure(sl++, (size_t)G * K * DX_HIST_BINS * 4, &d_hist)); HX(g_pipe.ensure(sl++, G * 4, &d_oc));
        HX(g_pipe.ensure(sl++, nnz * 4 + 16, &d_ov)); HX(g_pipe.ensure(sl++, nnz + 16, &d_og));
        HX(g_pipe.ensure(sl++, (size_t)K * p * 8, &d_xl)); HX(g_pipe.ensure(sl++, (size_t)K * ps * 8, &d_xs));
        HX(g_pipe.ensure(sl++, (size_t)K * 8, &d_off)); HX(g_pipe.ensure(sl++, (size_t)K * 8, &d_nk));
        HX(g_pipe.ensure(sl++, (size_t)K * p * 8, &d_pinv)); HX(g_pipe.ensure(sl++, (size_t)K * 4, &d_lg));
        HX(g_pipe.ensure(sl++, (size_t)G * p * 8, &d_a)); HX(g_pipe.ensure(sl++, (size_t)G * ps * 8, &d_b));
        HX(g_pipe.ensure(sl++, (size_t)G * p * p * 8, &d_fi)); HX(g_pipe.ensure(sl++, G * 8, &d_ll));
        HX(g_pipe.ensure(sl++, G * 8, &d_jac)); HX(g_pipe.ensure(sl++, G * 4, &d_ni)); HX(g_pipe.ensure(sl++, G * 4, &d_fl));
        HX(g_pipe.ensure(sl++, G * 8, &d_p)); HX(g_pipe.ensure(sl++, G * 8, &d_lp)); HX(g_pipe.ensure(sl++, G * 8, &d_sd));
        HX(g_pipe.ensure(sl++, G * 8, &d_pk));
        HX(g_pipe.ensure(sl++, G * 8, &d_q)); HX(g_pipe.ensure(sl++, dx_bh_work_bytes_impl(G) + 16, &d_bhw));
        // small inputs first
        HX(cudaMemcpyAsync(d_indptr, indptr, (G + 1) * 8, cudaMemcpyHostToDevice, sc));
        HX(cudaMemcpyAsync(d_grp, cell_group, n_cells, cudaMemcpyHostToDevice, sc));
        HX(cudaMemcpyAsync(d_xl, xu_loc, (size_t)K * p * 8, cudaMemcpyHostToDevice, sc));
        HX(cudaMemcpyAsync(d_xs, xu_scale, (size_t)K * ps * 8, cudaMemcpyHostToDevice, sc));
        if (offset) HX(cudaMemcpyAsync(d_off, offset, (size_t)K * 8, cudaMemcpyHostToDevice, sc));
        HX(cudaMemcpyAsync(d_nk, n_k, (size_t)K * 8, cudaMemcpyHostToDevice, sc));
        if (pinv_loc) HX(cudaMemcpyAsync(d_pinv, pinv_loc, (size_t)(n_loc_groups > 0 ? n_loc_groups : K) * p * 8, cudaMemcpyHostToDevice, sc));
        if (loc_group) HX(cudaMemcpyAsync(d_lg, loc_group, (size_t)K * 4, cudaMemcpyHostToDevice, sc));
        if (opts->init_a == DX_INIT_GIVEN) HX(cudaMemcpyAsync(d_a, a_var, (size_t)G * p * 8, cudaMemcpyHostToDevice, sc));
        if (opts->init_b == DX_INIT_GIVEN) HX(cudaMemcpyAsync(d_b, b_var, (size_t)G * ps * 8, cudaMemcpyHostToDevice, sc));
        for (int c = 0; c < n_chunks; ++c) {
            const int64_t g0 = cut[c], g1 = cut[c + 1];
            if (g1 <= g0) continue;
            const int64_t e0 = indptr[g0], e1 = indptr[g1];
            if (e1 > e0) {
                HX(cudaMemcpyAsync((int32_t*)d_rows + e0, rows + e0, (e1 - e0) * 4, cudaMemcpyHostToDevice, sc));
                HX(cudaMemcpyAsync((float*)d_vals + e0, vals + e0, (e1 - e0) * 4, cudaMemcpyHostToDevice, sc));
            }
            HX(cudaEventRecord(g_pipe.ev[c], sc));
            HX(cudaStreamWaitEvent(sx, g_pipe.ev[c], 0));
            const int64_t gc = g1 - g0;
            HX(dx_launch_ingest(gc, (const int64_t*)d_indptr + g0, (const int32_t*)d_rows, (const float*)d_vals,
                                (const uint8_t*)d_grp, K, (uint32_t*)d_hist + g0 * (int64_t)K * DX_HIST_BINS,
                                (int32_t*)d_oc + g0, (float*)d_ov, (uint8_t*)d_og, nullptr, 0, sm_count_cached(), sx));
            HX(dx_launch_pack_overflow(gc, K, DX_PACK_MIN_ENTRIES, (const int32_t*)d_oc + g0, (const int64_t*)d_indptr + g0,
                                       (float*)d_ov, (uint8_t*)d_og, (int32_t*)d_pk + 2 * g0, sm_count_cached(), sx));
            DxFitEpilogueHost epi;            // Wald row of the tested coefficient straight from the fit kernel
            epi.wald_coef = coef; epi.theta = nullptr; epi.sd = (double*)d_sd + g0; epi.pval = (double*)d_p + g0;
            epi.logp = (double*)d_lp + g0; epi.world = 1; epi.rank = 0; epi.gene_offset = 0; epi.slot_bytes = 0; epi.peer_bufs = nullptr;
            HX(dx_launch_fit_grouped(gc, K, p, ps, (const double*)d_xl, (const double*)d_xs,
                                     offset ? (const double*)d_off : nullptr, (const double*)d_nk,
                                     pinv_loc ? (const double*)d_pinv : nullptr, loc_group ? (const int32_t*)d_lg : nullptr,
                                     n_loc_groups, (const uint32_t*)d_hist + g0 * (int64_t)K * DX_HIST_BINS,
                                     (const int32_t*)d_oc + g0, (const int64_t*)d_indptr + g0, (const float*)d_ov,
                                     (const uint8_t*)d_og, opts, (double*)d_a + g0, (double*)d_b + g0,
                                     (double*)d_fi + g0 * (int64_t)p * p, (double*)d_ll + g0, (double*)d_jac + g0,
                                     (int32_t*)d_ni + g0, (int32_t*)d_fl + g0, G, (const int32_t*)d_pk + 2 * g0, &epi, sx));
        }
        if (qval) {      // Benjamini-Hochberg over all genes (correction.py:5-24), part of the Wald result (det.py:108-117)
            HX(dx_launch_bh(G, (const double*)d_p, (double*)d_q, (double*)d_bhw, sx));
            HX(cudaMemcpyAsync(qval, d_q, G * 8, cudaMemcpyDeviceToHost, sx));
        }
        HX(cudaMemcpyAsync(a_var, d_a, (size_t)G * p * 8, cudaMemcpyDeviceToHost, sx));
        HX(cudaMemcpyAsync(b_var, d_b, (size_t)G * ps * 8, cudaMemcpyDeviceToHost, sx));
        if (fim_inv) HX(cudaMemcpyAsync(fim_inv, d_fi, (size_t)G * p * p * 8, cudaMemcpyDeviceToHost, sx));
        if (ll) HX(cudaMemcpyAsync(ll, d_ll, G * 8, cudaMemcpyDeviceToHost, sx));
        if (jac) HX(cudaMemcpyAsync(jac, d_jac, G * 8, cudaMemcpyDeviceToHost, sx));
        if (niter) HX(cudaMemcpyAsync(niter, d_ni, G * 4, cudaMemcpyDeviceToHost, sx));
        if (flags) HX(cudaMemcpyAsync(flags, d_fl, G * 4, cudaMemcpyDeviceToHost, sx));
        if (pval) HX(cudaMemcpyAsync(pval, d_p, G * 8, cudaMemcpyDeviceToHost, sx));
        if (log_pval) HX(cudaMemcpyAsync(log_pval, d_lp, G * 8, cudaMemcpyDeviceToHost, sx));
        if (sd) HX(cudaMemcpyAsync(sd, d_sd, G * 8, cudaMemcpyDeviceToHost, sx));
        HX(cudaStreamSynchronize(sx));
        HX(cudaStreamSynchronize(sc));
    }
done:
#undef HX
    if (rc != DX_OK) { cudaStreamSynchronize(sx); cudaStreamSynchronize(sc); }
    return rc;
}

}  // extern "C"
