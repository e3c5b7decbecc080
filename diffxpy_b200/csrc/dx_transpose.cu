// dx_transpose.cu -- K8: cell-major CSR (what AnnData / scipy hold, /root/reference/diffxpy/testing/utils.py:20-33)
// -> gene-major shard (indptr int64, rows int32 ascending inside a gene, vals float32) of a gene range, on the device.
// batchglm densifies 128-gene column blocks from the row-major matrix for every chunk; here the
// matrix is transposed once:
//   pass 1  one CTA per chunk of consecutive cells counts the entries of every gene of the range in shared memory
//           (one counter per gene: up to ~55k genes) and stores the chunk's counts                 [reads the column indices]
//   scan    per gene an exclusive scan of the counts over the chunks (thread <-> gene, coalesced), then over the genes
//           (single CTA): gene pointers and, per chunk and gene, where the chunk's entries of that gene start
//   pass 2  the same CTAs walk their cells IN ORDER, one cell per step: a cell holds a gene at most once, so the cursor of
//           a gene is advanced by exactly one thread per step -- no atomics, and the entries of a gene come out in
//           ascending cell order (stable, deterministic)                      [reads indices + values, writes rows + values]
// HBM traffic: 4 nnz read + 16 nnz read/written, 20 bytes per non-zero in all (a sort by gene key moves several times that).
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int TP_T = 1024;

__global__ void __launch_bounds__(TP_T, 1)
dx_tp_count_kernel(int64_t n_cells, const int64_t* __restrict__ indptr, const int32_t* __restrict__ cols, int lo, int n_out,
                   int64_t cells_per_chunk, int32_t* __restrict__ counts) {
    extern __shared__ int tp_cur[];
    const int chunk = blockIdx.x;
    const int64_t r0 = chunk * cells_per_chunk, r1 = min(r0 + cells_per_chunk, n_cells);
    for (int c = threadIdx.x; c < n_out; c += TP_T) tp_cur[c] = 0;
    __syncthreads();
    const int64_t e0 = indptr[r0], e1 = indptr[r1];
    for (int64_t e = e0 + threadIdx.x; e < e1; e += TP_T) {
        const unsigned c = (unsigned)(__ldg(cols + e) - lo);
        if (c < (unsigned)n_out) atomicAdd(&tp_cur[c], 1);
    }
    __syncthreads();
    int32_t* out = counts + (int64_t)chunk * n_out;
    for (int c = threadIdx.x; c < n_out; c += TP_T) out[c] = tp_cur[c];
}

// counts[chunk][gene] -> exclusive prefix over the chunks (in place); totals[gene]
__global__ void __launch_bounds__(256)
dx_tp_scan_chunks_kernel(int n_out, int n_chunks, int32_t* __restrict__ counts, int64_t* __restrict__ totals) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n_out) return;
    int run = 0;
    for (int k = 0; k < n_chunks; ++k) {
        int32_t* p = counts + (int64_t)k * n_out + c;
        const int t = *p;
        *p = run;
        run += t;
    }
    totals[c] = run;
}

// out_ptr[0 .. n_out] = exclusive scan of totals (single CTA; totals may alias out_ptr + 1 ... no: separate arrays)
__global__ void __launch_bounds__(1024)
dx_tp_scan_genes_kernel(int n_out, const int64_t* __restrict__ totals, int64_t* __restrict__ out_ptr) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; out_ptr[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n_out; base += 1024) {
        const int c = base + tid;
        long long v = (c < n_out) ? totals[c] : 0;
        long long inc = dx_warp_scan_ll(v, lane);
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        long long before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (c < n_out) out_ptr[c + 1] = before + inc;
        __syncthreads();
        if (tid == 1023) s_carry = before + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TP_T, 1)
dx_tp_scatter_kernel(int64_t n_cells, const int64_t* __restrict__ indptr, const int32_t* __restrict__ cols,
                     const float* __restrict__ vals, int lo, int n_out, int64_t cells_per_chunk,
                     const int32_t* __restrict__ counts, const int64_t* __restrict__ out_ptr,
                     int32_t* __restrict__ out_rows, float* __restrict__ out_vals) {
    extern __shared__ int tp_cur[];          // cursor of every gene relative to the gene's start (fits 32 bits: <= n_cells)
    const int chunk = blockIdx.x;
    const int64_t r0 = chunk * cells_per_chunk, r1 = min(r0 + cells_per_chunk, n_cells);
    const int32_t* base = counts + (int64_t)chunk * n_out;
    for (int c = threadIdx.x; c < n_out; c += TP_T) tp_cur[c] = base[c];
    __syncthreads();
    int64_t e_next = indptr[r0];
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t e0 = e_next, e1 = indptr[r + 1];
        e_next = e1;
        for (int64_t e = e0 + threadIdx.x; e < e1; e += TP_T) {
            const unsigned c = (unsigned)(__ldg(cols + e) - lo);
            if (c < (unsigned)n_out) {
                const int k = tp_cur[c];          // the only entry of gene c in this cell: no other thread touches the cursor
                tp_cur[c] = k + 1;
                const int64_t pos = out_ptr[c] + k;
                out_rows[pos] = (int32_t)r;
                out_vals[pos] = __ldg(vals + e);
            }
        }
        __syncthreads();
    }
}

}  // namespace

size_t dx_transpose_work_bytes_impl(int64_t n_cells, int n_out, int sm_count) {
    const int64_t n_chunks = sm_count;
    (void)n_cells;
    return (size_t)n_chunks * n_out * 4 + (size_t)n_out * 8 + 64;
}

cudaError_t dx_launch_csr_to_csc(int64_t n_cells, const int64_t* indptr, const int32_t* cols, const float* vals, int lo, int n_out,
                                 int64_t* out_ptr, int32_t* out_rows, float* out_vals, void* work, int sm_count,
                                 cudaStream_t stream) {
    if (n_out <= 0) return cudaSuccess;
    const size_t smem = (size_t)n_out * 4;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;          // one counter per gene must fit the CTA's shared memory
    int n_chunks = sm_count;
    if (n_chunks > n_cells) n_chunks = (int)(n_cells > 0 ? n_cells : 1);
    const int64_t cells_per_chunk = (n_cells + n_chunks - 1) / n_chunks;
    n_chunks = (int)((n_cells + cells_per_chunk - 1) / (cells_per_chunk > 0 ? cells_per_chunk : 1));
    if (n_chunks < 1) n_chunks = 1;
    int32_t* counts = static_cast<int32_t*>(work);
    int64_t* totals = reinterpret_cast<int64_t*>(counts + (((int64_t)sm_count * n_out + 1) & ~1ll));
    cudaError_t e = cudaFuncSetAttribute(dx_tp_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dx_tp_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (n_cells <= 0) {
        return cudaMemsetAsync(out_ptr, 0, (size_t)(n_out + 1) * 8, stream);
    }
    dx_tp_count_kernel<<<n_chunks, TP_T, smem, stream>>>(n_cells, indptr, cols, lo, n_out, cells_per_chunk, counts);
    dx_tp_scan_chunks_kernel<<<(n_out + 255) / 256, 256, 0, stream>>>(n_out, n_chunks, counts, totals);
    dx_tp_scan_genes_kernel<<<1, 1024, 0, stream>>>(n_out, totals, out_ptr);
    dx_tp_scatter_kernel<<<n_chunks, TP_T, smem, stream>>>(n_cells, indptr, cols, vals, lo, n_out, cells_per_chunk, counts, out_ptr,
                                                            out_rows, out_vals);
    return cudaGetLastError();
}
