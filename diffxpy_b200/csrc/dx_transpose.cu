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
//
// Staged variant (dx_csr_to_csc_staged, the default of the host code): pass 2 above writes 4-byte elements to ~20k open
// positions per CTA, which the L2 cannot hold across 148 CTAs -- ncu shows 7x the algorithmic DRAM traffic (partial-sector
// writes turn into read-modify-write).  The staged variant keeps every open write stream long and few:
//   count      as pass 1, over 4 chunks of cells per SM; also the number of entries per (block of 32 genes, chunk)
//   partition  a CTA per chunk walks its cells in order and appends (cell << 5 | gene mod 32, value) records to the segment of
//              the entry's gene block: one 8-byte store per entry into ~600 append-only streams per CTA.  Column indices are sorted
//              inside a cell (canonical CSR), so the entries of a gene block are one run of the cell's row: positions are
//              run start + offset, without atomics -- the record order is (cell, gene), deterministic
//   place      a CTA per gene block streams its records (all chunks, in order: ascending cells) and appends them to the 32
//              genes of the block: stable multi-split per tile of 256 records (warp match + per-warp counts + scan)
// 36 bytes per non-zero in all (4 + 8 read, 8 written, 8 read, 8 written), every stream sector-friendly; rows come out ascending inside every gene, as above.
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int TP_T = 1024;

__global__ void __launch_bounds__(TP_T, 1)
dx_tp_count_kernel(int64_t n_cells, const int64_t* __restrict__ indptr, const int32_t* __restrict__ cols, int lo, int n_out,
                   int64_t cells_per_chunk, int32_t* __restrict__ counts, int32_t* __restrict__ segcnt, int n_blocks, int gb,
                   unsigned long long* __restrict__ totals64) {
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
    if (counts) {
        int32_t* out = counts + (int64_t)chunk * n_out;
        for (int c = threadIdx.x; c < n_out; c += TP_T) out[c] = tp_cur[c];
    }
    if (segcnt) {
        // staged variant: only the per-gene totals are needed (integer atomics: exact in any order)
        for (int c = threadIdx.x; c < n_out; c += TP_T)
            if (tp_cur[c]) atomicAdd(totals64 + c, (unsigned long long)tp_cur[c]);          // staged variant: entries of this chunk per block of `gb` genes
        for (int b = threadIdx.x; b < n_blocks; b += TP_T) {
            int sum = 0;
            for (int j = b * gb; j < min((b + 1) * gb, n_out); ++j) sum += tp_cur[j];
            segcnt[(int64_t)b * gridDim.x + chunk] = sum;
        }
    }
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

// ---- staged variant --------------------------------------------------------------------------------------------------------
constexpr int TS_T = 256;          // threads per CTA of the partition / place kernels
constexpr int TS_GB = 32;          // genes per block (5 bits of the record key)
constexpr int TS_CPS = 4;          // chunks of cells per SM

// exclusive scan of n int32 counts into n + 1 int64 offsets (single CTA; tiles of 8192 elements, eight consecutive ones per
// thread, with a running carry)
__global__ void __launch_bounds__(1024)
dx_ts_scan_kernel(int64_t n, const int32_t* __restrict__ cnt, int64_t* __restrict__ off) {
    constexpr int PER = 8;
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024 * PER) {
        const int64_t i0 = base + (int64_t)tid * PER;
        int v[PER];
        long long sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = (i0 + k < n) ? cnt[i0 + k] : 0; sum += v[k]; }
        const long long inc = dx_warp_scan_ll(sum, lane);
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        long long run = s_carry + inc - sum;
        for (int w = 0; w < warp; ++w) run += s_w[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) { if (i0 + k < n) off[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) off[n] = s_carry;
}

__global__ void __launch_bounds__(TS_T)
dx_ts_partition_kernel(int64_t n_cells, const int64_t* __restrict__ indptr, const int32_t* __restrict__ cols,
                       const float* __restrict__ vals, int lo, int n_out, int64_t cells_per_chunk, int n_blocks,
                       const int64_t* __restrict__ segoff, uint2* __restrict__ irec) {
    extern __shared__ long long ts_sm[];
    long long* seg = ts_sm;                                              // [n_blocks] start of this chunk's segment per block
    int* cur = reinterpret_cast<int*>(ts_sm + n_blocks);                // [n_blocks] records appended so far
    int* rstart = cur + n_blocks;                                        // [n_blocks] where the current cell's run started
    const int chunk = blockIdx.x, n_chunks = gridDim.x, tid = threadIdx.x;
    const int64_t r0 = chunk * cells_per_chunk, r1 = min(r0 + cells_per_chunk, n_cells);
    for (int b = tid; b < n_blocks; b += TS_T) { seg[b] = segoff[(int64_t)b * n_chunks + chunk]; cur[b] = 0; }
    __syncthreads();
    int64_t e_next = indptr[r0];
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t e0 = e_next, e1 = indptr[r + 1];
        e_next = e1;
        for (int64_t t0 = e0; t0 < e1; t0 += TS_T) {
            const int64_t e = t0 + tid;
            int c = -1, b = -1, pb = -1, nb = -1;
            if (e < e1) {
                c = __ldg(cols + e) - lo;
                if ((unsigned)c < (unsigned)n_out) {
                    b = c / TS_GB;
                    if (e > e0) { const int pc = __ldg(cols + e - 1) - lo; pb = ((unsigned)pc < (unsigned)n_out) ? pc / TS_GB : -1; }
                    if (e + 1 < e1) { const int nc = __ldg(cols + e + 1) - lo; nb = ((unsigned)nc < (unsigned)n_out) ? nc / TS_GB : -1; }
                }
            }
            if (b >= 0 && pb != b) rstart[b] = (int)(e - e0);          // first entry of the block's run in this cell
            __syncthreads();
            if (b >= 0) {
                const int64_t pos = seg[b] + cur[b] + ((int)(e - e0) - rstart[b]);
                irec[pos] = make_uint2(((uint32_t)r << 5) | (uint32_t)(c - b * TS_GB), __float_as_uint(__ldg(vals + e)));
            }
            __syncthreads();
            if (b >= 0 && nb != b) cur[b] += (int)(e - e0) - rstart[b] + 1;          // last entry of the run
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(TS_T)
dx_ts_place_kernel(int n_out, int n_chunks, const int64_t* __restrict__ segoff, const uint2* __restrict__ irec,
                   const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_rows, float* __restrict__ out_vals) {
    __shared__ int s_wh[TS_T / 32][TS_GB];
    __shared__ long long s_base[TS_GB];
    __shared__ int s_cursor[TS_GB];
    const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t s0 = segoff[(int64_t)blk * n_chunks], s1 = segoff[(int64_t)(blk + 1) * n_chunks];
    if (tid < TS_GB) {
        const int g = blk * TS_GB + tid;
        s_base[tid] = (g < n_out) ? out_ptr[g] : 0;
        s_cursor[tid] = 0;
    }
    __syncthreads();
    for (int64_t t0 = s0; t0 < s1; t0 += TS_T) {
        const int64_t i = t0 + tid;
        const bool valid = i < s1;
        const uint2 rec = valid ? irec[i] : make_uint2(0u, 0u);
        const uint32_t key = rec.x;
        const float v = __uint_as_float(rec.y);
        const int g = valid ? (int)(key & 31u) : -1;
        // stable multi-split of the tile: records of one gene keep their order (ascending cells)
        const unsigned m = __match_any_sync(DX_FULL_MASK, g);
        const int rank = __popc(m & ((1u << lane) - 1u));
        s_wh[warp][lane] = 0;
        __syncwarp();
        if (valid && rank == 0) s_wh[warp][g] = __popc(m);
        __syncthreads();
        if (tid < TS_GB) {
            int run = s_cursor[tid];
#pragma unroll
            for (int w = 0; w < TS_T / 32; ++w) { const int c = s_wh[w][tid]; s_wh[w][tid] = run; run += c; }
            s_cursor[tid] = run;
        }
        __syncthreads();
        if (valid) {
            const int64_t pos = s_base[g] + s_wh[warp][g] + rank;
            out_rows[pos] = (int32_t)(key >> 5);
            out_vals[pos] = v;
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
    dx_tp_count_kernel<<<n_chunks, TP_T, smem, stream>>>(n_cells, indptr, cols, lo, n_out, cells_per_chunk, counts, nullptr, 0, 1, nullptr);
    dx_tp_scan_chunks_kernel<<<(n_out + 255) / 256, 256, 0, stream>>>(n_out, n_chunks, counts, totals);
    dx_tp_scan_genes_kernel<<<1, 1024, 0, stream>>>(n_out, totals, out_ptr);
    dx_tp_scatter_kernel<<<n_chunks, TP_T, smem, stream>>>(n_cells, indptr, cols, vals, lo, n_out, cells_per_chunk, counts, out_ptr,
                                                            out_rows, out_vals);
    return cudaGetLastError();
}

// ---- staged variant: layout of the workspace ---------------------------------------------------------------------------
// totals int64[n_out] | segcnt int32[n_blocks n_chunks] | segoff int64[n_blocks n_chunks + 1] | records uint2[nnz]
static inline size_t ts_align(size_t x) { return (x + 255) & ~(size_t)255; }

size_t dx_transpose_staged_work_bytes_impl(int64_t n_cells, int n_out, int64_t nnz, int sm_count) {
    (void)n_cells;
    const size_t n_chunks = (size_t)sm_count * TS_CPS, n_blocks = ((size_t)n_out + TS_GB - 1) / TS_GB;
    return ts_align((size_t)n_out * 8) + ts_align(n_blocks * n_chunks * 4) + ts_align((n_blocks * n_chunks + 1) * 8) +
           ts_align((size_t)nnz * 8) + 256;
}

cudaError_t dx_launch_csr_to_csc_staged(int64_t n_cells, int64_t nnz, const int64_t* indptr, const int32_t* cols, const float* vals,
                                        int lo, int n_out, int64_t* out_ptr, int32_t* out_rows, float* out_vals, void* work,
                                        int64_t work_bytes, int sm_count, cudaStream_t stream) {
    if (n_out <= 0) return cudaSuccess;
    const size_t smem_cnt = (size_t)n_out * 4;
    if (smem_cnt > 220 * 1024 || n_cells >= (1ll << 27)) return cudaErrorInvalidConfiguration;
    if ((size_t)work_bytes < dx_transpose_staged_work_bytes_impl(n_cells, n_out, nnz, sm_count)) return cudaErrorInvalidValue;
    if (n_cells <= 0) return cudaMemsetAsync(out_ptr, 0, (size_t)(n_out + 1) * 8, stream);
    int n_chunks = sm_count * TS_CPS;
    if (n_chunks > n_cells) n_chunks = (int)n_cells;
    const int64_t cells_per_chunk = (n_cells + n_chunks - 1) / n_chunks;
    n_chunks = (int)((n_cells + cells_per_chunk - 1) / cells_per_chunk);
    const int n_blocks = (n_out + TS_GB - 1) / TS_GB;
    const size_t max_chunks = (size_t)sm_count * TS_CPS;
    char* w = static_cast<char*>(work);
    w = reinterpret_cast<char*>(ts_align(reinterpret_cast<size_t>(w)));
    int64_t* totals = reinterpret_cast<int64_t*>(w); w += ts_align((size_t)n_out * 8);
    int32_t* segcnt = reinterpret_cast<int32_t*>(w); w += ts_align((size_t)n_blocks * max_chunks * 4);
    int64_t* segoff = reinterpret_cast<int64_t*>(w); w += ts_align(((size_t)n_blocks * max_chunks + 1) * 8);
    uint2* irec = reinterpret_cast<uint2*>(w);
    cudaError_t e = cudaFuncSetAttribute(dx_tp_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cnt);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(totals, 0, (size_t)n_out * 8, stream);
    if (e != cudaSuccess) return e;
    dx_tp_count_kernel<<<n_chunks, TP_T, smem_cnt, stream>>>(n_cells, indptr, cols, lo, n_out, cells_per_chunk, nullptr, segcnt,
                                                             n_blocks, TS_GB, reinterpret_cast<unsigned long long*>(totals));
    dx_tp_scan_genes_kernel<<<1, 1024, 0, stream>>>(n_out, totals, out_ptr);
    dx_ts_scan_kernel<<<1, 1024, 0, stream>>>((int64_t)n_blocks * n_chunks, segcnt, segoff);
    const size_t smem_part = (size_t)n_blocks * (8 + 4 + 4);
    e = cudaFuncSetAttribute(dx_ts_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_part);
    if (e != cudaSuccess) return e;
    dx_ts_partition_kernel<<<n_chunks, TS_T, smem_part, stream>>>(n_cells, indptr, cols, vals, lo, n_out, cells_per_chunk, n_blocks,
                                                                   segoff, irec);
    dx_ts_place_kernel<<<n_blocks, TS_T, 0, stream>>>(n_out, n_chunks, segoff, irec, out_ptr, out_rows, out_vals);
    return cudaGetLastError();
}
