// dx_overflow.cu -- K_pack: multiplicity compression of long overflow lists.
//
// K_ingest keeps counts 1..64 as per-group tail counts and everything else as an "overflow" list of
// (value, group) entries, which K_fit walks once per likelihood / derivative evaluation.  For the usual
// single-cell shard (gene means below ~1) the lists are empty or a handful of entries.  A highly expressed gene
// (mean 30 .. 1000: mitochondrial, ribosomal, MALAT1 ...) puts most of its cells there, and one 16-lane
// sub-warp walking 10^5 entries ~40 times sets the duration of the whole fit launch.  Those lists hold few
// DISTINCT values, so this kernel rewrites the list of every gene with >= min_entries entries, in place, as
//
//   floats [0, 8K)                       per group k: { sum y, sum y^2, sum lgamma(y + 1), count } as 4 doubles
//   floats [8K, 8K + 2 nA)               records (value, multiplicity) per (group, value), group-major, value
//                                        ascending; the group byte of record i is ovf_group[8K + 2 i]
//   floats [8K + 2 nA, 8K + 2 nA + 2 nB) records (value, multiplicity) per value, all groups pooled
//
// (multiplicities are the uint32 bit pattern of the second float).  K_fit reads the header at set-up, the pooled
// records in the shared-dispersion loops and the per-group records otherwise.  Entries that are not integers in
// [65, 65 + R) (R = table columns that fit in shared memory) are kept verbatim, multiplicity 1, at the end of
// both record lists, in their original order.  A gene is left untouched (ovf_packed = {0, 0}) when packing would
// not at least fit in the space of the plain list or when more than PK_LMAX entries would have to be kept
// verbatim (non-integer data).  The record order depends only on the multiset of entries: packing is
// deterministic and independent of how the shard was cut.
//
// One CTA per packed gene; every CTA owns a contiguous range of genes and reads their counts coalesced, so the
// launch costs a few microseconds when no list is long (the usual shard).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int PK_T = 1024;           // threads per CTA
constexpr int PK_LMAX = 1024;        // entries kept verbatim, at most
constexpr int PK_FIRST = DX_HIST_BINS + 1;   // first value with a table column (65)

// exclusive block scan of one flag per thread (ordered by thread id); returns the position, adds the total to run
__device__ __forceinline__ unsigned pk_scan_flag(bool flag, unsigned* s_warp, unsigned& run) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(DX_FULL_MASK, flag);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PK_T / 32; ++w) {
        const unsigned c = s_warp[w];
        if (w < warp) before += c;
        total += c;
    }
    __syncthreads();
    const unsigned pos = run + before + __popc(bal & ((1u << lane) - 1u));
    run += total;
    return pos;
}

__device__ __forceinline__ void pk_store_double(float* dst, double v) {
    dst[0] = __int_as_float(__double2loint(v));
    dst[1] = __int_as_float(__double2hiint(v));
}

__global__ void __launch_bounds__(PK_T, 1)
dx_pack_overflow_kernel(int64_t n_genes, int K, int R, int min_entries, const int32_t* __restrict__ ovf_count,
                        const int64_t* __restrict__ indptr, float* ovf_val, uint8_t* ovf_group,
                        int32_t* __restrict__ ovf_packed) {
    extern __shared__ __align__(16) unsigned char pk_smem[];
    uint32_t* table = reinterpret_cast<uint32_t*>(pk_smem);                 // [K][R]
    float* l_val = reinterpret_cast<float*>(table + (size_t)K * R);         // [PK_LMAX]
    unsigned* s_warp = reinterpret_cast<unsigned*>(l_val + PK_LMAX);        // [32]
    unsigned* s_cnt = s_warp + 32;                                          // [4]: leftovers, nA, nB
    int* s_list = reinterpret_cast<int*>(s_cnt + 4);                        // [PK_T] long lists of the current batch
    uint8_t* l_grp = reinterpret_cast<uint8_t*>(s_list + PK_T);             // [PK_LMAX]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbins = K * R;

    // the CTA owns a contiguous range of genes: the counts are read coalesced, the (few) long lists collected in
    // shared memory and packed one after the other
    const int64_t per_cta = (n_genes + gridDim.x - 1) / gridDim.x;
    const int64_t g_lo = blockIdx.x * per_cta, g_hi = (g_lo + per_cta < n_genes) ? g_lo + per_cta : n_genes;
    for (int64_t g0 = g_lo; g0 < g_hi; g0 += PK_T) {
    const int64_t g_mine = g0 + tid;
    bool heavy = false;
    if (g_mine < g_hi) {
        const int n_mine = ovf_count[g_mine];
        heavy = n_mine >= min_entries && n_mine >= 8 * K + 8;
        if (!heavy) { ovf_packed[2 * g_mine] = 0; ovf_packed[2 * g_mine + 1] = 0; }
    }
    unsigned n_heavy = 0;
    {
        const unsigned pos = pk_scan_flag(heavy, s_warp, n_heavy);
        if (heavy) s_list[pos] = (int)(g_mine - g0);
    }
    __syncthreads();
    for (unsigned hi = 0; hi < n_heavy; ++hi) {
        const int64_t g = g0 + s_list[hi];
        const int n = ovf_count[g];
        float* val = ovf_val + indptr[g];
        uint8_t* grp = ovf_group + indptr[g];
        for (int i = tid; i < nbins; i += PK_T) table[i] = 0u;
        if (tid < 4) s_cnt[tid] = 0u;
        __syncthreads();
        // ---- pass 1: count the table entries, count what has to stay verbatim --------------------------------
        unsigned left = 0;
        for (int t = tid; t < n; t += PK_T) {
            const float v = val[t];
            const int k = grp[t];
            const int y = __float2int_rz(v);
            if (__int2float_rn(y) == v && y >= PK_FIRST && y < PK_FIRST + R) atomicAdd(&table[k * R + (y - PK_FIRST)], 1u);
            else ++left;
        }
        left = __reduce_add_sync(DX_FULL_MASK, left);
        if (lane == 0 && left) atomicAdd(&s_cnt[0], left);
        __syncthreads();
        unsigned n_a = 0, n_b = 0;
        for (int i = tid; i < nbins; i += PK_T) n_a += table[i] != 0u;
        for (int c = tid; c < R; c += PK_T) {
            unsigned any = 0;
            for (int k = 0; k < K; ++k) any |= table[k * R + c];
            n_b += any != 0u;
        }
        n_a = __reduce_add_sync(DX_FULL_MASK, n_a);
        n_b = __reduce_add_sync(DX_FULL_MASK, n_b);
        if (lane == 0) { if (n_a) atomicAdd(&s_cnt[1], n_a); if (n_b) atomicAdd(&s_cnt[2], n_b); }
        __syncthreads();
        const unsigned n_left = s_cnt[0];
        n_a = s_cnt[1] + n_left;
        n_b = s_cnt[2] + n_left;
        if (n_left > (unsigned)PK_LMAX || (uint64_t)8 * K + 2ull * n_a + 2ull * n_b > (uint64_t)n) {
            if (tid == 0) { ovf_packed[2 * g] = 0; ovf_packed[2 * g + 1] = 0; }
            __syncthreads();
            continue;
        }
        // ---- verbatim entries, in their original order -------------------------------------------------------
        if (n_left) {
            unsigned run = 0;
            for (int base = 0; base < n; base += PK_T) {
                const int t = base + tid;
                float v = 0.0f;
                int k = 0;
                bool keep = false;
                if (t < n) {
                    v = val[t];
                    k = grp[t];
                    const int y = __float2int_rz(v);
                    keep = !(__int2float_rn(y) == v && y >= PK_FIRST && y < PK_FIRST + R);
                }
                const unsigned pos = pk_scan_flag(keep, s_warp, run);
                if (keep) { l_val[pos] = v; l_grp[pos] = (uint8_t)k; }
            }
        }
        __syncthreads();       // every read of the plain list is done: the region is rewritten from here on
        // ---- header: per-group sums, one warp per group, fixed reduction order -----------------------------
        for (int k = warp; k < K; k += PK_T / 32) {
            double sy = 0.0, sq = 0.0, sg = 0.0, cn = 0.0;
            for (int c = lane; c < R; c += 32) {
                const unsigned m = table[k * R + c];
                if (m) {
                    const double y = (double)(PK_FIRST + c), w = (double)m;
                    sy += w * y; sq += w * y * y; sg += w * dx_lgamma(y + 1.0); cn += w;
                }
            }
            for (unsigned l = lane; l < n_left; l += 32) {
                if (l_grp[l] == k) {
                    const double y = (double)l_val[l];
                    sy += y; sq += y * y; sg += dx_lgamma(y + 1.0); cn += 1.0;
                }
            }
            sy = dx_warp_sum(sy); sq = dx_warp_sum(sq); sg = dx_warp_sum(sg); cn = dx_warp_sum(cn);
            if (lane == 0) {
                float* h = val + 8 * k;
                pk_store_double(h + 0, sy); pk_store_double(h + 2, sq); pk_store_double(h + 4, sg); pk_store_double(h + 6, cn);
            }
        }
        // ---- per-(group, value) records ------------------------------------------------------------------------
        {
            float* rec = val + 8 * K;
            uint8_t* rg = grp + 8 * K;
            unsigned run = 0;
            for (int base = 0; base < nbins; base += PK_T) {
                const int i = base + tid;
                const unsigned m = i < nbins ? table[i] : 0u;
                const unsigned pos = pk_scan_flag(m != 0u, s_warp, run);
                if (m) {
                    const int k = i / R, c = i - k * R;
                    rec[2 * pos] = (float)(PK_FIRST + c);
                    rec[2 * pos + 1] = __uint_as_float(m);
                    rg[2 * pos] = (uint8_t)k;
                }
            }
            for (unsigned l = tid; l < n_left; l += PK_T) {
                rec[2 * (run + l)] = l_val[l];
                rec[2 * (run + l) + 1] = __uint_as_float(1u);
                rg[2 * (run + l)] = l_grp[l];
            }
        }
        // ---- pooled per-value records ------------------------------------------------------------------------
        {
            float* rec = val + 8 * K + 2 * (size_t)n_a;
            unsigned run = 0;
            for (int base = 0; base < R; base += PK_T) {
                const int c = base + tid;
                unsigned m = 0;
                if (c < R)
                    for (int k = 0; k < K; ++k) m += table[k * R + c];
                const unsigned pos = pk_scan_flag(m != 0u, s_warp, run);
                if (m) {
                    rec[2 * pos] = (float)(PK_FIRST + c);
                    rec[2 * pos + 1] = __uint_as_float(m);
                }
            }
            for (unsigned l = tid; l < n_left; l += PK_T) {
                rec[2 * (run + l)] = l_val[l];
                rec[2 * (run + l) + 1] = __uint_as_float(1u);
            }
        }
        if (tid == 0) { ovf_packed[2 * g] = (int32_t)n_a; ovf_packed[2 * g + 1] = (int32_t)n_b; }
        __syncthreads();
    }
    __syncthreads();
    }
}

}  // namespace

cudaError_t dx_launch_pack_overflow(int64_t n_genes, int K, int min_entries, const int32_t* ovf_count,
                                    const int64_t* indptr, float* ovf_val, uint8_t* ovf_group, int32_t* ovf_packed,
                                    int sm_count, cudaStream_t stream) {
    if (n_genes <= 0) return cudaSuccess;
    const size_t fixed = PK_LMAX * 4 + 32 * 4 + 4 * 4 + PK_T * 4 + PK_LMAX;
    int R = (int)((200 * 1024 - fixed) / (4 * (size_t)K));
    if (R > 8192) R = 8192;
    R &= ~31;
    if (R < 32) {      // cannot happen for K <= 255; kept as a guard
        return cudaMemsetAsync(ovf_packed, 0, sizeof(int32_t) * 2 * n_genes, stream);
    }
    const size_t smem = (size_t)K * R * 4 + fixed;
    cudaError_t e = cudaFuncSetAttribute(dx_pack_overflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    int64_t grid = n_genes < sm_count ? n_genes : sm_count;
    dx_pack_overflow_kernel<<<(unsigned)grid, PK_T, smem, stream>>>(n_genes, K, R, min_entries < 1 ? 1 : min_entries,
                                                                    ovf_count, indptr, ovf_val, ovf_group, ovf_packed);
    return cudaGetLastError();
}
