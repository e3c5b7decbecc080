// dx_ingest.cu -- K_ingest: one streaming pass over the gene-major (CSC) count shard that turns every
// gene into per-design-group count histograms (the sufficient statistics of the NB likelihood for a
// grouped design, and of the Wilcoxon / Welch tests).  HBM-bound: reads 8 B per non-zero once.
//
// Replaces the dense per-iteration passes of batchglm's numpy backend that diffxpy triggers at
// /root/reference/diffxpy/testing/tests.py:242-245 and the x.mean / x.power(2).mean passes of
// /root/reference/diffxpy/testing/det.py:1567-1584, 1693-1709.
//
// One CTA (4 warps) per gene, grid-stride over genes.  Hot bins (y <= YP, K*YP <= 128, YP a power of two) are
// counted in THREAD-PRIVATE 8-bit counters packed four to a 32-bit shared-memory word; the words of one
// thread live in bank == lane (layout [warp][word][lane]), so the fire-and-forget shared atomic that bumps a
// counter is conflict-free and carries no read-modify-write dependency.  A warp folds its own counters into
// the CTA histogram with a skewed (conflict-free) column walk; no block barrier is needed for that.  The rare
// bins YP < y <= 64 go through shared atomics on the CTA histogram; everything else (non-integer, negative,
// > 64) is an "overflow" entry that is copied out in a deterministic order by a second, L2-resident pass.
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "dx_common.cuh"
#include "dx_internal.h"

namespace {

constexpr int T = DX_INGEST_THREADS;   // 128
#ifndef DX_INGEST_U
#define DX_INGEST_U 8
#endif
#ifndef DX_INGEST_MINB
#define DX_INGEST_MINB 6
#endif
constexpr int U = DX_INGEST_U;         // independent loads in flight per thread and array
constexpr int NWARP = T / 32;

// streaming loads: read-only path, do not allocate in L1 (the cell -> group table should stay there instead)
#ifndef DX_INGEST_NOALLOC
#define DX_INGEST_NOALLOC 1
#endif
__device__ __forceinline__ float dx_ld_stream(const float* p) {
#if DX_INGEST_NOALLOC
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ int dx_ld_stream(const int32_t* p) {
#if DX_INGEST_NOALLOC
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// fire-and-forget shared-memory add on a 32-bit shared-space address
__device__ __forceinline__ void dx_red_shared(uint32_t addr, unsigned inc) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
}

__device__ __forceinline__ bool dx_classify(float v, int k, int K, int& y) {
    // true -> direct bin y (1..64); false -> overflow candidate (or ignorable when v == 0 / skipped cell)
    y = (int)v;
    return (k < K) && (v >= 1.0f) && (v <= 64.0f) && ((float)y == v);
}

__global__ void __launch_bounds__(T, DX_INGEST_MINB)
dx_ingest_kernel(int64_t n_genes, const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                 const float* __restrict__ vals, const uint8_t* __restrict__ cell_group, int K, int ysh, int NW, int NWA,
                 uint32_t* __restrict__ hist_out, int32_t* __restrict__ ovf_count,
                 float* __restrict__ ovf_val, uint8_t* __restrict__ ovf_group) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);         // [K][64]
    uint32_t* s_scan = s_hist + K * DX_HIST_BINS;                     // [T + 4]
    uint32_t* s_priv = s_scan + T + 4;                                // [NWARP][NWA][32] packed 8-bit counters
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned YP = (ysh >= 0) ? (1u << ysh) : 0u;
    const unsigned ysh_u = (ysh >= 0) ? (unsigned)ysh : 0u;
    const unsigned n_priv = (unsigned)K * YP;
    uint32_t* priv_warp = s_priv + warp * NWA * 32;
    uint32_t* priv_t = priv_warp + lane;
    const unsigned row_mask = (unsigned)(NWA - 1) << 5;   // word-row offset (in words) of counter j: (j >> 2) * 32 == (8j) & row_mask
    const unsigned ysh3 = ysh_u + 3u;
    const uint32_t priv_addr = (uint32_t)__cvta_generic_to_shared(priv_t);

    for (int i = tid; i < NWARP * NWA * 32; i += T) s_priv[i] = 0u;    // every flush leaves the counters zero again
    for (int64_t g = blockIdx.x; g < n_genes; g += gridDim.x) {
        const int64_t start = indptr[g], end = indptr[g + 1];
        {
            uint4* h4 = reinterpret_cast<uint4*>(s_hist);
            for (int i = tid; i < K * (DX_HIST_BINS / 4); i += T) h4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        unsigned n_ovf = 0;
        // a thread sees at most 255 entries between two flushes of its 8-bit counters
        const int64_t chunk = (int64_t)255 * T;
        for (int64_t c0 = start; c0 < end; c0 += chunk) {
            const int64_t c1 = (c0 + chunk < end) ? c0 + chunk : end;
            // software pipeline: the next batch of U value/row loads is in flight while the current batch does
            // its group gathers and counter updates.  Full batches use immediate-offset loads without bounds
            // checks; only the last batch of a chunk is predicated (padding entries are zeros -> ignored).
            float va[U], vb[U];
            int ra[U], rb[U];
            const int n_here = (int)(c1 - c0);
            const int n_batches = (n_here + U * T - 1) / (U * T);
            const float* vp = vals + c0 + tid;
            const int32_t* rp = rows + c0 + tid;
            auto load = [&](int b, float (&v)[U], int (&r)[U]) {
                const float* vq = vp + b * (U * T);
                const int32_t* rq = rp + b * (U * T);
                if ((b + 1) * (U * T) <= n_here) {
#pragma unroll
                    for (int u = 0; u < U; ++u) { v[u] = dx_ld_stream(vq + u * T); r[u] = dx_ld_stream(rq + u * T); }
                } else if (b < n_batches) {
                    const int left = n_here - b * (U * T) - tid;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool ok = u * T < left;
                        v[u] = ok ? dx_ld_stream(vq + u * T) : 0.0f;
                        r[u] = ok ? dx_ld_stream(rq + u * T) : 0;
                    }
                }
            };
            auto consume = [&](const float (&v)[U], const int (&r)[U]) {
                unsigned k[U];
#pragma unroll
                for (int u = 0; u < U; ++u) k[u] = __ldg(cell_group + (unsigned)r[u]);
                bool slow = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int y = __float2int_rz(v[u]);
                    const unsigned ym1 = (unsigned)(y - 1);
                    const bool fast = (__int2float_rn(y) == v[u]) && (ym1 < YP) && (k[u] < (unsigned)K);
                    const unsigned j8 = (k[u] << ysh3) + (ym1 << 3);             // 8 * counter index
                    // branch-free: a non-private entry adds 0 to an in-range word
                    const unsigned inc = fast ? (1u << (j8 & 24u)) : 0u;
                    dx_red_shared(priv_addr + ((j8 & row_mask) << 2), inc);
                    slow = slow || !fast;
                }
                if (slow) {     // rare: counts above the private range, non-counts, padding
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int y = __float2int_rz(v[u]);
                        const unsigned ym1 = (unsigned)(y - 1);
                        const bool isbin = (ym1 < (unsigned)DX_HIST_BINS) && (__int2float_rn(y) == v[u]) && (k[u] < (unsigned)K);
                        if (isbin) {
                            if (ym1 >= YP) atomicAdd(&s_hist[k[u] * DX_HIST_BINS + ym1], 1u);
                        } else if (v[u] != 0.0f && k[u] < (unsigned)K) {
                            ++n_ovf;
                        }
                    }
                }
            };
            load(0, va, ra);
            for (int b = 0; b < n_batches; b += 2) {
                load(b + 1, vb, rb);
                consume(va, ra);
                load(b + 2, va, ra);
                if (b + 1 < n_batches) consume(vb, rb);
            }
            __syncwarp();
            // flush of the warp's own counters: lane l owns word-row l (bins 4l .. 4l+3) and walks the 32 columns
            // with a skew so that the 32 lanes always hit 32 different banks
            if (lane < NW) {
                uint32_t* row = priv_warp + lane * 32;
                unsigned lo = 0, hi = 0;
#pragma unroll 8
                for (int c = 0; c < 32; ++c) {
                    const int col = (c + lane) & 31;
                    const unsigned x = row[col];
                    row[col] = 0u;
                    lo += x & 0x00ff00ffu;
                    hi += (x >> 8) & 0x00ff00ffu;
                }
                const unsigned cnt[4] = {lo & 0xffffu, hi & 0xffffu, lo >> 16, hi >> 16};
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const unsigned idx = 4u * lane + b;
                    if (cnt[b] && idx < n_priv)
                        atomicAdd(&s_hist[(idx >> ysh_u) * DX_HIST_BINS + (idx & (YP - 1u))], cnt[b]);
                }
            }
            __syncwarp();
        }
        const int any_ovf = __syncthreads_or(n_ovf != 0u);
        // write tail counts c_k(j) = #{cells of group k with j < y <= 64}, j = 0..63 (suffix sums of the bins):
        // the NB likelihood needs sum_j c(j) f(r + j), the rank test recovers the bins as c(j-1) - c(j).
        uint32_t* out = hist_out + g * (int64_t)K * DX_HIST_BINS;
        for (int k = warp; k < K; k += NWARP) {
            const unsigned lo = s_hist[k * DX_HIST_BINS + lane], hi = s_hist[k * DX_HIST_BINS + 32 + lane];
            unsigned plo = lo, phi = hi;   // inclusive prefix sums
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned a = __shfl_up_sync(DX_FULL_MASK, plo, o), b = __shfl_up_sync(DX_FULL_MASK, phi, o);
                if (lane >= o) { plo += a; phi += b; }
            }
            const unsigned tot_lo = __shfl_sync(DX_FULL_MASK, plo, 31), tot_hi = __shfl_sync(DX_FULL_MASK, phi, 31);
            out[k * DX_HIST_BINS + 32 + lane] = tot_hi - phi + hi;
            out[k * DX_HIST_BINS + lane] = tot_hi + (tot_lo - plo + lo);
        }
        if (!any_ovf) {
            if (tid == 0) ovf_count[g] = 0;
            __syncthreads();
            continue;
        }
        // deterministic overflow compaction: exclusive scan of per-thread counts, then a second pass
        s_scan[tid] = n_ovf;
        __syncthreads();
        if (warp == 0) {
            unsigned run = 0;
            for (int base = 0; base < T; base += 32) {
                const unsigned v = s_scan[base + lane];
                unsigned inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(DX_FULL_MASK, inc, o); if (lane >= o) inc += t; }
                s_scan[base + lane] = run + inc - v;
                run += __shfl_sync(DX_FULL_MASK, inc, 31);
            }
            if (lane == 0) { s_scan[T] = run; ovf_count[g] = (int32_t)run; }
        }
        __syncthreads();
        {
            // thread t re-visits exactly the entries it counted: chunk by chunk, batch by batch, stride T
            int64_t pos = start + s_scan[tid];
            for (int64_t i = start + tid; i < end; i += T) {
                const float v = __ldg(vals + i);
                const int k = __ldg(cell_group + __ldg(rows + i));
                int y;
                if (!dx_classify(v, k, K, y) && v != 0.0f && k < K) {
                    ovf_val[pos] = v;
                    ovf_group[pos] = (uint8_t)k;
                    ++pos;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// K_ingest, TMA variant (K <= 32, 16-byte aligned shard): one WARP per gene (dynamic gene queue), persistent
// CTAs.  Each warp owns a ring of NS stages in shared memory that lane 0 keeps full with cp.async.bulk
// (TMA bulk copies completing on an mbarrier), across gene boundaries, so ~NS*2 KB per warp are always in
// flight from HBM without holding registers; the streamed data never enters L1, which is left to the
// cell -> group table.  The group gathers of tile t+1 are issued before tile t is counted (software pipeline),
// so their L1/L2 latency hides behind a full tile of work.  Counters / histogram / flush as in the CTA kernel,
// but private to the warp: no block barrier anywhere.
// ---------------------------------------------------------------------------------------------------------
#ifndef DX_TMA_NS
#define DX_TMA_NS 2
#endif
#ifndef DX_TMA_TILE
#define DX_TMA_TILE 256
#endif
#ifndef DX_TMA_TW
#define DX_TMA_TW 20
#endif
constexpr int TW = DX_TMA_TW;     // max warps per CTA (sets the register budget: 65536 / (32 TW))
constexpr int NS = DX_TMA_NS;     // ring stages per warp
constexpr int TILE = DX_TMA_TILE; // elements per stage
constexpr int TU = TILE / 32;     // elements per lane and tile
constexpr int NG = 8;             // gene descriptors in flight per warp (>= NS + 2)
constexpr int DX_INGEST_MODE_GENES = 0, DX_INGEST_MODE_SPLIT = 1, DX_INGEST_MODE_REDO = 2, DX_INGEST_MODE_PIECES = 3;

__device__ __forceinline__ uint32_t dx_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dx_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dx_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dx_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dx_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void dx_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct GeneDesc {          // written once per gene by the issue side, read once by the consume side
    long long start, end;  // the gene's [start, end) in the shard
    int gene, pad;
};
struct TileInfo {          // registers, per fetched tile
    int slot, last;        // gene descriptor slot; 1 when this is the last tile of the gene
};

// INT_ONLY: the caller guarantees that every stored value is a non-negative integer (< 2^24), which removes the
// integrality test from the hot loop.  gene_order (nullable): the sequence in which the queue hands out genes
// (largest first keeps the tail of the launch short).
// H16: the caller guarantees that no design group has more than 65535 cells, so the per-warp histogram fits 16-bit
// bins (half the shared memory -> more resident warps).
// MODE: how the work queue is read (DX_INGEST_MODE_*): a compile-time parameter, so the instantiation of one mode carries neither
// the branches nor the registers of the others through the streaming loop.
template <bool INT_ONLY, bool H16, int MODE>
__global__ void __launch_bounds__(TW * 32, 1)
dx_ingest_tma_kernel(int64_t n_genes, const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                     const float* __restrict__ vals, const uint8_t* __restrict__ cell_group, int K, int ysh, int NW, int NWA,
                     uint32_t* __restrict__ hist_out, int32_t* __restrict__ ovf_count,
                     float* __restrict__ ovf_val, uint8_t* __restrict__ ovf_group, unsigned int* __restrict__ queue,
                     const int32_t* __restrict__ gene_order, int warp_bytes, int /*mode: see MODE*/, int split_rpw,
                     const int4* __restrict__ pieces, int64_t n_pieces) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * warp_bytes;
    float* ring_v = reinterpret_cast<float*>(base);                               // [NS][TILE]
    int32_t* ring_r = reinterpret_cast<int32_t*>(base + NS * TILE * 4);           // [NS][TILE]
    using hist_t = typename std::conditional<H16, uint16_t, uint32_t>::type;
    hist_t* s_hist = reinterpret_cast<hist_t*>(base + 2 * NS * TILE * 4);         // [K][64]
    uint32_t* s_priv = reinterpret_cast<uint32_t*>(s_hist + K * DX_HIST_BINS);    // [NWA][32]
    GeneDesc* s_gene = reinterpret_cast<GeneDesc*>(s_priv + NWA * 32);            // [NG]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_gene + NG);   // [NS]
    const unsigned YP = (ysh >= 0) ? (1u << ysh) : 0u;
    const unsigned ysh_u = (ysh >= 0) ? (unsigned)ysh : 0u;
    const unsigned n_priv = (unsigned)K * YP;
    const unsigned row_mask = (unsigned)(NWA - 1) << 5;
    const unsigned ysh3 = ysh_u + 3u;
    const uint32_t priv_addr = dx_smem_u32(s_priv + lane);
    const uint32_t bar0 = dx_smem_u32(s_bar);
    const uint32_t ringv0 = dx_smem_u32(ring_v), ringr0 = dx_smem_u32(ring_r);

    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < NS; ++st) dx_mbar_init(bar0 + 8u * st, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = lane; i < NWA * 32; i += 32) s_priv[i] = 0u;
    for (int i = lane; i < K * DX_HIST_BINS; i += 32) s_hist[i] = 0;
    const uint32_t hist_addr = dx_smem_u32(s_hist);
    __syncwarp();

    // ---- issue side (warp-uniform state; lane 0 performs the side effects) ----------------------------------
    unsigned issued = 0, consumed = 0, g_head = 0, g_tail = 0;
    bool done_issue = false;
    int i_tiles = 0;                  // tiles of the current issue gene not yet issued
    const float* i_pv = vals;         // next tile of the issue gene
    const int32_t* i_pr = rows;
    int i_left = 0;                   // elements (padded to 4) of the issue gene not yet issued
    // MODE_SPLIT: the queue hands out fixed-size ranges of the shard instead of genes; a range is cut at the gene
    // boundaries inside it into pieces, every piece is counted like a (short) gene and its tail counts are ADDED to the
    // (zeroed) output.  Used when the shard has fewer genes than the GPU has warps (a gene range of a multi-GPU run).
    long long sp_pos = 0, sp_end = 0, sp_lo = 0, sp_hi = 0;     // current range [sp_pos, sp_end) of the shard [sp_lo, sp_hi)
    long long sp_seg = 0;
    long long sp_gene = 0;
    if (MODE == DX_INGEST_MODE_SPLIT) {
        sp_lo = __ldg(indptr);
        sp_hi = __ldg(indptr + n_genes);
        const long long span = sp_hi - sp_lo;
        const long long want = span / ((long long)gridDim.x * (blockDim.x >> 5) * split_rpw) + 1;      // ~split_rpw ranges per warp
        sp_seg = ((want + TILE - 1) / TILE) * TILE;
        if (sp_seg < 4 * TILE) sp_seg = 4 * TILE;
        if (sp_seg > 128 * TILE) sp_seg = 128 * TILE;
    }
    auto issue_one = [&]() {
        if (i_tiles == 0) {
            unsigned gq = 0;
            long long st_ = 0, en_ = 0;
            int piece_multi = (MODE == DX_INGEST_MODE_SPLIT) ? 1 : 0;      // 1: tail counts are added to the output (part of a gene)
            if (MODE == DX_INGEST_MODE_SPLIT) {
                for (;;) {
                    if (sp_pos >= sp_end) {                      // next range of the shard
                        unsigned q = 0;
                        if (lane == 0) q = atomicAdd(queue, 1u);
                        q = __shfl_sync(DX_FULL_MASK, q, 0);
                        sp_pos = sp_lo + (long long)q * sp_seg;
                        if (sp_pos >= sp_hi) { done_issue = true; return; }
                        sp_end = sp_pos + sp_seg < sp_hi ? sp_pos + sp_seg : sp_hi;
                        // the gene that holds position sp_pos: last g with indptr[g] <= sp_pos (pointers are cached in L1)
                        long long lo = 0, hi = n_genes;          // invariant: indptr[lo] <= sp_pos < indptr[hi]
                        while (hi - lo > 1) {
                            const long long mid = (lo + hi) >> 1;
                            if (__ldg(indptr + mid) <= sp_pos) lo = mid; else hi = mid;
                        }
                        sp_gene = lo;
                    }
                    const long long g_end = __ldg(indptr + sp_gene + 1);
                    st_ = sp_pos;
                    en_ = g_end < sp_end ? g_end : sp_end;
                    gq = (unsigned)sp_gene;
                    if (g_end <= sp_end) ++sp_gene;              // the gene ends inside this range: the next piece is the next gene's
                    sp_pos = en_;
                    if (en_ > st_) break;                        // (genes without entries yield no piece)
                }
            } else if (MODE == DX_INGEST_MODE_PIECES) {
                // the queue hands out the pieces of the resident shard's work list (dx_group_suffstats_pieces): whole genes,
                // and equal parts of the genes that are too large to be one unit of work
                unsigned q = 0;
                if (lane == 0) q = atomicAdd(queue, 1u);
                q = __shfl_sync(DX_FULL_MASK, q, 0);
                if ((int64_t)q >= n_pieces) { done_issue = true; return; }
                const int4 pc = __ldg(pieces + q);              // gene, first element (relative to the gene), elements, part of a gene?
                gq = (unsigned)pc.x;
                st_ = __ldg(indptr + gq) + pc.y;
                en_ = st_ + pc.z;
                piece_multi = pc.w;
            } else {
                for (;;) {
                    if (lane == 0) gq = atomicAdd(queue, 1u);
                    gq = __shfl_sync(DX_FULL_MASK, gq, 0);
                    if ((int64_t)gq >= n_genes) { done_issue = true; return; }
                    if (gene_order) gq = (unsigned)__ldg(gene_order + gq);
                    // MODE_REDO: second pass after a split launch, only the genes that reported overflow entries
                    if (MODE != DX_INGEST_MODE_REDO || ovf_count[gq] != 0) break;
                }
                st_ = __ldg(indptr + gq); en_ = __ldg(indptr + gq + 1);
            }
            if (lane == 0) {
                GeneDesc d;
                d.start = st_; d.end = en_; d.gene = (int)gq; d.pad = piece_multi;
                s_gene[g_head % NG] = d;
            }
            ++g_head;
            const long long a0 = st_ & ~3ll;
            i_left = (int)(((en_ + 3ll) & ~3ll) - a0);
            i_tiles = (i_left + TILE - 1) / TILE;
            if (i_tiles < 1) i_tiles = 1;
            i_pv = vals + a0;
            i_pr = rows + a0;
        }
        const unsigned st = issued % NS;
        const int n = i_left > TILE ? TILE : i_left;
        if (lane == 0) {
            const uint32_t bar = bar0 + 8u * st;
            if (n > 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of the stage
                dx_mbar_expect_tx(bar, (uint32_t)(8 * n));
                dx_bulk_g2s(ringv0 + st * (TILE * 4), i_pv, (uint32_t)(4 * n), bar);
                dx_bulk_g2s(ringr0 + st * (TILE * 4), i_pr, (uint32_t)(4 * n), bar);
            } else {
                dx_mbar_arrive(bar);
            }
        }
        i_pv += TILE; i_pr += TILE;
        i_left -= n;
        --i_tiles;
        ++issued;
    };

    // ---- consume side ---------------------------------------------------------------------------------------
    float v0[TU], v1[TU];
    int r0[TU], r1[TU];
    unsigned k0[TU], k1[TU];
    TileInfo m0, m1;
    unsigned n_ovf = 0;               // overflow entries this lane saw in the current tile
    unsigned ovf_base = 0;            // overflow entries of the current gene written so far (warp-uniform)
    int since_flush = 0;
    int c_tiles = 0;                  // tiles of the current consume gene not yet fetched
    int c_lo = 0, c_hi = 0;           // valid element range of the next tile, relative to the tile start
    int c_slot = 0;

    // waits for the next tile, copies it to registers (masked to the gene), frees the stage, refills the ring
    // and issues the group gathers
    auto fetch = [&](float (&v)[TU], int (&r)[TU], unsigned (&k)[TU], TileInfo& m) {
        const unsigned st = consumed % NS;
        dx_mbar_wait(bar0 + 8u * st, (consumed / NS) & 1u);     // acquire: also orders lane 0's descriptor write
        if (c_tiles == 0) {
            c_slot = (int)(g_tail % NG);
            const GeneDesc d = s_gene[c_slot];
            ++g_tail;
            const long long a0 = d.start & ~3ll;
            c_lo = (int)(d.start - a0);
            c_hi = (int)(d.end - a0);
            const int span = (int)(((d.end + 3ll) & ~3ll) - a0);
            c_tiles = (span + TILE - 1) / TILE;
            if (c_tiles < 1) c_tiles = 1;
        }
        const float* sv = ring_v + st * TILE + lane;
        const int32_t* sr = ring_r + st * TILE + lane;
        if (c_lo <= 0 && c_hi >= TILE) {                        // interior tile: no masking
#pragma unroll
            for (int u = 0; u < TU; ++u) { v[u] = sv[32 * u]; r[u] = sr[32 * u]; }
        } else {
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int pos = lane + 32 * u;
                const bool ok = (pos >= c_lo) && (pos < c_hi);
                v[u] = ok ? sv[32 * u] : 0.0f;
                r[u] = ok ? sr[32 * u] : 0;
            }
        }
        __syncwarp();
        ++consumed;
        --c_tiles;
        c_lo -= TILE; c_hi -= TILE;
        m.slot = c_slot;
        m.last = (c_tiles == 0) ? 1 : 0;
        if (!done_issue && issued - consumed < NS) issue_one();
#pragma unroll
        for (int u = 0; u < TU; ++u) k[u] = __ldg(cell_group + (unsigned)r[u]);
    };

    auto flush = [&]() {
        __syncwarp();
        if (lane < NW) {
            uint32_t* row = s_priv + lane * 32;
            unsigned lo = 0, hi = 0;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) {
                const int col = (c + lane) & 31;
                const unsigned x = row[col];
                row[col] = 0u;
                lo += x & 0x00ff00ffu;
                hi += (x >> 8) & 0x00ff00ffu;
            }
            const unsigned cnt[4] = {lo & 0xffffu, hi & 0xffffu, lo >> 16, hi >> 16};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const unsigned idx = 4u * lane + b;
                if (cnt[b] && idx < n_priv) s_hist[(idx >> ysh_u) * DX_HIST_BINS + (idx & (YP - 1u))] += (hist_t)cnt[b];
            }
        }
        __syncwarp();
        since_flush = 0;
    };

    auto process = [&](const float (&v)[TU], const int (&r)[TU], const unsigned (&k)[TU], const TileInfo& m) {
        if (since_flush + TU > 255) flush();
        since_flush += TU;
        // counted two elements at a time: the rare path (counts above the private range, non-counts, padding) is
        // entered per pair, so one unusual entry costs two re-classifications instead of TU
#pragma unroll
        for (int u0 = 0; u0 < TU; u0 += 2) {
            bool slow = false;
#pragma unroll
            for (int u = u0; u < u0 + 2; ++u) {
                const int y = __float2int_rz(v[u]);
                const unsigned ym1 = (unsigned)(y - 1);
                const bool fast = (INT_ONLY || __int2float_rn(y) == v[u]) && (ym1 < YP) && (k[u] < (unsigned)K);
                const unsigned j8 = (k[u] << ysh3) + (ym1 << 3);
                const unsigned inc = fast ? (1u << (j8 & 24u)) : 0u;
                dx_red_shared(priv_addr + ((j8 & row_mask) << 2), inc);
                slow = slow || !fast;
            }
            if (slow) {
#pragma unroll
                for (int u = u0; u < u0 + 2; ++u) {
                    const int y = __float2int_rz(v[u]);
                    const unsigned ym1 = (unsigned)(y - 1);
                    const bool isbin = (ym1 < (unsigned)DX_HIST_BINS) && (__int2float_rn(y) == v[u]) && (k[u] < (unsigned)K);
                    if (isbin) {
                        if (ym1 >= YP) {
                            const unsigned b = k[u] * DX_HIST_BINS + ym1;
                            if (H16) dx_red_shared(hist_addr + ((b >> 1) << 2), 1u << ((b & 1u) << 4));   // 16-bit bin inside its word
                            else dx_red_shared(hist_addr + (b << 2), 1u);
                        }
                    } else if (v[u] != 0.0f && k[u] < (unsigned)K) {
                        ++n_ovf;
                    }
                }
            }
        }
        if (__any_sync(DX_FULL_MASK, n_ovf != 0u)) {
            // Overflow entries of this tile, appended in ascending shard position (element lane + 32 u of the tile):
            // an order that does not depend on where the gene starts, so a sharded run writes the same list.
            const int64_t obase = s_gene[m.slot].start;
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int y = __float2int_rz(v[u]);
                const bool isbin = ((unsigned)(y - 1) < (unsigned)DX_HIST_BINS) && (__int2float_rn(y) == v[u]);
                const bool ovf = !isbin && v[u] != 0.0f && k[u] < (unsigned)K;
                const unsigned bal = __ballot_sync(DX_FULL_MASK, ovf);
                if (ovf && !s_gene[m.slot].pad) {      // (parts of a gene only count: the redo pass writes the list)
                    const int64_t pos = obase + ovf_base + __popc(bal & ((1u << lane) - 1u));
                    ovf_val[pos] = v[u];
                    ovf_group[pos] = (uint8_t)k[u];
                }
                ovf_base += __popc(bal);
            }
            n_ovf = 0;
        }
        if (!m.last) return;
        // ---- the gene is complete: tail counts, overflow count, reset ----------------------------------------
        flush();
        const GeneDesc gd = s_gene[m.slot];      // still valid: at most NS + 1 < NG genes are in flight
        const int64_t g = gd.gene;
        uint32_t* out = hist_out + g * (int64_t)K * DX_HIST_BINS;
#pragma unroll 1
        for (int kk = 0; kk < K; ++kk) {
            const unsigned lo = s_hist[kk * DX_HIST_BINS + lane], hi = s_hist[kk * DX_HIST_BINS + 32 + lane];
            s_hist[kk * DX_HIST_BINS + lane] = 0;
            s_hist[kk * DX_HIST_BINS + 32 + lane] = 0;
            unsigned plo = lo, phi = hi;   // inclusive prefix sums
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned a = __shfl_up_sync(DX_FULL_MASK, plo, o), b = __shfl_up_sync(DX_FULL_MASK, phi, o);
                if (lane >= o) { plo += a; phi += b; }
            }
            const unsigned tot_lo = __shfl_sync(DX_FULL_MASK, plo, 31), tot_hi = __shfl_sync(DX_FULL_MASK, phi, 31);
            const unsigned t_hi = tot_hi - phi + hi, t_lo = tot_hi + (tot_lo - plo + lo);
            if (gd.pad) {          // a part of the gene: tail counts are additive
                if (t_hi) atomicAdd(out + kk * DX_HIST_BINS + 32 + lane, t_hi);
                if (t_lo) atomicAdd(out + kk * DX_HIST_BINS + lane, t_lo);
            } else {
                out[kk * DX_HIST_BINS + 32 + lane] = t_hi;
                out[kk * DX_HIST_BINS + lane] = t_lo;
            }
        }
        if (gd.pad) {
            if (lane == 0 && ovf_base) atomicAdd(ovf_count + g, (int32_t)ovf_base);
        } else if (lane == 0) {
            ovf_count[g] = (int32_t)ovf_base;
        }
        ovf_base = 0;
        __syncwarp();
    };

    while (!done_issue && issued < (unsigned)NS) issue_one();
    if (issued == 0) return;
    fetch(v0, r0, k0, m0);
    for (;;) {      // ping-pong between the two register sets: tile t+1 is fetched (and its gathers issued) before t is counted
        bool more = consumed < issued;
        if (more) fetch(v1, r1, k1, m1);
        process(v0, r0, k0, m0);
        if (!more) break;
        more = consumed < issued;
        if (more) fetch(v0, r0, k0, m0);
        process(v1, r1, k1, m1);
        if (!more) break;
    }
}

// per-gene sum(x), sum(x^2) of the raw shard (optionally of x / size factor): init of the per-cell path
__global__ void __launch_bounds__(256)
dx_gene_moments_kernel(int64_t n_genes, const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                       const float* __restrict__ vals, const double* __restrict__ inv_sf,
                       double* __restrict__ sum1, double* __restrict__ sum2) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_genes) return;
    const int64_t start = indptr[g], end = indptr[g + 1];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = start + lane; i < end; i += 32) {
        double v = (double)__ldg(vals + i);
        if (inv_sf) v *= __ldg(inv_sf + __ldg(rows + i));
        s1 += v; s2 += v * v;
    }
    s1 = dx_warp_sum(s1); s2 = dx_warp_sum(s2);
    if (lane == 0) { sum1[g] = s1; sum2[g] = s2; }
}

}  // namespace

int dx_ingest_private_bins(int K) {
    // largest power of two YP <= 64 with K * YP <= 128; 0 when K > 128 (everything through shared atomics)
    int yp = 0;
    for (int c = 1; c <= DX_HIST_BINS && K * c <= 128; c <<= 1) yp = c;
    return yp;
}

// dynamic gene queues: a ring of counters, zeroed on the launch stream right before the kernel that uses one, so that
// concurrent launches (different streams, a replayed CUDA graph next to eager launches) do not share one.  A slot is handed
// out again after DX_QUEUE_SLOTS further launches of the process: more launches than that in flight at once would alias
// (the library itself keeps at most a few dozen in flight per device).
#define DX_QUEUE_SLOTS 4096
__device__ unsigned int g_dx_queue_slots[DX_QUEUE_SLOTS];

cudaError_t dx_queue_acquire(cudaStream_t stream, unsigned int** out) {
    static std::atomic<unsigned> seq{0};
    void* base = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&base, g_dx_queue_slots);
    if (e != cudaSuccess) return e;
    *out = static_cast<unsigned int*>(base) + (seq.fetch_add(1u) % DX_QUEUE_SLOTS);
    return cudaMemsetAsync(*out, 0, sizeof(unsigned int), stream);
}

// rows of the tail-count table (and the overflow counts) of the genes that are counted in parts
__global__ void __launch_bounds__(128)
dx_zero_genes_kernel(const int32_t* __restrict__ genes, int K, uint32_t* __restrict__ hist, int32_t* __restrict__ ovf_count) {
    const int64_t g = genes[blockIdx.x];
    uint4* row = reinterpret_cast<uint4*>(hist + g * (int64_t)K * DX_HIST_BINS);
    for (int i = threadIdx.x; i < K * DX_HIST_BINS / 4; i += 128) row[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) ovf_count[g] = 0;
}

cudaError_t dx_launch_ingest(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                             const uint8_t* cell_group, int K, uint32_t* hist, int32_t* ovf_count,
                             float* ovf_val, uint8_t* ovf_group, const int32_t* gene_order, int flags, int sm_count,
                             cudaStream_t stream, const int32_t* pieces, int64_t n_pieces, const int32_t* multi_genes,
                             int64_t n_multi) {
    if (n_genes <= 0) return cudaSuccess;
    const int YP = dx_ingest_private_bins(K);
    int ysh = -1;
    for (int c = YP; c > 0; c >>= 1) ++ysh;
    const int NW = (K * YP + 3) / 4;        // packed counter words per thread (<= 32)
    int NWA = 1;                            // rows allocated: next power of two (index masking keeps every update in range)
    while (NWA < NW) NWA <<= 1;
    cudaError_t e;
    // ---- TMA variant: warp per gene, per-warp bulk-copy ring (needs 16-byte aligned arrays, modest K) ----------
    static const char* impl_env = getenv("DXB200_INGEST_IMPL");          // tuning hooks (development only)
    static const char* warps_env = getenv("DXB200_INGEST_WARPS");
    const bool aligned = ((reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(rows)) & 15u) == 0;
    const bool h16 = (flags & DX_GROUPS_FIT_U16) != 0;
    int warp_bytes = 2 * NS * TILE * 4 + K * DX_HIST_BINS * (h16 ? 2 : 4) + NWA * 32 * 4 + NG * (int)sizeof(GeneDesc) + NS * 8;
    warp_bytes = (warp_bytes + 15) & ~15;
    int warps = (227 * 1024) / warp_bytes;
    if (warps > TW) warps = TW;
    if (warps_env && atoi(warps_env) > 0 && atoi(warps_env) < warps) warps = atoi(warps_env);
    const bool want_tma = !(impl_env && impl_env[0] == 'c');
    if (want_tma && aligned && warps >= 6 && n_genes < 2147483647LL) {
        unsigned int* queue = nullptr;
        e = dx_queue_acquire(stream, &queue);
        if (e != cudaSuccess) return e;
        const size_t smem = (size_t)warps * warp_bytes;
        const bool io = (flags & DX_SHARD_INTEGER_COUNTS) != 0;
        // one instantiation per (integer counts, 16-bit bins, queue mode)
#define DX_TMA_PICK(M) (io ? (h16 ? dx_ingest_tma_kernel<true, true, M> : dx_ingest_tma_kernel<true, false, M>)      \
                           : (h16 ? dx_ingest_tma_kernel<false, true, M> : dx_ingest_tma_kernel<false, false, M>))
        auto kern = DX_TMA_PICK(DX_INGEST_MODE_GENES);
        auto kern_pieces = DX_TMA_PICK(DX_INGEST_MODE_PIECES);
        auto kern_redo = DX_TMA_PICK(DX_INGEST_MODE_REDO);
#undef DX_TMA_PICK
        for (auto kfn : {kern, kern_pieces, kern_redo}) {
            e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        // few genes (a gene range of a multi-GPU run): one warp per gene leaves warps idle and the longest gene sets the
        // duration, so the shard is cut into equal ranges of non-zeros instead (pieces of genes, tail counts added with
        // atomics into a zeroed table); genes with overflow entries are then redone gene-wise, which writes their lists in
        // the canonical order.  DXB200_INGEST_SPLIT=0 / 1 forces the mode (development hook).
        if (pieces && n_pieces > 0 && (reinterpret_cast<uintptr_t>(pieces) & 15u) == 0) {
            // work list prepared with the resident shard: genes that are parts of the list are zeroed first, added to by
            // their parts, and -- if they turn out to have overflow entries -- counted once more as whole genes, which
            // writes their lists in the canonical order
            if (n_multi > 0) {
                dx_zero_genes_kernel<<<(unsigned)n_multi, 128, 0, stream>>>(multi_genes, K, hist, ovf_count);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
            }
            kern_pieces<<<(unsigned)sm_count, warps * 32, smem, stream>>>(n_genes, indptr, rows, vals, cell_group, K, ysh, NW, NWA,
                                                                          hist, ovf_count, ovf_val, ovf_group, queue, nullptr,
                                                                          warp_bytes, DX_INGEST_MODE_PIECES, 0,
                                                                          reinterpret_cast<const int4*>(pieces), n_pieces);
            e = cudaGetLastError();
            if (e != cudaSuccess || n_multi <= 0 || (flags & DX_SHARD_NO_OVERFLOW)) return e;
            unsigned int* queue2 = nullptr;
            e = dx_queue_acquire(stream, &queue2);
            if (e != cudaSuccess) return e;
            int64_t grid2 = (n_multi + warps - 1) / warps;
            if (grid2 > sm_count) grid2 = sm_count;
            kern_redo<<<(unsigned)grid2, warps * 32, smem, stream>>>(n_multi, indptr, rows, vals, cell_group, K, ysh, NW, NWA, hist,
                                                                     ovf_count, ovf_val, ovf_group, queue2, multi_genes, warp_bytes,
                                                                     DX_INGEST_MODE_REDO, 0, nullptr, 0);
            return cudaGetLastError();
        }
        const char* split_env = getenv("DXB200_INGEST_SPLIT");
        bool split = n_genes < 2 * (int64_t)sm_count * warps;
        if (split_env) split = split_env[0] == '1';
        if (split) {
            const char* rpw_env = getenv("DXB200_INGEST_RPW");          // tuning hook: ranges per warp of the split mode
            const int split_rpw = (rpw_env && atoi(rpw_env) >= 1 && atoi(rpw_env) <= 64) ? atoi(rpw_env) : 2;
            // a piece never holds more than 128 tiles: its counts always fit 16-bit on-chip bins (the shared-memory
            // layout sized for 32-bit bins holds them as well)
            auto kern_split = io ? dx_ingest_tma_kernel<true, true, DX_INGEST_MODE_SPLIT>
                                 : dx_ingest_tma_kernel<false, true, DX_INGEST_MODE_SPLIT>;
            e = cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(hist, 0, (size_t)n_genes * K * DX_HIST_BINS * sizeof(uint32_t), stream);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(ovf_count, 0, (size_t)n_genes * sizeof(int32_t), stream);
            if (e != cudaSuccess) return e;
            kern_split<<<(unsigned)sm_count, warps * 32, smem, stream>>>(n_genes, indptr, rows, vals, cell_group, K, ysh, NW, NWA,
                                                                         hist, ovf_count, ovf_val, ovf_group, queue, nullptr,
                                                                         warp_bytes, DX_INGEST_MODE_SPLIT, split_rpw, nullptr, 0);
            e = cudaGetLastError();
            if (e != cudaSuccess || (flags & DX_SHARD_NO_OVERFLOW)) return e;
            unsigned int* queue2 = nullptr;
            e = dx_queue_acquire(stream, &queue2);
            if (e != cudaSuccess) return e;
            int64_t grid2 = sm_count < 16 ? sm_count : 16;          // usually nothing to redo: a small grid walks the counts
            if (grid2 > n_genes) grid2 = n_genes;
            kern_redo<<<(unsigned)grid2, warps * 32, smem, stream>>>(n_genes, indptr, rows, vals, cell_group, K, ysh, NW, NWA, hist,
                                                                     ovf_count, ovf_val, ovf_group, queue2, nullptr, warp_bytes,
                                                                     DX_INGEST_MODE_REDO, 0, nullptr, 0);
            return cudaGetLastError();
        }
        int64_t grid = sm_count;                    // persistent CTAs; the warps pull genes from the queue
        if (grid > n_genes) grid = n_genes;
        kern<<<(unsigned)grid, warps * 32, smem, stream>>>(n_genes, indptr, rows, vals, cell_group, K, ysh, NW, NWA, hist,
                                                           ovf_count, ovf_val, ovf_group, queue, gene_order, warp_bytes,
                                                           DX_INGEST_MODE_GENES, 0, nullptr, 0);
        return cudaGetLastError();
    }
    // ---- CTA-per-gene variant (any K, any alignment) -------------------------------------------------------------
    const size_t smem = (size_t)K * DX_HIST_BINS * 4 + (T + 4) * 4 + (size_t)NWARP * NWA * 32 * 4 + 16;
    e = cudaFuncSetAttribute(dx_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dx_ingest_kernel, T, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count * per_sm;
    if (grid > n_genes) grid = n_genes;
    dx_ingest_kernel<<<(unsigned)grid, T, smem, stream>>>(n_genes, indptr, rows, vals, cell_group, K, ysh, NW, NWA,
                                                           hist, ovf_count, ovf_val, ovf_group);
    return cudaGetLastError();
}

cudaError_t dx_launch_gene_moments(int64_t n_genes, const int64_t* indptr, const int32_t* rows, const float* vals,
                                   const double* inv_sf, double* sum1, double* sum2, cudaStream_t stream) {
    if (n_genes <= 0) return cudaSuccess;
    const int wpb = 8;
    const unsigned grid = (unsigned)((n_genes + wpb - 1) / wpb);
    dx_gene_moments_kernel<<<grid, wpb * 32, 0, stream>>>(n_genes, indptr, rows, vals, inv_sf, sum1, sum2);
    return cudaGetLastError();
}
