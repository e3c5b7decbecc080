// dx_fit_args.cuh -- launch arguments shared by the two grouped K_fit kernels (dx_fit_grouped.cu: any design /
// dispersion model; dx_fit_lane.cu: lane-resident variant for <= 32 groups, <= 9 coefficients, constant
// dispersion) and the per-gene epilogue both of them run after finalize: the single-coefficient Wald row
// (/root/reference/diffxpy/testing/det.py:783-802 + stats.py:181-218) and, on several GPUs, the store of the p-value
// into every rank's exchange buffer (the one exchange of the path), so that neither a separate Wald launch
// nor a separate gather launch is needed.
#pragma once
#include "dx_common.cuh"
#include "dx_internal.h"

constexpr int DX_EX_HEADER = 512;          // exchange buffer header bytes (layout in dx_exchange.cu)
constexpr int DX_EX_MAX_WORLD = 16;
constexpr int DX_LANE_MAX_OVF = 0;

struct FitEpilogue {
    int wald_coef;                        // < 0: no Wald row
    double *theta, *sd, *pval, *logp;     // [n_genes] each, nullable
    // multi-GPU: p-values go straight into slot (seq & 1) of every rank's exchange buffer at global gene index
    // gene_offset + g; the last CTA of the launch publishes the sequence number in every rank's flag row
    int world, rank;
    int publish;                          // 0: this launch only stores p-values, a later launch of the same fit publishes
    int64_t gene_offset, slot_bytes;
    unsigned char* peer[DX_EX_MAX_WORLD];
    unsigned int* done_ctas;              // device counter of the launch (zeroed by the launcher)
};

struct FitArgs {
    int64_t n_genes, stride;          // stride: distance (in genes) between coefficient rows of a_var / b_var
    int K, p, ps, KL;
    const double *xu_loc, *xu_scale, *offset, *n_k, *pinv_loc;
    const int32_t* loc_group;
    const uint32_t* tail_all;
    const int32_t* ovf_count;
    const int64_t* indptr;
    const float* ovf_val;
    const uint8_t* ovf_group;
    const int32_t* ovf_packed;        // [n_genes][2] record counts of dx_pack_overflow (0, 0 = plain list); may be null
    dx_fit_opts opts;
    double *a_var, *b_var, *fim_inv, *ll_out, *jac_out;
    int32_t *niter_out, *flags_out;
    int slot_doubles, table_doubles, nent, nz_max, ainv_extra;
    unsigned int* queue;              // gene queue of this launch (dx_queue_acquire)
    // Genes with overflow entries (counts above 64 or non-integer values: highly expressed genes, normalised data) are not
    // fitted by the lane-resident kernel -- a single thread would walk the whole list at every likelihood evaluation -- but
    // marked with niter = -1 and fitted by the general kernel, whose 16 lanes share the list, in a second launch
    // (only_deferred = 1).
    int only_deferred;
    FitEpilogue epi;
};

__device__ __forceinline__ void dx_ex_st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t dx_ex_ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// Wald row of gene g from its coefficient and the diagonal element of the inverse Fisher information (one lane).
// det.py:795-797: variance floored at the smallest subnormal (NaN stays NaN); stats.py:215-217.
__device__ __forceinline__ void dx_fit_epilogue_row(const FitEpilogue& E, int64_t g, double theta, double var_cc) {
    double v = var_cc;
    if (v < DX_TINY) v = DX_TINY;
    const double s = sqrt(v);
    double pv, lp;
    dx_two_sided_normal(theta / ((s < DX_TINY) ? DX_TINY : s), &pv, &lp);
    if (isnan(s) || isnan(theta)) { pv = NAN; lp = NAN; }
    if (E.theta) E.theta[g] = theta;
    if (E.sd) E.sd[g] = s;
    if (E.pval) E.pval[g] = pv;
    if (E.logp) E.logp[g] = lp;
    if (E.world > 1) {
        const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(E.peer[E.rank] + 264) + 1u;
        const int64_t off = DX_EX_HEADER + (int64_t)(seq & 1u) * E.slot_bytes + 8 * (E.gene_offset + g);
#pragma unroll 1
        for (int r = 0; r < E.world; ++r) *reinterpret_cast<double*>(E.peer[r] + off) = pv;
    }
}

// End of the launch (every thread of the CTA calls it): the last CTA publishes this rank's sequence number in every
// rank's flag row -- all p-value stores of the launch are fenced before it.
__device__ __forceinline__ void dx_fit_epilogue_publish(const FitEpilogue& E) {
    if (E.world <= 1 || !E.publish) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(E.done_ctas, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(E.peer[E.rank] + 264) + 1u;
            for (int r = 0; r < E.world; ++r) dx_ex_st_release_sys(reinterpret_cast<uint32_t*>(E.peer[r]) + E.rank, seq);
        }
    }
}

// launcher of the lane-resident variant (dx_fit_lane.cu); A.queue / A.epi.done_ctas are filled in by it
bool dx_fit_lane_eligible(int K, int p, int ps, const dx_fit_opts* opts);
cudaError_t dx_launch_fit_lane(FitArgs& A, int sm_count, cudaStream_t stream);
