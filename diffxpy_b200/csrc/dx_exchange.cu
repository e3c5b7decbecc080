// dx_exchange.cu -- all-gather of per-gene result rows over NVLink peer memory (one node, one process per GPU).
//
// The only exchange on the path is the vector of per-gene p-values (8 bytes per gene and rank) that the
// Benjamini-Hochberg step needs in full on every rank.  At that size an all-gather is pure latency
// (ring steps, protocol hand-shakes), so the gather is one kernel here (20-40 us on 8 GPUs): every rank stores
// its rows straight into the exchange buffer of every peer (posted NVLink writes), the last CTA publishes a sequence
// number in each peer's flag row with a system-scope release, and the same kernel acquires the flags of all ranks
// before it copies the gathered vector out.  Data slots are double-buffered by sequence parity:
// a rank can only be one exchange ahead of a peer (it needs the peer's flag of exchange s to finish s), so the slot
// written in exchange s + 1 is never the one a peer still reads in exchange s.
//
// Buffer layout (every rank allocates the same):  [0, 256): uint32 flags[64] (flag[r] = last sequence number rank r
// published here);  [256, 260): error word;  [260, 264): CTA ticket (stores fenced);  [264, 268): number of finished
// exchanges;  [268, 272): CTA ticket (leaving);  [512, ...): two data
// slots of slot_bytes each.  The buffers are cudaMalloc allocations shared through CUDA IPC handles, which the host
// side exchanges with its own plumbing (torch.distributed all_gather_object in diffxpy_b200/dist.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dx_internal.h"

namespace {

constexpr int EX_HEADER = 512;
constexpr int EX_MAX_WORLD = 16;

struct PeerPtrs { unsigned char* p[EX_MAX_WORLD]; };

__device__ __forceinline__ void ex_st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ex_ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ex_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One launch: store to every peer, publish, wait for every peer, copy the gathered slot out.  The grid never exceeds
// the SM count, so every CTA is resident while it spins on the flags.
__global__ void __launch_bounds__(256)
dx_peer_gather_kernel(int64_t n, const double* __restrict__ src, PeerPtrs peers, int world, int rank, int64_t slot_bytes,
                      double* __restrict__ dst, unsigned long long timeout_ns) {
    // the sequence number lives in the buffer header (so a captured CUDA graph can be replayed): this exchange is
    // number `done + 1`; the last CTA to leave the kernel stores it back
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(peers.p[rank] + 264) + 1u;
    const int64_t slot_off = EX_HEADER + (int64_t)(seq & 1u) * slot_bytes;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double v = src[i];
        for (int r = 0; r < world; ++r)
            reinterpret_cast<double*>(peers.p[r] + slot_off)[(int64_t)rank * n + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    unsigned char* local = peers.p[rank];
    if (threadIdx.x == 0) {
        uint32_t* ticket = reinterpret_cast<uint32_t*>(local + 260);
        const uint32_t t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {              // every CTA of this rank has fenced its stores
            *ticket = 0u;
            __threadfence_system();
            for (int r = 0; r < world; ++r) ex_st_release_sys(reinterpret_cast<uint32_t*>(peers.p[r]) + rank, seq);
        }
    }
    if ((int)threadIdx.x < world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(local) + threadIdx.x;
        const unsigned long long t0 = ex_globaltimer();
        // sequence numbers only grow; the signed difference keeps the comparison valid across the 2^32 wrap
        while ((int32_t)(ex_ld_acquire_sys(flag) - seq) < 0) {
            if (ex_globaltimer() - t0 > timeout_ns) {          // a peer died: report instead of hanging the GPU
                atomicExch(reinterpret_cast<uint32_t*>(local + 256), 1u + threadIdx.x);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (dst) {
        const double* slot = reinterpret_cast<const double*>(local + slot_off);
        const int64_t total = (int64_t)world * n;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) dst[i] = __ldcg(slot + i);   // not through L1
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* leave = reinterpret_cast<uint32_t*>(local + 268);
        if (atomicAdd(leave, 1u) == gridDim.x - 1) {           // every CTA has read `done`
            *leave = 0u;
            *reinterpret_cast<volatile uint32_t*>(local + 264) = seq;
        }
    }
}

}  // namespace

cudaError_t dx_launch_peer_allgather(int64_t n, const double* src, void* const* peer_bufs, int world, int rank,
                                     int64_t slot_bytes, double* dst, int sm_count, cudaStream_t stream) {
    if (world < 1 || world > EX_MAX_WORLD || rank < 0 || rank >= world) return cudaErrorInvalidValue;
    if ((int64_t)world * n * 8 > slot_bytes) return cudaErrorInvalidValue;
    PeerPtrs pp;
    for (int r = 0; r < EX_MAX_WORLD; ++r) pp.p[r] = r < world ? static_cast<unsigned char*>(peer_bufs[r]) : nullptr;
    int64_t grid = (n + 255) / 256;          // an empty shard still publishes its flag
    if (grid < 1) grid = 1;
    if (grid > sm_count) grid = sm_count;
    dx_peer_gather_kernel<<<(unsigned)grid, 256, 0, stream>>>(n, src, pp, world, rank, slot_bytes, dst, 5000000000ull);
    cudaError_t e = cudaGetLastError();
    return e;
}
