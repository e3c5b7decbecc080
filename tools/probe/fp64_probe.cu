// Probe: DFMA dependent-issue latency and throughput on this GPU (one CTA on one SM, W warps).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) x[u] = threadIdx.x * 1e-3 + u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < ILP; ++u) x[u] = fma(x[u], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u) s += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP> void run(int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    dfma_chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    dfma_chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * ILP;          // DFMA per thread
    printf("ILP %d warps/CTA %2d (per SMSP %.1f): %.2f cycles per DFMA per warp, %.2f cycles per warp-DFMA per SMSP\n",
           ILP, warps, warps / 4.0, h / n, h / (n * warps / 4.0 < n ? n : n * warps / 4.0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
