// dx_fit_percell.cu -- K_fit (per-cell design): placeholder until the streaming kernel lands.
#include "dx_common.cuh"
#include "dx_internal.h"

cudaError_t dx_launch_fit_percell(int64_t, int64_t, int, int, const double*, const double*, const double*,
                                  const int64_t*, const int32_t*, const float*, const dx_fit_opts*,
                                  double*, double*, double*, double*, double*, int32_t*, int32_t*, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
