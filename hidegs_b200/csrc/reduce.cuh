// reduce.cuh — deterministic CTA-wide reductions over per-block partials (second stage of the two-stage
// reductions of the loss kernels).  The order of additions is fixed by (n, blockDim), never by timing.
#pragma once
#include <cuda_runtime.h>

namespace hg {

__device__ __forceinline__ double warp_sum_double(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of p[offset + i*stride], i in [0, n), over the whole CTA (blockDim.x a multiple of 32, <= 1024).
// Four independent accumulators per thread keep the loads in flight.  Result valid in EVERY thread.
// `sm` must hold 32 doubles; the function ends with a barrier, so it can be called back to back.
__device__ __forceinline__ double cta_sum_strided(const double* __restrict__ p, int n, int stride, int offset,
                                                  double* sm) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * T < n; i += 4 * T) {
    a0 += p[(size_t)i * stride + offset];
    a1 += p[(size_t)(i + T) * stride + offset];
    a2 += p[(size_t)(i + 2 * T) * stride + offset];
    a3 += p[(size_t)(i + 3 * T) * stride + offset];
  }
  for (; i < n; i += T) a0 += p[(size_t)i * stride + offset];
  double v = warp_sum_double((a0 + a1) + (a2 + a3));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  const int nw = (T + 31) >> 5;
  v = warp_sum_double(lane < nw ? sm[lane] : 0.0);
  __syncthreads();
  return v;
}

}  // namespace hg
