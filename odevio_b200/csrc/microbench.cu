// Diagnostics: a dense FFMA loop to measure this GPU's fp32 FMA peak, the roofline denominator of
// the fp32 (CUDA-core) mode of the vector-field GEMMs.  Not on the product path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "../../include/odevio_debug.h"

namespace odevio {

__global__ void __launch_bounds__(512, 2) ffma_peak_kernel(int iters, float seed, float* sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
  const float m = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;   // never true; keeps the chain alive
}

}  // namespace odevio

extern "C" ODEVIO_API int32_t odevio_microbench_ffma(int32_t iters, int32_t blocks, float* sink, double* flops_out,
                                                     void* stream) {
  if (!sink || iters <= 0 || blocks <= 0) return ODEVIO_E_NULL;
  odevio::ffma_peak_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(stream)>>>(iters, 1.0f, sink);
  if (flops_out) *flops_out = 2.0 * 16 * 8 * static_cast<double>(iters) * 512.0 * blocks;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}
