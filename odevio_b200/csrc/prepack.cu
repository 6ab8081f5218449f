// Weight pre-packing: PyTorch [out, in] row-major -> K-major [K][ldN] blocks that the TMA ring
// of tile_gemm.cuh streams as contiguous KC-row chunks.  Runs on the caller's stream at the top
// of every forward (a few MB, microseconds), so the library stays stateless.
#include <cuda_runtime.h>

namespace odevio {

// dst[(k_off + k) * ldN + n_off + n] = src[n * K + k],  n < N, k < K   (32x32 smem transpose)
__global__ void transpose_pack_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst,
                                      int ldN, int k_off, int n_off) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int n = n0 + j, k = k0 + threadIdx.x;
    tile[j][threadIdx.x] = (n < N && k < K) ? src[static_cast<size_t>(n) * K + k] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, n = n0 + threadIdx.x;
    if (n < N && k < K) dst[static_cast<size_t>(k_off + k) * ldN + n_off + n] = tile[threadIdx.x][j];
  }
}

// dst[i] = a[i] + (b ? b[i] : 0)
__global__ void bias_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = b ? __fadd_rn(a[i], b[i]) : a[i];
}

cudaError_t transpose_pack(const float* src, int N, int K, float* dst, int ldN, int k_off, int n_off,
                           cudaStream_t stream) {
  dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
  transpose_pack_kernel<<<grid, block, 0, stream>>>(src, N, K, dst, ldN, k_off, n_off);
  return cudaGetLastError();
}

cudaError_t bias_sum(const float* a, const float* b, float* dst, int n, cudaStream_t stream) {
  bias_sum_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, dst, n);
  return cudaGetLastError();
}

}  // namespace odevio
