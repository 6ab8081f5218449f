// Deferred weight-gradient reduction of the fused backward.
//
//   dW[n][k] = sum_m G[m][n] * A[m][k]        (PyTorch [out][in] layout of nn.Linear.weight.grad)
//   db[n]    = sum_m G[m][n]
//
// A [M][K] / G [M][N] are the row-major record streams written by odernn_bwd.cu: one row per
// (sequence row, vector-field evaluation) with the Linear's input and its pre-activation
// gradient.  M is 1e5..1e6+ (every stage of every accepted solver step of every row), so unlike
// the 16-row tile GEMMs of the solver this IS a genuinely dense contraction: 128x128 output
// tiles, split over M across CTAs, partials reduced in a fixed order (deterministic).
// fp32 FFMA (autograd parity of the reference's fp32 training, scripts/train_model.py:63-66,78).
#include <cuda_runtime.h>
#include <stdint.h>

namespace odevio {

namespace {

constexpr int WG_T = 128;     // output tile edge
constexpr int WG_MC = 16;     // m-rows per pipeline stage
constexpr int WG_ST = 3;      // cp.async stages

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;            // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid = (tiles_k, tiles_n, splits); block = 256.  part[z][n][k] (ld = K).
__global__ void __launch_bounds__(256, 2)
wgrad_gemm_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ A, int lda,
                  long long M, int N, int K, long long rows_per_split, float* __restrict__ part) {
  __shared__ __align__(16) float sG[WG_ST][WG_MC][WG_T];
  __shared__ __align__(16) float sA[WG_ST][WG_MC][WG_T];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * WG_T, k0 = blockIdx.x * WG_T;
  const long long m_begin = static_cast<long long>(blockIdx.z) * rows_per_split;
  long long m_end = m_begin + rows_per_split;
  if (m_end > M) m_end = M;
  const int nchunks = m_end > m_begin ? static_cast<int>((m_end - m_begin + WG_MC - 1) / WG_MC) : 0;

  // loader: 16 rows x 32 float4 per operand = 512 float4, 2 per thread
  auto load_stage = [&](int st, int chunk) {
    const long long mb = m_begin + static_cast<long long>(chunk) * WG_MC;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = tid + h * 256;
      const int r = idx >> 5, c4 = (idx & 31) * 4;
      const long long m = mb + r;
      const bool okm = m < m_end;
      const bool okg = okm && (n0 + c4 < N);
      const bool oka = okm && (k0 + c4 < K);
      cp_async16(&sG[st][r][c4], okg ? G + m * ldg + n0 + c4 : G, okg);
      cp_async16(&sA[st][r][c4], oka ? A + m * lda + k0 + c4 : A, oka);
    }
  };

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;       // n = {ty*4 + 0..3, 64 + ty*4 + 0..3}, k likewise with tx
  for (int s = 0; s < WG_ST - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait<WG_ST - 2>();
    __syncthreads();
    const int nx = ch + WG_ST - 1;
    if (nx < nchunks) load_stage(nx % WG_ST, nx);
    cp_async_commit();
    const int st = ch % WG_ST;
#pragma unroll
    for (int r = 0; r < WG_MC; ++r) {
      const float4 g0 = *reinterpret_cast<const float4*>(&sG[st][r][ty * 4]);
      const float4 g1 = *reinterpret_cast<const float4*>(&sG[st][r][64 + ty * 4]);
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[st][r][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[st][r][64 + tx * 4]);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(g[u], a[v], acc[u][v]);
    }
  }
  cp_async_wait<0>();

  float* out = part + static_cast<size_t>(blockIdx.z) * N * K;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int n = n0 + (u < 4 ? ty * 4 + u : 64 + ty * 4 + (u - 4));
    if (n >= N) continue;
#pragma unroll
    for (int hv = 0; hv < 2; ++hv) {
      const int k = k0 + hv * 64 + tx * 4;
      if (k + 3 < K) {
        *reinterpret_cast<float4*>(out + static_cast<size_t>(n) * K + k) =
            make_float4(acc[u][hv * 4], acc[u][hv * 4 + 1], acc[u][hv * 4 + 2], acc[u][hv * 4 + 3]);
      } else {
        for (int v = 0; v < 4; ++v)
          if (k + v < K) out[static_cast<size_t>(n) * K + k + v] = acc[u][hv * 4 + v];
      }
    }
  }
}

// out[i (+ column remap)] = sum_z part[z][i]; optional split of the K axis into two destinations
// (rnn: [dW_ih | dW_hh] from the [x ; h] record).  ld_split = K of the first destination (0: none).
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int N, int K,
                                    float* __restrict__ out0, float* __restrict__ out1, int k_split,
                                    int accumulate = 0) {
  const size_t total = static_cast<size_t>(N) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[static_cast<size_t>(z) * total + i];
    if (!out1) {
      out0[i] = accumulate ? out0[i] + s : s;
    } else {
      const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<size_t>(n) * K);
      if (k < k_split) out0[static_cast<size_t>(n) * k_split + k] = s;
      else out1[static_cast<size_t>(n) * (K - k_split) + (k - k_split)] = s;
    }
  }
}

// part[z][n] = sum over the z-th slab of rows of G[m][n]; grid = (ceil(N/128), splits), block = 128
__global__ void colsum_kernel(const float* __restrict__ G, int ldg, long long M, int N, long long rows_per_split,
                              float* __restrict__ part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.y) * rows_per_split;
  long long m1 = m0 + rows_per_split;
  if (m1 > M) m1 = M;
  if (n >= N) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long long m = m0;
  for (; m + 3 < m1; m += 4) {
    s0 += G[m * ldg + n]; s1 += G[(m + 1) * ldg + n]; s2 += G[(m + 2) * ldg + n]; s3 += G[(m + 3) * ldg + n];
  }
  for (; m < m1; ++m) s0 += G[m * ldg + n];
  part[static_cast<size_t>(blockIdx.y) * N + n] = (s0 + s1) + (s2 + s3);
}

}  // namespace

// Number of M-splits used for an [N][K] output (host; also sizes the partial buffer).
int wgrad_splits(long long M, int N, int K, int nsm) {
  const long long tiles = static_cast<long long>((N + WG_T - 1) / WG_T) * ((K + WG_T - 1) / WG_T);
  long long s = (4LL * nsm + tiles - 1) / tiles;          // ~2 waves at 2 CTAs/SM
  const long long max_by_rows = (M + 4 * WG_MC - 1) / (4 * WG_MC);
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return static_cast<int>(s);
}

// dW (and optionally db) of one Linear from its record streams.  `part` must hold
// wgrad_splits(...) * N * K floats (and >= 256 * N for the bias pass).
cudaError_t wgrad_linear_ex(const float* G, int ldg, const float* A, int lda, long long M, int N, int K,
                            float* dW0, float* dW1, int k_split, float* db0, float* db1, float* part, int nsm,
                            int accumulate, cudaStream_t stream);

cudaError_t wgrad_linear(const float* G, int ldg, const float* A, int lda, long long M, int N, int K,
                         float* dW0, float* dW1, int k_split, float* db0, float* db1, float* part, int nsm,
                         cudaStream_t stream) {
  return wgrad_linear_ex(G, ldg, A, lda, M, N, K, dW0, dW1, k_split, db0, db1, part, nsm, 0, stream);
}

// accumulate != 0 (single destination only): dW0 += ..., db0 += ...   (record streams reduced chunk by chunk)
cudaError_t wgrad_linear_ex(const float* G, int ldg, const float* A, int lda, long long M, int N, int K,
                            float* dW0, float* dW1, int k_split, float* db0, float* db1, float* part, int nsm,
                            int accumulate, cudaStream_t stream) {
  if (accumulate && (dW1 || db1)) return cudaErrorInvalidValue;
  if (M <= 0 && accumulate) return cudaSuccess;
  if (M <= 0) {
    cudaError_t e = cudaMemsetAsync(dW0, 0, sizeof(float) * static_cast<size_t>(N) * (dW1 ? k_split : K), stream);
    if (e != cudaSuccess) return e;
    if (dW1) { e = cudaMemsetAsync(dW1, 0, sizeof(float) * static_cast<size_t>(N) * (K - k_split), stream); if (e != cudaSuccess) return e; }
    if (db0) { e = cudaMemsetAsync(db0, 0, sizeof(float) * N, stream); if (e != cudaSuccess) return e; }
    if (db1) { e = cudaMemsetAsync(db1, 0, sizeof(float) * N, stream); if (e != cudaSuccess) return e; }
    return cudaSuccess;
  }
  const int splits = wgrad_splits(M, N, K, nsm);
  long long rps = (M + splits - 1) / splits;
  rps = (rps + WG_MC - 1) / WG_MC * WG_MC;
  dim3 grid((K + WG_T - 1) / WG_T, (N + WG_T - 1) / WG_T, splits);
  wgrad_gemm_kernel<<<grid, 256, 0, stream>>>(G, ldg, A, lda, M, N, K, rps, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t total = static_cast<size_t>(N) * K;
  int rb = static_cast<int>((total + 255) / 256);
  if (rb > 4 * nsm) rb = 4 * nsm;
  wgrad_reduce_kernel<<<rb, 256, 0, stream>>>(part, splits, N, K, dW0, dW1, k_split, accumulate);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (db0) {
    int bs = static_cast<int>((M + 2047) / 2048);
    if (bs > 256) bs = 256;
    if (bs < 1) bs = 1;
    const long long brps = (M + bs - 1) / bs;
    dim3 g2((N + 127) / 128, bs);
    colsum_kernel<<<g2, 128, 0, stream>>>(G, ldg, M, N, brps, part);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    wgrad_reduce_kernel<<<(N + 255) / 256, 256, 0, stream>>>(part, bs, 1, N, db0, nullptr, 0, accumulate);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (db1) {
      e = cudaMemcpyAsync(db1, db0, sizeof(float) * N, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

}  // namespace odevio
