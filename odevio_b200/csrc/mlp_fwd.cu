// Generic small-MLP evaluation on CUDA cores: out = act_n(W_n ... act_0(W_0 x + b_0) ... + b_n).
// Serves the module interfaces that are called directly and are NOT on the fused paths:
//   CDEFunc.forward(t, z) -> [B, Hc, C]        reference src/models/ODEFunc.py:81-84
//   ODEFunc.forward(t, x) for shapes the tcgen05 kernel (odefunc_tc.cu) does not cover
// One CTA per 8 rows; the layer input lives in shared memory, every thread owns output columns and
// accumulates sequentially over k in fp32 (same operation order for every row -> deterministic).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "common.cuh"

namespace odevio {

namespace {

constexpr int ML_ROWS = 8;
constexpr int ML_MAX_LAYERS = ODEVIO_MAX_ODE_LINEARS;

struct MlpParams {
  int M, NL;
  int K[ML_MAX_LAYERS], N[ML_MAX_LAYERS], act[ML_MAX_LAYERS];
  const float* W[ML_MAX_LAYERS];      // PyTorch [N][K]
  const float* b[ML_MAX_LAYERS];
  const float* x; float* out;
  int kmax;                           // max hidden width kept in shared memory
};

__global__ void __launch_bounds__(256) mlp_fwd_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ float sm[];
  float* cur = sm;                    // [ML_ROWS][kmax]
  float* nxt = sm + ML_ROWS * p.kmax;
  const int row0 = blockIdx.x * ML_ROWS;
  for (int e = threadIdx.x; e < ML_ROWS * p.K[0]; e += blockDim.x) {
    const int r = e / p.K[0], k = e - r * p.K[0];
    cur[r * p.kmax + k] = (row0 + r < p.M) ? p.x[static_cast<size_t>(row0 + r) * p.K[0] + k] : 0.f;
  }
  __syncthreads();
  for (int l = 0; l < p.NL; ++l) {
    const int K = p.K[l], N = p.N[l];
    const bool last = l == p.NL - 1;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float* w = p.W[l] + static_cast<size_t>(n) * K;
      float acc[ML_ROWS];
#pragma unroll
      for (int r = 0; r < ML_ROWS; ++r) acc[r] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float wk = w[k];
#pragma unroll
        for (int r = 0; r < ML_ROWS; ++r) acc[r] = fmaf(cur[r * p.kmax + k], wk, acc[r]);
      }
      const float bn = p.b[l][n];
#pragma unroll
      for (int r = 0; r < ML_ROWS; ++r) {
        const float v = apply_act(acc[r] + bn, p.act[l]);
        if (last) { if (row0 + r < p.M) p.out[static_cast<size_t>(row0 + r) * N + n] = v; }
        else nxt[r * p.kmax + n] = v;
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace
}  // namespace odevio

using namespace odevio;

extern "C" ODEVIO_API int32_t odevio_mlp_forward(int32_t M, int32_t n_linears, const int32_t* dims, const int32_t* acts,
                                      const float* const* weights, const float* const* biases,
                                      const float* x, float* out, void* stream_) {
  if (!dims || !acts || !weights || !biases || !x || !out) return ODEVIO_E_NULL;
  if (M <= 0 || n_linears < 1 || n_linears > ML_MAX_LAYERS) return ODEVIO_E_SHAPE;
  MlpParams p;
  p.M = M; p.NL = n_linears; p.x = x; p.out = out; p.kmax = 0;
  for (int l = 0; l < n_linears; ++l) {
    if (!weights[l] || !biases[l]) return ODEVIO_E_NULL;
    if (dims[l] <= 0 || dims[l + 1] <= 0) return ODEVIO_E_SHAPE;
    if (acts[l] < 0 || acts[l] > ACT_SIGMOID) return ODEVIO_E_ENUM;
    p.K[l] = dims[l]; p.N[l] = dims[l + 1]; p.act[l] = acts[l]; p.W[l] = weights[l]; p.b[l] = biases[l];
    if (dims[l] > p.kmax) p.kmax = dims[l];            // inputs of every layer live in shared memory
  }
  const size_t smem = static_cast<size_t>(2) * ML_ROWS * p.kmax * sizeof(float);
  if (smem > 200 * 1024) return ODEVIO_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int32_t>(e);
  mlp_fwd_kernel<<<(M + ML_ROWS - 1) / ML_ROWS, 256, smem, static_cast<cudaStream_t>(stream_)>>>(p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}
