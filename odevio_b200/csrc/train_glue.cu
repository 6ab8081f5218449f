// Training-loop glue of the regressor path, the step immediately downstream of the fused
// backward -- reference scripts/train_model.py:72-86 and src/utils/utils.py:143-157:
//   * pose loss  100 * MSE(pose[:, :, :3], gt[:, :, :3]) + MSE(pose[:, :, 3:], gt[:, :, 3:])  and its gradient w.r.t. the
//     poses in one pass (the reference's three autograd ops + two reductions);
//   * global-norm clip (torch.nn.utils.clip_grad_norm_, max_norm 5) + Adam (torch.optim.Adam: betas 0.9 / 0.999, eps 1e-8,
//     L2 weight_decay 5e-5) on ONE flat fp32 bucket -- the same bucket the NCCL gradient all-reduce uses
//     (odevio_b200/distributed.py), so the optimiser step is the all-reduce's epilogue: no per-tensor launches, the
//     clip coefficient stays on the device (no host synchronisation).
// HBM-bound elementwise work: 16 B read + 12 B written per parameter and step (p, g, m, v -> p, m, v), grid sized
// in multiples of the SM count, 128-bit accesses, fixed-order two-level reductions (bit-reproducible run to run).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/odevio.h"

namespace odevio {
namespace {

constexpr int TG_THREADS = 256;
constexpr int TG_BLOCKS = 148 * 4;

__device__ __forceinline__ float block_sum(float v, float* sh) {      // fixed order: lanes, then warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < TG_THREADS / 32; ++w) t += sh[w];
  __syncthreads();
  return t;          // valid in thread 0
}

// partial[b] = (sum of squared angle errors, sum of squared translation errors) of block b's rows; optional gradient
__global__ void __launch_bounds__(TG_THREADS) pose_loss_kernel(long long n_rows, const float* __restrict__ pose,
                                                               const float* __restrict__ gts, float w_angle, float grad_scale,
                                                               float* __restrict__ grad, float2* __restrict__ partial) {
  __shared__ float sh[TG_THREADS / 32];
  float sa = 0.f, st = 0.f;
  const float ga = grad_scale * w_angle * 2.f / (3.f * static_cast<float>(n_rows));     // d/dpose of the mean over n*3 elements
  const float gt_ = grad_scale * 2.f / (3.f * static_cast<float>(n_rows));
  for (long long r = blockIdx.x * static_cast<long long>(TG_THREADS) + threadIdx.x; r < n_rows;
       r += static_cast<long long>(gridDim.x) * TG_THREADS) {
    const float2* p2 = reinterpret_cast<const float2*>(pose + r * 6);
    const float2* g2 = reinterpret_cast<const float2*>(gts + r * 6);
    const float2 pa = p2[0], pb = p2[1], pc = p2[2], qa = g2[0], qb = g2[1], qc = g2[2];
    const float d0 = pa.x - qa.x, d1 = pa.y - qa.y, d2 = pb.x - qb.x, d3 = pb.y - qb.y, d4 = pc.x - qc.x, d5 = pc.y - qc.y;
    sa += d0 * d0 + d1 * d1 + d2 * d2;
    st += d3 * d3 + d4 * d4 + d5 * d5;
    if (grad) {
      float2* o2 = reinterpret_cast<float2*>(grad + r * 6);
      o2[0] = make_float2(ga * d0, ga * d1); o2[1] = make_float2(ga * d2, gt_ * d3); o2[2] = make_float2(gt_ * d4, gt_ * d5);
    }
  }
  const float ta = block_sum(sa, sh), tt = block_sum(st, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_float2(ta, tt);
}
__global__ void pose_loss_final_kernel(int nblocks, const float2* __restrict__ partial, long long n_rows, float w_angle,
                                       float* __restrict__ loss3) {
  if (threadIdx.x || blockIdx.x) return;
  float a = 0.f, t = 0.f;
  for (int b = 0; b < nblocks; ++b) { a += partial[b].x; t += partial[b].y; }
  const float ma = a / (3.f * static_cast<float>(n_rows)), mt = t / (3.f * static_cast<float>(n_rows));
  loss3[0] = w_angle * ma + mt; loss3[1] = ma; loss3[2] = mt;
}

__global__ void __launch_bounds__(TG_THREADS) sumsq_kernel(long long n, const float* __restrict__ g, float* __restrict__ partial) {
  __shared__ float sh[TG_THREADS / 32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(TG_THREADS) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * TG_THREADS) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += TG_THREADS) s += g[i] * g[i];
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// norm_coef[0] = total gradient norm, [1] = clip coefficient min(1, max_norm / (norm + 1e-6))  (clip_grad_norm_)
__global__ void clip_coef_kernel(int nblocks, const float* __restrict__ partial, float max_norm, float* __restrict__ norm_coef) {
  if (threadIdx.x || blockIdx.x) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[b];
  const float norm = sqrtf(s);
  norm_coef[0] = norm;
  norm_coef[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
}

struct AdamArgs { float lr, lr_rest, beta1, beta2, eps, wd, bc1, bc2_sqrt; long long split; };   // elements >= split step with lr_rest
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float coef, const AdamArgs& a, float lr) {
  g = g * coef;                                   // clip_grad_norm_ scales the gradients in place
  g = fmaf(a.wd, p, g);                           // Adam's L2 weight decay: grad + wd * param
  m = m + (1.f - a.beta1) * (g - m);              // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf((1.f - a.beta2) * g, g, a.beta2 * v);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - (lr / a.bc1) * (m / denom);             // param.addcdiv_(exp_avg, denom, value = -step_size)
}
__global__ void __launch_bounds__(TG_THREADS) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const float* __restrict__ norm_coef, AdamArgs a) {
  const float coef = norm_coef ? norm_coef[1] : 1.f;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * static_cast<long long>(TG_THREADS) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * TG_THREADS) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    const float lr = (i << 2) < a.split ? a.lr : a.lr_rest;          // split is a multiple of 4 (checked by the host)
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, a, lr); adam_one(pp.y, gg.y, mm.y, vv.y, coef, a, lr);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, a, lr); adam_one(pp.w, gg.w, mm.w, vv.w, coef, a, lr);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += TG_THREADS) adam_one(p[i], g[i], m[i], v[i], coef, a, i < a.split ? a.lr : a.lr_rest);
}

}  // namespace
}  // namespace odevio

using namespace odevio;

extern "C" {

size_t odevio_train_glue_workspace_bytes(void) { return static_cast<size_t>(TG_BLOCKS) * sizeof(float2) + 256; }

int32_t odevio_pose_loss(int64_t n_rows, const float* pose, const float* gts, float w_angle, float grad_scale,
                         float* loss3, float* grad_pose, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!pose || !gts || !loss3 || !workspace) return ODEVIO_E_NULL;
  if (n_rows <= 0) return ODEVIO_E_SHAPE;
  if (workspace_bytes < odevio_train_glue_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 15)) return ODEVIO_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  long long nb = (n_rows + TG_THREADS - 1) / TG_THREADS;
  if (nb > TG_BLOCKS) nb = TG_BLOCKS;
  float2* partial = static_cast<float2*>(workspace);
  pose_loss_kernel<<<static_cast<unsigned>(nb), TG_THREADS, 0, stream>>>(n_rows, pose, gts, w_angle, grad_scale, grad_pose, partial);
  pose_loss_final_kernel<<<1, 32, 0, stream>>>(static_cast<int>(nb), partial, n_rows, w_angle, loss3);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}

int32_t odevio_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t step,
                         float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                         float* norm_coef, void* workspace, size_t workspace_bytes, void* stream_) {
  return odevio_adam_step_groups(n, n, params, grads, exp_avg, exp_avg_sq, step, lr, lr, beta1, beta2, eps, weight_decay, max_norm,
                                 norm_coef, workspace, workspace_bytes, stream_);
}

int32_t odevio_adam_step_groups(int64_t n, int64_t split, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                int32_t step, float lr_first, float lr_rest, float beta1, float beta2, float eps,
                                float weight_decay, float max_norm, float* norm_coef, void* workspace, size_t workspace_bytes,
                                void* stream_) {
  const float lr = lr_first;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !workspace) return ODEVIO_E_NULL;
  if (n <= 0 || step < 1 || split < 0 || split > n || (split & 3)) return ODEVIO_E_SHAPE;
  if (max_norm > 0.f && !norm_coef) return ODEVIO_E_NULL;
  if (workspace_bytes < odevio_train_glue_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 15)) return ODEVIO_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) return ODEVIO_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  long long nb = ((n >> 2) + TG_THREADS - 1) / TG_THREADS;
  if (nb > TG_BLOCKS) nb = TG_BLOCKS;
  if (nb < 1) nb = 1;
  if (norm_coef) {
    float* partial = static_cast<float*>(workspace);
    sumsq_kernel<<<static_cast<unsigned>(nb), TG_THREADS, 0, stream>>>(n, grads, partial);
    clip_coef_kernel<<<1, 32, 0, stream>>>(static_cast<int>(nb), partial, max_norm, norm_coef);
  }
  AdamArgs a;
  a.lr = lr; a.lr_rest = lr_rest; a.split = split; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));          // torch: Python doubles
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  adam_kernel<<<static_cast<unsigned>(nb), TG_THREADS, 0, stream>>>(n, params, grads, exp_avg, exp_avg_sq, norm_coef, a);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}

}  // extern "C"
