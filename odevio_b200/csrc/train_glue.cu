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
#include <string.h>

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

// ------------------------------------------------------------------------------------------------------------------
// Gradient all-reduce + global-norm clip + Adam in ONE kernel over NVLink peer memory (reference scripts/train_model.py:78-86
// with src/utils/utils.py:143-157 under data parallelism: all-reduce of the Pose_net gradients, clip_grad_norm_(5), Adam).
// Every rank's gradient bucket, parameter bucket and a small flag pad live in symmetric memory (mapped into every process);
// rank r owns the r-th slice of the bucket:
//   A  cross-GPU barrier: every rank's gradients are final;
//   1  reduce-scatter by pull: g[i] = grad_scale * sum_p grads_p[i] for i in my slice (P2P loads over NVLink, fixed order over
//      p: bit-reproducible), sum of squares of my slice -> every peer's pad (P2P store) + flag B;
//   2  wait for all partial sums -> total norm, clip coefficient (the same value on every rank: same numbers, same order);
//   3  Adam on my slice (its moments exist only here: the optimiser state is sharded), all-gather by push: the new parameters
//      of my slice are stored into every rank's parameter bucket;
//   C  cross-GPU barrier: every slice has landed everywhere.
// One launch instead of NCCL all-reduce + 3 launches; 1/world of the moment traffic; the bucket crosses NVLink once as
// gradients (pull) and once as parameters (push).  Flags are monotonic call counters (no reset); spins are bounded by a trap.
namespace odevio {
namespace {

constexpr int PEER_MAX = 16;
struct PeerArgs {
  float* params[PEER_MAX]; const float* grads[PEER_MAX]; uint32_t* pad[PEER_MAX];
  int rank, world; uint32_t epoch; float grad_scale, max_norm;
  long long n, lo, hi;
  float* m; float* v; float* gred; float* partial; unsigned int* gridbar; float* norm_coef;
  AdamArgs a;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// flag `which` (0 = A, 1 = B, 2 = C) of this rank := epoch on every peer's pad, then wait for every peer's flag on my pad
__device__ __forceinline__ void peer_signal(const PeerArgs& a, int which) {
  __threadfence_system();
  for (int p = 0; p < a.world; ++p) st_release_sys(a.pad[p] + which * PEER_MAX + a.rank, a.epoch);
}
__device__ __forceinline__ void peer_wait(const PeerArgs& a, int which) {
  for (int p = 0; p < a.world; ++p) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(a.pad[a.rank] + which * PEER_MAX + p) < a.epoch)
      if (++spins > (1ull << 31)) __trap();             // a rank that never arrives must not hang the GPU
  }
}
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    target += gridDim.x;
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 28)) __trap();
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TG_THREADS) allreduce_adam_peer_kernel(const PeerArgs a) {
  __shared__ float sh[TG_THREADS / 32];
  unsigned int gtarget = 0;
  // ---- A: all gradients final (kernel-boundary writes of the producers are visible to system scope)
  if (blockIdx.x == 0 && threadIdx.x == 0) { peer_signal(a, 0); peer_wait(a, 0); }
  grid_barrier(a.gridbar, gtarget);
  // ---- 1: reduce my slice (pull), sum of squares
  const long long lo4 = a.lo >> 2, hi4 = a.hi >> 2;
  float s = 0.f;
  // (four float4 per thread and trip = 4 x world P2P loads in flight measured the same 126-135 us at world 8: the step is
  // bound by its three cross-GPU and four grid barriers, not by the NVLink round trips)
  for (long long i = lo4 + blockIdx.x * static_cast<long long>(TG_THREADS) + threadIdx.x; i < hi4;
       i += static_cast<long long>(gridDim.x) * TG_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < a.world; ++p) {
      const float4 g = __ldcg(reinterpret_cast<const float4*>(a.grads[p]) + i);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    acc.x *= a.grad_scale; acc.y *= a.grad_scale; acc.z *= a.grad_scale; acc.w *= a.grad_scale;
    reinterpret_cast<float4*>(a.gred)[i - lo4] = acc;
    s += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) a.partial[blockIdx.x] = t;
  grid_barrier(a.gridbar, gtarget);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float tot = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) tot += a.partial[b];
    for (int p = 0; p < a.world; ++p) reinterpret_cast<float*>(a.pad[p] + 3 * PEER_MAX)[a.rank] = tot;
    peer_signal(a, 1);
    // ---- 2: total norm of the averaged gradient, clip coefficient (clip_grad_norm_)
    peer_wait(a, 1);
    float ss = 0.f;
    for (int p = 0; p < a.world; ++p) ss += reinterpret_cast<volatile float*>(a.pad[a.rank] + 3 * PEER_MAX)[p];
    const float norm = sqrtf(ss);
    a.norm_coef[0] = norm;
    a.norm_coef[1] = a.max_norm > 0.f ? fminf(1.f, a.max_norm / (norm + 1e-6f)) : 1.f;
  }
  grid_barrier(a.gridbar, gtarget);
  // ---- 3: Adam on my slice, new parameters pushed to every rank
  const float coef = __ldcg(a.norm_coef + 1);
  for (long long i = lo4 + blockIdx.x * static_cast<long long>(TG_THREADS) + threadIdx.x; i < hi4;
       i += static_cast<long long>(gridDim.x) * TG_THREADS) {
    float4 pp = reinterpret_cast<const float4*>(a.params[a.rank])[i];
    float4 mm = reinterpret_cast<float4*>(a.m)[i], vv = reinterpret_cast<float4*>(a.v)[i];
    const float4 gg = reinterpret_cast<const float4*>(a.gred)[i - lo4];
    const float lr = (i << 2) < a.a.split ? a.a.lr : a.a.lr_rest;
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, a.a, lr); adam_one(pp.y, gg.y, mm.y, vv.y, coef, a.a, lr);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, a.a, lr); adam_one(pp.w, gg.w, mm.w, vv.w, coef, a.a, lr);
    reinterpret_cast<float4*>(a.m)[i] = mm; reinterpret_cast<float4*>(a.v)[i] = vv;
    for (int p = 0; p < a.world; ++p) reinterpret_cast<float4*>(a.params[p])[i] = pp;
  }
  // ---- C: every slice has landed everywhere
  __threadfence_system();
  grid_barrier(a.gridbar, gtarget);
  if (blockIdx.x == 0 && threadIdx.x == 0) { peer_signal(a, 2); peer_wait(a, 2); }
}

}  // namespace
}  // namespace odevio

extern "C" {

size_t odevio_allreduce_adam_peer_workspace_bytes(int64_t n, int32_t world) {
  if (n <= 0 || world < 1 || world > PEER_MAX) return 0;
  const long long per = ((n / 4 + world - 1) / world) * 4;
  return static_cast<size_t>(per) * sizeof(float) + static_cast<size_t>(TG_BLOCKS) * sizeof(float) + 512;
}

int32_t odevio_allreduce_adam_peer(int64_t n, int64_t split, int32_t rank, int32_t world,
                                   float* const* params_peers, const float* const* grads_peers, uint32_t* const* pad_peers,
                                   float* exp_avg, float* exp_avg_sq, int32_t step, uint32_t epoch, float grad_scale,
                                   float lr_first, float lr_rest, float beta1, float beta2, float eps, float weight_decay,
                                   float max_norm, float* norm_coef, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!params_peers || !grads_peers || !pad_peers || !exp_avg || !exp_avg_sq || !norm_coef || !workspace) return ODEVIO_E_NULL;
  if (n <= 0 || (n & 3) || step < 1 || epoch < 1 || split < 0 || split > n || (split & 3)) return ODEVIO_E_SHAPE;
  if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world) return ODEVIO_E_SHAPE;
  if (workspace_bytes < odevio_allreduce_adam_peer_workspace_bytes(n, world) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return ODEVIO_E_WORKSPACE;
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!params_peers[p] || !grads_peers[p] || !pad_peers[p]) return ODEVIO_E_NULL;
    a.params[p] = params_peers[p]; a.grads[p] = grads_peers[p]; a.pad[p] = pad_peers[p];
  }
  const long long per = ((n / 4 + world - 1) / world) * 4;
  a.rank = rank; a.world = world; a.epoch = epoch; a.grad_scale = grad_scale; a.max_norm = max_norm; a.n = n;
  a.lo = per * rank < n ? per * rank : n;
  a.hi = per * (rank + 1) < n ? per * (rank + 1) : n;
  a.m = exp_avg; a.v = exp_avg_sq; a.norm_coef = norm_coef;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  a.gred = reinterpret_cast<float*>(ws);
  a.partial = reinterpret_cast<float*>(ws + static_cast<size_t>(per) * sizeof(float));
  a.gridbar = reinterpret_cast<unsigned int*>(ws + static_cast<size_t>(per) * sizeof(float) + static_cast<size_t>(TG_BLOCKS) * sizeof(float) + 256);
  a.a.lr = lr_first; a.a.lr_rest = lr_rest; a.a.split = split; a.a.beta1 = beta1; a.a.beta2 = beta2; a.a.eps = eps; a.a.wd = weight_decay;
  a.a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  a.a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(a.gridbar, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return static_cast<int32_t>(e);
  // cooperative launch: the in-kernel grid barriers need every CTA resident; the slice is small (1/world of the bucket)
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  long long nb = ((a.hi - a.lo) / 4 + TG_THREADS - 1) / TG_THREADS;
  if (nb > nsm) nb = nsm;
  if (nb < 1) nb = 1;
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(allreduce_adam_peer_kernel), dim3(static_cast<unsigned>(nb)),
                                  dim3(TG_THREADS), args, 0, stream);
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}

}  // extern "C"
