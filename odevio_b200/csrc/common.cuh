// Shared device helpers: mbarrier / bulk-TMA PTX wrappers, activations, rounding-exact
// scalar arithmetic.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace odevio {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or
// the hint (ns) elapses, instead of returning after the short default window.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if (++spins > (1u << 26)) __trap();
  }
}
// Producer-side wait: parked in hardware so the lane does not steal issue slots from the two
// consumer warps that share its scheduler (ncu: 828M polling TRYWAITs per launch without it).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one lane of a converged warp (operands stay warp-uniform -> uniform registers, no per-lane waterfall)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// ---------------------------------------------------------------- activations
enum { ACT_TANH = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SOFTPLUS = 3, ACT_NONE = 4, ACT_LEAKY01 = 5,
       ACT_SIGMOID = 6 };

// Deliberately NOT inlined: epilogues call it once per output element (a few dozen calls per
// GEMM per thread), and inlining the tanh/exp/log1p bodies at every call site multiplied the
// kernel's code size ~10x for no measurable gain.
static __device__ __noinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_TANH: return tanhf(v);
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LEAKY: return v > 0.f ? v : 0.01f * v;          // nn.LeakyReLU() default slope
    case ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v)); // nn.Softplus(beta=1, threshold=20)
    case ACT_LEAKY01: return v > 0.f ? v : 0.1f * v;          // regressor LeakyReLU(0.1)
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// 4-wide variant used by the GEMM epilogues: one ABI call per float4, the four transcendental
// chains inside interleave (ncu: the scalar calls were ~11 % of the fused kernel's warp time).
static __device__ __noinline__ float4 apply_act4(float4 v, int act) {
  switch (act) {
    case ACT_TANH: return make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    case ACT_RELU: return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    case ACT_NONE: return v;
    default:
      return make_float4(apply_act(v.x, act), apply_act(v.y, act), apply_act(v.z, act), apply_act(v.w, act));
  }
}

// Two-rounding arithmetic that nvcc may not contract into FMA (mirrors the oracle's op order).
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace odevio
