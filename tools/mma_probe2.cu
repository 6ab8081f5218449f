// Development probe (not product code): why do the MMAs of odernn_h3.cu run at ~109 clk instead of the 48 clk
// tools/mma_probe.cu measured for M = 128, N = 64?  Replays h3_issue_chunk<1, 64> (A K-major no-swizzle fp16, B MN-major
// no-swizzle fp16, 3 MMAs per k-step) back to back from one elected lane and varies one thing at a time:
//   bmajor  0 = B K-major (the layout mma_probe.cu timed), 1 = B MN-major (odernn_h3.cu)
//   tma     0 = no concurrent copies, 1 = a second warp streams 48 KB bulk copies into the other ring stages meanwhile
//   N       32 / 64 / 128
//   nctas   1 or 128
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_probe2 tools/mma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../odevio_b200/csrc/common.cuh"
using namespace odevio;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo) {
  return static_cast<uint64_t>(((saddr >> 4) & 0x3fffu) | ((128u >> 4) << 16)) | (static_cast<uint64_t>((sbo >> 4) | (1u << 14)) << 32);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
struct Prm { int N, bmajor, tma, iters, pattern, poll, style; const unsigned char* src; long long* out; };

constexpr int WCH = 32768, XCH = 16384, NST = 4;

__global__ void __launch_bounds__(352, 1) k(const Prm p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done, tbar[NST], never, cbar[2 * NST];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 2 * NST; ++i) mbar_init(&cbar[i], 1); mbar_init(&never, 1); mbar_init(&done, 1); for (int i = 0; i < NST; ++i) mbar_init(&tbar[i], 1); stop = 0; fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < NST * (WCH + XCH) / 16; i += 352) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t N = p.N;
  const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(p.bmajor) << 16) | ((N >> 3) << 17) | (8u << 24);
  const uint32_t sbo = 8 * 128;          // KCH = 64: 8 k-groups of 128 B per 8-row group
  if (warp == 1) {
    __syncwarp();
    const long long t0 = clock64();
    if (p.style < 2) {
      if (elect_one()) {
        for (int it = 0; it < p.iters; ++it) {
          const uint32_t st = it & (NST - 1);
          const uint32_t wb = smem_u32(smem) + st * WCH, xb = smem_u32(smem) + NST * WCH + st * XCH;
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 256u;
            const uint64_t ah = desc(wb + o, sbo), al = desc(wb + 16384 + o, sbo);
            const uint64_t xh = desc(xb + o, sbo), xl = desc(xb + 8192 + o, sbo);
            mma(tmem, ah, xh, idesc, 1); mma(tmem + 256, al, xh, idesc, 1); mma(tmem + 256, ah, xl, idesc, 1);
          }
          if (p.style == 1) { commit(&cbar[st]); commit(&cbar[NST + st]); }
        }
        commit(&done);
      }
    } else {
      // the loop shape of odernn_h3.cu: per chunk a converged warp, fence, elect, 12 MMAs, 2 commits, syncwarp
      uint32_t g = 0;
      for (int it = 0; it < p.iters; ++it, ++g) {
        const uint32_t st = g & (NST - 1);
        if (p.style == 3 && g >= NST) mbar_wait(&cbar[st], ((g / NST) - 1) & 1u);       // an already complete barrier
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wb = smem_u32(smem) + st * WCH, xb = smem_u32(smem) + NST * WCH + st * XCH;
        if (elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 256u;
            const uint64_t ah = desc(wb + o, sbo), al = desc(wb + 16384 + o, sbo);
            const uint64_t xh = desc(xb + o, sbo), xl = desc(xb + 8192 + o, sbo);
            mma(tmem, ah, xh, idesc, 1); mma(tmem + 256, al, xh, idesc, 1); mma(tmem + 256, ah, xl, idesc, 1);
          }
          commit(&cbar[st]); commit(&cbar[NST + st]);
        }
        __syncwarp();
      }
      if (elect_one()) commit(&done);
    }
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    stop = 1;
    if (lane == 0) p.out[blockIdx.x] = t1 - t0;
  } else if (warp == 2 && p.tma) {
    // concurrent 48 KB copies into the ring (timing only: races with the MMAs' reads are harmless here)
    uint32_t ph[NST] = {0, 0, 0, 0};
    int n = 0;
    while (!stop) {
      const int st = n & (NST - 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&tbar[st], WCH + XCH);
        tma_load_1d(smem + st * WCH, p.src + static_cast<size_t>(n % 12) * WCH, WCH, &tbar[st]);
        tma_load_1d(smem + NST * WCH + st * XCH, p.src + (12 + blockIdx.x * 12 + n % 12) * static_cast<size_t>(XCH), XCH, &tbar[st]);
      }
      __syncwarp();
      mbar_wait(&tbar[st], ph[st]); ph[st] ^= 1u;
      ++n;
    }
    if (lane == 0) p.out[gridDim.x + blockIdx.x] = n;
  } else if (warp >= 3 && p.poll) {
    // the epilogue warps of odernn_h3.cu wait for the accumulators during the whole MMA phase
    if (p.poll == 1) { while (!stop) { if (mbar_try_wait(&never, 0)) break; } }
    else { while (!stop) { if (mbar_try_wait_hint(&never, 0, 2000u)) break; } }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  unsigned char* src; long long* out;
  const size_t src_bytes = 12 * (size_t)WCH + (12 + 148 * 12 + 12) * (size_t)XCH;
  CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 0, src_bytes));
  CK(cudaMalloc(&out, 2 * 148 * sizeof(long long)));
  const int smem = NST * (WCH + XCH) + 1024;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 200;
  printf("nctas N bmajor tma pattern : clk per MMA-equivalent k-step-third (chunk = 12 MMAs), copies done\n");
  for (int nctas : {128})
    for (int N : {32, 64})
      for (int bmajor : {1})
        for (int tma : {0, 1})
          for (int poll : {0})
          for (int style : {0, 1, 2, 3})
          for (int pattern : {0}) {
            Prm p{N, bmajor, tma, iters, pattern, poll, style, src, out};
            CK(cudaMemset(out, 0, 2 * 148 * sizeof(long long)));
            k<<<nctas, 352, smem>>>(p);
            CK(cudaDeviceSynchronize());
            long long h[2 * 148];
            CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
            long long mx = 0, copies = 0;
            for (int i = 0; i < nctas; ++i) { if (h[i] > mx) mx = h[i]; copies += h[nctas + i]; }
            printf("nctas %4d N %3d tma %d style %d : %.1f clk per MMA, %.1f per k-step, copies/cta %.1f (%.0f B/clk/SM ingest)\n", nctas, N, tma, style,
                   double(mx) / (iters * 12), double(mx) / (iters * 4), double(copies) / nctas, double(copies) / nctas * (WCH + XCH) / double(mx));
          }
  return 0;
}
