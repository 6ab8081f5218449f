// Development probe (not product code): measures on a B200
//   (1) the cost of back-to-back SS-mode tcgen05.mma instructions for kind::tf32 / kind::f16, M = 64 / 128,
//       N in {32..256}, operand layouts K-major no-swizzle (the layout ft_layer.cuh uses) and K-major SWIZZLE_128B;
//   (2) the per-SM ingest rate of 1-D bulk TMA copies out of L2 as a function of chunk size and CTA count;
//   (3) both together: a ring of operand stages filled by bulk TMA and drained by MMAs (the inner loop of the solver).
// Operand contents are arbitrary (zeros): only timing is observed.  All loops are tight (no integer division, unrolled
// MMA groups, descriptors advanced by adding to the address field): the first version of this probe measured its own
// scalar overhead (~200 clk per loop iteration of a single warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_probe tools/mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../odevio_b200/csrc/common.cuh"

using namespace odevio;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Prm {
  int M, N;
  int layout;    // 0 = K-major no swizzle (core matrices, LBO 128), 2 = K-major SWIZZLE_128B
  int nmma;      // MMAs per measurement (mode 1) / per chunk (mode 3; multiple of 3)
  int mode;      // 1 = MMA only, 2 = TMA ingest only, 3 = ring: TMA -> MMA
  int chunk_a, chunk_b;   // bytes per stage (A part, B part)
  int nstages, nchunks;
  int ncopies;            // bulk copies per stage (the stage's bytes split evenly)
  int shared_src;         // 1: all CTAs read the same global region (weights), 0: per-CTA regions
  const unsigned char* src; int src_chunks_per_cta;
  long long* out;         // [grid] clocks
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int layout, int rowgroup_bytes) {
  if (layout == 0) {
    // K-major, no swizzle: LBO = 128 B between k core matrices, SBO = bytes between 8-row groups
    return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (static_cast<uint64_t>(128u >> 4) << 16) |
           (static_cast<uint64_t>((uint32_t)rowgroup_bytes >> 4) << 32) | (1ull << 46);
  }
  // K-major SWIZZLE_128B: 8-row atoms of 1024 B, LBO unused (1)
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (1ull << 16) | (static_cast<uint64_t>(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) probe_kernel(const Prm p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[16];
  __shared__ __align__(8) uint64_t empty[16];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int stage_bytes = p.chunk_a + p.chunk_b;
  for (int i = tid; i < (p.nstages * stage_bytes) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 0u) << 7) | ((KIND == 0 ? 2u : 0u) << 10) |
                         (static_cast<uint32_t>(p.N >> 3) << 17) | (static_cast<uint32_t>(p.M >> 4) << 24);
  // bytes between 8-row groups in the no-swizzle layout: the stage holds chunk_a / 128 bytes of k per row
  const int rowgroup_a = (p.chunk_a / 128) * 8, rowgroup_b = (p.chunk_b / p.N) * 8;
  const uint32_t kstep = (p.layout == 0 ? 256u : 32u) >> 4;         // descriptor address units per MMA k-step (32 B of k)
  long long t0 = 0, t1 = 0;

  if (p.mode == 1) {
    if (warp == 1) {
      // two stages' descriptors, 4 k-steps each
      const uint64_t a0 = make_desc(smem_u32(smem), p.layout, rowgroup_a), b0 = make_desc(smem_u32(smem) + p.chunk_a, p.layout, rowgroup_b);
      const uint64_t a1 = make_desc(smem_u32(smem) + stage_bytes, p.layout, rowgroup_a);
      const uint64_t b1 = make_desc(smem_u32(smem) + stage_bytes + p.chunk_a, p.layout, rowgroup_b);
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int it = 0; it < p.nmma / 8; ++it) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mma<KIND>(tmem, a0 + u * kstep, b0 + u * kstep, idesc, 1);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma<KIND>(tmem, a1 + u * kstep, b1 + u * kstep, idesc, 1);
        }
        commit(&done);
      }
      __syncwarp();
      mbar_wait(&done, 0);
      t1 = clock64();
      if (lane == 0) p.out[blockIdx.x] = t1 - t0;
    }
  } else {
    const unsigned char* src0 = p.src + (p.shared_src ? 0 : static_cast<size_t>(blockIdx.x) * p.src_chunks_per_cta * stage_bytes);
    if (warp == 2) {
      // producer
      uint32_t st = 0, ph = 0;
      int sc = 0;
      const uint32_t cb = static_cast<uint32_t>(stage_bytes / p.ncopies);
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait(&empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[st], stage_bytes);
          const unsigned char* s = src0 + static_cast<size_t>(sc) * stage_bytes;
          unsigned char* d = smem + st * stage_bytes;
          for (int c = 0; c < p.ncopies; ++c) tma_load_1d(d + c * cb, s + c * cb, cb, &full[st]);
        }
        __syncwarp();
        if (++st == static_cast<uint32_t>(p.nstages)) { st = 0; ph ^= 1u; }
        if (++sc == p.src_chunks_per_cta) sc = 0;
      }
    } else if (warp == 1) {
      uint32_t st = 0, ph = 0;
      const int ksteps = p.nmma / 3;
      t0 = clock64();
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait(&full[st], ph);
        if (p.mode == 2) {
          if (lane == 0) mbar_arrive(&empty[st]);
          __syncwarp();
        } else {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem) + st * stage_bytes;
          // stage = [A hi | A lo | B hi | B lo]
          const uint64_t ah = make_desc(base, p.layout, rowgroup_a / 2), al = make_desc(base + p.chunk_a / 2, p.layout, rowgroup_a / 2);
          const uint64_t bh = make_desc(base + p.chunk_a, p.layout, rowgroup_b / 2), bl = make_desc(base + p.chunk_a + p.chunk_b / 2, p.layout, rowgroup_b / 2);
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t ko = k * kstep;
              mma<KIND>(tmem, ah + ko, bh + ko, idesc, 1);
              mma<KIND>(tmem + 256, al + ko, bh + ko, idesc, 1);
              mma<KIND>(tmem + 256, ah + ko, bl + ko, idesc, 1);
            }
            commit(&empty[st]);
          }
          __syncwarp();
        }
        if (++st == static_cast<uint32_t>(p.nstages)) { st = 0; ph ^= 1u; }
      }
      if (p.mode == 3) {
        if (elect_one()) commit(&done);
        __syncwarp();
        mbar_wait(&done, 0);
      }
      t1 = clock64();
      if (lane == 0) p.out[blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static long long* d_out;
static unsigned char* d_src;
static const size_t kSrcBytes = 256u << 20;
static double run(Prm p, int kind, int grid, const char* label) {
  p.out = d_out; p.src = d_src;
  const int stage_bytes = p.chunk_a + p.chunk_b;
  size_t smem = static_cast<size_t>(p.nstages) * stage_bytes + 40 * 1024;   // slack: swizzled descriptors on short stages read past the stage
  if (smem > 227 * 1024 - 2048) { printf("%s: smem too big\n", label); return 0; }
  // per-CTA source regions of ~512 KB (L2 resident for 148 CTAs)
  p.src_chunks_per_cta = (512 * 1024) / stage_bytes;
  if (p.src_chunks_per_cta < 1) p.src_chunks_per_cta = 1;
  auto kern = kind == 0 ? probe_kernel<0> : probe_kernel<1>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  for (int rep = 0; rep < 3; ++rep) {
    kern<<<grid, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
  }
  static long long h[1024];
  CK(cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mean = 0; long long mx = 0;
  for (int i = 0; i < grid; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
  mean /= grid;
  if (p.mode == 1)
    printf("%-52s grid %3d  clk/MMA mean %7.1f max %7.1f\n", label, grid, mean / p.nmma, (double)mx / p.nmma);
  else {
    const double bytes = static_cast<double>(p.nchunks) * stage_bytes;
    printf("%-52s grid %3d  clk/chunk mean %7.1f max %7.1f  B/clk/SM %6.1f  chip B/clk %8.0f\n", label, grid, mean / p.nchunks,
           (double)mx / p.nchunks, bytes / mean, bytes / mean * grid);
  }
  fflush(stdout);
  return mean;
}

int main() {
  CK(cudaMalloc(&d_out, sizeof(long long) * 1024));
  CK(cudaMalloc(&d_src, kSrcBytes));
  CK(cudaMemset(d_src, 0, kSrcBytes));
  char label[160];
  // (1) MMA cost
  for (int kind = 0; kind < 2; ++kind)
    for (int layout = 0; layout <= 2; layout += 2)
      for (int M = 64; M <= 128; M += 64) {
        const int Ns[] = {32, 64, 96, 128, 192, 256};
        for (int ni = 0; ni < 6; ++ni) {
          Prm p; memset(&p, 0, sizeof(p));
          p.M = M; p.N = Ns[ni]; p.layout = layout; p.nmma = 2048; p.mode = 1;
          p.chunk_a = 128 * 128; p.chunk_b = p.N * 128; p.nstages = 2; p.ncopies = 1;   // 128 B of k per row = 4 k-steps
          snprintf(label, sizeof(label), "mma %s M%d N%d %s", kind ? "f16 " : "tf32", M, p.N, layout ? "sw128" : "noswz");
          run(p, kind, 1, label);
          if (M == 128 && (p.N == 64 || p.N == 128 || p.N == 256)) run(p, kind, 148, label);
        }
      }
  // (2) ingest
  const int chunks[] = {4096, 8192, 16384, 32768, 49152};
  const int grids[] = {1, 32, 64, 128, 148};
  for (int sh = 0; sh < 2; ++sh)
    for (int ci = 0; ci < 5; ++ci)
      for (int nc = 1; nc <= 4; nc *= 4)
        for (int gi = 0; gi < 5; ++gi) {
          if (sh == 1 && gi < 3) continue;
          Prm p; memset(&p, 0, sizeof(p));
          p.mode = 2; p.chunk_a = chunks[ci]; p.chunk_b = 0; p.nstages = chunks[ci] > 16384 ? (chunks[ci] > 32768 ? 3 : 4) : 8; p.nchunks = 4096; p.shared_src = sh;
          p.ncopies = nc; p.N = 64; p.M = 128;
          snprintf(label, sizeof(label), "ingest %5d B x%d stages %d copies/stage %s", p.chunk_a, p.nstages, nc, sh ? "shared src" : "per-CTA src");
          run(p, 0, grids[gi], label);
        }
  // (3) ring: stage = ks k-steps (32 B of k per row each) of [A hi | A lo | B hi | B lo]
  for (int kind = 0; kind < 2; ++kind)
    for (int layout = 0; layout <= 2; layout += 2) {
      const int Ns[] = {64, 128, 256};
      for (int ni = 0; ni < 3; ++ni)
        for (int ks = 1; ks <= 4; ks *= 2)
          for (int gi = 3; gi < 5; ++gi) {
            Prm p; memset(&p, 0, sizeof(p));
            p.M = 128; p.N = Ns[ni]; p.layout = layout; p.mode = 3;
            p.chunk_a = 2 * 128 * ks * 32; p.chunk_b = 2 * p.N * ks * 32; p.nmma = 3 * ks; p.nchunks = 2048; p.ncopies = 2;
            p.nstages = (180 * 1024) / (p.chunk_a + p.chunk_b);
            if (p.nstages > 8) p.nstages = 8;
            if (p.nstages < 2) continue;
            snprintf(label, sizeof(label), "ring %s N%d %s (%d B/stage x%d, %d MMA)", kind ? "f16 " : "tf32", p.N, layout ? "sw128" : "noswz",
                     p.chunk_a + p.chunk_b, p.nstages, p.nmma);
            run(p, kind, grids[gi], label);
          }
    }
  printf("done\n");
  return 0;
}
