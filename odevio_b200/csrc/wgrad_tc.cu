// Weight gradients of the ODEFunc Linears on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   dW[n][k] = sum_m G[m][n] * A[m][k]      m = every (row, vector-field evaluation) of the backward
//
// This is the one genuinely dense contraction of the training path (reduction length M ~ 1e5..1e6,
// output 512 x 768): it runs as `tcgen05.mma.kind::tf32` with fp32 accumulators in TMEM.  fp32
// parity is kept with the 3xTF32 split: the backward kernel writes every operand as a TF32-exact
// high part and the exact residual (tile_gemm.cuh: tf32_hi), and each k-step issues
//   D += G_hi A_hi + G_lo A_hi + G_hi A_lo        (the dropped lo*lo term is ~2^-22 relative).
// Operands are stored by the producer kernel directly in the tensor core's canonical K-major
// no-swizzle shared-memory image (core matrices of 8 features x 4 rows = 128 contiguous bytes,
// [feature/8][R/4][8][4] per block of R rows), so one 1-D bulk TMA copy per operand tile brings a
// ready-to-use stage: no tensor maps, no swizzle, no register staging.
//
// CTA = one 128 x BN output tile over a slab of blocks (split-M); 128 threads:
//   warp 0 lane 0  TMA producer (cp.async.bulk -> mbarrier ring)
//   warp 1 lane 0  MMA issuer (tcgen05.mma, tcgen05.commit frees ring slots / signals the epilogue)
//   warps 0-3      epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> fp32 partial tile in HBM
// Split-M partials are summed in a fixed order by wgrad.cu's reduce kernel (deterministic).
// Replaces autograd's weight-gradient GEMMs of scripts/train_model.py:78 for ode_func.net.*.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace odevio {

namespace {

constexpr int TC_BM = 128;            // output rows (G features) per CTA = MMA M
constexpr int TC_MAX_STAGES = 4;

__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), SWIZZLE_NONE
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// grid = (k_tiles, n_tiles, splits), block = 128, dynamic smem = nst * stage_bytes + 1024
__global__ void __launch_bounds__(128, 1)
wgrad_tc_kernel(const float* __restrict__ Ghi, const float* __restrict__ Glo,
                const float* __restrict__ Ahi, const float* __restrict__ Alo,
                int N, int K, int R, int BN, long long nblocks, long long blocks_per_split,
                int nst, float* __restrict__ part) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * TC_BM, k0 = blockIdx.x * BN;
  const long long b0 = static_cast<long long>(blockIdx.z) * blocks_per_split;
  long long b1 = b0 + blocks_per_split;
  if (b1 > nblocks) b1 = nblocks;
  const int nb = b1 > b0 ? static_cast<int>(b1 - b0) : 0;

  const uint32_t g_tile_bytes = static_cast<uint32_t>(TC_BM) * R * 4;      // 128 features x R rows
  const uint32_t a_tile_bytes = static_cast<uint32_t>(BN) * R * 4;
  const uint32_t stage_bytes = 2 * (g_tile_bytes + a_tile_bytes);
  // two accumulators of BN columns: hi*hi products | cross terms (see the note at the MMA issue loop)
  const uint32_t tmem_cols = 2 * BN <= 64 ? 64u : 2 * BN <= 128 ? 128u : 2 * BN <= 256 ? 256u : 512u;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {            // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: 4 bulk copies per block (G hi/lo tile, A hi/lo tile)
    const size_t g_blk = static_cast<size_t>(N) * R, a_blk = static_cast<size_t>(K) * R;    // floats per block
    const size_t g_off = static_cast<size_t>(n0) * R, a_off = static_cast<size_t>(k0) * R;
    uint32_t stage = 0, phase = 0;
    for (int b = 0; b < nb; ++b) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      unsigned char* dst = smem + static_cast<size_t>(stage) * stage_bytes;
      mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
      const size_t blk = static_cast<size_t>(b0 + b);
      tma_load_1d(dst, Ghi + blk * g_blk + g_off, g_tile_bytes, &full_bar[stage]);
      tma_load_1d(dst + g_tile_bytes, Glo + blk * g_blk + g_off, g_tile_bytes, &full_bar[stage]);
      tma_load_1d(dst + 2 * g_tile_bytes, Ahi + blk * a_blk + a_off, a_tile_bytes, &full_bar[stage]);
      tma_load_1d(dst + 2 * g_tile_bytes + a_tile_bytes, Alo + blk * a_blk + a_off, a_tile_bytes, &full_bar[stage]);
      if (++stage == static_cast<uint32_t>(nst)) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform operands, one elected lane issues
    // (a single active lane makes ptxas move every descriptor through an ELECT / R2UR waterfall)
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, K-major both, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                           (static_cast<uint32_t>(TC_BM >> 4) << 24);
    const uint32_t lbo = 128, sbo = static_cast<uint32_t>(R >> 2) * 128;
    // The tensor core accumulates with truncation (measured in odefunc_tc.cu: ~n * 2^-24 drift after n
    // accumulations into one accumulator), so (a) the hi*hi products and the 2^-11 smaller cross terms
    // go to separate accumulators, added in fp32 in the epilogue, and (b) a CTA only reduces
    // kBlocksPerSplit blocks -- longer reductions are split-M partials summed by the reduce kernel.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
    const uint32_t d_main = tmem_u, d_corr = tmem_u + static_cast<uint32_t>(BN);
    uint32_t stage = 0, phase = 0, acc = 0;
    for (int b = 0; b < nb; ++b) {
      mbar_wait(&full_bar[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
      const uint32_t g_hi = base, g_lo = base + g_tile_bytes;
      const uint32_t a_hi = base + 2 * g_tile_bytes, a_lo = a_hi + a_tile_bytes;
      if (elect_one()) {
        for (int ks = 0; ks < (R >> 3); ++ks) {        // 8 rows (2 core matrices along K) per tf32 MMA
          const uint32_t o = static_cast<uint32_t>(ks) * 256;
          const uint64_t dgh = smem_desc_kmajor(g_hi + o, lbo, sbo), dgl = smem_desc_kmajor(g_lo + o, lbo, sbo);
          const uint64_t dah = smem_desc_kmajor(a_hi + o, lbo, sbo), dal = smem_desc_kmajor(a_lo + o, lbo, sbo);
          const uint32_t a0 = (acc | static_cast<uint32_t>(ks)) ? 1u : 0u;
          umma_tf32(d_main, dgh, dah, idesc, a0);
          umma_tf32(d_corr, dgl, dah, idesc, a0);
          umma_tf32(d_corr, dgh, dal, idesc, 1);
        }
        umma_commit(&empty_bar[stage]);                // frees the slot when these MMAs have read it
      }
      acc = 1;
      __syncwarp();
      if (++stage == static_cast<uint32_t>(nst)) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(&accum_bar);          // accumulator complete
    __syncwarp();
  }
  __syncwarp();

  // ===== epilogue: TMEM -> registers -> partial tile
  float* out = part + static_cast<size_t>(blockIdx.z) * N * K;
  const int n = n0 + warp * 32 + lane;                 // warp w owns TMEM lanes 32w .. 32w + 31
  if (nb > 0) {
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      float sum[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) sum[q] = 0.f;
      for (int part_i = 1; part_i >= 0; --part_i) {          // cross terms first, then the main accumulator
      const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(part_i * BN + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; ++q) sum[q] += __uint_as_float(v[q]);
      }
      if (n < N) {
        float* row = out + static_cast<size_t>(n) * K + k0 + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(row + 4 * q) = make_float4(sum[4 * q], sum[4 * q + 1], sum[4 * q + 2], sum[4 * q + 3]);
      }
    }
  } else if (n < N) {
    for (int c = 0; c < BN; ++c) out[static_cast<size_t>(n) * K + k0 + c] = 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// db[n] partials over block-format G (hi + lo): grid = (ceil(N/128), splits), block = 128
__global__ void colsum_blocks_kernel(const float* __restrict__ Ghi, const float* __restrict__ Glo, int N, int R,
                                     long long nblocks, long long blocks_per_split, float* __restrict__ part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const long long b0 = static_cast<long long>(blockIdx.y) * blocks_per_split;
  long long b1 = b0 + blocks_per_split;
  if (b1 > nblocks) b1 = nblocks;
  if (n >= N) return;
  const int nj = R >> 2;
  float s = 0.f;
  for (long long b = b0; b < b1; ++b) {
    const size_t base = static_cast<size_t>(b) * N * R + (static_cast<size_t>(n >> 3) * nj * 8 + (n & 7)) * 4;
    for (int j = 0; j < nj; ++j) {
      const float4 h = *reinterpret_cast<const float4*>(Ghi + base + static_cast<size_t>(j) * 32);
      const float4 l = *reinterpret_cast<const float4*>(Glo + base + static_cast<size_t>(j) * 32);
      s += ((h.x + l.x) + (h.y + l.y)) + ((h.z + l.z) + (h.w + l.w));
    }
  }
  part[static_cast<size_t>(blockIdx.y) * N + n] = s;
}

}  // namespace

__global__ void wgrad_reduce_kernel_tc(const float* __restrict__ part, int splits, size_t total, float* __restrict__ out,
                                       int accumulate = 0) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[static_cast<size_t>(z) * total + i];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// Largest MMA N (output columns per CTA) in {256, 128, 64, 32} dividing K; 0 if none.
int wgrad_tc_bn(int K) {
  for (int bn = 256; bn >= 32; bn >>= 1)
    if (K % bn == 0) return bn;
  return 0;
}

// Blocks one CTA reduces: bounds the accumulation chain (2 * R/8 MMAs per block into the main
// accumulator) so that the tensor core's truncating accumulate stays below ~5e-6 relative.
constexpr long long kBlocksPerSplit = 48;

int wgrad_tc_splits(long long nblocks, int N, int K, int nsm) {
  (void)N; (void)nsm;
  if (!wgrad_tc_bn(K)) return 0;
  long long s = (nblocks + kBlocksPerSplit - 1) / kBlocksPerSplit;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

// dW [N][K] and db [N] of one ODEFunc Linear from its block-format streams (R rows per block).
// `part` must hold max(splits * N * K, 256 * N) floats.
cudaError_t wgrad_linear_tc_ex(const float* Ghi, const float* Glo, const float* Ahi, const float* Alo,
                               long long nblocks, int N, int K, int R, float* dW, float* db, float* part, int nsm,
                               int accumulate, cudaStream_t stream);

cudaError_t wgrad_linear_tc(const float* Ghi, const float* Glo, const float* Ahi, const float* Alo,
                            long long nblocks, int N, int K, int R, float* dW, float* db, float* part, int nsm,
                            cudaStream_t stream) {
  return wgrad_linear_tc_ex(Ghi, Glo, Ahi, Alo, nblocks, N, K, R, dW, db, part, nsm, 0, stream);
}

// accumulate != 0: dW += ..., db += ...  (the record streams of a training step reduced interval range by interval range)
cudaError_t wgrad_linear_tc_ex(const float* Ghi, const float* Glo, const float* Ahi, const float* Alo,
                               long long nblocks, int N, int K, int R, float* dW, float* db, float* part, int nsm,
                               int accumulate, cudaStream_t stream) {
  const int bn = wgrad_tc_bn(K);
  if (!bn || N % TC_BM || R % 8 || R > 32) return cudaErrorInvalidValue;
  if (nblocks <= 0 && accumulate) return cudaSuccess;
  if (nblocks <= 0) {
    cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * static_cast<size_t>(N) * K, stream);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(db, 0, sizeof(float) * N, stream);
  }
  const int splits = wgrad_tc_splits(nblocks, N, K, nsm);
  const long long bps = kBlocksPerSplit;
  const size_t stage_bytes = static_cast<size_t>(2) * (TC_BM + bn) * R * 4;
  int nst = static_cast<int>((200 * 1024) / stage_bytes);
  if (nst > TC_MAX_STAGES) nst = TC_MAX_STAGES;
  if (nst < 2) return cudaErrorInvalidValue;
  const size_t smem_bytes = nst * stage_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e != cudaSuccess) return e;
  dim3 grid(K / bn, N / TC_BM, splits);
  wgrad_tc_kernel<<<grid, 128, smem_bytes, stream>>>(Ghi, Glo, Ahi, Alo, N, K, R, bn, nblocks, bps, nst, part);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t total = static_cast<size_t>(N) * K;
  int rb = static_cast<int>((total + 255) / 256);
  if (rb > 4 * nsm) rb = 4 * nsm;
  wgrad_reduce_kernel_tc<<<rb, 256, 0, stream>>>(part, splits, total, dW, accumulate);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  int bs = static_cast<int>((nblocks + 255) / 256);
  if (bs > 256) bs = 256;
  if (bs < 1) bs = 1;
  const long long bbps = (nblocks + bs - 1) / bs;
  dim3 g2((N + 127) / 128, bs);
  colsum_blocks_kernel<<<g2, 128, 0, stream>>>(Ghi, Glo, N, R, nblocks, bbps, part);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  wgrad_reduce_kernel_tc<<<(N + 255) / 256, 256, 0, stream>>>(part, bs, static_cast<size_t>(N), db, accumulate);
  return cudaGetLastError();
}

}  // namespace odevio
