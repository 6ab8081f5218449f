// Tensor-core layer routine shared by the tcgen05 kernels of the path (odefunc_tc.cu: ODEFunc.forward;
// odernn_tc.cu: the tensor-core ODE solver): one Linear (+ bias + activation) of a 128-row tile computed by the
// CTAs of a cluster as an nN x nK grid, 3xTF32 (fp32-accurate), see odefunc_tc.cu for the design notes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "common.cuh"

namespace odevio {

namespace {

// development timeline (clock64 of cluster 0 / CTA 0, first tile): [layer][event]; one copy per translation unit
__device__ long long g_ft_dbg[64];

#ifdef ODEVIO_FT_TIMELINE
#define FT_STAMP(idx) do { if (blockIdx.x == 0 && c.tile == 0) g_ft_dbg[idx] = clock64(); } while (0)
#else
#define FT_STAMP(idx) do { } while (0)
#endif

constexpr int FT_ROWS = 128;          // rows per tile = MMA M
// cluster size NC (column slices per 128-row tile) and k per ring stage KCH are template parameters:
//   <8, 8>   few tiles (B ~ 1024): 8 CTAs per tile keep 128 SMs busy; per layer they form an nN x nK grid
//            (2 x 4 / 4 x 2 at the default shapes) so that every MMA is 192..256 columns wide, the k-split
//            partials are reduced through L2
//   <4, 16>  many tiles: N = 128 / 192 / 256 per MMA amortises the ~125 clk A-operand read of every MMA
// FT_SPLIT: true = one fp32 operand copy crosses L2, hi/lo produced in shared memory by the splitter warps;
//           false = the epilogues write hi and lo images (twice the bytes, no split stage: wide slices do not
//           leave shared memory for a separate operand ring)
constexpr int FT_RAW_STAGES = 8;      // fp32 chunks in flight (bulk TMA -> splitter); 4 when shared memory is short
constexpr int FT_OP_STAGES = 8;       // operand stages (activations: splitter -> tensor core; weights: TMA -> tensor core), a
                                      // multiple of the 4 splitter warps: chunk g uses raw stage g % nraw and operand stage
                                      // g % 8, both always served by splitter warp g % 4, so every parity wait is at most one
                                      // phase behind.  8 (was 4): the producer waits for the operand stage, so its run-ahead is
                                      // the operand ring depth and must cover the ~1.4 k clk TMA landing latency
constexpr int FT_MAX_LAYERS = ODEVIO_MAX_ODE_LINEARS;
// warps 0-7 epilogue (thread = row, two column halves), warps 8-11 hi/lo splitter, warp 12 TMA producer,
// warp 13 MMA issuer
constexpr int FT_EPI_WARPS = 8, FT_SPLIT_WARPS = 4;
constexpr int FT_WARP_SPLIT = FT_EPI_WARPS, FT_WARP_TMA = FT_EPI_WARPS + FT_SPLIT_WARPS, FT_WARP_MMA = FT_WARP_TMA + 1;
constexpr int FT_THREADS = 32 * (FT_WARP_MMA + 1);


template <int KCH>
__device__ __forceinline__ uint64_t ft_desc(uint32_t saddr) {
  // K-major, no swizzle: LBO (k core matrices) = 128 B, SBO (8-row groups) = (KCH / 4) * 128 B; version 1
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (static_cast<uint64_t>(128u >> 4) << 16) |
         (static_cast<uint64_t>((KCH / 4 * 128u) >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void ft_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ft_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// accumulators for the hi*hi products (K segments); one more holds the cross terms: (nseg + 1) * Nc <= 512
__device__ __forceinline__ int ft_nseg(int Nc) { const int n = 512 / Nc - 1; return n > 4 ? 4 : (n < 1 ? 1 : n); }
__device__ __forceinline__ float ft_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// float offset of (row r, feature k) in a 128-row operand buffer: [k/KCH][r/8][(k%KCH)/4][r%8][k%4]
template <int KCH>
__device__ __forceinline__ size_t xa_offset(int r, int k) {
  return ((static_cast<size_t>(k / KCH) * 16 + (r >> 3)) * (KCH / 4) + ((k % KCH) >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

// thread = row: 32 consecutive features (one k-chunk) -> fp32 operand image (the hi / lo parts are
// produced in shared memory by the splitter warps: one copy crosses L2 and the SM boundary, not two)
template <int KCH>
__device__ __forceinline__ void store_chunk(float* dst, int r, int k0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(dst + xa_offset<KCH>(r, k0 + 4 * q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <int KCH>
__device__ __forceinline__ void store_chunk_hilo(float* hi, float* lo, int r, int k0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const size_t o = xa_offset<KCH>(r, k0 + 4 * q);
    const float4 h = make_float4(ft_hi(v[4 * q]), ft_hi(v[4 * q + 1]), ft_hi(v[4 * q + 2]), ft_hi(v[4 * q + 3]));
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = make_float4(v[4 * q] - h.x, v[4 * q + 1] - h.y, v[4 * q + 2] - h.z, v[4 * q + 3] - h.w);
  }
}

// arrivals that complete op_ready[]: split mode = splitter warp + the producer's expect_tx for the weight images
constexpr uint32_t ft_op_ready_arrivals(bool split) { return split ? 2u : 1u; }

// where a layer's epilogue puts its result
enum { FT_OUT_OPERAND = 0,         // next layer's operand image (L.nx)
       FT_OUT_ROWS = 1,            // fp32 row-major L.out[(row0 + r) * N + n], rows < M
       FT_OUT_FEATURE_MAJOR = 2 }; // fp32 L.out[n * 128 + r] (per-cluster stage vector of the solver kernel)

// kernel-scope state of the layer routine (one per thread, identical in all threads of the CTA)
struct FtCtx {
  unsigned char* smem;             // dynamic shared memory: raw stages, then operand stages
  unsigned char* op_base;
  uint64_t* raw_full; uint64_t* raw_empty; uint64_t* op_ready; uint64_t* op_empty; uint64_t* accum_bar;
  uint32_t tmem_d, crank, nraw, raw_stage_bytes, op_stage_bytes;
  size_t xa_buf_floats;            // lo image = hi image + xa_buf_floats (only !FT_SPLIT)
  float* part;                     // this cluster's k-split partial sums
  uint32_t g0, accum_phase;        // running chunk count / accumulator barrier parity
  int tile;
};

struct FtLayer {
  int K, N, nN, nK, act, out_mode, stamp;
  const float* Wp; const float* Wlo; const float* bias;
  const float* a_src;              // operand image of the layer input (128 x K)
  float* nx;                       // FT_OUT_OPERAND destination
  float* out; int M, row0;         // FT_OUT_ROWS / FT_OUT_FEATURE_MAJOR destination
};

// All threads of all CTAs of the cluster call this with identical (per-CTA) arguments; ends with a cluster barrier
// after which every CTA's slice of the result is visible cluster-wide (and to bulk TMA).
template <int FT_NC, int FT_KCH, bool FT_SPLIT>
__device__ __forceinline__ void ft_layer(FtCtx& c, const FtLayer& L) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // CTA (cn, ck) of the layer's nN x nK grid: Nc output columns starting at cn * Nc, k range [ck * Ksl, (ck + 1) * Ksl)
  const int K = L.K, N = L.N, nN = L.nN, nK = L.nK;
  const int cn = static_cast<int>(c.crank) % nN, ck = static_cast<int>(c.crank) / nN;
  const int Nc = N / nN, Ksl = K / nK, nch = Ksl / FT_KCH, ch0 = ck * nch;
  const float* a_src_buf = L.a_src;
  float* nx = L.nx;
  const uint32_t a_bytes = FT_ROWS * FT_KCH * 4, w_bytes = static_cast<uint32_t>(Nc) * FT_KCH * 4;

  if (warp == FT_WARP_TMA) {
    // ===== TMA producer: ONE fp32 copy of the A chunk and of the weight chunk per raw stage.  Whole warp,
    // warp-uniform operands, one elected lane issues (a lone active lane costs an R2UR waterfall per copy).
    const float* wsrc = L.Wp + static_cast<size_t>(cn) * Nc * K + static_cast<size_t>(ch0) * Nc * FT_KCH;
    const float* asrc = a_src_buf + static_cast<size_t>(ch0) * FT_ROWS * FT_KCH;
    for (int ch = 0; ch < nch; ++ch) {
      const uint32_t g = c.g0 + ch, rs = g % c.nraw, rph = (g / c.nraw) & 1u;
      const uint32_t os_p = g % FT_OP_STAGES, oph_p = (g / FT_OP_STAGES) & 1u;
      mbar_wait(&c.raw_empty[rs], rph ^ 1u);
      if (FT_SPLIT) mbar_wait(&c.op_empty[os_p], oph_p ^ 1u);     // the tensor core is done with the stage's previous chunk
      unsigned char* dst = c.smem + static_cast<size_t>(rs) * c.raw_stage_bytes;
      if (elect_one()) {
        if (FT_SPLIT) {
          // the activation chunk goes to a raw stage (split hi / lo by a splitter warp); the weights are static, so their
          // hi / lo images were split once on the host side of the launch and land straight in the operand stage:
          // op_ready[os] completes on the splitter's arrive + this arrive + the weight bytes
          unsigned char* ob = c.op_base + static_cast<size_t>(os_p) * c.op_stage_bytes;
          const float* wlo = L.Wlo + static_cast<size_t>(cn) * Nc * K + static_cast<size_t>(ch0) * Nc * FT_KCH;
          mbar_arrive_expect_tx(&c.raw_full[rs], a_bytes);
          tma_load_1d(dst, asrc + static_cast<size_t>(ch) * FT_ROWS * FT_KCH, a_bytes, &c.raw_full[rs]);
          mbar_arrive_expect_tx(&c.op_ready[os_p], 2 * w_bytes);
          tma_load_1d(ob + 2 * a_bytes, wsrc + static_cast<size_t>(ch) * Nc * FT_KCH, w_bytes, &c.op_ready[os_p]);
          tma_load_1d(ob + 2 * a_bytes + w_bytes, wlo + static_cast<size_t>(ch) * Nc * FT_KCH, w_bytes, &c.op_ready[os_p]);
        } else {       // stage = [A_hi | A_lo | W_hi | W_lo], consumed by the MMA issuer directly
          const float* wlo = L.Wlo + static_cast<size_t>(cn) * Nc * K + static_cast<size_t>(ch0) * Nc * FT_KCH;
          mbar_arrive_expect_tx(&c.raw_full[rs], 2 * (a_bytes + w_bytes));
          tma_load_1d(dst, asrc + static_cast<size_t>(ch) * FT_ROWS * FT_KCH, a_bytes, &c.raw_full[rs]);
          tma_load_1d(dst + a_bytes, asrc + c.xa_buf_floats + static_cast<size_t>(ch) * FT_ROWS * FT_KCH, a_bytes, &c.raw_full[rs]);
          tma_load_1d(dst + 2 * a_bytes, wsrc + static_cast<size_t>(ch) * Nc * FT_KCH, w_bytes, &c.raw_full[rs]);
          tma_load_1d(dst + 2 * a_bytes + w_bytes, wlo + static_cast<size_t>(ch) * Nc * FT_KCH, w_bytes, &c.raw_full[rs]);
        }
      }
      __syncwarp();
    }
  } else if (FT_SPLIT && warp >= FT_WARP_SPLIT && warp < FT_WARP_TMA) {
    // ===== splitter: raw fp32 chunk -> operand stage (TF32-exact high part | exact residual).  Every warp
    // owns every 4th chunk by itself, so four chunks are in the split stage at once: one chunk costs
    // ~1 k clk of latency (two barrier waits, LDS -> STS, fence.proxy.async), which bounded the kernel
    // when all four warps worked on the same chunk.
    const int sw = warp - FT_WARP_SPLIT;
    const int a_vec = FT_ROWS * FT_KCH / 4;
    for (int ch = sw; ch < nch; ch += FT_SPLIT_WARPS) {
      const uint32_t g = c.g0 + ch, rs = g % c.nraw, rph = (g / c.nraw) & 1u;
      const uint32_t os = g % FT_OP_STAGES, oph = (g / FT_OP_STAGES) & 1u;
      mbar_wait(&c.raw_full[rs], rph);
      mbar_wait(&c.op_empty[os], oph ^ 1u);
      const unsigned char* raw = c.smem + static_cast<size_t>(rs) * c.raw_stage_bytes;
      unsigned char* ob = c.op_base + static_cast<size_t>(os) * c.op_stage_bytes;
      const float4* ra = reinterpret_cast<const float4*>(raw);
      float4* ahi = reinterpret_cast<float4*>(ob); float4* alo = reinterpret_cast<float4*>(ob + a_bytes);
#pragma unroll 8
      for (int e = lane; e < a_vec; e += 32) {
        const float4 x = ra[e];
        const float4 h = make_float4(ft_hi(x.x), ft_hi(x.y), ft_hi(x.z), ft_hi(x.w));
        ahi[e] = h; alo[e] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive(&c.op_ready[os]); mbar_arrive(&c.raw_empty[rs]); }
    }
  } else if (warp == FT_WARP_MMA) {
    // ===== MMA issuer: D[128 x Nc] = sum_k A[128 x k] W[Nc x k]^T, 3xTF32.  The WHOLE warp runs the loop
    // with warp-uniform operands and one elected lane issues: with a single active lane ptxas moves
    // every descriptor through an ELECT / R2UR.BROADCAST waterfall (~120 clk per MMA, measured).
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(Nc >> 3) << 17) |
                           (static_cast<uint32_t>(FT_ROWS >> 4) << 24);
    // The tensor core accumulates with truncation, so a long chain into ONE accumulator drifts by
    // ~n * 2^-24 (measured: 2.6e-5 after 288 accumulations).  The K range is therefore split over
    // `nseg` accumulators for the hi*hi products plus one for the (2^-11 smaller) cross terms; the
    // epilogue adds them in fp32 with round-to-nearest.
    const int nseg = ft_nseg(Nc);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, c.tmem_d, 0);
    uint32_t corr_started = 0;
    for (int ch = 0; ch < nch; ++ch) {
      const uint32_t g = c.g0 + ch;
      const uint32_t os = FT_SPLIT ? g % FT_OP_STAGES : g % c.nraw, oph = FT_SPLIT ? (g / FT_OP_STAGES) & 1u : (g / c.nraw) & 1u;
      uint64_t* wait_bar = FT_SPLIT ? &c.op_ready[os] : &c.raw_full[os];
      uint64_t* free_bar = FT_SPLIT ? &c.op_empty[os] : &c.raw_empty[os];
      mbar_wait(wait_bar, oph);
      if (ch == 0 && lane == 0) FT_STAMP(8 + L.stamp * 8 + 0);          // first chunk landed and split
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = FT_SPLIT ? smem_u32(c.op_base + static_cast<size_t>(os) * c.op_stage_bytes)
                                     : smem_u32(c.smem + static_cast<size_t>(os) * c.raw_stage_bytes);
      const uint32_t sa_hi = base, sa_lo = base + a_bytes, sw_hi = base + 2 * a_bytes, sw_lo = sw_hi + w_bytes;
      const int seg = ch * nseg / nch;
      const bool seg_first = ch == (seg * nch + nseg - 1) / nseg;
      const uint32_t d_main = tmem_u + static_cast<uint32_t>(seg * Nc), d_corr = tmem_u + static_cast<uint32_t>(nseg * Nc);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FT_KCH / 8; ++ks) {
          const uint32_t o = ks * 256;
          ft_mma(d_main, ft_desc<FT_KCH>(sa_hi + o), ft_desc<FT_KCH>(sw_hi + o), idesc, (seg_first && ks == 0) ? 0u : 1u);
          ft_mma(d_corr, ft_desc<FT_KCH>(sa_lo + o), ft_desc<FT_KCH>(sw_hi + o), idesc, (corr_started | ks) ? 1u : 0u);
          ft_mma(d_corr, ft_desc<FT_KCH>(sa_hi + o), ft_desc<FT_KCH>(sw_lo + o), idesc, 1);
        }
        ft_commit(free_bar);
      }
      corr_started = 1;
      __syncwarp();
    }
    if (elect_one()) ft_commit(c.accum_bar);
    __syncwarp();
    if (lane == 0) FT_STAMP(8 + L.stamp * 8 + 1);                         // all MMAs issued
  } else if (warp < FT_EPI_WARPS) {
    // ===== epilogue warps: thread = row; warps w and w + 4 share TMEM lane quarter w and split the columns
    mbar_wait(c.accum_bar, c.accum_phase);
    if (tid == 0) FT_STAMP(8 + L.stamp * 8 + 2);           // accumulators complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = tid & 127;
    const int half = tid >> 7, nchunks = Nc / 32;
    const int cbeg = half == 0 ? 0 : (nchunks + 1) / 2, cend = half == 0 ? (nchunks + 1) / 2 : nchunks;
    const int omode = L.out_mode;
    const int act = L.act;
    const int nseg = ft_nseg(Nc);
    for (int c0 = 32 * cbeg; c0 < 32 * cend; c0 += 32) {
      float accv[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) accv[q] = 0.f;
      uint32_t u[32];
      for (int sgm = nseg; sgm >= 0; --sgm) {        // cross terms first (smallest), then the K segments
      const uint32_t taddr = c.tmem_d + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(sgm * Nc + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
            "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
            "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
            "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; ++q) accv[q] += __uint_as_float(u[q]);
      }
      if (nK > 1) {
        // k-split layer: fp32 partial sums of this CTA's k range -> L2; reduced after the cluster barrier
        // column-major [slice][col][row]: the 32 rows of a warp are contiguous -> coalesced both ways
        float* dstp = c.part +
                      (static_cast<size_t>(ck * nN + cn) * Nc + c0) * FT_ROWS + r;
#pragma unroll
        for (int q = 0; q < 32; ++q) __stcg(dstp + static_cast<size_t>(q) * FT_ROWS, accv[q]);
        continue;
      }
      const int n0 = cn * Nc + c0;
      float v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(L.bias + n0 + 4 * q);
        const float4 a4 = apply_act4(make_float4(accv[4 * q] + b4.x, accv[4 * q + 1] + b4.y,
                                                 accv[4 * q + 2] + b4.z, accv[4 * q + 3] + b4.w), act);
        v[4 * q] = a4.x; v[4 * q + 1] = a4.y; v[4 * q + 2] = a4.z; v[4 * q + 3] = a4.w;
      }
      if (omode == FT_OUT_ROWS) {
        if (L.row0 + r < L.M) {
          float4* dst = reinterpret_cast<float4*>(L.out + static_cast<size_t>(L.row0 + r) * N + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      } else if (omode == FT_OUT_FEATURE_MAJOR) {
        float* dst = L.out + static_cast<size_t>(n0) * FT_ROWS + r;       // [feature][128 rows]: coalesced over the warp
#pragma unroll
        for (int q = 0; q < 32; ++q) __stcg(dst + static_cast<size_t>(q) * FT_ROWS, v[q]);
      } else {
        if (FT_SPLIT) store_chunk<FT_KCH>(nx, r, n0, v);
        else store_chunk_hilo<FT_KCH>(nx, nx + c.xa_buf_floats, r, n0, v);
      }
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (tid == 0) FT_STAMP(8 + L.stamp * 8 + 3);           // epilogue done
  }
  c.accum_phase ^= 1u;
  c.g0 += static_cast<uint32_t>(nch);
  __syncwarp();
  if (nK > 1) {
    // every CTA's partial is in L2: CTA (cn, ck) finishes columns [ck * Nc / nK, (ck + 1) * Nc / nK) of slice cn,
    // summing the nK partials in a fixed order (deterministic), then bias + activation + operand store
    cluster_sync_all();
    if (warp < FT_EPI_WARPS) {
      const int r = tid & 127, half = tid >> 7;
      const int ncols = Nc / nK, nchunks = ncols / 32;
      const int cbeg = half == 0 ? 0 : (nchunks + 1) / 2, cend = half == 0 ? (nchunks + 1) / 2 : nchunks;
      const int omode = L.out_mode;
      const int act = L.act;
      const float* pb = c.part;
      for (int cc = cbeg; cc < cend; ++cc) {
        const int c0 = ck * ncols + 32 * cc;                      // column inside slice cn
        const int n0 = cn * Nc + c0;
        float v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
        for (int kk = 0; kk < nK; ++kk) {
          const float* src = pb + (static_cast<size_t>(kk * nN + cn) * Nc + c0) * FT_ROWS + r;
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += __ldcg(src + static_cast<size_t>(q) * FT_ROWS);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(L.bias + n0 + 4 * q);
          const float4 a4 = apply_act4(make_float4(v[4 * q] + b4.x, v[4 * q + 1] + b4.y, v[4 * q + 2] + b4.z, v[4 * q + 3] + b4.w), act);
          v[4 * q] = a4.x; v[4 * q + 1] = a4.y; v[4 * q + 2] = a4.z; v[4 * q + 3] = a4.w;
        }
        if (omode == FT_OUT_ROWS) {
          if (L.row0 + r < L.M) {
            float4* dst = reinterpret_cast<float4*>(L.out + static_cast<size_t>(L.row0 + r) * N + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else if (omode == FT_OUT_FEATURE_MAJOR) {
          float* dst = L.out + static_cast<size_t>(n0) * FT_ROWS + r;
#pragma unroll
          for (int q = 0; q < 32; ++q) __stcg(dst + static_cast<size_t>(q) * FT_ROWS, v[q]);
        } else {
          if (FT_SPLIT) store_chunk<FT_KCH>(nx, r, n0, v);
          else store_chunk_hilo<FT_KCH>(nx, nx + c.xa_buf_floats, r, n0, v);
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    __syncwarp();
  }
  // every CTA's slice of the next operand is written (and this CTA's accumulator drained)
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) FT_STAMP(8 + L.stamp * 8 + 4);             // cluster barrier passed
}

// W [N][K] (PyTorch) -> fp32 operand image [c][K/KCH][Nc/8][KCH/4 kq][8 n][4 k]
__global__ void ft_pack_weight_kernel(const float* __restrict__ W, int N, int K, int NC, int KCH, float* __restrict__ dst,
                                      float* __restrict__ dst_lo) {
  const int Nc = N / NC;
  const size_t total = static_cast<size_t>(N) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<size_t>(n) * K);
    const int c = n / Nc, nl = n - c * Nc;
    const size_t o = static_cast<size_t>(c) * Nc * K +
                     ((static_cast<size_t>(k / KCH) * (Nc >> 3) + (nl >> 3)) * (KCH / 4) + ((k % KCH) >> 2)) * 32 +
                     (nl & 7) * 4 + (k & 3);
    const float w = W[i];
    if (dst_lo) { const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u); dst[o] = h; dst_lo[o] = w - h; }
    else dst[o] = w;
  }
}

struct FtPlan {
  int NL, ntiles, nclusters, kmax, ncmax, nraw, NC, KCH, split;
  int nN[FT_MAX_LAYERS], nK[FT_MAX_LAYERS];
  size_t off_part, part_floats;
  size_t off_w[FT_MAX_LAYERS], off_wlo[FT_MAX_LAYERS], off_xa, xa_buf_floats, total_bytes, smem_bytes;
  uint32_t raw_stage_bytes, op_stage_bytes;
  int K[FT_MAX_LAYERS], N[FT_MAX_LAYERS];
};

// mode: 0 = by tile count (ODEFunc.forward), 1 = clusters of 8 / split operands, 2 = clusters of 4 / pre-split wide MMAs
int ft_plan(int M, int D, int H, int n_hidden, FtPlan& pl, int mode = 0) {
  if (M <= 0 || n_hidden < 1 || n_hidden + 1 > FT_MAX_LAYERS) return ODEVIO_E_SHAPE;
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) {
    cudaGetLastError();
    nsm = 148;
  }
  pl.ntiles = (M + FT_ROWS - 1) / FT_ROWS;
  // column slices of 32 .. 256 (one MMA, <= 512 TMEM columns with two accumulators), tcgen05.ld in 32-column chunks
  auto ok = [](int n, int nc_) { const int nc = n / nc_; return n % nc_ == 0 && nc % 32 == 0 && nc >= 32 && nc <= 256; };
  // few tiles: 8 CTAs per tile (latency, SM count); many tiles: 4 CTAs per tile (wider MMAs, fewer A re-reads)
  if (mode == 2 && !(ok(D, 4) && ok(H, 4))) return ODEVIO_E_SHAPE;
  const bool wide = mode != 1 && (mode == 2 || pl.ntiles > nsm / 8) && ok(D, 4) && ok(H, 4);
  pl.NC = wide ? 4 : 8;
  pl.KCH = wide ? 16 : 8;
  pl.split = wide ? 0 : 1;
  pl.NL = n_hidden + 1;
  pl.kmax = D > H ? D : H;
  if (D % 64 || H % 64 || (D / pl.NC) % 32) return ODEVIO_E_SHAPE;     // per-CTA input-conversion slices of 32 features
  pl.ncmax = 0;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (off + n + 63) / 64 * 64; return o; };
  for (int l = 0; l < pl.NL; ++l) {
    const int K = l == 0 ? D : H, N = l == pl.NL - 1 ? D : H;
    pl.K[l] = K; pl.N[l] = N;
    if (wide) {
      pl.nN[l] = 4; pl.nK[l] = 1;
    } else {
      // 8 CTAs as an nN x nK grid with >= 192-column MMAs where the shape allows (the A-operand read of an
      // SS-mode MMA costs ~125 clk regardless of N): widest slice whose count divides the cluster
      pl.nN[l] = 0;
      const int cand[] = {256, 192, 128, 96, 64, 32};
      for (int ci = 0; ci < 6 && !pl.nN[l]; ++ci) {
        const int nsl = cand[ci];
        if (N % nsl) continue;
        const int nn = N / nsl;
        if (nn > 8 || 8 % nn) continue;
        const int nk = 8 / nn;
        if (K % (nk * pl.KCH) || (nsl / nk) % 32) continue;
        pl.nN[l] = nn; pl.nK[l] = nk;
      }
      if (!pl.nN[l]) return ODEVIO_E_SHAPE;
    }
    const int nsl = N / pl.nN[l];
    if (nsl % 32 || nsl < 32 || nsl > 256 || K % pl.KCH) return ODEVIO_E_SHAPE;
    if (nsl > pl.ncmax) pl.ncmax = nsl;
    pl.off_w[l] = take(static_cast<size_t>(K) * N);
    pl.off_wlo[l] = take(static_cast<size_t>(K) * N);
  }
  pl.nclusters = nsm / pl.NC;
  if (pl.nclusters > pl.ntiles) pl.nclusters = pl.ntiles;
  pl.xa_buf_floats = static_cast<size_t>(FT_ROWS) * pl.kmax;
  pl.off_xa = take(static_cast<size_t>(pl.nclusters) * 4 * pl.xa_buf_floats);
  pl.part_floats = static_cast<size_t>(8) * FT_ROWS * 256;
  pl.off_part = take(wide ? 64 : static_cast<size_t>(pl.nclusters) * pl.part_floats);
  pl.total_bytes = off * sizeof(float);
  if (pl.split) {
    pl.raw_stage_bytes = static_cast<uint32_t>(FT_ROWS) * pl.KCH * 4u;                 // fp32 activation chunk
    pl.op_stage_bytes = 2u * static_cast<uint32_t>(FT_ROWS + pl.ncmax) * pl.KCH * 4u;   // A hi | A lo | W hi | W lo
    pl.nraw = FT_RAW_STAGES;
    pl.smem_bytes = static_cast<size_t>(pl.nraw) * pl.raw_stage_bytes + static_cast<size_t>(FT_OP_STAGES) * pl.op_stage_bytes + 1024;
    if (pl.smem_bytes > 218u * 1024u) {          // leave room for the kernels' static shared memory (row state, barriers)
      pl.nraw = 4;
      pl.smem_bytes = static_cast<size_t>(pl.nraw) * pl.raw_stage_bytes + static_cast<size_t>(FT_OP_STAGES) * pl.op_stage_bytes + 1024;
    }
  } else {
    pl.raw_stage_bytes = 2u * static_cast<uint32_t>(FT_ROWS + pl.ncmax) * pl.KCH * 4u;      // hi + lo images
    pl.op_stage_bytes = 0;
    pl.nraw = static_cast<int>((224u * 1024u) / pl.raw_stage_bytes);
    if (pl.nraw > FT_RAW_STAGES) pl.nraw = FT_RAW_STAGES;
    pl.smem_bytes = static_cast<size_t>(pl.nraw) * pl.raw_stage_bytes + 1024;
    if (pl.nraw < 2) return ODEVIO_E_SHAPE;
  }
  if (pl.smem_bytes > 227u * 1024u) return ODEVIO_E_SHAPE;
  return 0;
}


}  // namespace
}  // namespace odevio
