// Tensor-core ODE solver, second generation (cfg.precision = ODEVIO_PRECISION_FP16X3): PoseODERNN.evolve_state
// (reference src/models/PoseODERNN.py:70-75) for every (sequence, rnn layer) row of one observation interval, the
// ODEFunc GEMMs (src/models/ODEFunc.py:38-39) on tcgen05 as 3xFP16 and the whole solver loop of a tile -- stage
// combines, Butcher tableau, per-row error norm, per-row step-size controller (torchode semantics as restated in
// oracle/torchode_like.py), FSAL -- inside ONE cluster of 4 CTAs, no host round trip between solver steps.
//
// What changed against odernn_tc.cu (3xTF32, clusters of 8 around 128-row tiles) and why -- every number below is from
// profiles/r02_mma_tma_probe.md:
//   * WEIGHTS ON THE M SIDE: D^T[feature][row] = W[feature][k] . X^T[k][row].  A CTA owns Fc = N_out / 4 output features
//     (one or two 128-feature MMA tiles) of every Linear for the tile's NR = 64 rows; 2048 rows = 32 clusters = 128 SMs
//     in ONE round (37 clusters of 4 are co-resident; only 15 clusters of 8 were, which forced a concurrent FFMA side
//     launch), no k-split, no partial-sum reduce through L2.
//   * 3xFP16 instead of 3xTF32: fp16 and tf32 both carry 11 significant bits, but kind::f16 takes K = 16 per MMA at the
//     cost of a kind::tf32 K = 8 MMA and the operands are half the bytes.  x = hi + lo * 2^-11 with hi = fp16(x),
//     lo = fp16((x - hi) * 2^11) (the scaling keeps the residual out of fp16's subnormal range); products hi*hi into
//     the main accumulators, lo*hi + hi*lo into a separate one that the epilogue scales by 2^-11: relative product
//     error <= 2^-22, the same as 3xTF32.  Operands must stay below 65504 in magnitude (tanh-bounded states and
//     activations do; an overflow surfaces as a non-finite error norm -> ODEVIO_STATUS_INFINITE_NORM).
//   * the accumulate of tcgen05.mma truncates (~n 2^-24 drift after n accumulations, measured in round 1): the K range
//     is split over up to 4 main accumulators per tile, added in fp32 round-to-nearest by the epilogue.
//   * 48 KB ring stages (32 KB of weights + 16 KB of activations = 64 k of a 128-feature tile, 12 MMAs = 576 clk): a
//     ring iteration costs ~450 clk of serial mbarrier / bulk-copy instruction latency whatever the stage size, so small
//     stages were the bound of the old kernel, not the L2 fabric.  Separate weight / activation rings with their own
//     producer warps; the weight producer runs ahead into the next Linear while the epilogue and the cluster barrier of
//     the current one are in flight (weights are static, activations are not).
//   * no integer division in any role loop (the old MMA issuer spent ~125 clk per MMA on them).
// Layouts: activations are exchanged between the CTAs of a cluster through L2 in the tensor core's canonical
// MN-major / no-swizzle fp16 image [k-chunk][hi|lo][row/8][k/8][k%8][row%8] -- feature-major, so that epilogue
// thread = feature writes 16-byte pieces; weights are pre-packed per CTA in the K-major image
// [k-chunk][tile][hi|lo][feature/8][k/8][feature%8][k%8].  One 1-D bulk TMA copy per chunk and ring.
// Stage vectors K0..K6, Y, Y1 of a tile: per-cluster L2-resident scratch, feature-major [D][NR] fp32; every element is
// only ever touched by the CTA that owns its feature.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/odevio.h"
#include "common.cuh"
#include "odernn_h3.h"
#include "odernn_params.h"

namespace odevio {

namespace {

// development timeline (-DODEVIO_H3_TIMELINE): clock64 stamps / wait sums of cluster 0 / CTA 0, last solver iteration
__device__ long long g_h3_dbg[96];
#ifdef ODEVIO_H3_TIMELINE
#ifndef H3_TL_STAGE
#define H3_TL_STAGE -1                    // record only this stage of the iteration (-1: every stage, the last one survives)
#endif
#define H3_STAMP(idx) do { if (blockIdx.x == 0 && c.tl_on) g_h3_dbg[idx] = clock64(); } while (0)
#define H3_ADD(idx, v) do { if (blockIdx.x == 0 && c.tl_on) g_h3_dbg[idx] += (v); } while (0)
#define H3_SET(idx, v) do { if (blockIdx.x == 0 && c.tl_on) g_h3_dbg[idx] = (v); } while (0)
#define H3_TL_SELECT(st) do { c.tl_on = (H3_TL_STAGE < 0 || (st) == H3_TL_STAGE) ? 1u : 0u; } while (0)
#define H3_CLOCK() clock64()
#else
#define H3_STAMP(idx) do { } while (0)
#define H3_ADD(idx, v) do { } while (0)
#define H3_SET(idx, v) do { } while (0)
#define H3_TL_SELECT(st) do { } while (0)
#define H3_CLOCK() 0ll
#endif

// A/B switch (measured, NOT kept -- DESIGN.md 4.9): 1 forms the next stage's argument in the epilogue of the ODEFunc's last
// Linear (H3Fuse) from partial sums prepared in the MMA shadows of the earlier Linears (H3Pre).  Same results (35 parity
// tests green), the 15.7 k clk stage-argument pass and its cluster barrier disappear -- but every MMA stream that shares
// the SM with a preparation pass gets ~2 k clk slower (operand chunks land later) and the last epilogue ~8 k clk longer
// (warps of lane quarters 2 / 3 own two blocks; the second one's fetch is exposed): stage 74.7 k vs 75.9 k clk on the
// timeline, configs[1] forward 16.2 ms vs 15.74 ms.
#ifndef H3_XJ_ALIAS
#define H3_XJ_ALIAS 1                     // jump-input images alias the stage vectors K0 .. K(2L-1) (needs a cluster barrier before they are written)
#endif
#ifndef H3_SKIP_LAST_BARRIER
#define H3_SKIP_LAST_BARRIER 0            // A/B: no cluster barrier after the ODEFunc's last Linear (the following pass is CTA-local): 14.95 vs 14.85 ms, not kept
#endif
#ifndef H3_TMEM_PAIR
#define H3_TMEM_PAIR 0                    // A/B: the two cross-term accumulators of an epilogue block fetched with one tcgen05.wait::ld
#endif
#ifndef H3_COMMIT_BATCH
#define H3_COMMIT_BATCH 1                 // A/B: commit pass fetches both row blocks (and the FSAL pair) before its first store
#endif
#ifndef H3_FUSE_STAGE_ARG
#define H3_FUSE_STAGE_ARG 0
#endif
constexpr int H3_NC = 4;                  // CTAs per cluster = feature slices of every Linear
constexpr int H3_MAXL = ODEVIO_MAX_ODE_LINEARS;
constexpr int H3_EPI_WARPS = 8;           // warps 0-7: epilogue + elementwise solver passes
constexpr int H3_WARP_WPROD = 8, H3_WARP_XPROD = 9, H3_WARP_MMA = 10, H3_WARP_MMA2 = 11;
constexpr int H3_THREADS = 32 * 12;
constexpr int H3_EPI_THREADS = 32 * H3_EPI_WARPS;
constexpr int H3_WCHUNK = 32768;          // bytes of a weight stage: T tiles x (hi | lo) x 128 features x KCH k (KCH = 64 / T)
constexpr int H3_NST = 4;                 // ring depth (power of two): stage s holds weight chunk + activation chunk g, s = g % 4

struct H3Layer {
  int K, N;                 // input / output features
  int Fc;                   // output features per CTA = N / nctas, a multiple of 32 in [128, 256]
  int T;                    // 128-feature MMA tiles per CTA: tile 0 = local features [0, 128), tile 1 = [Fc - 128, Fc)
  int KCH;                  // k per ring chunk = 64 / T
  int nch;                  // K / KCH
  int nseg;                 // main accumulators per tile (K split; + 1 accumulator for the cross terms)
  int act;
  int nctas;                // CTAs of the cluster that take part (4; 1 for the pose head's 128-feature Linear)
  const unsigned char* Wimg; // [nctas][nch][H3_WCHUNK]
  const float* bias;
  const float* bias2;       // second bias added in the epilogue (rnn: b_ih + b_hh) or nullptr
};

// one call of the layer routine: which columns (rows of the tile) it computes and where the result goes
struct H3Call {
  const unsigned char* xsrc;   // activation image of the input (chunks of L.KCH k)
  int col0, ncol;              // tile rows [col0, col0 + ncol) = the N side of the MMAs (ncol = 32 or 64)
  float* out;                  // fp32 [feature][NR] destination (stage vector / state / head activations) or nullptr
  unsigned char* xdst;         // activation image that receives act(W x + b) as fp16 hi / lo, or nullptr
  int xdst_kshift;             // log2 of that image's chunk size
  int xdst_colshift;           // the image row that receives tile row n is n + xdst_colshift
};

// Next stage's argument fused into the epilogue of the ODEFunc's last Linear (which produces k_st): the epilogue thread
// = feature also forms  y + dt * (S + a_st k_st),  S = sum_{j < st} a_j k_j,  for its rows and writes it as the activation
// image of the next stage's first Linear -- the separate stage-argument pass and its cluster barrier go away.  S and y are
// prepared by the epilogue warps in the shadow of the MMA streams of the stage's earlier Linears (H3Pre: coalesced pass,
// same operation order as h3_stage_input) in an "epilogue layout" -- EL[((cb * 8 + i) * own_nf + fl) * 4 + n % 4] for
// feature fl of the CTA's slice and tile row n = 32 cb + 4 i + n % 4 -- so that the 8 float4 loads of epilogue thread =
// feature are coalesced across the warp; they are issued before the wait on the accumulators.
struct H3Fuse {
  const float* S;                      // EL scratch of the CTA's slice (nprev > 0)
  const float* Yel;                    // y in the same layout
  int own_nf;
  float coef_last;                     // a[st + 1][st]
  int nprev;                           // = st: earlier stages in S
  const float* dt_rows;                // per-row step size (shared memory)
  unsigned char* xa; int kshift;       // destination image
};
struct H3Pre {                         // the preparation pass: features 32 * it of the slice, it in [it0, it1)
  const float* base; size_t arr; int own_f0, warp, fs, g, own_nf;      // the elementwise passes' thread mapping (H3Slice)
  const float* coef; int nprev;        // a[st + 1][0 .. nprev)
  float* S; float* Yel; int write_y;
  int it0, it1;
};

struct H3Params {
  int B, L, D, NL;
  int SPT;                   // sequences per tile: row r of a tile = (layer r / SPT, sequence tile * SPT + r % SPT), SPT = NR / L;
                             // 0 (L does not divide NR): kernel row g = l * B + b, a tile = 64 consecutive g
  H3Layer lay[H3_MAXL];
  H3Layer jump[2];           // rnn jump of layer l: [W_ih | W_hh] (K = 2 D), tanh
  H3Layer head;              // regressor.0 (D -> 128, LeakyReLU(0.1)); regressor.2 runs on the CUDA cores
  unsigned char* xa; size_t xa_buf_bytes;       // per cluster: 2 activation-image buffers of xa_buf_bytes
  unsigned char* xj; size_t xj_buf_bytes;       // per cluster: L jump-input images [x ; h] of 2 D x NR
  float* state; size_t state_floats;            // per cluster: (kMaxStages + 2) x [D][NR] + [4][NR] norm partials + [128][NR] head
  int ntiles;
  DevTableau tab;
  int adaptive, substeps;
  float atol, rtol, dt0, safety, fmin, fmax;
  int accept_strict, floor_factor, max_steps, exact_landing, endpoint_dense;
  const float* h0;           // [L][B][D] initial hidden state or nullptr (zeros); may alias hT
  float* hT;                 // [L][B][D] final hidden state
  const float* ts; int ts_ld;
  int interval0, nI;         // observation intervals [interval0, interval0 + nI) are integrated by this launch
  int do_jump;               // after every interval: rnn jump on the fused features + pose head (PoseODERNN.py:112-122)
  const float* fv; const float* fi; int Dv, S_io;   // features of interval i: fv[(b * S_io + i) * Dv + k], fi[.. * (D - Dv) + k - Dv]
  const float* reg_w1; const float* reg_b1;         // regressor.2: [6][128], [6]
  float* pose;               // [B][S_io][6]
  int* stats; int* status;
  // training: checkpoints for odevio_odernn_backward in the FMA kernels' layout (odernn_params.h: per tile of RTf sequences
  // and interval [Yend | Ypost | CK x (Y0 | dt[R] | upd[R])], T-layout [d][R], R = RTf * L, row r = l * RTf + m); nullptr: none
  float* ckpt; int* nloops; size_t ckpt_floats_per_tile; int CK, RTf, ntiles_f, S_total;
};

template <int NR>
struct H3Rows {      // per-row solver state, replicated in every CTA of the cluster
  __align__(16) float dt[NR];
  float t[NR], tend[NR], tmin[NR], tmax[NR];
  int run[NR], upd[NR], nsteps[NR], nacc[NR], status[NR];
  float dtstep[NR];                   // the step size the current iteration was taken with (checkpoints)
  float x[NR];                        // dense end point: (t_end - t) / dt of the step that reached t_end
  int toeval[NR], noteval[NR];        // this iteration's step reached t_end / t_end not evaluated yet
  int nsaved[NR / 4];                 // stored solver iterations of every checkpoint sub-tile in the current interval
  long long grow[NR];                 // state row of the tile's row (-1: beyond M)
  int bidx[NR], lyr[NR];
  float psum[H3_EPI_WARPS][NR];
};

struct H3Ctx {
  unsigned char* wring; unsigned char* xring;
  // full[s]: 2 arrivals (weight producer + activation producer, each with its bytes); empty[s]: 1 arrival (tcgen05.commit
  // after the chunk's MMAs).  One wait + one commit per chunk in the MMA issuer: every instruction between two MMAs of
  // the single issuing thread delays the tensor pipe (profiles/r02_mma_tma_probe.md, second part).
  uint64_t* full; uint64_t* empty; uint64_t* accum_bar;
  uint32_t tmem, crank;
  uint32_t count;            // chunks of all previous layers (identical in every role)
  uint32_t accum_phase;
  uint32_t w_ahead;          // weight producer: chunks of the coming layer already issued
  uint32_t tl_on;            // development timeline: stamps enabled
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint64_t h3_desc(uint32_t saddr, uint32_t hi_word) {
  // no-swizzle canonical layouts, LBO = 128 B (k groups of 8), SBO in hi_word (8-row / 8-column groups), version 1
  return (static_cast<uint64_t>(hi_word) << 32) | static_cast<uint64_t>(((saddr >> 4) & 0x3fffu) | ((128u >> 4) << 16));
}
__device__ __forceinline__ void h3_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void h3_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void h3_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void h3_tmem_ld32(uint32_t taddr, uint32_t (&u)[32]) {
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
}
__device__ __forceinline__ void h3_tmem_ld32_nowait(uint32_t taddr, uint32_t (&u)[32]) {
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
}
__device__ __forceinline__ void h3_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = hi + lo * 2^-11: hi = fp16(x), lo = fp16((x - hi) * 2^11)  (x - hi is exact in fp32)
__device__ __forceinline__ void h3_split(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}
__device__ __forceinline__ uint32_t h3_pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// byte offset of (feature k, row n0 = multiple of 4) in the activation image of a layer whose chunks hold KCH = 1 << kshift
// k: [k / KCH][hi | lo][n / 8][(k % KCH) / 8][k % 8][n % 8] fp16
template <int NR>
__device__ __forceinline__ size_t h3_x_offset(int k, int n0, int kshift) {
  const int KCH = 1 << kshift;
  return (static_cast<size_t>(k >> kshift) * (4u * KCH * NR)) + static_cast<size_t>(n0 >> 3) * ((KCH >> 3) * 128u) +
         static_cast<size_t>((k & (KCH - 1)) >> 3) * 128u + static_cast<size_t>(k & 7) * 16u + static_cast<size_t>(n0 & 7) * 2u;
}

// MMAs of one ring chunk: T tiles x (KCH / 16) k-steps x {hi*hi -> main accumulator of segment `seg`, lo*hi + hi*lo -> the
// tile's cross-term accumulator}; KCH = 64 / T.  Weight stage [tile][hi | lo][feature/8][k/8][feature%8][k%8] (K-major A),
// activation stage [hi | lo][row/8][k/8][k%8][row%8] (MN-major B); both no-swizzle with LBO = 128 B, SBO = (KCH / 8) * 128 B.
// The N side is the tile rows [col0, col0 + ncol): the B descriptors start col0 / 8 row groups into the stage, the
// accumulators of a tile are ncol columns apart.
template <int T, int NR>
__device__ __forceinline__ void h3_issue_chunk(uint32_t tmem, uint32_t wbase, uint32_t xbase, uint32_t nseg, uint32_t seg,
                                               uint32_t cross, bool acc_main, bool acc_cross, uint32_t col0, uint32_t ncol) {
  constexpr uint32_t KCH = 64u / T, KS = KCH / 16u;
  constexpr uint32_t sbo = (KCH >> 3) * 128u;
  constexpr uint32_t hi_word = (sbo >> 4) | (1u << 14);                        // SBO | descriptor version 1 (bit 46)
  constexpr uint32_t tile_bytes = 128u * KCH * 2u, ximg_bytes = KCH * NR * 2u;
  // instruction descriptor: D fp32, A / B fp16, A K-major, B MN-major, N = ncol, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 16) | ((ncol >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  const uint32_t acc_per_tile = (nseg + 2u) * ncol;           // main segments + one cross-term accumulator per issuer
  const uint32_t xcol = xbase + (col0 >> 3) * sbo;
#pragma unroll
  for (uint32_t t = 0; t < static_cast<uint32_t>(T); ++t) {
    const uint32_t a_hi = wbase + t * 2u * tile_bytes, a_lo = a_hi + tile_bytes;
    const uint32_t d_main = tmem + t * acc_per_tile + seg * ncol, d_cross = tmem + t * acc_per_tile + (nseg + cross) * ncol;
#pragma unroll
    for (uint32_t ks = 0; ks < KS; ++ks) {
      const uint32_t o = ks * 256u;
      const uint64_t ah = h3_desc(a_hi + o, hi_word), al = h3_desc(a_lo + o, hi_word);
      const uint64_t xh = h3_desc(xcol + o, hi_word), xl = h3_desc(xcol + ximg_bytes + o, hi_word);
      h3_mma(d_main, ah, xh, idesc, (acc_main || ks) ? 1u : 0u);
      h3_mma(d_cross, al, xh, idesc, (acc_cross || ks) ? 1u : 0u);
      h3_mma(d_cross, ah, xl, idesc, 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------- stage-argument preparation
template <int NR>
struct H3Slice {
  float* base; size_t arr;       // K[j] = base + j * arr, Y = base + 7 * arr, Y1 = base + 8 * arr
  int own_f0, nit, warp, fs, g;
};
__device__ __forceinline__ float4 h3_ld4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void h3_st4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

template <int N>
__device__ __forceinline__ float4 h3_wsum(const float4 (&k)[N > 0 ? N : 1], const float (&cf)[kMaxStages]) {
  float4 a = make_float4(mul_(k[0].x, cf[0]), mul_(k[0].y, cf[0]), mul_(k[0].z, cf[0]), mul_(k[0].w, cf[0]));
#pragma unroll
  for (int j = 1; j < N; ++j) {
    a.x = add_(a.x, mul_(k[j].x, cf[j])); a.y = add_(a.y, mul_(k[j].y, cf[j]));
    a.z = add_(a.z, mul_(k[j].z, cf[j])); a.w = add_(a.w, mul_(k[j].w, cf[j]));
  }
  return a;
}

// S = sum_{j < N} a_j k_j (left to right, as h3_stage_input) and, when asked, y -> epilogue layout; the CTA's feature slice,
// pass iterations [it0, it1)
template <int N, int NR>
__device__ __forceinline__ void h3_pre_pass_n(const H3Pre& pr) {
  const H3Pre& sl = pr;
  float cf[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) cf[j] = j < N ? pr.coef[j] : 0.f;
  const float* Y = sl.base + kMaxStages * sl.arr;
  const int own_nf = pr.own_nf;
  for (int it = pr.it0; it < pr.it1; ++it) {
    const int fl = 32 * it + 4 * sl.warp + sl.fs, f = sl.own_f0 + fl;
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      const size_t off = static_cast<size_t>(f) * NR + n0;
      const size_t el = ((static_cast<size_t>(m) * 8 + sl.g) * own_nf + fl) * 4;
      float4 k[N > 0 ? N : 1];
#pragma unroll
      for (int j = 0; j < N; ++j) k[j] = h3_ld4(sl.base + j * sl.arr + off);
      if (pr.write_y) h3_st4(pr.Yel + el, h3_ld4(Y + off));
      if (N > 0) h3_st4(pr.S + el, h3_wsum<N>(k, cf));
    }
  }
}
template <int NR>
__device__ __forceinline__ void h3_pre_pass(const H3Pre& pr) {
  switch (pr.nprev) {
    case 0: h3_pre_pass_n<0, NR>(pr); break;
    case 1: h3_pre_pass_n<1, NR>(pr); break;
    case 2: h3_pre_pass_n<2, NR>(pr); break;
    case 3: h3_pre_pass_n<3, NR>(pr); break;
    case 4: h3_pre_pass_n<4, NR>(pr); break;
    case 5: h3_pre_pass_n<5, NR>(pr); break;
    default: h3_pre_pass_n<6, NR>(pr); break;
  }
}

// ---------------------------------------------------------------------------------------------- one Linear
// All threads of all 4 CTAs call this with identical arguments.  Reads the cluster's activation image `cl.xsrc` (complete
// and visible: the caller's previous step ended with a cluster barrier) and computes act(W x + b) for the tile rows
// [cl.col0, cl.col0 + cl.ncol); the result goes to the activation image `cl.xdst` of a following Linear (fp16 hi / lo) and /
// or to fp32 [feature][NR] rows at `cl.out`.  Ends with a cluster barrier.  CTAs with rank >= L.nctas only take part in
// the barrier.  `next` (may be nullptr): the Linear that certainly follows -- its first weight chunks are issued before
// the barrier.
template <int NR>
__device__ __forceinline__ void h3_layer(H3Ctx& c, const H3Layer& L, const H3Layer* next, const H3Call& cl, int stamp = 0,
                                         const H3Fuse* fz = nullptr, const H3Pre* pre = nullptr, bool end_barrier = true) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sb = 16 + 10 * stamp;          // timeline slots of this layer
  (void)sb;
  constexpr uint32_t XSTAGE = 4u * 64u * NR;                 // bytes of an activation stage (KCH = 64: hi | lo)
  const int nch = L.nch;
  const bool active = static_cast<int>(c.crank) < L.nctas;
  const bool next_active = next != nullptr && static_cast<int>(c.crank) < next->nctas;

  if (warp == H3_WARP_WPROD) {
    // ===== weight producer: one 32 KB bulk copy per chunk; whole warp, one elected lane issues
    uint32_t g = c.count + c.w_ahead;
    if (active) {
      const unsigned char* src = L.Wimg + static_cast<size_t>(c.crank) * nch * H3_WCHUNK;
      for (int ch = static_cast<int>(c.w_ahead); ch < nch; ++ch, ++g) {
        const uint32_t s = g & (H3_NST - 1), ph = (g / H3_NST) & 1u;
        mbar_wait(&c.empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&c.full[s], H3_WCHUNK);
          tma_load_1d(c.wring + s * H3_WCHUNK, src + static_cast<size_t>(ch) * H3_WCHUNK, H3_WCHUNK, &c.full[s]);
        }
        __syncwarp();
      }
    }
    uint32_t pre = 0;
    if (next_active) {
      const unsigned char* nsrc = next->Wimg + static_cast<size_t>(c.crank) * next->nch * H3_WCHUNK;
      pre = next->nch < H3_NST ? next->nch : H3_NST;
      for (uint32_t ch = 0; ch < pre; ++ch, ++g) {
        const uint32_t s = g & (H3_NST - 1), ph = (g / H3_NST) & 1u;
        mbar_wait(&c.empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&c.full[s], H3_WCHUNK);
          tma_load_1d(c.wring + s * H3_WCHUNK, nsrc + static_cast<size_t>(ch) * H3_WCHUNK, H3_WCHUNK, &c.full[s]);
        }
        __syncwarp();
      }
    }
    c.w_ahead = pre;
  } else if (warp == H3_WARP_XPROD) {
    // ===== activation producer: one bulk copy (hi | lo images of KCH k x NR rows) per chunk
    if (active) {
      const uint32_t xcb = 4u * static_cast<uint32_t>(L.KCH) * NR;
      uint32_t g = c.count;
      for (int ch = 0; ch < nch; ++ch, ++g) {
        const uint32_t s = g & (H3_NST - 1), ph = (g / H3_NST) & 1u;
        mbar_wait(&c.empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&c.full[s], xcb);
          tma_load_1d(c.xring + s * XSTAGE, cl.xsrc + static_cast<size_t>(ch) * xcb, xcb, &c.full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == H3_WARP_MMA || warp == H3_WARP_MMA2) {
    // ===== MMA issuers: D^T[128 features x ncol rows] += W[128 x 16] X^T[16 x ncol], three products per k-step.  In each
    // of the two issuer warps ONE elected lane runs a whole chunk loop (wait, 12 MMAs, commit): measured, a per-chunk warp
    // reconvergence (elect + fence + syncwarp) costs ~200 clk and every wait / commit ~55 clk, none of it overlapped with
    // the 12 x 50 clk of MMAs of the same thread -- so the chunks alternate between two threads, each hiding the other's
    // wait + commit.  Issuer i takes the chunks ch = i (mod 2) and owns its own accumulators (main segments
    // [i * nseg / 2 ...) and cross-term accumulator i): the result does not depend on how the tensor pipe interleaves
    // the two streams.  The chunk body is fully unrolled (h3_issue_chunk), no division anywhere in the loop.
    if (active) {
      const uint32_t issuer = warp == H3_WARP_MMA ? 0u : 1u;
      const uint32_t nseg = static_cast<uint32_t>(L.nseg);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, c.tmem, 0);
      const uint32_t col0 = static_cast<uint32_t>(cl.col0), ncol = static_cast<uint32_t>(cl.ncol);
      // this issuer's chunks: ch = issuer, issuer + 2, ... (nmine of them) over its segments [seg0, seg0 + nsegi)
      const uint32_t nmine = (static_cast<uint32_t>(nch) + 1u - issuer) >> 1;
      const uint32_t nseg0 = (nseg + 1u) >> 1;
      const uint32_t seg0 = issuer ? nseg0 : 0u, nsegi = issuer ? nseg - nseg0 : nseg0;
      uint32_t cps = 1;                                               // chunks per segment (ceil)
      while (cps * nsegi < nmine) ++cps;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        uint32_t seg = seg0, in_seg = 0, g = c.count + issuer;
        const uint32_t wring_u = smem_u32(c.wring), xring_u = smem_u32(c.xring);
        for (uint32_t j = 0; j < nmine; ++j, g += 2) {
          const uint32_t st = g & (H3_NST - 1), ph = (g / H3_NST) & 1u;
          const long long tw0 = H3_CLOCK();
          mbar_wait(&c.full[st], ph);
          const long long tw1 = H3_CLOCK();
          if (issuer == 0) {
            if (j == 0) { H3_STAMP(sb + 0); H3_SET(sb + 6, 0); H3_SET(sb + 7, 0); }
            else H3_ADD(sb + 6, tw1 - tw0);                          // starvation after the first chunk
          } else if (j + 1 == nmine) H3_STAMP(sb + 5);
          (void)tw0; (void)tw1;
          const uint32_t wbase = wring_u + st * H3_WCHUNK, xbase = xring_u + st * XSTAGE;
          if (L.T == 1) h3_issue_chunk<1, NR>(tmem_u, wbase, xbase, nseg, seg, issuer, in_seg != 0, j != 0, col0, ncol);
          else h3_issue_chunk<2, NR>(tmem_u, wbase, xbase, nseg, seg, issuer, in_seg != 0, j != 0, col0, ncol);
          h3_commit(&c.empty[st]);
          if (++in_seg == cps) { in_seg = 0; ++seg; }
        }
        h3_commit(c.accum_bar);
        if (issuer == 0) H3_STAMP(sb + 1);
      }
      __syncwarp();
    }
  } else if (warp < H3_EPI_WARPS && active) {
    // ===== epilogue: thread = output feature (TMEM lane); warps w and w + 4 share lane quarter w & 3 and alternate over
    // the 32-column blocks of the tile rows
    const int q = warp & 3, h = warp >> 2;
    const int nseg = L.nseg, act = L.act, ncol = cl.ncol;
    // fused next-stage argument: partial sum over the earlier stages and y for this thread's FIRST block (tile 0, column block
    // h), fetched while the MMAs run; later blocks fetch theirs on the fly
    float fS[32], fY[32];
    auto fz_fetch = [&](int fl, int cb) {
      const size_t o = (static_cast<size_t>(cb) * 8 * fz->own_nf + fl) * 4;
      const float4* ys = reinterpret_cast<const float4*>(fz->Yel + o);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float4 v = __ldcg(ys + static_cast<size_t>(i) * fz->own_nf); fY[4 * i] = v.x; fY[4 * i + 1] = v.y; fY[4 * i + 2] = v.z; fY[4 * i + 3] = v.w; }
      if (fz->nprev > 0) {
        const float4* ss = reinterpret_cast<const float4*>(fz->S + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float4 v = __ldcg(ss + static_cast<size_t>(i) * fz->own_nf); fS[4 * i] = v.x; fS[4 * i + 1] = v.y; fS[4 * i + 2] = v.z; fS[4 * i + 3] = v.w; }
      }
    };
    if (pre) h3_pre_pass<NR>(*pre);
    const bool fz_first = fz != nullptr && h < (ncol >> 5);
    if (fz_first) fz_fetch(32 * q + lane, h);
    mbar_wait(c.accum_bar, c.accum_phase);
    if (tid == 0) H3_STAMP(sb + 2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int t = 0; t < L.T; ++t) {
      const int m = 32 * q + lane;                                  // row of the MMA tile = TMEM lane
      // tile 1 overlaps tile 0 when Fc < 256: only its upper Fc - 128 features are new (warp-uniform: Fc % 32 == 0)
      if (t == 1 && 32 * q < 256 - L.Fc) continue;
      const int fl = (t == 0 ? 0 : L.Fc - 128) + m;                 // feature inside the CTA's slice
      const int f = static_cast<int>(c.crank) * L.Fc + fl;          // output feature of the Linear
      const float bias = L.bias2 ? add_(L.bias[f], L.bias2[f]) : L.bias[f];
      const uint32_t tbase = c.tmem + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(t * (nseg + 2) * ncol);
#pragma unroll 1
      for (int cb = h; cb < (ncol >> 5); cb += 2) {
        const int colb = 32 * cb;                                   // column inside the accumulators
        const int col = cl.col0 + colb;                             // tile row
        uint32_t u[32];
        float acc[32];
#if H3_TMEM_PAIR
        {
          // both cross-term accumulators with one wait (acc is not live yet: no extra registers)
          uint32_t u2[32];
          h3_tmem_ld32_nowait(tbase + static_cast<uint32_t>(nseg * ncol + colb), u);
          h3_tmem_ld32_nowait(tbase + static_cast<uint32_t>((nseg + 1) * ncol + colb), u2);
          h3_tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = (__uint_as_float(u[i]) + __uint_as_float(u2[i])) * (1.0f / 2048.0f);
        }
#else
        h3_tmem_ld32(tbase + static_cast<uint32_t>(nseg * ncol + colb), u);             // cross terms first (smallest)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(u[i]);
        h3_tmem_ld32(tbase + static_cast<uint32_t>((nseg + 1) * ncol + colb), u);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (acc[i] + __uint_as_float(u[i])) * (1.0f / 2048.0f);
#endif
        for (int sgm = nseg - 1; sgm >= 0; --sgm) {
          h3_tmem_ld32(tbase + static_cast<uint32_t>(sgm * ncol + colb), u);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(u[i]);
        }
        if (act == ACT_TANH) {
          // inlined: the out-of-line apply_act4 call per 4 elements was a scheduling barrier between the 32 tanh chains and
          // the image stores (measured: 2.8 k of the epilogue's 6.7 k clk in the activation, 2.5 k in the stores)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = tanhf(acc[i] + bias);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a4 = apply_act4(make_float4(acc[4 * i] + bias, acc[4 * i + 1] + bias, acc[4 * i + 2] + bias, acc[4 * i + 3] + bias), act);
            acc[4 * i] = a4.x; acc[4 * i + 1] = a4.y; acc[4 * i + 2] = a4.z; acc[4 * i + 3] = a4.w;
          }
        }
        if (cl.out) {
          float4* dst = reinterpret_cast<float4*>(cl.out + static_cast<size_t>(f) * NR + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) __stcg(dst + i, make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]));
        }
        unsigned char* xdst = cl.xdst;
        int xdst_kshift = cl.xdst_kshift, xdst_colshift = cl.xdst_colshift;
        if (fz) {
          // y + dt * (sum_j a_j k_j), the sum left to right with the fresh k_st last (operation order of h3_stage_input)
          if (!(t == 0 && cb == h)) fz_fetch(fl, cb);
          const float cl_ = fz->coef_last;
          const bool first = fz->nprev == 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float sacc = first ? mul_(acc[i], cl_) : add_(fS[i], mul_(acc[i], cl_));
            acc[i] = add_(fY[i], mul_(fz->dt_rows[col + i], sacc));
          }
          xdst = fz->xa; xdst_kshift = fz->kshift; xdst_colshift = 0;
        }
        if (xdst) {
          const size_t lo_off = static_cast<size_t>(2u * NR) << xdst_kshift;       // KCH * NR * 2 bytes
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __half hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h3_split(acc[8 * j + i], hi[i], lo[i]);
            unsigned char* dst = xdst + h3_x_offset<NR>(f, col + xdst_colshift + 8 * j, xdst_kshift);
            __stcg(reinterpret_cast<uint4*>(dst),
                   make_uint4(h3_pack2(hi[0], hi[1]), h3_pack2(hi[2], hi[3]), h3_pack2(hi[4], hi[5]), h3_pack2(hi[6], hi[7])));
            __stcg(reinterpret_cast<uint4*>(dst + lo_off),
                   make_uint4(h3_pack2(lo[0], lo[1]), h3_pack2(lo[2], lo[3]), h3_pack2(lo[4], lo[5]), h3_pack2(lo[6], lo[7])));
          }
        }
      }
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");             // generic-proxy global stores -> bulk copies of all 4 CTAs
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (tid == 0) H3_STAMP(sb + 3);
  }
  if (active) {
    c.accum_phase ^= 1u;
    c.count += static_cast<uint32_t>(nch);
  }
  __syncwarp();
  // end_barrier == false (uniform over the cluster): the caller's next step only touches this CTA's own feature slice and
  // ends with a cluster barrier of its own, which then also orders this Linear's results and its TMEM reads
  if (end_barrier) {
    h3_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) H3_STAMP(sb + 4);
}

// ---------------------------------------------------------------------------------------------- elementwise passes
// 256 threads walk the CTA's own feature slice [own_f0, own_f0 + own_nf) of the [D][NR] stage vectors: lane = (fs, g),
// feature = own_f0 + 32 * it + 4 * warp + fs, rows 32 * m + 4 * g + {0..3} (m < NR / 32): every float4 load of 8 lanes
// covers 128 contiguous bytes.  Weighted sums run left to right over j in the oracle's order (oracle/_weighted_sum); zero
// coefficients contribute an exact +-0.
// stage argument y + dt * sum_{j<N} a_j k_j -> fp16 hi / lo activation image of the first Linear
template <int N, int NR>
__device__ __forceinline__ void h3_stage_input(const H3Slice<NR>& sl, const float* coef, const float* dt_rows, unsigned char* xa,
                                               int kshift) {
  float cf[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) cf[j] = j < N ? coef[j] : 0.f;
  const float* Y = sl.base + kMaxStages * sl.arr;
  const size_t lo_off = static_cast<size_t>(2u * NR) << kshift;
  for (int it = 0; it < sl.nit; ++it) {
    const int f = sl.own_f0 + 32 * it + 4 * sl.warp + sl.fs;
    // both row blocks' loads before the first store: the pass is bound by dependent L2 round trips, and the image stores
    // (through a char pointer) would otherwise order block 1's loads behind block 0's stores
    float4 yv[NR / 32];
    float4 kv[NR / 32][N > 0 ? N : 1];
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const size_t off = static_cast<size_t>(f) * NR + 32 * m + 4 * sl.g;
      yv[m] = h3_ld4(Y + off);
#pragma unroll
      for (int j = 0; j < N; ++j) kv[m][j] = h3_ld4(sl.base + j * sl.arr + off);
    }
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      float4 y = yv[m];
      if (N > 0) {
        const float4 dt = *reinterpret_cast<const float4*>(dt_rows + n0);
        const float4 s = h3_wsum<N>(kv[m], cf);
        y.x = add_(y.x, mul_(dt.x, s.x)); y.y = add_(y.y, mul_(dt.y, s.y));
        y.z = add_(y.z, mul_(dt.z, s.z)); y.w = add_(y.w, mul_(dt.w, s.w));
      }
      __half hi[4], lo[4];
      h3_split(y.x, hi[0], lo[0]); h3_split(y.y, hi[1], lo[1]); h3_split(y.z, hi[2], lo[2]); h3_split(y.w, hi[3], lo[3]);
      unsigned char* dst = xa + h3_x_offset<NR>(f, n0, kshift);
      __stcg(reinterpret_cast<uint2*>(dst), make_uint2(h3_pack2(hi[0], hi[1]), h3_pack2(hi[2], hi[3])));
      __stcg(reinterpret_cast<uint2*>(dst + lo_off), make_uint2(h3_pack2(lo[0], lo[1]), h3_pack2(lo[2], lo[3])));
    }
  }
}

// y1 -> Y1 and this thread's share of sum_d (err_d / bound_d)^2 for its rows (odernn_fwd.cu:error_pass)
template <int NS, int NR>
__device__ __forceinline__ void h3_error_pass(const H3Slice<NR>& sl, const DevTableau& tb, const float* dt_rows, float atol, float rtol,
                                              float (&sum)[NR / 32][4]) {
  float cy[kMaxStages], ce[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) {
    cy[j] = j < NS ? (tb.ssal ? (j < NS - 1 ? tb.a[NS - 1][j] : 0.f) : tb.b[j]) : 0.f;
    ce[j] = j < NS ? tb.e[j] : 0.f;
  }
  const float* Y = sl.base + kMaxStages * sl.arr;
  float* Y1 = sl.base + (kMaxStages + 1) * sl.arr;
#pragma unroll
  for (int m = 0; m < NR / 32; ++m) { sum[m][0] = 0.f; sum[m][1] = 0.f; sum[m][2] = 0.f; sum[m][3] = 0.f; }
  for (int it = 0; it < sl.nit; ++it) {
    const int f = sl.own_f0 + 32 * it + 4 * sl.warp + sl.fs;
    // both row blocks' loads before the first store (see h3_stage_input)
    float4 y0v[NR / 32], kv[NR / 32][NS];
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const size_t off = static_cast<size_t>(f) * NR + 32 * m + 4 * sl.g;
      y0v[m] = h3_ld4(Y + off);
#pragma unroll
      for (int j = 0; j < NS; ++j) kv[m][j] = h3_ld4(sl.base + j * sl.arr + off);
    }
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      const size_t off = static_cast<size_t>(f) * NR + n0;
      const float4 y0 = y0v[m];
      const float4 (&k)[NS] = kv[m];
      const float4 dt = *reinterpret_cast<const float4*>(dt_rows + n0);
      const float4 sy = h3_wsum<NS>(k, cy);
      const float4 y1 = make_float4(add_(y0.x, mul_(dt.x, sy.x)), add_(y0.y, mul_(dt.y, sy.y)), add_(y0.z, mul_(dt.z, sy.z)),
                                    add_(y0.w, mul_(dt.w, sy.w)));
      h3_st4(Y1 + off, y1);
      if (tb.has_err) {
        const float4 se = h3_wsum<NS>(k, ce);
        const float e0 = mul_(dt.x, se.x), e1 = mul_(dt.y, se.y), e2 = mul_(dt.z, se.z), e3 = mul_(dt.w, se.w);
        const float q0 = __fdiv_rn(e0, add_(atol, mul_(rtol, fmaxf(fabsf(y0.x), fabsf(y1.x)))));
        const float q1 = __fdiv_rn(e1, add_(atol, mul_(rtol, fmaxf(fabsf(y0.y), fabsf(y1.y)))));
        const float q2 = __fdiv_rn(e2, add_(atol, mul_(rtol, fmaxf(fabsf(y0.z), fabsf(y1.z)))));
        const float q3 = __fdiv_rn(e3, add_(atol, mul_(rtol, fmaxf(fabsf(y0.w), fabsf(y1.w)))));
        sum[m][0] = fmaf(q0, q0, sum[m][0]); sum[m][1] = fmaf(q1, q1, sum[m][1]);
        sum[m][2] = fmaf(q2, q2, sum[m][2]); sum[m][3] = fmaf(q3, q3, sum[m][3]);
      }
    }
  }
}

// fixed step: Y <- y0 + dt * sum b_j k_j (odernn_fwd.cu:fixed_commit)
template <int NS, int NR>
__device__ __forceinline__ void h3_fixed_commit(const H3Slice<NR>& sl, const DevTableau& tb, const float* dt_rows) {
  float cb[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) cb[j] = j < NS ? tb.b[j] : 0.f;
  float* Y = sl.base + kMaxStages * sl.arr;
  for (int it = 0; it < sl.nit; ++it) {
    const int f = sl.own_f0 + 32 * it + 4 * sl.warp + sl.fs;
#pragma unroll
    for (int m = 0; m < NR / 32; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      const size_t off = static_cast<size_t>(f) * NR + n0;
      const float4 y0 = h3_ld4(Y + off);
      float4 k[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) k[j] = h3_ld4(sl.base + j * sl.arr + off);
      const float4 dt = *reinterpret_cast<const float4*>(dt_rows + n0);
      const float4 s = h3_wsum<NS>(k, cb);
      h3_st4(Y + off, make_float4(add_(y0.x, mul_(dt.x, s.x)), add_(y0.y, mul_(dt.y, s.y)), add_(y0.z, mul_(dt.z, s.z)),
                                  add_(y0.w, mul_(dt.w, s.w))));
    }
  }
}

// accepted rows: Y <- Y1, FSAL carry K0 <- K[ns - 1]; other rows keep both
template <int NR>
__device__ __forceinline__ void h3_commit_rows(const H3Slice<NR>& sl, const DevTableau& tb, const int* upd_rows,
                                               const int* toeval_rows, const float* x_rows, const float* dt_rows) {
  const int ns = tb.n_stages, fsal = tb.fsal;
  float* Y = sl.base + kMaxStages * sl.arr;
  const float* Y1 = sl.base + (kMaxStages + 1) * sl.arr;
  const float* Kl = sl.base + static_cast<size_t>(ns - 1) * sl.arr;
  constexpr int MB = NR / 32;
  for (int it = 0; it < sl.nit; ++it) {
    const int f = sl.own_f0 + 32 * it + 4 * sl.warp + sl.fs;
    // all loads of both row blocks (y1, y and the FSAL pair k_last, k_0) before the first store: the pass is bound by L2
    // round trips, and it used to make two dependent ones per row block
    int4 uv[MB];
    float4 av[MB], ov[MB], bv[MB], o0v[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      uv[m] = *reinterpret_cast<const int4*>(upd_rows + n0);
      if (!(uv[m].x | uv[m].y | uv[m].z | uv[m].w)) continue;
#if H3_COMMIT_BATCH
      const size_t off = static_cast<size_t>(f) * NR + n0;
      av[m] = h3_ld4(Y1 + off); ov[m] = h3_ld4(Y + off);
      if (fsal) { bv[m] = h3_ld4(Kl + off); o0v[m] = h3_ld4(sl.base + off); }
#endif
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const int n0 = 32 * m + 4 * sl.g;
      const int4 u = uv[m];
      if (!(u.x | u.y | u.z | u.w)) continue;
      const size_t off = static_cast<size_t>(f) * NR + n0;
#if !H3_COMMIT_BATCH
      av[m] = h3_ld4(Y1 + off); ov[m] = h3_ld4(Y + off);
#endif
      const float4 a = av[m], o = ov[m];
      float4 v = make_float4(u.x ? a.x : o.x, u.y ? a.y : o.y, u.z ? a.z : o.z, u.w ? a.w : o.w);
      const int4 te = *reinterpret_cast<const int4*>(toeval_rows + n0);
      if (te.x | te.y | te.z | te.w) {
        // dense output of the accepted step at x = (t_end - t) / dt: quartic through (y0, y1, dt f0, dt f1, y_mid), or the
        // linear interpolant for the tableaus without b_mid -- operation order of odernn_fwd.cu: commit_pass / the oracle
        const float4 dtv = *reinterpret_cast<const float4*>(dt_rows + n0);
        const float4 xv = *reinterpret_cast<const float4*>(x_rows + n0);
        const float y0[4] = {o.x, o.y, o.z, o.w}, y1[4] = {a.x, a.y, a.z, a.w};
        const float dts[4] = {dtv.x, dtv.y, dtv.z, dtv.w}, xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const int tev[4] = {te.x, te.y, te.z, te.w};
        float out[4] = {v.x, v.y, v.z, v.w};
        if (tb.has_mid) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          bool any = false;
          for (int j = 0; j < ns; ++j) {
            const float bm = tb.bmid[j];
            if (bm == 0.f) continue;
            const float4 k = h3_ld4(sl.base + static_cast<size_t>(j) * sl.arr + off);
            const float kk[4] = {k.x, k.y, k.z, k.w};
            for (int q = 0; q < 4; ++q) acc[q] = any ? add_(acc[q], mul_(kk[q], bm)) : mul_(kk[q], bm);
            any = true;
          }
          const float4 k0 = h3_ld4(sl.base + off), kl = h3_ld4(Kl + off);
          const float k0a[4] = {k0.x, k0.y, k0.z, k0.w}, kla[4] = {kl.x, kl.y, kl.z, kl.w};
          for (int q = 0; q < 4; ++q) {
            if (!tev[q]) continue;
            const float ymid = add_(y0[q], mul_(dts[q], acc[q]));
            const float f0 = mul_(dts[q], k0a[q]), f1 = mul_(dts[q], kla[q]);
            const float ca = add_(sub_(mul_(2.0f, sub_(f1, f0)), mul_(8.0f, add_(y1[q], y0[q]))), mul_(16.0f, ymid));
            const float cb = sub_(add_(add_(sub_(mul_(5.0f, f0), mul_(3.0f, f1)), mul_(18.0f, y0[q])), mul_(14.0f, y1[q])),
                                  mul_(32.0f, ymid));
            const float cc = add_(sub_(sub_(sub_(f1, mul_(4.0f, f0)), mul_(11.0f, y0[q])), mul_(5.0f, y1[q])), mul_(16.0f, ymid));
            const float x = xs[q];
            out[q] = add_(mul_(add_(mul_(add_(mul_(add_(mul_(ca, x), cb), x), cc), x), f0), x), y0[q]);
          }
        } else {
          for (int q = 0; q < 4; ++q) if (tev[q]) out[q] = add_(y0[q], mul_(xs[q], sub_(y1[q], y0[q])));
        }
        v = make_float4(out[0], out[1], out[2], out[3]);
      }
      h3_st4(Y + off, v);
      if (fsal) {
#if !H3_COMMIT_BATCH
        bv[m] = h3_ld4(Kl + off); o0v[m] = h3_ld4(sl.base + off);
#endif
        const float4 b = bv[m], o0 = o0v[m];
        h3_st4(sl.base + off, make_float4(u.x ? b.x : o0.x, u.y ? b.y : o0.y, u.z ? b.z : o0.z, u.w ? b.w : o0.w));
      }
    }
  }
}

#define H3_DISPATCH_STAGES(n, CALL)                                                                                 \
  switch (n) {                                                                                                      \
    case 1: { constexpr int NSV = 1; CALL; break; }                                                                 \
    case 2: { constexpr int NSV = 2; CALL; break; }                                                                 \
    case 3: { constexpr int NSV = 3; CALL; break; }                                                                 \
    case 4: { constexpr int NSV = 4; CALL; break; }                                                                 \
    case 5: { constexpr int NSV = 5; CALL; break; }                                                                 \
    case 6: { constexpr int NSV = 6; CALL; break; }                                                                 \
    default: { constexpr int NSV = 7; CALL; break; }                                                                \
  }

__device__ __forceinline__ int h3_log2(int v) { return 31 - __clz(v); }

// Input images of the rnn jump (PoseODERNN.py:112-117): layer l's image holds [x ; h] for its tile rows, x = the fused
// features of the interval (layer 0: cat(fv, fi), FusionModule "cat") or the new hidden state of layer l - 1 (written by
// that jump's epilogue), h = the evolved state.  This pass writes, for the CTA's own feature slice, the h part of every
// layer (k = D + f) and the feature part of layer 0 (k = f).
template <int NR>
__device__ __forceinline__ void h3_jump_input(const H3Params& p, const H3Rows<NR>& rs, const float* Yc, unsigned char* xj,
                                              int own_f0, int own_nf, int interval, int tid) {
  const int D = p.D, kshift = h3_log2(p.jump[0].KCH);
  const size_t lo_off = static_cast<size_t>(2u * NR) << kshift;
  // h part: task = (feature, group of 8 rows); 8 consecutive lanes cover the NR = 64 rows of one feature (256 B)
  for (int task = tid; task < own_nf * (NR / 8); task += H3_EPI_THREADS) {
    const int f = own_f0 + task / (NR / 8), n0 = 8 * (task % (NR / 8));
    const float4 a = h3_ld4(Yc + static_cast<size_t>(f) * NR + n0), b = h3_ld4(Yc + static_cast<size_t>(f) * NR + n0 + 4);
    __half hi[8], lo[8];
    h3_split(a.x, hi[0], lo[0]); h3_split(a.y, hi[1], lo[1]); h3_split(a.z, hi[2], lo[2]); h3_split(a.w, hi[3], lo[3]);
    h3_split(b.x, hi[4], lo[4]); h3_split(b.y, hi[5], lo[5]); h3_split(b.z, hi[6], lo[6]); h3_split(b.w, hi[7], lo[7]);
    const int lyr = p.SPT ? n0 / p.SPT : 0;
    unsigned char* dst = xj + static_cast<size_t>(lyr) * p.xj_buf_bytes + h3_x_offset<NR>(D + f, n0, kshift);
    __stcg(reinterpret_cast<uint4*>(dst), make_uint4(h3_pack2(hi[0], hi[1]), h3_pack2(hi[2], hi[3]), h3_pack2(hi[4], hi[5]), h3_pack2(hi[6], hi[7])));
    __stcg(reinterpret_cast<uint4*>(dst + lo_off),
           make_uint4(h3_pack2(lo[0], lo[1]), h3_pack2(lo[2], lo[3]), h3_pack2(lo[4], lo[5]), h3_pack2(lo[6], lo[7])));
  }
  // feature part of layer 0 (tile rows [0, SPT)): task = (group of 8 sequences, feature); consecutive lanes read
  // consecutive features of one sequence (coalesced) and write consecutive 16-byte pieces of the image
  const int ngrp = p.SPT / 8;
  for (int task = tid; task < own_nf * ngrp; task += H3_EPI_THREADS) {
    const int grp = task / own_nf, f = own_f0 + task - grp * own_nf;
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int n = 8 * grp + e;
      float v = 0.f;
      if (rs.grow[n] >= 0) {
        const size_t row = static_cast<size_t>(rs.bidx[n]) * p.S_io + interval;
        // read once per forward: streaming (evict-first) loads keep the 31 MB of features from displacing the L2-resident scratch
        v = f < p.Dv ? __ldcs(p.fv + row * p.Dv + f) : __ldcs(p.fi + row * (D - p.Dv) + (f - p.Dv));
      }
      h3_split(v, hi[e], lo[e]);
    }
    unsigned char* dst = xj + h3_x_offset<NR>(f, 8 * grp, kshift);
    __stcg(reinterpret_cast<uint4*>(dst), make_uint4(h3_pack2(hi[0], hi[1]), h3_pack2(hi[2], hi[3]), h3_pack2(hi[4], hi[5]), h3_pack2(hi[6], hi[7])));
    __stcg(reinterpret_cast<uint4*>(dst + lo_off),
           make_uint4(h3_pack2(lo[0], lo[1]), h3_pack2(lo[2], lo[3]), h3_pack2(lo[4], lo[5]), h3_pack2(lo[6], lo[7])));
  }
}

// Checkpoints (training): the 64-row tile is nsub = SPT / RTf sub-tiles of the FMA kernels' geometry (RTf sequences x L
// layers, T-layout [d][R] with r = l * RTf + m).  Copies the CTA's feature slice of Yc into the slot of every sub-tile whose
// bit is set in `mask`: stored iteration slot_of[u] (slot_of != nullptr) or the fixed array `fixed_idx` (0 = Yend, 1 = Ypost).
template <int NR>
__device__ __forceinline__ void h3_ckpt_copy(const H3Params& p, int tile, int interval, const float* Yc, int own_f0, int own_nf, int tid,
                                             unsigned mask, const int* slot_of, int fixed_idx) {
  const int RTf = p.RTf, nsub = p.SPT / RTf, R = RTf * p.L;
  const size_t arr = static_cast<size_t>(p.D) * R, ivf = ckpt_interval_floats(p.D, R, p.CK);
  const int pieces = RTf / 4;                                 // float4 pieces of one (feature, layer, sub-tile) run of rows
  const int ntask = own_nf * p.L * nsub * pieces;
  for (int task = tid; task < ntask; task += H3_EPI_THREADS) {
    int r = task;
    const int pc = r % pieces; r /= pieces;
    const int u = r % nsub; r /= nsub;
    const int l = r % p.L; r /= p.L;
    const int d = own_f0 + r;
    const int tf = tile * nsub + u;
    if (!((mask >> u) & 1u) || tf >= p.ntiles_f) continue;
    const float4 v = h3_ld4(Yc + static_cast<size_t>(d) * NR + l * p.SPT + u * RTf + 4 * pc);
    float* dst = p.ckpt + static_cast<size_t>(tf) * p.ckpt_floats_per_tile + static_cast<size_t>(interval) * ivf +
                 (slot_of ? 2 * arr + static_cast<size_t>(slot_of[u]) * (arr + 2 * R) : static_cast<size_t>(fixed_idx) * arr);
    *reinterpret_cast<float4*>(dst + static_cast<size_t>(d) * R + l * RTf + 4 * pc) = v;
  }
}

// One solver iteration's record (odernn_fwd.cu:PH_STEP_END): for every sub-tile with an accepted row, the state the iteration
// started from (Yc before the commit), the step sizes it used and the accept flags.  Called by ALL threads of the CTA after
// the controller's __syncthreads; ends with a __syncthreads.
template <int NR>
__device__ __forceinline__ void h3_ckpt_iteration(const H3Params& p, H3Rows<NR>& rs, int tile, int interval, const float* Yc, int own_f0,
                                                  int own_nf, uint32_t crank) {
  const int tid = threadIdx.x;
  const int RTf = p.RTf, nsub = p.SPT / RTf, R = RTf * p.L;
  unsigned mask = 0;
  for (int u = 0; u < nsub; ++u) {
    int any = 0;
    for (int l = 0; l < p.L; ++l)
      for (int m = 0; m < RTf; ++m) any |= rs.upd[l * p.SPT + u * RTf + m];
    if (any && rs.nsaved[u] < p.CK) mask |= 1u << u;
    else if (any && tid < NR && (tid % p.SPT) / RTf == u) rs.status[tid] = max(rs.status[tid], 3);      // checkpoint overflow
  }
  if (tid < H3_EPI_THREADS) h3_ckpt_copy<NR>(p, tile, interval, Yc, own_f0, own_nf, tid, mask, rs.nsaved, 0);
  if (crank == 0 && tid < NR && rs.grow[tid] >= 0) {
    const int l = tid / p.SPT, j = tid - l * p.SPT, u = j / RTf, m = j - u * RTf, tf = tile * nsub + u;
    if (((mask >> u) & 1u) && tf < p.ntiles_f) {
      const size_t arr = static_cast<size_t>(p.D) * R;
      float* slot = p.ckpt + static_cast<size_t>(tf) * p.ckpt_floats_per_tile + static_cast<size_t>(interval) * ckpt_interval_floats(p.D, R, p.CK) +
                    2 * arr + static_cast<size_t>(rs.nsaved[u]) * (arr + 2 * R);
      slot[arr + l * RTf + m] = rs.dtstep[tid];
      reinterpret_cast<int*>(slot + arr + R)[l * RTf + m] = rs.upd[tid];
    }
  }
  __syncthreads();
  if (tid < nsub && ((mask >> tid) & 1u)) rs.nsaved[tid] += 1;
  __syncthreads();
}

template <int NR>
__global__ void __launch_bounds__(H3_THREADS, 1) odernn_h3_kernel(const __grid_constant__ H3Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[H3_NST];
  __shared__ __align__(8) uint64_t empty_bar[H3_NST];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) H3Rows<NR> rs;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x / H3_NC, nclusters = gridDim.x / H3_NC;
  const DevTableau& tb = p.tab;

  if (tid == 0) {
    for (int i = 0; i < H3_NST; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    mbar_init(&accum_bar, 2);                     // both MMA issuers commit
    fence_barrier_init();
  }
  if (warp == H3_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  H3Ctx c;
  c.wring = smem; c.xring = smem + H3_NST * H3_WCHUNK;
  c.full = full_bar; c.empty = empty_bar; c.accum_bar = &accum_bar;
  c.tmem = tmem_slot; c.crank = crank; c.count = 0; c.accum_phase = 0; c.w_ahead = 0; c.tl_on = 1;

  const int D = p.D, NL = p.NL;
  const int own_nf = D / H3_NC, own_f0 = static_cast<int>(crank) * own_nf;
  const size_t arr = static_cast<size_t>(D) * NR;
  float* const st_base = p.state + static_cast<size_t>(cluster_id) * p.state_floats;
  float* const Yc = st_base + kMaxStages * arr;
  float* const normpart = st_base + (kMaxStages + 2) * arr;            // [4][NR]
  float* const headbuf = normpart + H3_NC * NR;                        // [128][NR] activations of regressor.0
  unsigned char* const xa0 = p.xa + static_cast<size_t>(cluster_id) * 2 * p.xa_buf_bytes;
  unsigned char* const xa1 = xa0 + p.xa_buf_bytes;
  // the jump-input images [x ; h] of the L layers live in the stage vectors K0 .. K(2L-1), which are dead between the end of an
  // interval's solves and the next interval's first stage (FSAL does not carry across the jump): 25 MB less L2-resident
  // scratch at configs[1] (2 D x NR x 4 B = 2 stage vectors per layer)
#if H3_XJ_ALIAS
  unsigned char* const xj = reinterpret_cast<unsigned char*>(st_base);
#else
  unsigned char* const xj = p.xj + static_cast<size_t>(cluster_id) * p.L * p.xj_buf_bytes;
#endif

  const bool epi = warp < H3_EPI_WARPS;
  H3Slice<NR> sl;
  sl.base = st_base; sl.arr = arr; sl.own_f0 = own_f0; sl.nit = own_nf / 32; sl.warp = warp; sl.fs = lane >> 3; sl.g = lane & 7;
  const int ns = tb.n_stages;
  const int kshift0 = h3_log2(p.lay[0].KCH);
  const long long Mrows = static_cast<long long>(p.L) * p.B;

  for (int tile = cluster_id; tile < p.ntiles; tile += nclusters) {
    // ---- which (layer, sequence) each tile row is
    if (tid < NR) {
      const int r = tid;
      bool valid; int lyr, b;
      if (p.SPT) { lyr = r / p.SPT; b = tile * p.SPT + (r - lyr * p.SPT); valid = b < p.B; }
      else { const long long g = static_cast<long long>(tile) * NR + r; valid = g < Mrows; lyr = valid ? static_cast<int>(g / p.B) : 0; b = valid ? static_cast<int>(g - static_cast<long long>(lyr) * p.B) : 0; }
      rs.grow[r] = valid ? static_cast<long long>(lyr) * p.B + b : -1;
      rs.bidx[r] = valid ? b : 0; rs.lyr[r] = lyr;
    }
    __syncthreads();
    // ---- the tile's state rows (row-major [L][B][D]; zeros without h0) -> feature-major scratch, own feature slice
    if (epi) {
      const int nf4 = own_nf / 4;
      for (int item = tid; item < NR * nf4; item += H3_EPI_THREADS) {
        const int n = item / nf4, f4 = item - n * nf4;
        const long long gr = rs.grow[n];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr >= 0 && p.h0) v = *reinterpret_cast<const float4*>(p.h0 + static_cast<size_t>(gr) * D + own_f0 + 4 * f4);
        float* dst = Yc + static_cast<size_t>(own_f0 + 4 * f4) * NR + n;
        __stcg(dst, v.x); __stcg(dst + NR, v.y); __stcg(dst + 2 * NR, v.z); __stcg(dst + 3 * NR, v.w);
      }
      named_bar_sync(1, H3_EPI_THREADS);      // the elementwise passes walk the slice with another thread mapping
    }

    for (int ii = 0; ii < p.nI; ++ii) {
      const int interval = p.interval0 + ii;
      // ---- per-row solver state (PoseODERNN.py:70-75; odernn_fwd.cu:interval_begin)
      int run = 0;
      if (tid < NR) {
        const int r = tid;
        const bool valid = rs.grow[r] >= 0;
        float t0 = 0.f, t1 = 0.f;
        if (valid) {
          t0 = p.ts[static_cast<size_t>(rs.bidx[r]) * p.ts_ld + interval];
          t1 = p.ts[static_cast<size_t>(rs.bidx[r]) * p.ts_ld + interval + 1];
        }
        rs.t[r] = t0; rs.tend[r] = t1;
        rs.tmin[r] = fminf(t0, t1); rs.tmax[r] = fmaxf(t0, t1);
        rs.nsteps[r] = 0; rs.nacc[r] = 0; rs.upd[r] = 0; rs.status[r] = 0;
        rs.toeval[r] = 0; rs.noteval[r] = 1; rs.x[r] = 1.0f;
        if (p.adaptive) {
          rs.dt[r] = fminf(fmaxf(p.dt0, sub_(rs.tmin[r], t0)), sub_(rs.tmax[r], t0));
          run = (valid && t0 < t1) ? 1 : 0;
        } else {
          rs.dt[r] = __fdiv_rn(sub_(t1, t0), static_cast<float>(p.substeps));
          run = valid ? 1 : 0;
        }
        rs.run[r] = run;
        if (r < NR / 4) rs.nsaved[r] = 0;
      }
      int any_running = __syncthreads_or(run);
      int loops = 0;
      bool have_k0 = false;

      while (any_running) {
        ++loops;
        if (tid == 0) H3_STAMP(0);
        bool y_stale = true;             // y in epilogue layout is rewritten by the first preparation pass of every iteration
        bool arg_ready = false;          // this stage's argument was written by the previous stage's last epilogue (H3Fuse)
        for (int st = (tb.fsal && have_k0) ? 1 : 0; st < ns; ++st) {
          // ---- stage argument -> activation image of the first Linear (own feature slice, all rows)
          H3_TL_SELECT(st);
          if (tid == 0) H3_STAMP(1);
          if (!arg_ready) {
            if (epi) {
              if (st == 0) h3_stage_input<0, NR>(sl, tb.a[0], rs.dt, xa0, kshift0);
              else H3_DISPATCH_STAGES(st, (h3_stage_input<(NSV < kMaxStages ? NSV : kMaxStages - 1), NR>(sl, tb.a[st], rs.dt, xa0, kshift0)))
              asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            if (tid == 0) H3_STAMP(2);
            __syncwarp();
            h3_cluster_sync();
#if H3_SKIP_LAST_BARRIER
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
          }
          if (tid == 0) H3_STAMP(3);
          // the last Linear reads xa1 when NL is even: xa0 is free for the next stage's argument
          const bool fuse_next = H3_FUSE_STAGE_ARG && st + 1 < ns && (NL & 1) == 0;
          // scratch in epilogue layout, the CTA's slice of two stage-vector slots that are dead during the stages: Y1 (written by
          // the error pass) for S, K[ns-1] (written by the last stage's last epilogue, after the last fused argument) for y
          float* const el_S = st_base + (kMaxStages + 1) * arr + static_cast<size_t>(own_f0) * NR;
          float* const el_Y = st_base + static_cast<size_t>(ns - 1) * arr + static_cast<size_t>(own_f0) * NR;
          const float* const coef_next = tb.a[st + 1 < kMaxStages ? st + 1 : 0];
          H3Fuse fz;
          fz.S = el_S; fz.Yel = el_Y; fz.own_nf = own_nf; fz.coef_last = coef_next[st]; fz.nprev = st;
          fz.dt_rows = rs.dt; fz.xa = xa0; fz.kshift = kshift0;
          H3Pre pre;
          pre.base = sl.base; pre.arr = sl.arr; pre.own_f0 = sl.own_f0; pre.warp = sl.warp; pre.fs = sl.fs; pre.g = sl.g; pre.own_nf = own_nf;
          pre.coef = coef_next; pre.nprev = st; pre.S = el_S; pre.Yel = el_Y; pre.write_y = y_stale ? 1 : 0;
          // ---- ODEFunc on the tensor cores; last Linear (+ Tanh) -> K[st]
          for (int l = 0; l < NL; ++l) {
            const bool last = l == NL - 1;
            const H3Layer* next = !last ? &p.lay[l + 1] : (st + 1 < ns ? &p.lay[0] : nullptr);
            H3Call cl;
            cl.xsrc = (l & 1) ? xa1 : xa0; cl.col0 = 0; cl.ncol = NR;
            cl.out = last ? st_base + static_cast<size_t>(st) * arr : nullptr;
            cl.xdst = last ? nullptr : ((l & 1) ? xa0 : xa1);
            cl.xdst_kshift = last ? 0 : h3_log2(p.lay[l + 1].KCH); cl.xdst_colshift = 0;
            // the preparation pass of the next argument is spread over the MMA shadows of the Linears before the last one
            pre.it0 = l * sl.nit / (NL - 1); pre.it1 = (l + 1) * sl.nit / (NL - 1);
            // The barrier after the ODEFunc's last Linear is dropped (H3_SKIP_LAST_BARRIER): what follows -- the next stage's
            // argument pass, or the error / commit pass -- reads and writes only this CTA's feature slice of the stage vectors
            // and ends with a cluster barrier of its own; the argument image xa0 it overwrites was last read by the Linear
            // before the last one (the last Linear reads xa1 when NL is even), i.e. before every CTA's previous barrier.
            const bool skip_end = H3_SKIP_LAST_BARRIER && last && (NL & 1) == 0 && !fuse_next;
            h3_layer<NR>(c, p.lay[l], next, cl, l, (last && fuse_next) ? &fz : nullptr, (!last && fuse_next) ? &pre : nullptr, !skip_end);
            // the passes walk the slice with another thread mapping than the epilogue: CTA-level barrier of the epilogue warps
            if (skip_end && epi) named_bar_sync(1, H3_EPI_THREADS);
          }
          arg_ready = fuse_next;
          if (fuse_next) y_stale = false;
          if (tid == 0) H3_STAMP(4);
        }
        H3_TL_SELECT(H3_TL_STAGE);
        have_k0 = true;

        if (p.adaptive) {
          // ---- y1, embedded error, this CTA's share of the per-row error norm
          if (epi) {
            float sum[NR / 32][4];
            H3_DISPATCH_STAGES(ns, (h3_error_pass<NSV, NR>(sl, tb, rs.dt, p.atol, p.rtol, sum)))
#pragma unroll
            for (int m = 0; m < NR / 32; ++m)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float v = sum[m][i];
                v = add_(v, __shfl_xor_sync(0xffffffffu, v, 8));
                v = add_(v, __shfl_xor_sync(0xffffffffu, v, 16));
                if (sl.fs == 0) rs.psum[warp][32 * m + 4 * sl.g + i] = v;
              }
            if (tid == 0) H3_STAMP(5);
            named_bar_sync(1, H3_EPI_THREADS);
            if (tid < NR) {
              float tot = rs.psum[0][tid];
#pragma unroll
              for (int w = 1; w < H3_EPI_WARPS; ++w) tot = add_(tot, rs.psum[w][tid]);
              __stcg(normpart + static_cast<size_t>(crank) * NR + tid, tot);
            }
          }
          __syncwarp();
          if (tid == 0) H3_STAMP(6);
          h3_cluster_sync();
          if (tid == 0) H3_STAMP(7);
          // ---- per-row controller, identical in every CTA (odernn_fwd.cu:controller; torchode IntegralController)
          run = 0;
          if (tid < NR) {
            const int r = tid;
            run = rs.run[r];
            const float dt = rs.dt[r];
            float t = rs.t[r];
            const float tend = rs.tend[r];
            bool accept = true, finite = true;
            float dt_next = dt;
            if (tb.has_err) {
              float total = 0.f;
#pragma unroll
              for (int k = 0; k < H3_NC; ++k) total = add_(total, __ldcg(normpart + k * NR + r));
              const float ratio = sqrtf(__fdiv_rn(total, static_cast<float>(D)));
              finite = isfinite(ratio);
              accept = p.accept_strict ? (ratio < 1.0f) : (ratio <= 1.0f);
              float factor = mul_(p.safety, powf(ratio, tb.exponent));
              factor = fminf(fmaxf(factor, p.fmin), p.fmax);
              if (p.floor_factor && accept) factor = fmaxf(factor, 1.0f);
              dt_next = mul_(dt, factor);
            }
            const int upd = (accept && run) ? 1 : 0;
            rs.nsteps[r] += run;
            rs.nacc[r] += upd;
            const bool lands = p.exact_landing && dt >= sub_(tend, t);
            const float t_new = upd ? (lands ? tend : add_(t, dt)) : t;
            // literal torchode end point (cfg.endpoint_dense): the state at t_end is the dense output of the step that
            // reached it (odernn_fwd.cu: controller / commit_pass)
            const int toeval = (upd && t_new >= tend && rs.noteval[r]) ? 1 : 0;
            if (toeval) { rs.x[r] = __fdiv_rn(sub_(tend, t), dt); rs.noteval[r] = 0; }
            rs.toeval[r] = toeval && p.endpoint_dense;
            t = t_new;
            rs.upd[r] = upd;
            if (run && !finite) rs.status[r] = max(rs.status[r], 2);
            run = (run && t < tend && finite) ? 1 : 0;
            if (run && loops >= p.max_steps) { rs.status[r] = max(rs.status[r], 1); run = 0; }
            float dtn = run ? dt_next : dt;
            dtn = fminf(fmaxf(dtn, sub_(rs.tmin[r], t)), sub_(rs.tmax[r], t));
            rs.t[r] = t;
            rs.dtstep[r] = dt;
            rs.dt[r] = dtn;
            rs.run[r] = run;
          }
          any_running = __syncthreads_or(run);
          if (p.ckpt) h3_ckpt_iteration<NR>(p, rs, tile, interval, Yc, own_f0, own_nf, crank);
          if (tid == 0) H3_STAMP(8);
          // ---- commit: accepted rows take y1; FSAL carry (end point rule "y1": exact landing makes y1 the value at t_end)
          if (epi) h3_commit_rows<NR>(sl, tb, rs.upd, rs.toeval, rs.x, rs.dtstep);
          if (tid == 0) H3_STAMP(9);
        } else {
          if (p.ckpt) {
            if (tid < NR) { rs.dtstep[tid] = rs.dt[tid]; rs.upd[tid] = rs.grow[tid] >= 0 ? 1 : 0; }
            __syncthreads();
            h3_ckpt_iteration<NR>(p, rs, tile, interval, Yc, own_f0, own_nf, crank);
          }
          if (epi) {
            H3_DISPATCH_STAGES(ns, (h3_fixed_commit<NSV, NR>(sl, tb, rs.dt)))
            if (tid < NR && rs.grow[tid] >= 0) { rs.nsteps[tid] += 1; rs.nacc[tid] += 1; }
          }
          any_running = loops < p.substeps;
        }
        // the elementwise passes of the next iteration read what this one wrote with a different thread mapping only
        // within the epilogue warps of this CTA
        if (epi) named_bar_sync(1, H3_EPI_THREADS);
      }

      // ---- stats / status of the interval
      __syncthreads();
      if (crank == 0 && tid < NR && rs.grow[tid] >= 0) {
        if (p.stats) {
          int* sp = p.stats + ((static_cast<size_t>(interval) * p.L + rs.lyr[tid]) * p.B + rs.bidx[tid]) * 2;
          sp[0] = rs.nsteps[tid]; sp[1] = rs.nacc[tid];
        }
        if (p.status && rs.status[tid]) atomicMax(p.status + rs.bidx[tid], rs.status[tid]);
      }

      if (p.ckpt) {
        // state at the end of the interval's solves + the number of stored iterations of every sub-tile
        if (epi) h3_ckpt_copy<NR>(p, tile, interval, Yc, own_f0, own_nf, tid, 0xffffffffu, nullptr, 0);
        const int nsub = p.SPT / p.RTf;
        if (crank == 0 && tid < nsub && tile * nsub + tid < p.ntiles_f)
          p.nloops[static_cast<size_t>(tile * nsub + tid) * p.S_total + interval] = rs.nsaved[tid];
      }
      if (p.do_jump) {
        // ---- rnn jump at the observation (PoseODERNN.py:112-117) and pose head (:119-122), all on this cluster
        // The jump-input images alias the stage vectors K0 .. K(2L-1) of ALL four feature slices: every CTA must be through
        // its commit pass (which reads and rewrites its slice of K0 for the FSAL carry) before any CTA writes them.
#if H3_XJ_ALIAS
        __syncwarp();
        h3_cluster_sync();
#endif
        if (epi) {
          h3_jump_input<NR>(p, rs, Yc, xj, own_f0, own_nf, interval, tid);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        __syncwarp();
        h3_cluster_sync();
        const int kshj = h3_log2(p.jump[0].KCH), kshh = h3_log2(p.head.KCH);
        const int ncolj = p.SPT;                 // SPT = NR / L: each layer's rows are a block of SPT tile rows
        for (int l = 0; l < p.L; ++l) {
          const bool top = l == p.L - 1;
          H3Call cl;
          cl.xsrc = xj + static_cast<size_t>(l) * p.xj_buf_bytes; cl.col0 = l * ncolj; cl.ncol = ncolj;
          cl.out = Yc;                            // the new hidden state of the layer's rows
          // ... which is also the input x of the next layer's jump (same sequences, SPT rows further) / of the pose head
          cl.xdst = top ? xa0 : xj + static_cast<size_t>(l + 1) * p.xj_buf_bytes;
          cl.xdst_kshift = top ? kshh : kshj; cl.xdst_colshift = top ? 0 : ncolj;
          h3_layer<NR>(c, p.jump[l], top ? &p.head : &p.jump[l + 1], cl, 0);
        }
        if (p.ckpt && epi) h3_ckpt_copy<NR>(p, tile, interval, Yc, own_f0, own_nf, tid, 0xffffffffu, nullptr, 1);      // post-jump state
        {
          H3Call cl;
          cl.xsrc = xa0; cl.col0 = (p.L - 1) * ncolj; cl.ncol = ncolj;
          cl.out = headbuf; cl.xdst = nullptr; cl.xdst_kshift = 0; cl.xdst_colshift = 0;
          h3_layer<NR>(c, p.head, nullptr, cl, 0);
        }
        // regressor.2 (128 -> 6) on the CUDA cores: thread = (pose component, sequence of the tile)
        if (crank == 0) {
          for (int idx = tid; idx < kPoseDim * ncolj; idx += H3_THREADS) {
            const int o = idx / ncolj, j = idx - o * ncolj, n = (p.L - 1) * ncolj + j;
            if (rs.grow[n] < 0) continue;
            float acc = 0.f;
            for (int m = 0; m < kRegHidden; ++m) acc = fmaf(__ldg(p.reg_w1 + o * kRegHidden + m), __ldcg(headbuf + static_cast<size_t>(m) * NR + n), acc);
            p.pose[(static_cast<size_t>(rs.bidx[n]) * p.S_io + interval) * kPoseDim + o] = add_(acc, __ldg(p.reg_b1 + o));
          }
        }
      }
    }

    // ---- final state back to [L][B][D]
    __syncthreads();
    if (epi) {
      const int nf4 = own_nf / 4;
      for (int item = tid; item < NR * nf4; item += H3_EPI_THREADS) {
        const int n = item / nf4, f4 = item - n * nf4;
        const long long gr = rs.grow[n];
        if (gr < 0) continue;
        const float* src = Yc + static_cast<size_t>(own_f0 + 4 * f4) * NR + n;
        const float4 v = make_float4(__ldcg(src), __ldcg(src + NR), __ldcg(src + 2 * NR), __ldcg(src + 3 * NR));
        *reinterpret_cast<float4*>(p.hT + static_cast<size_t>(gr) * D + own_f0 + 4 * f4) = v;
      }
    }
    __syncthreads();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  h3_cluster_sync();
  if (warp == H3_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(512u) : "memory");
  }
}

// W [N][K1] (and W2 [N][K2], the columns K1 .. K1 + K2 of the concatenated operand; PyTorch layout) -> per-CTA fp16
// hi / lo operand images [cta][k-chunk][tile][hi | lo][feature / 8][k / 8][feature % 8][k % 8]; one thread per 16-byte piece
// (8 k of one feature; K1 is a multiple of 8)
__global__ void h3_pack_weight_kernel(const float* __restrict__ W, int K1, const float* __restrict__ W2, int K2, int nctas, int Fc,
                                      int T, int KCH, unsigned char* __restrict__ dst) {
  const int nch = (K1 + K2) / KCH, k8n = KCH / 8;
  const size_t total = static_cast<size_t>(nctas) * nch * T * 128 * k8n;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int k8 = static_cast<int>(r % k8n); r /= k8n;
    const int m = static_cast<int>(r % 128); r /= 128;
    const int t = static_cast<int>(r % T); r /= T;
    const int kc = static_cast<int>(r % nch); r /= nch;
    const int cta = static_cast<int>(r);
    const int f = cta * Fc + (t == 0 ? 0 : Fc - 128) + m;
    const int k = kc * KCH + 8 * k8;
    const float* src = k < K1 ? W + static_cast<size_t>(f) * K1 + k : W2 + static_cast<size_t>(f) * K2 + (k - K1);
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h3_split(src[e], hi[e], lo[e]);
    const size_t tile_bytes = static_cast<size_t>(128) * KCH * 2;
    unsigned char* o = dst + (static_cast<size_t>(cta) * nch + kc) * H3_WCHUNK + static_cast<size_t>(t) * 2 * tile_bytes +
                       static_cast<size_t>(m >> 3) * (k8n * 128) + static_cast<size_t>(k8) * 128 + static_cast<size_t>(m & 7) * 16;
    *reinterpret_cast<uint4*>(o) = make_uint4(h3_pack2(hi[0], hi[1]), h3_pack2(hi[2], hi[3]), h3_pack2(hi[4], hi[5]), h3_pack2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(o + tile_bytes) =
        make_uint4(h3_pack2(lo[0], lo[1]), h3_pack2(lo[2], lo[3]), h3_pack2(lo[4], lo[5]), h3_pack2(lo[6], lo[7]));
  }
}

// ------------------------------------------------------------------------------------------ host side
struct H3LayerPlan { int K, N, Fc, T, KCH, nch, nseg, nctas; size_t off_w; };

struct H3Plan {
  int NL, NR, SPT, ntiles, nclusters;
  bool can_jump;                         // the rnn jump + pose head can run inside the cluster kernel
  H3LayerPlan lay[H3_MAXL], jump[2], head;
  size_t off_xa, xa_buf_bytes, off_xj, xj_buf_bytes, off_state, state_floats, total_bytes, smem_bytes;
};

// geometry of one Linear on the cluster: `nctas` CTAs x Fc features, MMAs over `ncol` tile rows
int h3_plan_layer(int K, int N, int nctas, int ncol, H3LayerPlan& y) {
  if (N % nctas) return ODEVIO_E_SHAPE;
  const int Fc = N / nctas;
  if (Fc < 128 || Fc > 256 || Fc % 32) return ODEVIO_E_SHAPE;
  const int T = Fc > 128 ? 2 : 1, KCH = 64 / T;
  if (K % 64) return ODEVIO_E_SHAPE;
  y.K = K; y.N = N; y.Fc = Fc; y.T = T; y.KCH = KCH; y.nch = K / KCH; y.nctas = nctas;
  // the accumulate of tcgen05.mma truncates: at most ~12 k-steps of 16 go into one accumulator
  int nseg = (K / 16 + 11) / 12;
  const int fit = 512 / (T * ncol) - 2;                     // + one cross-term accumulator per MMA issuer
  if (nseg < 2) nseg = 2;                                   // each issuer needs a main accumulator of its own
  if (nseg > fit) nseg = fit;
  if (nseg > y.nch) nseg = y.nch;
  if (nseg < 2) return ODEVIO_E_SHAPE;
  y.nseg = nseg;
  return 0;
}

int h3_plan(const odevio_odernn_cfg& c, H3Plan& pl) {
  const long long M = static_cast<long long>(c.L) * c.B;
  if (M <= 0 || M > 0x7fffffffLL || c.n_hidden < 1 || c.n_hidden + 1 > H3_MAXL || c.L < 1) return ODEVIO_E_SHAPE;
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) {
    cudaGetLastError();
    nsm = 148;
  }
  pl.NL = c.n_hidden + 1;
  pl.NR = 64;
  pl.SPT = (pl.NR % c.L == 0 && (pl.NR / c.L) % 8 == 0) ? pl.NR / c.L : 0;
  pl.ntiles = pl.SPT ? (c.B + pl.SPT - 1) / pl.SPT : static_cast<int>((M + pl.NR - 1) / pl.NR);
  pl.nclusters = nsm / H3_NC;
  if (pl.nclusters > pl.ntiles) pl.nclusters = pl.ntiles;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (off + n + 1023) / 1024 * 1024; return o; };
  int kmax = 0;
  for (int l = 0; l < pl.NL; ++l) {
    const int K = l == 0 ? c.D : c.H, N = l == pl.NL - 1 ? c.D : c.H;
    const int rc = h3_plan_layer(K, N, H3_NC, pl.NR, pl.lay[l]);
    if (rc != 0) return rc;
    pl.lay[l].off_w = take(static_cast<size_t>(H3_NC) * pl.lay[l].nch * H3_WCHUNK);
    if (K > kmax) kmax = K;
  }
  if ((c.D / H3_NC) % 32 || c.D % 64) return ODEVIO_E_SHAPE;
  // jump + head inside the kernel: tanh rnn, 1 or 2 layers (MMAs over SPT = 64 / L rows: a multiple of 32)
  pl.can_jump = !c.evolve_only && c.rnn_type == ODEVIO_RNN_TANH && pl.SPT >= 32;
  if (pl.can_jump) {
    for (int l = 0; l < c.L; ++l) {
      if (h3_plan_layer(2 * c.D, c.D, H3_NC, pl.SPT, pl.jump[l]) != 0) { pl.can_jump = false; break; }
      pl.jump[l].off_w = take(static_cast<size_t>(H3_NC) * pl.jump[l].nch * H3_WCHUNK);
    }
  }
  if (pl.can_jump) {
    if (h3_plan_layer(c.D, kRegHidden, 1, pl.SPT, pl.head) != 0) pl.can_jump = false;
    else pl.head.off_w = take(static_cast<size_t>(pl.head.nch) * H3_WCHUNK);
  }
  pl.xa_buf_bytes = (static_cast<size_t>(kmax) * pl.NR * 4 + 1023) / 1024 * 1024;       // hi + lo fp16 images of kmax x NR
  pl.off_xa = take(static_cast<size_t>(pl.nclusters) * 2 * pl.xa_buf_bytes);
  pl.xj_buf_bytes = pl.can_jump ? static_cast<size_t>(2 * c.D) * pl.NR * 4 : 0;        // aliases 2 stage vectors per layer (kernel)
  if (pl.can_jump && 2 * c.L > kMaxStages) pl.can_jump = false;
#if H3_XJ_ALIAS
  pl.off_xj = off;
#else
  pl.off_xj = take(static_cast<size_t>(pl.nclusters) * c.L * pl.xj_buf_bytes);
#endif
  pl.state_floats = (static_cast<size_t>(kMaxStages + 2) * c.D + H3_NC + kRegHidden) * pl.NR;
  pl.state_floats = (pl.state_floats + 255) / 256 * 256;
  pl.off_state = take(static_cast<size_t>(pl.nclusters) * pl.state_floats * sizeof(float));
  pl.total_bytes = off;
  pl.smem_bytes = static_cast<size_t>(H3_NST) * H3_WCHUNK + static_cast<size_t>(H3_NST) * 4 * 64 * pl.NR + 1024;
  return 0;
}

typedef void (*H3Kernel)(H3Params);
H3Kernel h3_kernel_of(const H3Plan&) { return static_cast<H3Kernel>(odernn_h3_kernel<64>); }

cudaError_t h3_launch_config(const H3Plan& pl, cudaLaunchConfig_t& lc, cudaLaunchAttribute& at, int nclusters, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(h3_kernel_of(pl), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem_bytes));
  if (e != cudaSuccess) return e;
  memset(&lc, 0, sizeof(lc));
  lc.blockDim = dim3(H3_THREADS); lc.dynamicSmemBytes = pl.smem_bytes; lc.stream = stream;
  at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = H3_NC; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  lc.attrs = &at; lc.numAttrs = 1;
  lc.gridDim = dim3(nclusters * H3_NC);
  return cudaSuccess;
}

int g_h3_last_clusters = 0, g_h3_last_max_clusters = 0, g_h3_last_rows = 0;
constexpr int kH3TimingSlots = 512;
bool g_h3_timing = false;
int g_h3_timing_n = 0;
cudaEvent_t g_h3_ev[kH3TimingSlots][2];
bool g_h3_ev_made = false;

}  // namespace

struct H3Evolve::Impl {
  H3Plan pl;
  H3Params prm;
  int maxc = 0;
};

size_t odernn_h3_workspace_bytes(const odevio_odernn_cfg& c) {
  H3Plan pl;
  if (h3_plan(c, pl) != 0) return 0;
  return pl.total_bytes;
}

void odernn_h3_last_geometry(int* clusters, int* max_clusters, int* rows) {
  *clusters = g_h3_last_clusters; *max_clusters = g_h3_last_max_clusters; *rows = g_h3_last_rows;
}
int odernn_h3_debug_timeline(long long* host_dst) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_dst, g_h3_dbg, sizeof(long long) * 96));
}
void odernn_h3_timing_enable(bool on) {
  g_h3_timing = on; g_h3_timing_n = 0;
  if (on && !g_h3_ev_made) {
    for (int i = 0; i < kH3TimingSlots; ++i) { cudaEventCreate(&g_h3_ev[i][0]); cudaEventCreate(&g_h3_ev[i][1]); }
    g_h3_ev_made = true;
  }
}
int odernn_h3_timing_read(float* total_ms, int* launches) {
  float tot = 0.f;
  for (int i = 0; i < g_h3_timing_n; ++i) {
    cudaError_t e = cudaEventSynchronize(g_h3_ev[i][1]);
    if (e != cudaSuccess) return static_cast<int>(e);
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_h3_ev[i][0], g_h3_ev[i][1]);
    if (e != cudaSuccess) return static_cast<int>(e);
    tot += ms;
  }
  *total_ms = tot; *launches = g_h3_timing_n;
  g_h3_timing_n = 0;
  return 0;
}

H3Evolve::H3Evolve() : impl(nullptr) {}
H3Evolve::~H3Evolve() { delete impl; }

namespace {
void h3_fill_layer(H3Layer& y, const H3LayerPlan& lp, int act, const unsigned char* ws, const float* bias, const float* bias2) {
  y.K = lp.K; y.N = lp.N; y.Fc = lp.Fc; y.T = lp.T; y.KCH = lp.KCH; y.nch = lp.nch; y.nseg = lp.nseg; y.act = act;
  y.nctas = lp.nctas; y.Wimg = ws + lp.off_w; y.bias = bias; y.bias2 = bias2;
}
cudaError_t h3_pack(const float* W, int K1, const float* W2, int K2, const H3LayerPlan& lp, unsigned char* ws, cudaStream_t stream) {
  h3_pack_weight_kernel<<<296, 256, 0, stream>>>(W, K1, W2, K2, lp.nctas, lp.Fc, lp.T, lp.KCH, ws + lp.off_w);
  return cudaGetLastError();
}
}  // namespace

bool odernn_h3_can_fuse_jump(const odevio_odernn_cfg& c) {
  H3Plan pl;
  return h3_plan(c, pl) == 0 && pl.can_jump;
}

int H3Evolve::prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const odevio_odernn_weights* w, bool with_jump,
                      bool pack_weights, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  delete impl;
  impl = new Impl();
  H3Plan& pl = impl->pl;
  const int rc = h3_plan(c, pl);
  if (rc != 0) return rc;
  if (with_jump && !pl.can_jump) return ODEVIO_E_SHAPE;
  if (workspace_bytes < pl.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  H3Params& p = impl->prm;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.L = c.L; p.D = c.D; p.NL = pl.NL; p.SPT = pl.SPT;
  for (int l = 0; l < pl.NL; ++l) {
    if (!w->ode_w[l] || !w->ode_b[l]) return ODEVIO_E_NULL;
    h3_fill_layer(p.lay[l], pl.lay[l], l == pl.NL - 1 ? ACT_TANH : c.activation, ws, w->ode_b[l], nullptr);
    if (pack_weights) {
      const cudaError_t e = h3_pack(w->ode_w[l], pl.lay[l].K, nullptr, 0, pl.lay[l], ws, stream);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
  }
  if (with_jump) {
    for (int l = 0; l < c.L; ++l) {
      if (!w->rnn_w_ih[l] || !w->rnn_w_hh[l] || !w->rnn_b_ih[l] || !w->rnn_b_hh[l]) return ODEVIO_E_NULL;
      h3_fill_layer(p.jump[l], pl.jump[l], ACT_TANH, ws, w->rnn_b_ih[l], w->rnn_b_hh[l]);
      if (pack_weights) {
        const cudaError_t e = h3_pack(w->rnn_w_ih[l], c.D, w->rnn_w_hh[l], c.D, pl.jump[l], ws, stream);
        if (e != cudaSuccess) return static_cast<int>(e);
      }
    }
    if (!w->reg_w0 || !w->reg_b0 || !w->reg_w1 || !w->reg_b1) return ODEVIO_E_NULL;
    h3_fill_layer(p.head, pl.head, ACT_LEAKY01, ws, w->reg_b0, nullptr);
    if (pack_weights) {
      const cudaError_t e = h3_pack(w->reg_w0, c.D, nullptr, 0, pl.head, ws, stream);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    p.reg_w1 = w->reg_w1; p.reg_b1 = w->reg_b1;
  }
  p.xa = ws + pl.off_xa; p.xa_buf_bytes = pl.xa_buf_bytes;
  p.xj = ws + pl.off_xj; p.xj_buf_bytes = pl.xj_buf_bytes;
  p.state = reinterpret_cast<float*>(ws + pl.off_state); p.state_floats = pl.state_floats;
  p.ntiles = pl.ntiles;
  p.tab = tab; p.adaptive = adaptive ? 1 : 0; p.substeps = c.substeps;
  p.atol = c.atol; p.rtol = c.rtol; p.dt0 = c.dt0; p.safety = c.safety; p.fmin = c.factor_min; p.fmax = c.factor_max;
  p.accept_strict = c.accept_strict; p.floor_factor = c.floor_factor; p.max_steps = c.max_steps; p.exact_landing = c.exact_landing; p.endpoint_dense = c.endpoint_dense;
  p.do_jump = with_jump ? 1 : 0;
  return 0;
}

int H3Evolve::max_clusters() {
  if (!impl) return 0;
  if (impl->maxc > 0) return impl->maxc;
  cudaLaunchConfig_t lc; cudaLaunchAttribute at;
  if (h3_launch_config(impl->pl, lc, at, impl->pl.nclusters, nullptr) != cudaSuccess) { cudaGetLastError(); return impl->pl.nclusters; }
  int maxc = 0;
  if (cudaOccupancyMaxActiveClusters(&maxc, h3_kernel_of(impl->pl), &lc) != cudaSuccess || maxc <= 0) {
    cudaGetLastError();
    maxc = impl->pl.nclusters;
  }
  if (maxc > impl->pl.nclusters) maxc = impl->pl.nclusters;       // scratch is sized for pl.nclusters
  impl->maxc = maxc;
  return maxc;
}

void H3Evolve::set_checkpoints(float* ckpt, int* nloops, size_t ckpt_floats_per_tile, int CK, int RTf, int ntiles_f, int S_total) {
  if (!impl) return;
  H3Params& p = impl->prm;
  p.ckpt = ckpt; p.nloops = nloops; p.ckpt_floats_per_tile = ckpt_floats_per_tile; p.CK = CK; p.RTf = RTf; p.ntiles_f = ntiles_f;
  p.S_total = S_total;
}

int H3Evolve::run(const float* h0, float* hT, const float* ts, int ts_ld, int interval0, int n_intervals, const float* fv,
                  const float* fi, int Dv, int S_io, float* pose, int* stats, int* status, cudaStream_t stream) {
  if (!impl) return ODEVIO_E_NULL;
  H3Params p = impl->prm;
  H3Plan& pl = impl->pl;
  if (!hT || !ts || n_intervals < 1) return ODEVIO_E_NULL;
  if (p.do_jump && (!fv || !pose || Dv <= 0 || Dv > p.D || (Dv < p.D && !fi))) return ODEVIO_E_NULL;
  if (p.ckpt && (p.SPT == 0 || p.RTf < 4 || p.RTf % 4 || p.SPT % p.RTf || p.SPT / p.RTf > 64 / 4)) return ODEVIO_E_SHAPE;
  p.h0 = h0; p.hT = hT; p.ts = ts; p.ts_ld = ts_ld; p.interval0 = interval0; p.nI = n_intervals;
  p.fv = fv; p.fi = fi; p.Dv = Dv; p.S_io = S_io; p.pose = pose; p.stats = stats; p.status = status;
  int nclusters = max_clusters();
  if (nclusters > p.ntiles) nclusters = p.ntiles;
  cudaLaunchConfig_t lc; cudaLaunchAttribute at;
  cudaError_t e = h3_launch_config(pl, lc, at, nclusters, stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  g_h3_last_clusters = nclusters; g_h3_last_max_clusters = impl->maxc; g_h3_last_rows = p.L * p.B;
  const bool timed = g_h3_timing && g_h3_timing_n < kH3TimingSlots;
  if (timed) cudaEventRecord(g_h3_ev[g_h3_timing_n][0], stream);
  e = cudaLaunchKernelEx(&lc, h3_kernel_of(pl), p);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (timed) { cudaEventRecord(g_h3_ev[g_h3_timing_n][1], stream); ++g_h3_timing_n; }
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace odevio
