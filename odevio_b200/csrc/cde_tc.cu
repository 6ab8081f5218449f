// Neural-CDE forward on the tensor cores (cfg.precision = ODEVIO_PRECISION_FP16X3): the same solve as cde_fwd.cu --
//   PoseCDE.forward                                  src/models/PoseCDE.py:76-103
//   torchcde cdeint -> torchdiffeq dopri5 | rk4      (:94-101; semantics: oracle/torchdiffeq_like.py)
//   CDEFunc.forward                                  src/models/ODEFunc.py:81-84
// -- with the one dense contraction of the path, CDEFunc's final Linear Hc -> Hc x (Hc + 1) (ODEFunc.py:55; 4.3 GFLOP per
// vector-field evaluation at B = 1024, Hc = 128), on tcgen05 as 3xFP16 (x = hi + lo 2^-11, products hi.hi + (lo.hi + hi.lo) 2^-11:
// relative product error <= 2^-22, fp32 accumulation in TMEM; same scheme as odernn_h3.cu).
//
// WEIGHT-STATIONARY: CTA h (one per hidden unit, a cooperative grid of Hc CTAs) keeps the Hc x Hc block W[h, 1.., :] of the
// final Linear -- the weights of the Hc value channels of its hidden unit, 64 KB as fp16 hi / lo images at Hc = 128 -- in
// shared memory for the WHOLE solve; the FFMA kernel re-streams the 8.3 MB weight per 8-row tile and evaluation.  Every
// vector-field evaluation has two phases separated by grid-wide barriers:
//   row phase      CTA r owns rows [r RP, (r + 1) RP): stage argument z (Butcher combination in the oracle's operation
//                  order), the n Hc x Hc Linears on CUDA cores (same sequential-k FMA order as tile_gemm.cuh), the time
//                  channel tanh(w_{h,0} . a + b) dX_0/dt -> K_out[h][row], the last activation a as the fp16 hi / lo
//                  K-major canonical image of its row tile, and dX/dt of the value channels [row][c];
//   feature phase  CTA h, for every 128-row tile: one bulk TMA copy of the tile's image (64 KB), Hc / 16 k-steps x 3
//                  MMAs (M = 128 rows, N = Hc channels) into a double-buffered TMEM accumulator, epilogue thread = row:
//                  sum_c tanh(acc + b[h,c]) dX_c/dt in registers -- the [B, Hc, C] tensor never exists anywhere --
//                  added to the time channel's term in K_out[h][row].
// Segments of the rectilinear path on which only the time channel moves skip the feature phase altogether.  The solver
// logic (batch-joint controller in fp64, Hairer initial step, knot landings + re-evaluation, dense output, pose head)
// is cde_fwd.cu's; state vectors Z, Y1, K0..K6 are feature-major [Hc][Bpad] fp32 in L2-resident global memory, FSAL and
// commit are index rotations.
#include <cuda_fp16.h>

#include "cde_params.h"
#include "common.cuh"

namespace odevio {

namespace {

constexpr int TC_WORK_THREADS = 256;      // warps 0-7: row phase + epilogue
constexpr int TC_THREADS = 320;           // + warp 8 (TMA producer) + warp 9 (MMA issuer)
constexpr int TC_WARP_PROD = 8, TC_WARP_MMA = 9;

// development timeline: clock64 sums of CTA 0 over all evaluations of the last launch (slot meanings: tools/cde_tc_timeline.py)
__device__ long long g_cde_tc_dbg[32];
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(slot, t0) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { \
    if (threadIdx.x == (slot >= 16 ? ((slot >= 24) ? 288 : 256) : 0)) g_cde_tc_dbg[slot] += clock64() - (t0); } } while (0)

__constant__ float kDpCt[7] = {0.0f, 0.2f, 0.3f, 0.8f, static_cast<float>(8.0 / 9.0), 1.0f, 1.0f};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t sbo_bytes) {
  // no-swizzle K-major canonical layout [m/8][k/8][m%8][k%8]: LBO = 128 B (k groups of 8), SBO = (K/8) * 128 B, version 1
  const uint32_t hi_word = (sbo_bytes >> 4) | (1u << 14);
  return (static_cast<uint64_t>(hi_word) << 32) | static_cast<uint64_t>(((saddr >> 4) & 0x3fffu) | ((128u >> 4) << 16));
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&u)[32]) {
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
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_split(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}
__device__ __forceinline__ uint32_t tc_pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

struct TCtx {
  const CdeParams* prm;
  int tid, warp, lane;
  unsigned char* wimg; unsigned char* xring;
  float* bufA; float* bufB; float* dX0; float* biasv; float* part; float* dxc;
  uint32_t oA, oB, oRing, oBias;        // byte offsets of bufA / bufB / xring / biasv from the dynamic shared-memory base
  double* redsm;
  uint64_t* full; uint64_t* empty; uint64_t* tfull; uint64_t* tempty; uint64_t* wbar; uint64_t* wfull; uint64_t* wfree;
  uint32_t tmem, tcount, wcount;
  unsigned int bar_target, red_count;
  int row0, RP;                         // rows of this CTA in the row phase: [row0, row0 + RP) (RP = 0: none)
  int ix[2 + kMaxStages];               // array slots: ix[0] = Z, ix[1] = Y1, ix[2 + j] = K_j   (rotated by commit)
};

__device__ __forceinline__ float* arr_ptr(const TCtx& c, int which) {
  return c.prm->state + static_cast<size_t>(c.ix[which]) * c.prm->Hc * c.prm->Bpad;
}

// ---- grid-wide barrier / sum of two doubles (deterministic: fixed tree in the block, CTA order across)
__device__ __forceinline__ void grid_wait(TCtx& c) {
  const CdeParams& p = *c.prm;
  if (c.tid == 0) {
    __threadfence();
    atomicAdd(p.bar, 1u);
    const unsigned int target = c.bar_target + gridDim.x;
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.bar) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 27)) __trap();       // a lost CTA must not hang the GPU
    }
    __threadfence();
  }
  c.bar_target += gridDim.x;
}
__device__ __forceinline__ void grid_sync(TCtx& c) {
  __syncthreads();
  grid_wait(c);
  __syncthreads();
}
__device__ __forceinline__ void grid_reduce2(TCtx& c, double a, double b, double& sa, double& sb) {
  const CdeParams& p = *c.prm;
  const int nwarps = TC_THREADS / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  if (c.lane == 0) { c.redsm[2 * c.warp] = a; c.redsm[2 * c.warp + 1] = b; }
  __syncthreads();
  const int slot = c.red_count & 1;
  double* mine = p.red + (static_cast<size_t>(slot) * gridDim.x + blockIdx.x) * 2;
  if (c.tid == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < nwarps; ++w) { ta += c.redsm[2 * w]; tb += c.redsm[2 * w + 1]; }
    __stcg(mine, ta); __stcg(mine + 1, tb);
  }
  grid_wait(c);
  if (c.tid == 0) {
    double ga = 0.0, gb = 0.0;
    const double* all = p.red + static_cast<size_t>(slot) * gridDim.x * 2;
    for (unsigned int g = 0; g < gridDim.x; ++g) { ga += __ldcg(all + 2 * g); gb += __ldcg(all + 2 * g + 1); }
    c.redsm[64] = ga; c.redsm[65] = gb;
  }
  c.red_count += 1;
  __syncthreads();
  sa = c.redsm[64]; sb = c.redsm[65];
  __syncthreads();
}

__device__ __forceinline__ int seg_index_t(float t, int nk) {
  int cnt = static_cast<int>(ceilf(t));
  cnt = max(0, min(cnt, nk));
  return max(0, min(cnt - 1, nk - 2));
}
__device__ __forceinline__ float obs_val_t(const CdeParams& p, int b, int o, int ch) {
  if (ch == 0) return p.tobs[static_cast<size_t>(b) * p.So + o];
  const int f = ch - 1;
  const size_t row = static_cast<size_t>(b) * p.So + o;
  return (f < p.Dv) ? p.fv[row * p.Dv + f] : p.fi[row * (p.Hc - p.Dv) + (f - p.Dv)];
}

// ---- row phase: out[n][r] = act(sum_k Wt[k][n] x[k][r] + bias[n]),  x / out shared [.][RP], Wt global K-major [K][N];
// N in {32, 64, 128}: thread = (output n, group of 4 rows); sequential-k FMA chain from 0, bias added last (tile_gemm order)
// Operands are addressed as byte offsets from the dynamic shared-memory base so that the compiler emits LDS / STS with
// 32-bit addresses (through generic pointers kept in a struct it fell back to LD.E with 64-bit address arithmetic per
// load: measured 7.5 k clk per Linear instead of ~2 k).
template <bool W_SHARED>
__device__ __forceinline__ void row_linear(const float* __restrict__ Wg, uint32_t w_off, const float* __restrict__ bias,
                                           int K, int N, uint32_t x_off, uint32_t out_off, int RP, int act, int tid) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const float* __restrict__ x = reinterpret_cast<const float*>(smem + x_off);
  const float* __restrict__ ws = reinterpret_cast<const float*>(smem + w_off);
  float* __restrict__ out = reinterpret_cast<float*>(smem + out_off);
  const int G = TC_WORK_THREADS / N;
  const int n = tid % N, rg = tid / N;
  for (int ci = rg; ci < (RP >> 2); ci += G) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* xp = x + 4 * ci;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float w = W_SHARED ? ws[k * N + n] : __ldg(Wg + static_cast<size_t>(k) * N + n);
      const float4 v = *reinterpret_cast<const float4*>(xp + k * RP);
      a0 = fmaf(v.x, w, a0); a1 = fmaf(v.y, w, a1); a2 = fmaf(v.z, w, a2); a3 = fmaf(v.w, w, a3);
    }
    const float b = bias ? bias[n] : 0.f;
    const float4 o4 = apply_act4(make_float4(a0 + b, a1 + b, a2 + b, a3 + b), act);
    *reinterpret_cast<float4*>(out + n * RP + 4 * ci) = o4;
  }
}

// recipes for the argument of a vector-field evaluation / an output value (cde_fwd.cu: build_vector), -> bufA [Hc][RP]
enum { RC_Z = 0, RC_HAIRER, RC_STAGE, RC_RK4_1, RC_RK4_2, RC_RK4_3, RC_INTERP, RC_Y1, RC_LERP };
struct Recipe { int kind; int stage; float dt_s; float x; };

__device__ __forceinline__ void build_rows(TCtx& c, const Recipe& rc, const DevTableau& tab, bool save_y1) {
  const CdeParams& p = *c.prm;
  const int rq4 = c.RP >> 2;
  const float dt = rc.dt_s;
  const float* Z = arr_ptr(c, 0);
  float* Y1 = arr_ptr(c, 1);
  for (int e = c.tid; e < p.Hc * rq4; e += TC_WORK_THREADS) {
    const int k = e / rq4, r4 = e - k * rq4;
    const size_t off = static_cast<size_t>(k) * p.Bpad + c.row0 + 4 * r4;
    const float4 y4 = ldcg4(Z + off);
    const float y[4] = {y4.x, y4.y, y4.z, y4.w};
    float out[4];
    float kk[kMaxStages][4];
    auto ldk = [&](int j) { const float4 v = ldcg4(arr_ptr(c, 2 + j) + off); kk[j][0] = v.x; kk[j][1] = v.y; kk[j][2] = v.z; kk[j][3] = v.w; };
    switch (rc.kind) {
      case RC_Z:
        for (int q = 0; q < 4; ++q) out[q] = y[q];
        break;
      case RC_HAIRER:
        ldk(0);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(dt, kk[0][q]));
        break;
      case RC_STAGE: {
        const int i = rc.stage;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        bool any = false;
        for (int j = 0; j < i; ++j) {
          const float a = tab.a[i][j];
          if (a == 0.f) continue;
          ldk(j);
          const float w = mul_(a, dt);
          for (int q = 0; q < 4; ++q) acc[q] = any ? add_(acc[q], mul_(kk[j][q], w)) : mul_(kk[j][q], w);
          any = true;
        }
        for (int q = 0; q < 4; ++q) out[q] = any ? add_(y[q], acc[q]) : y[q];
        break;
      }
      case RC_RK4_1:
        ldk(0);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(mul_(dt, kk[0][q]), static_cast<float>(1.0 / 3.0)));
        break;
      case RC_RK4_2:
        ldk(0); ldk(1);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(dt, sub_(kk[1][q], mul_(kk[0][q], static_cast<float>(1.0 / 3.0)))));
        break;
      case RC_RK4_3:
        ldk(0); ldk(1); ldk(2);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(dt, add_(sub_(kk[0][q], kk[1][q]), kk[2][q])));
        break;
      case RC_Y1: {
        const float4 v = ldcg4(Y1 + off);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        break;
      }
      case RC_LERP: {
        const float4 v = ldcg4(Y1 + off);
        const float y1[4] = {v.x, v.y, v.z, v.w};
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(rc.x, sub_(y1[q], y[q])));
        break;
      }
      default: {          // RC_INTERP: quartic dense output of the accepted step (oracle interp_fit / interp_evaluate)
        const float4 v = ldcg4(Y1 + off);
        const float y1[4] = {v.x, v.y, v.z, v.w};
        float ym[4] = {0.f, 0.f, 0.f, 0.f};
        bool any = false;
        for (int j = 0; j < tab.n_stages; ++j) {
          const float bm = tab.bmid[j];
          if (bm == 0.f) { if (j == 0 || j == tab.n_stages - 1) ldk(j); continue; }
          ldk(j);
          const float w = mul_(bm, dt);
          for (int q = 0; q < 4; ++q) ym[q] = any ? add_(ym[q], mul_(kk[j][q], w)) : mul_(kk[j][q], w);
          any = true;
        }
        const int last = tab.n_stages - 1;
        const float x = rc.x;
        for (int q = 0; q < 4; ++q) {
          const float f0 = kk[0][q], f1 = kk[last][q], y0 = y[q], ymid = add_(y0, ym[q]);
          const float a = add_(sub_(mul_(mul_(2.0f, dt), sub_(f1, f0)), mul_(8.0f, add_(y1[q], y0))), mul_(16.0f, ymid));
          const float b = sub_(add_(add_(mul_(dt, sub_(mul_(5.0f, f0), mul_(3.0f, f1))), mul_(18.0f, y0)),
                                    mul_(14.0f, y1[q])), mul_(32.0f, ymid));
          const float cc = add_(sub_(sub_(mul_(dt, sub_(f1, mul_(4.0f, f0))), mul_(11.0f, y0)), mul_(5.0f, y1[q])),
                                mul_(16.0f, ymid));
          const float d = mul_(dt, f0);
          float total = add_(y0, mul_(x, d));
          float xp = x;
          xp = mul_(xp, x); total = add_(total, mul_(xp, cc));
          xp = mul_(xp, x); total = add_(total, mul_(xp, b));
          xp = mul_(xp, x); total = add_(total, mul_(xp, a));
          out[q] = total;
        }
        break;
      }
    }
    const float4 o4 = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(c.bufA + k * c.RP + 4 * r4) = o4;
    if (save_y1) stcg4(Y1 + off, o4);
  }
  named_bar_sync(1, TC_WORK_THREADS);
}

enum { PC_INIT = 0, PC_POSE0, PC_F0, PC_HAIRER_A, PC_HAIRER_B, PC_STEP_BEGIN, PC_STAGE, PC_STEP_END,
       PC_OUTPUTS, PC_COMMIT, PC_AFTER_JUMP, PC_RK4_END, PC_END };

}  // namespace

__global__ void __launch_bounds__(TC_THREADS, 1)
cde_tc_kernel(const __grid_constant__ CdeParams prm, const __grid_constant__ DevTableau tab) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[13];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) double redsm[72];
  const CdeParams& p = prm;
  TCtx c;
  c.prm = &prm;
  c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
  const int Hc = p.Hc, S = p.S, Bpad = p.Bpad;
  const uint32_t wbytes = 4u * Hc * Hc;                  // hi | lo images of W[h, 1.., :]
  const uint32_t ximg = 256u * Hc;                        // one fp16 image of a 128-row tile
  const uint32_t xstage = 2u * ximg;
  c.wimg = smem;
  c.xring = smem + wbytes;
  float* fsm = reinterpret_cast<float*>(c.xring + 2 * xstage);
  const int rows_buf = max(max(Hc, p.Cpad), kRegHidden);
  c.bufA = fsm; fsm += static_cast<size_t>(rows_buf) * p.RP;
  c.bufB = fsm; fsm += static_cast<size_t>(rows_buf) * p.RP;
  c.dX0 = fsm; fsm += p.RP;
  c.biasv = fsm; fsm += Hc;
  c.part = fsm; fsm += 2 * 2 * 128;
  c.dxc = fsm;                                          // [2][RP][C] when p.dx_cache
  c.redsm = redsm;
  c.oA = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(c.bufA) - smem);
  c.oB = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(c.bufB) - smem);
  c.oRing = static_cast<uint32_t>(c.xring - smem);
  c.oBias = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(c.biasv) - smem);
  c.full = bars; c.empty = bars + 2; c.tfull = bars + 4; c.tempty = bars + 6; c.wbar = bars + 8; c.wfull = bars + 9; c.wfree = bars + 11;
  c.tcount = 0; c.wcount = 0; c.bar_target = 0; c.red_count = 0;
  c.row0 = static_cast<int>(blockIdx.x) * p.RP;
  c.RP = p.RP;
  const bool has_rows = c.row0 < Bpad;               // Bpad is a multiple of RP's granule: a CTA owns RP rows or none
  for (int j = 0; j < 2 + kMaxStages; ++j) c.ix[j] = j;

  if (c.tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&c.full[i], 1); mbar_init(&c.empty[i], 1); mbar_init(&c.tfull[i], 1); mbar_init(&c.tempty[i], 8); }
    mbar_init(c.wbar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&c.wfull[i], 1); mbar_init(&c.wfree[i], 8); }
    fence_barrier_init();
  }
  if (c.warp == TC_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  c.tmem = tmem_slot;
  const bool worker = c.tid < TC_WORK_THREADS;
  const int h_own = blockIdx.x;                        // hidden unit of the feature phase
  // resident weights of my hidden unit + their biases
  if (c.warp == TC_WARP_PROD && elect_one()) {
    mbar_arrive_expect_tx(c.wbar, wbytes);
    tma_load_1d(c.wimg, p.Wimg + static_cast<size_t>(h_own) * wbytes, wbytes, c.wbar);
  }
  if (worker) for (int i = c.tid; i < Hc; i += TC_WORK_THREADS) c.biasv[i] = p.bval[static_cast<size_t>(h_own) * Hc + i];
  __syncthreads();

  if (blockIdx.x == 0 && c.tid < 32) g_cde_tc_dbg[c.tid] = 0;
  __syncthreads();
  const long long t_kernel0 = clock64();
  const int nk = p.interp == CDE_INTERP_LINEAR ? 2 * p.So - 1 : p.So;
  const bool adaptive = p.solver == CDE_SOLVER_DOPRI5;
  const double elems = static_cast<double>(p.B) * Hc;
  const uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(Hc) >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  const uint32_t sbo = (static_cast<uint32_t>(Hc) >> 3) * 128u;

  // ---- solver scalars: identical in every thread of every CTA
  double t_cur = p.tout[0], dt = 0.0, step = 0.0, t_b = 0.0, h0 = 0.0, d1 = 0.0;
  float dt_s = 0.f, ta_s = 0.f, tb_s = 0.f;
  int on_jump = 0, next_jump = 0, i_out = 1, st = 0;
  int n_steps = 0, n_acc = 0, n_f = 0, status = 0;
  {
    long long cnt = static_cast<long long>(floor(t_cur)) + 1;
    if (cnt < 0) cnt = 0;
    if (cnt > nk) cnt = nk;
    next_jump = static_cast<int>(cnt < nk - 1 ? cnt : nk - 1);
  }
  int grid_n = 0, grid_i = 0;
  if (!adaptive) {
    if (p.step_size > 0.0) grid_n = static_cast<int>(ceil((p.tout[S - 1] - p.tout[0]) / p.step_size + 1.0));
    else grid_n = S;
  }
  auto grid_time = [&](int k) -> double {
    if (p.step_size > 0.0) return k == grid_n - 1 ? p.tout[S - 1] : p.tout[0] + p.step_size * k;
    return p.tout[k];
  };

  int cached_seg = -1;                                  // knot segment whose (m, d) are in shared memory
  // training: checkpoint of an accepted step -- Z, Y1, K0..K6 of my hidden unit for all sequences, in LOGICAL order
  // [9][Hc][Bpad] (the index rotations resolved) -- and its log entry (cde_bwd.cu reads both; same log as cde_fwd.cu)
  int vjp_total = 0;
  auto save_step = [&](int step_idx, int onj) {
    if (worker) {
      float* dst0 = p.ckpt + static_cast<size_t>(step_idx) * (2 + kMaxStages) * Hc * Bpad + static_cast<size_t>(h_own) * Bpad;
      for (int which = 0; which < 2 + kMaxStages; ++which) {
        const float* src = arr_ptr(c, which) + static_cast<size_t>(h_own) * Bpad;
        float* dst = dst0 + static_cast<size_t>(which) * Hc * Bpad;
        for (int e = c.tid; e < (Bpad >> 2); e += TC_WORK_THREADS) stcg4(dst + 4 * e, ldcg4(src + 4 * e));
      }
    }
    int cnt = 0;
    for (int i = i_out; i < S && !(p.tout[i] > t_b); ++i) ++cnt;
    if (blockIdx.x == 0 && c.tid == 0) {
      CdeStepRec r;
      r.ta = t_cur; r.tb = t_b; r.dt_s = dt_s; r.ta_s = ta_s; r.tb_s = tb_s; r.on_jump = onj;
      r.out_first = i_out; r.out_count = cnt; r.vjp_base = vjp_total; r.pad = 0;
      p.log[1 + step_idx] = r;
    }
    vjp_total += cde_step_vjps(adaptive ? tab.n_stages : 4, adaptive ? 1 : 0, onj, step_idx);
    grid_sync(c);          // my copy reads ALL rows of my hidden unit: the next evaluation's row phases (other CTAs) overwrite them
  };
  // ================================================================ one vector-field evaluation -> K[out]
  auto eval_now = [&](int kind, int stage, float dts, float t, int perturb, int out, bool save_y1) {
    ++n_f;
    float tt = t;
    if (perturb > 0) tt = nextafterf(tt, tt + 1.0f);
    else if (perturb < 0) tt = nextafterf(tt, tt - 1.0f);
    const int seg = seg_index_t(tt, nk);
    const bool time_only = p.interp == CDE_INTERP_LINEAR && (seg & 1) == 0;
    float* Kout = arr_ptr(c, 2 + out);
    // ---------------- row phase
    const long long t_ev0 = clock64();
    if (worker && has_rows) {
      const int RP = c.RP;
      Recipe rc{kind, stage, dts, 0.f};
      build_rows(c, rc, tab, save_y1);
      TC_ACC(1, t_ev0);
      // dX/dt of my rows: channel 0 -> dX0 (shared), value channels -> dXg   (cde_fwd.cu: control_derivative)
      const float s = sub_(tt, static_cast<float>(seg));
      const int nch = time_only ? 1 : p.C;
      if (p.dx_cache) {
        // m and d of a knot segment do not change between evaluations: keep them in shared memory (the observation
        // loads were 7 k clk of latency per evaluation), v = m + (d - m) ((4 - 3 s) s) as before
        float* mS = c.dxc; float* dS = c.dxc + p.C * RP;
        if (seg != cached_seg) {
          for (int e = c.tid; e < RP * p.C; e += TC_WORK_THREADS) {
            const int r = e / p.C, ch = e - r * p.C;
            const int b = c.row0 + r;
            float m = 0.f, d = 0.f;
            if (b < p.B) {
              if (p.interp == CDE_INTERP_LINEAR) {
                const int mm = seg >> 1;
                if (((seg & 1) == 0) == (ch == 0)) d = sub_(obs_val_t(p, b, mm + 1, ch), obs_val_t(p, b, mm, ch));
                m = d;
              } else {
                const float x0 = obs_val_t(p, b, seg, ch), x1 = obs_val_t(p, b, seg + 1, ch);
                d = sub_(x1, x0);
                m = seg == 0 ? d : sub_(x0, obs_val_t(p, b, seg - 1, ch));
              }
            }
            mS[e] = m; dS[e] = d;
          }
          cached_seg = seg;
          named_bar_sync(1, TC_WORK_THREADS);
        }
        const float wq = mul_(sub_(4.0f, mul_(3.0f, s)), s);
        for (int e = c.tid; e < RP * nch; e += TC_WORK_THREADS) {
          const int r = e / nch, ch = e - r * nch;
          const int b = c.row0 + r;
          const float m = mS[r * p.C + ch], d = dS[r * p.C + ch];
          const float v = p.interp == CDE_INTERP_LINEAR ? d : add_(m, mul_(sub_(d, m), wq));
          if (ch == 0) c.dX0[r] = v;
          else {
            const int cc = ch - 1;
            __stcg(p.dXg + (static_cast<size_t>(b >> 7) * (Hc >> 2) + (cc >> 2)) * 512 + static_cast<size_t>(b & 127) * 4 + (cc & 3), v);
          }
        }
      } else
      for (int e = c.tid; e < RP * nch; e += TC_WORK_THREADS) {
        const int r = e / nch, ch = e - r * nch;
        const int b = c.row0 + r;
        float v = 0.f;
        if (b < p.B) {
          if (p.interp == CDE_INTERP_LINEAR) {
            const int m = seg >> 1;
            if ((seg & 1) == 0) {
              if (ch == 0) v = sub_(obs_val_t(p, b, m + 1, 0), obs_val_t(p, b, m, 0));
            } else if (ch > 0) {
              v = sub_(obs_val_t(p, b, m + 1, ch), obs_val_t(p, b, m, ch));
            }
          } else {
            const float x0 = obs_val_t(p, b, seg, ch), x1 = obs_val_t(p, b, seg + 1, ch);
            const float d = sub_(x1, x0);
            const float m = seg == 0 ? d : sub_(x0, obs_val_t(p, b, seg - 1, ch));
            v = add_(m, mul_(sub_(d, m), mul_(sub_(4.0f, mul_(3.0f, s)), s)));
          }
        }
        if (ch == 0) c.dX0[r] = v;
        else {
          // [row tile][c / 4][row in tile][c % 4]: the epilogue's LDG.128 of a warp (32 rows, 4 channels) is 512 contiguous bytes
          const int cc = ch - 1;
          __stcg(p.dXg + (static_cast<size_t>(b >> 7) * (Hc >> 2) + (cc >> 2)) * 512 + static_cast<size_t>(b & 127) * 4 + (cc & 3), v);
        }
      }
      // the Hc x Hc weights come through the X ring, idle in this phase (the producer warp streams them, see below)
      TC_ACC(2, t_ev0);
      float* lin = c.bufA; float* lout = c.bufB;
      uint32_t oin = c.oA, oout = c.oB;
      for (int l = 0; l <= p.NM; ++l) {
        const uint32_t g = c.wcount + l, s2 = g & 1u, ph = (g >> 1) & 1u;
        { TC_T0(); mbar_wait(&c.wfull[s2], ph); TC_ACC(9, _t0); }
        const long long t_l0 = clock64();
        const uint32_t ow = c.oRing + s2 * xstage;
        // l == NM: the time channel  tanh(w_{h,0} . a + b_{h,0})  (its dX_0/dt factor follows below)
        if (l < p.NM) row_linear<true>(nullptr, ow, p.bmlp[l], Hc, Hc, oin, oout, RP, p.act, c.tid);
        else row_linear<true>(nullptr, ow, p.b0, Hc, Hc, oin, oout, RP, ACT_TANH, c.tid);
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.wfree[s2]);
        TC_ACC(10, t_l0);
        named_bar_sync(1, TC_WORK_THREADS);
        TC_ACC(11, t_l0);
        if (l < p.NM) { float* tsw = lin; lin = lout; lout = tsw; const uint32_t to = oin; oin = oout; oout = to; }
      }
      TC_ACC(3, t_ev0);
      const int rq4 = RP >> 2;
      for (int e = c.tid; e < Hc * rq4; e += TC_WORK_THREADS) {
        const int hh = e / rq4, r4 = e - hh * rq4;
        const float4 t4 = *reinterpret_cast<const float4*>(lout + hh * RP + 4 * r4);
        const float4 d4 = *reinterpret_cast<const float4*>(c.dX0 + 4 * r4);
        stcg4(Kout + static_cast<size_t>(hh) * Bpad + c.row0 + 4 * r4,
              make_float4(mul_(t4.x, d4.x), mul_(t4.y, d4.y), mul_(t4.z, d4.z), mul_(t4.w, d4.w)));
      }
      if (!time_only) {
        // the last activation as the fp16 hi / lo K-major canonical image of its row tile: 16-byte pieces (row, 8 k)
        const int kg_n = Hc >> 3;
        for (int e = c.tid; e < RP * kg_n; e += TC_WORK_THREADS) {
          const int r = e / kg_n, kg = e - r * kg_n;
          const int row = c.row0 + r, rt = row >> 7, rl = row & 127;
          __half hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) tc_split(lin[(8 * kg + j) * RP + r], hi[j], lo[j]);
          unsigned char* dst = p.Ximg + static_cast<size_t>(rt) * xstage +
                               (static_cast<size_t>(rl >> 3) * kg_n + kg) * 128u + static_cast<size_t>(rl & 7) * 16u;
          __stcg(reinterpret_cast<uint4*>(dst),
                 make_uint4(tc_pack2(hi[0], hi[1]), tc_pack2(hi[2], hi[3]), tc_pack2(hi[4], hi[5]), tc_pack2(hi[6], hi[7])));
          __stcg(reinterpret_cast<uint4*>(dst + ximg),
                 make_uint4(tc_pack2(lo[0], lo[1]), tc_pack2(lo[2], lo[3]), tc_pack2(lo[4], lo[5]), tc_pack2(lo[6], lo[7])));
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");        // generic-proxy global stores -> the bulk copies of every CTA
    } else if (c.warp == TC_WARP_PROD && has_rows) {
      // weight stream of the row phase: Linear l -> ring stage (count & 1), one bulk copy each (4 Hc^2 <= stage bytes)
      if (elect_one()) {
        for (int l = 0; l <= p.NM; ++l) {
          const uint32_t g = c.wcount + l, s2 = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(&c.wfree[s2], ph ^ 1u);
          mbar_arrive_expect_tx(&c.wfull[s2], wbytes);
          tma_load_1d(c.xring + s2 * xstage, l < p.NM ? p.Wmlp[l] : p.W0t, wbytes, &c.wfull[s2]);
        }
      }
      __syncwarp();
    }
    if (has_rows) c.wcount += static_cast<uint32_t>(p.NM + 1);
    TC_ACC(4, t_ev0);
    grid_sync(c);
    TC_ACC(5, t_ev0);
    if (blockIdx.x == 0 && threadIdx.x == 0) g_cde_tc_dbg[0] += 1;
    if (time_only) return;
    const long long t_f0 = clock64();
    // ---------------- feature phase: K_out[h_own][row] += sum_c tanh(W[h,c] . a[row] + b[h,c]) dX_c[row]
    // CTA h walks the row tiles starting at tile h mod nrt: at any time the grid reads nrt different tiles (spreads the L2 load)
    const int nrt = p.nrt;
    const int tile_skew = h_own % nrt;
    auto tile_of = [&](int i) { const int t2 = i + tile_skew; return t2 >= nrt ? t2 - nrt : t2; };
    if (c.warp == TC_WARP_PROD) {
      if (elect_one()) {
        for (int rt = 0; rt < nrt; ++rt) {
          const uint32_t t = c.tcount + rt, s2 = t & 1u, ph = (t >> 1) & 1u;
          mbar_wait(&c.empty[s2], ph ^ 1u);
          mbar_arrive_expect_tx(&c.full[s2], xstage);
          tma_load_1d(c.xring + s2 * xstage, p.Ximg + static_cast<size_t>(tile_of(rt)) * xstage, xstage, &c.full[s2]);
        }
      }
      __syncwarp();
    } else if (c.warp == TC_WARP_MMA) {
      if (elect_one()) {
        if (c.tcount == 0) mbar_wait(c.wbar, 0);
        const uint32_t w_hi = smem_u32(c.wimg), w_lo = w_hi + 2u * Hc * Hc;
        for (int rt = 0; rt < nrt; ++rt) {
          const uint32_t t = c.tcount + rt, s2 = t & 1u, ph = (t >> 1) & 1u;
          { TC_T0(); mbar_wait(&c.full[s2], ph); if (blockIdx.x == 0) g_cde_tc_dbg[25] += clock64() - _t0; }
          { TC_T0(); mbar_wait(&c.tempty[s2], ph ^ 1u); if (blockIdx.x == 0) g_cde_tc_dbg[26] += clock64() - _t0; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t x_hi = smem_u32(c.xring) + s2 * xstage, x_lo = x_hi + ximg;
          const uint32_t d_main = c.tmem + s2 * 2u * Hc, d_cross = d_main + Hc;
          for (int ks = 0; ks < (Hc >> 4); ++ks) {
            const uint32_t o = ks * 256u;
            const uint64_t ah = tc_desc(x_hi + o, sbo), al = tc_desc(x_lo + o, sbo);
            const uint64_t bh = tc_desc(w_hi + o, sbo), bl = tc_desc(w_lo + o, sbo);
            tc_mma(d_main, ah, bh, idesc, ks ? 1u : 0u);
            tc_mma(d_cross, al, bh, idesc, ks ? 1u : 0u);
            tc_mma(d_cross, ah, bl, idesc, 1u);
          }
          tc_commit(&c.empty[s2]);
          tc_commit(&c.tfull[s2]);
        }
      }
      __syncwarp();
    } else {
      const int q = c.warp & 3, hf = c.warp >> 2;
      const int rl = 32 * q + c.lane;
      const float* __restrict__ biasv = reinterpret_cast<const float*>(smem + c.oBias);
      float dxn[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) dxn[i] = 0.f;
      if (hf < (Hc >> 5)) {
        const float* dp = p.dXg + (static_cast<size_t>(tile_of(0)) * (Hc >> 2) + 8 * hf) * 512 + rl * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = ldcg4(dp + 512 * i);
          dxn[4 * i] = v.x; dxn[4 * i + 1] = v.y; dxn[4 * i + 2] = v.z; dxn[4 * i + 3] = v.w;
        }
      }
      for (int rt = 0; rt < nrt; ++rt) {
        const uint32_t t = c.tcount + rt, s2 = t & 1u, ph = (t >> 1) & 1u;
        const int row = tile_of(rt) * 128 + rl;
        // the time channel's term of my row, fetched now: its L2 round trip (1.5-2 k clk under load) sat on the critical path
        // of every tile when it was loaded after the channel-half barrier below
        float* const kp = Kout + static_cast<size_t>(h_own) * Bpad + row;
        float kprev = 0.f;
        if (hf == 0) kprev = __ldcg(kp);
        { TC_T0(); mbar_wait(&c.tfull[s2], ph); TC_ACC(8, _t0); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float sum = 0.f;
        for (int cb = hf; cb < (Hc >> 5); cb += 2) {
          const uint32_t taddr = c.tmem + (static_cast<uint32_t>(32 * q) << 16) + s2 * 2u * Hc + 32u * cb;
          float dx[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) dx[i] = dxn[i];
          // dX/dt of the NEXT block of 32 channels (this tile or the next one) while this one is in the tanh chains
          {
            int ncb = cb + 2, nrt2 = rt;
            if (ncb >= (Hc >> 5)) { ncb = hf; nrt2 = rt + 1; }
            if (nrt2 < nrt && ncb < (Hc >> 5)) {
              const float* dp = p.dXg + (static_cast<size_t>(tile_of(nrt2)) * (Hc >> 2) + 8 * ncb) * 512 + rl * 4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 v = ldcg4(dp + 512 * i);
                dxn[4 * i] = v.x; dxn[4 * i + 1] = v.y; dxn[4 * i + 2] = v.z; dxn[4 * i + 3] = v.w;
              }
            }
          }
          // (two x16 loads with a wait each were measured slower than one x32 pair: 51 k vs 41 k clk per evaluation)
          uint32_t um[32], ux[32];
          tc_tmem_ld32(taddr + Hc, ux);
          tc_tmem_ld32(taddr, um);
          tc_tmem_ld_wait();
          if (p.fast_tanh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = (__uint_as_float(um[i]) + __uint_as_float(ux[i]) * (1.0f / 2048.0f)) + biasv[32 * cb + i];
              // 1 - 2 / (1 + e^{2v}) on the two SFU approximations: absolute error ~2e-7
              const float t = 1.0f - __fdividef(2.0f, __expf(2.0f * v) + 1.0f);
              sum = fmaf(t, dx[i], sum);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = (__uint_as_float(um[i]) + __uint_as_float(ux[i]) * (1.0f / 2048.0f)) + biasv[32 * cb + i];
              sum = fmaf(tanhf(v), dx[i], sum);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.tempty[s2]);
        float* part = c.part + (t & 1u) * 256;
        part[hf * 128 + rl] = sum;
        named_bar_sync(2, TC_WORK_THREADS);
        if (hf == 0) __stcg(kp, kprev + (part[rl] + part[128 + rl]));
      }
    }
    c.tcount += static_cast<uint32_t>(nrt);
    TC_ACC(6, t_f0);
    TC_ACC(16, t_f0);
    TC_ACC(24, t_f0);
    grid_sync(c);
    TC_ACC(7, t_f0);
  };

  // ================================================================ output i: hidden state of my rows -> pose head
  auto pose_now = [&](int kind, float dts, float x, int i) {
    if (worker && has_rows) {
      const int RP = c.RP;
      Recipe rc{kind, 0, dts, x};
      build_rows(c, rc, tab, false);
      for (int e = c.tid; e < Hc * RP; e += TC_WORK_THREADS) {
        const int r = e / Hc, hh = e - r * Hc;
        const int b = c.row0 + r;
        if (b < p.B) {
          const float v = c.bufA[hh * RP + r];
          if (p.hout) p.hout[(static_cast<size_t>(b) * S + i) * Hc + hh] = v;
          if (i == 0) p.z0_out[static_cast<size_t>(b) * Hc + hh] = v;
        }
      }
      row_linear<false>(p.Wreg0, 0, p.breg0, Hc, kRegHidden, c.oA, c.oB, RP, ACT_LEAKY01, c.tid);
      named_bar_sync(1, TC_WORK_THREADS);
      for (int e = c.tid; e < RP * kPoseDim; e += TC_WORK_THREADS) {
        const int r = e / kPoseDim, o = e - r * kPoseDim;
        const int b = c.row0 + r;
        float acc = 0.f;
        for (int k = 0; k < kRegHidden; ++k) acc = fmaf(c.bufB[k * RP + r], p.Wreg1[o * kRegHidden + k], acc);
        if (b < p.B) p.pose[(static_cast<size_t>(b) * S + i) * kPoseDim + o] = acc + p.breg1[o];
      }
      named_bar_sync(1, TC_WORK_THREADS);
    }
  };

  // sum over the elements (my hidden unit, all sequences) of an expression of the state arrays -> fp64 partial
  auto for_my_feature = [&](auto&& fn) -> double {
    double a = 0.0;
    if (worker) {
      const size_t base = static_cast<size_t>(h_own) * Bpad;
      for (int b = c.tid; b < p.B; b += TC_WORK_THREADS) a += fn(base + b);
    }
    return a;
  };

  // ================================================================ z0
  if (worker && has_rows) {
    const int RP = c.RP;
    float* Z = arr_ptr(c, 0);
    if (p.z0_in) {
      for (int e = c.tid; e < Hc * RP; e += TC_WORK_THREADS) {
        const int r = e / Hc, hh = e - r * Hc;
        const int b = c.row0 + r;
        __stcg(Z + static_cast<size_t>(hh) * Bpad + b, b < p.B ? p.z0_in[static_cast<size_t>(b) * Hc + hh] : 0.f);
      }
    } else {
      for (int e = c.tid; e < p.Cpad * RP; e += TC_WORK_THREADS) {
        const int r = e / p.Cpad, ch = e - r * p.Cpad;
        const int b = c.row0 + r;
        c.bufA[ch * RP + r] = (b < p.B && ch < p.C) ? obs_val_t(p, b, 0, ch) : 0.f;
      }
      named_bar_sync(1, TC_WORK_THREADS);
      row_linear<false>(p.Winit, 0, p.binit, p.Cpad, Hc, c.oA, c.oB, RP, ACT_TANH, c.tid);
      named_bar_sync(1, TC_WORK_THREADS);
      const int rq4 = RP >> 2;
      for (int e = c.tid; e < Hc * rq4; e += TC_WORK_THREADS) {
        const int hh = e / rq4, r4 = e - hh * rq4;
        stcg4(Z + static_cast<size_t>(hh) * Bpad + c.row0 + 4 * r4, *reinterpret_cast<const float4*>(c.bufB + hh * RP + 4 * r4));
      }
    }
    named_bar_sync(1, TC_WORK_THREADS);
  }

  // every case below at most REQUESTS one evaluation / output as its last action; the single call sites after the switch
  // keep one copy of the (large) role code in the kernel
  struct { int kind, stage; float dts, t; int perturb, out; bool save_y1; } ev{};
  struct { int kind; float dts, x; int i; } po{};
  bool want_eval = false, want_pose = false;
  auto eval = [&](int kind, int stage, float dts, float t, int perturb, int out, bool save_y1) {
    ev.kind = kind; ev.stage = stage; ev.dts = dts; ev.t = t; ev.perturb = perturb; ev.out = out; ev.save_y1 = save_y1;
    want_eval = true;
  };
  auto pose_out = [&](int kind, float dts, float x, int i) { po.kind = kind; po.dts = dts; po.x = x; po.i = i; want_pose = true; };

  int pc = PC_POSE0;
  while (pc != PC_END) {
    want_eval = want_pose = false;
    switch (pc) {
      case PC_POSE0:
        pose_out(RC_Z, 0.f, 0.f, 0);
        pc = PC_F0;
        break;
      case PC_F0:
        if (S == 1) { pc = PC_END; break; }
        if (adaptive) {
          eval(RC_Z, 0, 0.f, static_cast<float>(t_cur), 0, 0, false);
          pc = PC_HAIRER_A;
        } else {
          grid_i = 0;
          pc = PC_STEP_BEGIN;
        }
        break;
      case PC_HAIRER_A: {
        const float* Z = arr_ptr(c, 0); const float* K0 = arr_ptr(c, 2);
        double b2 = 0.0;
        const double a = for_my_feature([&](size_t o) {
          const float y = __ldcg(Z + o), f = __ldcg(K0 + o);
          const float sc = add_(p.atol, mul_(fabsf(y), p.rtol));
          const float q0 = __fdiv_rn(y, sc), q1 = __fdiv_rn(f, sc);
          b2 += static_cast<double>(q1) * q1;
          return static_cast<double>(q0) * q0;
        });
        double sa, sb;
        grid_reduce2(c, a, b2, sa, sb);
        const double d0 = sqrt(sa / elems);
        d1 = sqrt(sb / elems);
        h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        eval(RC_HAIRER, 0, static_cast<float>(h0), static_cast<float>(t_cur + h0), 0, 1, false);
        pc = PC_HAIRER_B;
        break;
      }
      case PC_HAIRER_B: {
        const float* Z = arr_ptr(c, 0); const float* K0 = arr_ptr(c, 2); const float* K1 = arr_ptr(c, 3);
        const double a = for_my_feature([&](size_t o) {
          const float sc = add_(p.atol, mul_(fabsf(__ldcg(Z + o)), p.rtol));
          const float q = __fdiv_rn(sub_(__ldcg(K1 + o), __ldcg(K0 + o)), sc);
          return static_cast<double>(q) * q;
        });
        double sa, sb;
        grid_reduce2(c, a, 0.0, sa, sb);
        const double d2 = sqrt(sa / elems) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        dt = fmin(100.0 * h0, h1);
        pc = PC_STEP_BEGIN;
        break;
      }
      case PC_STEP_BEGIN: {
        if (adaptive) {
          if (i_out >= S) { pc = PC_END; break; }
          if (n_steps >= p.max_steps) { status = 1; pc = PC_END; break; }
          step = dt; t_b = t_cur + dt; on_jump = 0;
          const double nj = static_cast<double>(next_jump);
          if (t_cur < nj && nj < t_cur + step) { on_jump = 1; step = nj - t_cur; t_b = nj; }
          dt_s = static_cast<float>(step); ta_s = static_cast<float>(t_cur); tb_s = static_cast<float>(t_b);
          st = 1;
        } else {
          if (grid_i >= grid_n - 1 || i_out >= S) { pc = PC_END; break; }
          t_cur = grid_time(grid_i); t_b = grid_time(grid_i + 1);
          dt_s = static_cast<float>(t_b - t_cur); ta_s = static_cast<float>(t_cur); tb_s = static_cast<float>(t_b);
          st = 0;
        }
        pc = PC_STAGE;
        break;
      }
      case PC_STAGE: {
        if (adaptive) {
          if (st >= tab.n_stages) { pc = PC_STEP_END; break; }
          const bool last = kDpCt[st] == 1.0f;
          const float ts = last ? tb_s : add_(ta_s, mul_(kDpCt[st], dt_s));
          const int s_now = st++;
          eval(RC_STAGE, s_now, dt_s, ts, last ? -1 : 0, s_now, s_now == tab.n_stages - 1);
        } else {
          if (st >= 4) { pc = PC_RK4_END; break; }
          const float third = static_cast<float>(1.0 / 3.0);
          const int s_now = st++;
          if (s_now == 0) eval(RC_Z, 0, dt_s, ta_s, 0, 0, false);
          else if (s_now == 1) eval(RC_RK4_1, 0, dt_s, add_(ta_s, mul_(dt_s, third)), 0, 1, false);
          else if (s_now == 2) eval(RC_RK4_2, 0, dt_s, add_(ta_s, mul_(dt_s, mul_(2.0f, third))), 0, 2, false);
          else eval(RC_RK4_3, 0, dt_s, tb_s, -1, 3, false);
        }
        break;
      }
      case PC_STEP_END: {
        const float* Z = arr_ptr(c, 0); const float* Y1 = arr_ptr(c, 1);
        const float* K[kMaxStages];
        for (int j = 0; j < kMaxStages; ++j) K[j] = arr_ptr(c, 2 + j);
        const double a = for_my_feature([&](size_t o) {
          float err = 0.f;
          bool any = false;
          for (int j = 0; j < tab.n_stages; ++j) {
            const float ej = tab.e[j];
            if (ej == 0.f) continue;
            const float term = mul_(__ldcg(K[j] + o), mul_(ej, dt_s));
            err = any ? add_(err, term) : term;
            any = true;
          }
          const float tol = add_(p.atol, mul_(p.rtol, fmaxf(fabsf(__ldcg(Z + o)), fabsf(__ldcg(Y1 + o)))));
          const float q = __fdiv_rn(err, tol);
          return static_cast<double>(q) * q;
        });
        double sa, sb;
        grid_reduce2(c, a, 0.0, sa, sb);
        const double ratio = sqrt(sa / elems);
        ++n_steps;
        if (!(ratio == ratio) || isinf(ratio)) { status = 2; pc = PC_END; break; }
        const bool accept = ratio <= 1.0;
        if (ratio == 0.0) dt = step * 10.0;
        else {
          const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
          dt = step * fmin(10.0, fmax(0.9 / pow(ratio, 0.2), dfactor));
        }
        if (accept) {
          if (p.ckpt) {
            if (n_acc >= p.ckpt_cap) { status = 3; pc = PC_END; break; }
            save_step(n_acc, on_jump);
          }
          ++n_acc; pc = PC_OUTPUTS;
        } else pc = PC_STEP_BEGIN;
        break;
      }
      case PC_OUTPUTS: {
        if (i_out < S && !(p.tout[i_out] > t_b)) {
          const float x = static_cast<float>((p.tout[i_out] - t_cur) / (t_b - t_cur));
          const int i = i_out++;
          pose_out(RC_INTERP, dt_s, x, i);
        } else {
          pc = PC_COMMIT;
        }
        break;
      }
      case PC_COMMIT: {
        // z <- y1, k0 <- k6 (FSAL): index rotations, identical in every thread of every CTA
        { const int t0 = c.ix[0]; c.ix[0] = c.ix[1]; c.ix[1] = t0; }
        { const int t0 = c.ix[2]; c.ix[2] = c.ix[2 + tab.n_stages - 1]; c.ix[2 + tab.n_stages - 1] = t0; }
        t_cur = t_b;
        if (on_jump) {
          if (next_jump != nk - 1) ++next_jump;
          eval(RC_Z, 0, 0.f, tb_s, +1, 0, false);               // vector field just after the knot
        }
        pc = PC_STEP_BEGIN;
        break;
      }
      case PC_RK4_END: {
        // y1 = y + dt (k1 + 3 (k2 + k3) + k4) * 0.125 (my hidden unit, all sequences)
        if (worker) {
          const float* Z = arr_ptr(c, 0); float* Y1 = arr_ptr(c, 1);
          const float* K0 = arr_ptr(c, 2); const float* K1 = arr_ptr(c, 3); const float* K2 = arr_ptr(c, 4); const float* K3 = arr_ptr(c, 5);
          const size_t base = static_cast<size_t>(h_own) * Bpad;
          for (int b = c.tid; b < Bpad; b += TC_WORK_THREADS) {
            const size_t o = base + b;
            const float s = add_(add_(__ldcg(K0 + o), mul_(3.0f, add_(__ldcg(K1 + o), __ldcg(K2 + o)))), __ldcg(K3 + o));
            __stcg(Y1 + o, add_(__ldcg(Z + o), mul_(mul_(dt_s, s), 0.125f)));
          }
        }
        grid_sync(c);
        if (p.ckpt) {
          if (n_acc >= p.ckpt_cap) { status = 3; pc = PC_END; break; }
          save_step(n_acc, 0);
        }
        ++n_steps; ++n_acc;
        pc = PC_AFTER_JUMP;
        break;
      }
      case PC_AFTER_JUMP: {      // the rk4 output loop
        if (i_out < S && !(t_b < p.tout[i_out])) {
          const int i = i_out++;
          if (t_b == p.tout[i]) pose_out(RC_Y1, dt_s, 0.f, i);
          else pose_out(RC_LERP, dt_s, static_cast<float>((p.tout[i] - t_cur) / (t_b - t_cur)), i);
        } else {
          { const int t0 = c.ix[0]; c.ix[0] = c.ix[1]; c.ix[1] = t0; }      // z <- y1
          ++grid_i;
          pc = PC_STEP_BEGIN;
        }
        break;
      }
      default:
        pc = PC_END;
        break;
    }
    if (want_eval) eval_now(ev.kind, ev.stage, ev.dts, ev.t, ev.perturb, ev.out, ev.save_y1);
    if (want_pose) pose_now(po.kind, po.dts, po.x, po.i);
  }

  if (blockIdx.x == 0 && c.tid == 0) g_cde_tc_dbg[15] = clock64() - t_kernel0;
  if (blockIdx.x == 0 && c.tid == 0 && p.log) {
    CdeLogHead* hd = reinterpret_cast<CdeLogHead*>(p.log);
    hd->n_acc = n_acc; hd->n_vjp = vjp_total; hd->status = status;
  }
  if (blockIdx.x == 0 && c.tid == 0 && p.stats) {
    p.stats[0] = n_steps; p.stats[1] = n_acc; p.stats[2] = n_f; p.stats[3] = status;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (c.warp == TC_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(512u) : "memory");
  }
}

cudaError_t cde_tc_debug_timeline(long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, g_cde_tc_dbg, sizeof(g_cde_tc_dbg));
}

cudaError_t launch_cde_tc(const CdeParams& prm, const DevTableau& tab, int grid, size_t smem_bytes, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(cde_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
  if (err != cudaSuccess) return err;
  void* args[] = {const_cast<CdeParams*>(&prm), const_cast<DevTableau*>(&tab)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(cde_tc_kernel), dim3(grid), dim3(TC_THREADS), args,
                                     smem_bytes, stream);
}

// ---- weight images of the final Linear W [Hc*C][Hc] (row h*C + c):
//   Wimg[h][hi|lo][(c/8)*(Hc/8) + k/8][c%8][k%8] fp16 for the value channels c' = 1 + c;  bval[h][c] = b[h*C + 1 + c];
//   W0t[k][h] = W[h*C][k];  b0[h] = b[h*C]
__global__ void cde_tc_pack_kernel(const float* __restrict__ W, const float* __restrict__ b, int Hc, int C,
                                   unsigned char* __restrict__ Wimg, float* __restrict__ bval, float* __restrict__ W0t,
                                   float* __restrict__ b0) {
  const size_t total = static_cast<size_t>(Hc) * Hc * Hc;
  const size_t img = static_cast<size_t>(Hc) * Hc * 2;        // bytes of one fp16 image
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Hc);
    const int cc = static_cast<int>((i / Hc) % Hc);
    const int h = static_cast<int>(i / (static_cast<size_t>(Hc) * Hc));
    const float w = W[(static_cast<size_t>(h) * C + 1 + cc) * Hc + k];
    __half hi, lo;
    tc_split(w, hi, lo);
    const size_t off = ((static_cast<size_t>(cc >> 3) * (Hc >> 3) + (k >> 3)) * 64 + (cc & 7) * 8 + (k & 7)) * 2;
    unsigned char* base = Wimg + static_cast<size_t>(h) * 2 * img;
    *reinterpret_cast<__half*>(base + off) = hi;
    *reinterpret_cast<__half*>(base + img + off) = lo;
    if (k == 0) bval[static_cast<size_t>(h) * Hc + cc] = b[static_cast<size_t>(h) * C + 1 + cc];
    if (cc == 0) {
      W0t[static_cast<size_t>(k) * Hc + h] = W[(static_cast<size_t>(h) * C) * Hc + k];
      if (k == 0) b0[h] = b[static_cast<size_t>(h) * C];
    }
  }
}

cudaError_t cde_tc_pack(const float* W, const float* b, int Hc, int C, unsigned char* Wimg, float* bval, float* W0t, float* b0,
                        cudaStream_t stream) {
  cde_tc_pack_kernel<<<592, 256, 0, stream>>>(W, b, Hc, C, Wimg, bval, W0t, b0);
  return cudaGetLastError();
}

}  // namespace odevio
