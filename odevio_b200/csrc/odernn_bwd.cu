// Fused ODE-RNN regressor backward (discretise-then-optimise, step sizes constant).
//
// Replaces the arithmetic behind `loss.backward()` through (reference file:line)
//   PoseODERNN.forward / evolve_state      src/models/PoseODERNN.py:88-123, :70-75
//   torchode AutoDiffAdjoint (plain autograd through the solver loop), call site :58-60,74
//   ODEFunc.forward                        src/models/ODEFunc.py:38-39
//   nn.RNN single step                     src/models/PoseODERNN.py:114
//   regressor head                         src/models/PoseODERNN.py:64-68,122
// as driven by scripts/train_model.py:72-78.
//
// One persistent CTA per sequence tile, same geometry as the forward (R = RT x L rows, T-layout
// [feature][R] tile arrays, weights streamed through the TMA ring of tile_gemm.cuh).  The forward
// left, per interval, the tile state at the start of every solver iteration that accepted a step
// (odernn_params.h: ckpt layout).  For every such iteration, in reverse:
//   1. re-evaluate the step's stages k_j = f(z_j), keeping the hidden activations (per-CTA scratch);
//   2. for j = ns-1 .. 0:  gk_j = dt (b_j gY + sum_{m>j} a_mj gz_m),  g = gk_j (1 - k_j^2),
//      back through the MLP with the PyTorch-layout weights as K-major operand (W^T g needs no
//      transposed copy), giving gz_j;  3. gY <- upd ? gY + sum_j gz_j : gY.
// Weight gradients are NOT accumulated here: every Linear's (input row a, pre-activation gradient
// row g) pair is appended to row-major record streams and reduced by wgrad.cu's dense GEMM.
#include "odernn_params.h"
#include "tile_gemm.cuh"

namespace odevio {

namespace {

template <int RT>
struct BCtx {
  const BwdParams* prm;
  TileThread th;
  WeightRing ring;
  RingPos pos;
  float* bufA; float* bufB;
  float* dt; int* upd;                 // shared [R]
  float* K[kMaxStages]; float* GZ[kMaxStages]; float* GY; float* HS;   // per-CTA global scratch
  int R, rq4, rq, tile;
};

__device__ __forceinline__ float4 mul4s(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 fma4s(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}

// one float4 (rows 4q .. 4q+3 of feature d) into block `blk` of a tcgen05 operand stream (hi / lo)
__device__ __forceinline__ void rec_put4(float* hi, float* lo, long long blk, int F, int R, int d, int q, float4 v) {
  const size_t o = static_cast<size_t>(blk) * F * R + rec_block_offset(d, 4 * q, R);
  const float4 h4 = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  st4(hi + o, h4);
  st4(lo + o, make_float4(v.x - h4.x, v.y - h4.y, v.z - h4.z, v.w - h4.w));
}

// z_j = y0 + dt * sum_{m<j} a_jm k_m  -> bufA (T-layout) and the layer-0 input record
template <int RT>
__device__ __forceinline__ void stage_input_b(BCtx<RT>& c, const float* Y0, int j, long long row0) {
  if (c.th.producer) return;
  const BwdParams& p = *c.prm;
  const int nvec = p.D * c.rq4;
  const float4 dt = ld4(c.dt + 4 * c.rq);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    float4 y = ld4(Y0 + off);
    if (j > 0) {
      // same operation order as the forward (odernn_fwd.cu: wsum4 / axpy4) so z_j is bit-identical
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      for (int m = 0; m < j; ++m) {
        const float cj = p.tab.a[j][m];
        if (cj == 0.f) continue;
        const float4 k = ld4(c.K[m] + off);
        if (!any) {
          acc = make_float4(mul_(k.x, cj), mul_(k.y, cj), mul_(k.z, cj), mul_(k.w, cj));
          any = true;
        } else {
          acc = make_float4(add_(acc.x, mul_(k.x, cj)), add_(acc.y, mul_(k.y, cj)),
                            add_(acc.z, mul_(k.z, cj)), add_(acc.w, mul_(k.w, cj)));
        }
      }
      if (any) y = make_float4(add_(y.x, mul_(dt.x, acc.x)), add_(y.y, mul_(dt.y, acc.y)),
                               add_(y.z, mul_(dt.z, acc.z)), add_(y.w, mul_(dt.w, acc.w)));
    }
    st4(c.bufA + off, y);
    const int d = e / c.rq4;
    rec_put4(p.recA_ode[0], p.recA_ode_lo[0], row0, p.D, p.Rb, d, c.rq, y);
  }
  named_bar_sync(1, c.th.ncons);
}

// g = upd ? dt (b_j gY + sum_{m>j} a_mj gz_m) (1 - k_j^2) : 0   -> bufA and the last layer's G record
template <int RT>
__device__ __forceinline__ void stage_grad_b(BCtx<RT>& c, int j, long long row0) {
  if (c.th.producer) return;
  const BwdParams& p = *c.prm;
  const int nvec = p.D * c.rq4;
  const float4 dt = ld4(c.dt + 4 * c.rq);
  const int4 up = *reinterpret_cast<const int4*>(c.upd + 4 * c.rq);
  const int NLm1 = p.NL - 1;
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    float4 acc = mul4s(ld4(c.GY + off), p.tab.b[j]);
    for (int m = j + 1; m < p.ns; ++m) {
      const float a = p.tab.a[m][j];
      if (a != 0.f) acc = fma4s(ld4(c.GZ[m] + off), a, acc);
    }
    const float4 k = ld4(c.K[j] + off);
    float4 g;
    g.x = up.x ? dt.x * acc.x * (1.f - k.x * k.x) : 0.f;
    g.y = up.y ? dt.y * acc.y * (1.f - k.y * k.y) : 0.f;
    g.z = up.z ? dt.z * acc.z * (1.f - k.z * k.z) : 0.f;
    g.w = up.w ? dt.w * acc.w * (1.f - k.w * k.w) : 0.f;
    st4(c.bufA + off, g);
    const int d = e / c.rq4;
    rec_put4(p.recG_ode[NLm1], p.recG_ode_lo[NLm1], row0, p.D, p.Rb, d, c.rq, g);
  }
  named_bar_sync(1, c.th.ncons);
}

// gY <- upd ? gY + sum_j gz_j : gY
template <int RT>
__device__ __forceinline__ void iter_end_b(BCtx<RT>& c) {
  if (c.th.producer) return;
  const BwdParams& p = *c.prm;
  const int nvec = p.D * c.rq4;
  const int4 up = *reinterpret_cast<const int4*>(c.upd + 4 * c.rq);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const float4 g0 = ld4(c.GY + off);
    float4 s = g0;
    for (int j = 0; j < p.ns; ++j) {
      const float4 z = ld4(c.GZ[j] + off);
      s.x += z.x; s.y += z.y; s.z += z.z; s.w += z.w;
    }
    st4(c.GY + off, make_float4(up.x ? s.x : g0.x, up.y ? s.y : g0.y, up.z ? s.z : g0.z, up.w ? s.w : g0.w));
  }
  named_bar_sync(1, c.th.ncons);
}

struct GemmOpB {
  const float* W; int K; int N;
  const float* in; bool ode_layout;
  Epilogue epi;
};

enum { BP_HEAD1 = 0, BP_HEAD2, BP_JUMP_A, BP_JUMP_B, BP_GRU_GATES, BP_GRU_ELEM, BP_GRU_BWD, BP_ITER, BP_RSTAGE, BP_RLAYER, BP_BSTAGE, BP_BLAYER,
       BP_ITER_END, BP_TILE_END };

}  // namespace

template <int RT, int LL>
__global__ void __launch_bounds__(128 * LL + 32, 1)
odernn_bwd_kernel(const __grid_constant__ BwdParams prm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BCtx<RT> c;
  c.prm = &prm;
  const BwdParams& p = prm;
  const int tid = threadIdx.x;
  constexpr int ncons = 128 * LL;
  constexpr int R = RT * LL;
  c.th.ncons = ncons;
  c.th.lane = tid & 31;
  c.th.producer = tid >= ncons;
  c.th.ctid = c.th.producer ? 0 : tid;
  c.R = R; c.rq4 = R / 4; c.rq = c.th.ctid % (R / 4);

  float* sm = reinterpret_cast<float*>(smem_raw);
  c.bufA = sm; sm += prm.buf_floats;
  c.bufB = sm; sm += prm.buf_floats;
  float* stages = sm; sm += static_cast<size_t>(prm.nst) * prm.stage_floats;
  c.dt = sm; sm += R;
  c.upd = reinterpret_cast<int*>(sm); sm += R;
  uintptr_t bp = (reinterpret_cast<uintptr_t>(sm) + 7) & ~static_cast<uintptr_t>(7);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bp);
  c.ring.buf = stages;
  c.ring.buf_off = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(stages) - smem_raw);
  c.ring.full = bars;
  c.ring.empty = bars + MAX_STAGES;
  c.ring.stage_floats = prm.stage_floats;
  c.ring.nst = prm.nst;
  c.ring.kc = static_cast<uint32_t>(prm.kc);
  c.pos.stage = 0; c.pos.phase = 0; c.pos.ready = 0;
  if (tid == 0) {
    for (int s = 0; s < prm.nst; ++s) {
      mbar_init(&c.ring.full[s], 1);
      mbar_init(&c.ring.empty[s], ncons / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int D = prm.D, H = prm.H, NL = prm.NL, S = prm.S;
  const size_t arr = static_cast<size_t>(D) * R;
  const size_t harr = static_cast<size_t>(H) * R;
  float* sc = prm.scratch + static_cast<size_t>(blockIdx.x) * prm.scratch_floats_per_cta;
  for (int j = 0; j < kMaxStages; ++j) { c.K[j] = sc + j * arr; c.GZ[j] = sc + (kMaxStages + j) * arr; }
  c.GY = sc + 2 * kMaxStages * arr;
  c.HS = c.GY + arr;                       // HS[(j * (NL-1) + lam) * harr]
  const size_t ivf = ckpt_interval_floats(D, R, prm.CK);

  // Tiles are taken from a work queue in order of decreasing cost (bwd_tile_order_kernel): with 3.46 tiles per SM at
  // B = 4096 a static round-robin costs 4 tile times on 68 of the 148 SMs; longest-first keeps the makespan at the mean.
  // Records, carries and gradients are addressed by the tile, so the result does not depend on which CTA ran it.
  int* const s_next_tile = reinterpret_cast<int*>(bars + 2 * MAX_STAGES);      // inside the plan's 128 B of slack after the barriers
  for (int q = blockIdx.x;; q += gridDim.x) {
    if (prm.tile_order) {
      if (tid == 0) *s_next_tile = atomicAdd(prm.tile_counter, 1);
      __syncthreads();
      q = *s_next_tile;
      __syncthreads();
    }
    if (q >= prm.ntiles) break;
    const int tile = prm.tile_order ? prm.tile_order[q] : q;
    c.tile = tile;
    const int nvalid = min(RT, prm.B - tile * RT);
    const float* ck_tile = prm.ckpt + static_cast<size_t>(tile) * prm.ckpt_floats_per_tile;
    // ---- gY <- grad of the final hidden state
    if (!c.th.producer) {
      for (int e = c.th.ctid; e < D * R; e += ncons) {
        const int r = e / D, d = e - r * D;
        const int l = r / RT, b = tile * RT + (r % RT);
        float v = 0.f;
        if (prm.i_hi == S - 1) {
          if (prm.ghT && b < prm.B) v = prm.ghT[(static_cast<size_t>(l) * prm.B + b) * D + d];
        } else {
          v = prm.tile_gy[static_cast<size_t>(tile) * D * R + static_cast<size_t>(d) * R + r];     // carried from the later range
        }
        c.GY[static_cast<size_t>(d) * R + r] = v;
      }
    }
    __syncthreads();

    int ph = BP_HEAD1;
    int i = prm.i_hi, l = 0, it = 0, j = 0, lam = 0, gg = 0;
    float* lin = c.bufA; float* lout = c.bufB;
    const float* Y0 = nullptr;
    long long row_it = 0;                 // first ODE-stream block of the current iteration (one block per stage)
    while (ph != BP_TILE_END) {
      GemmOpB op{};
      bool do_gemm = false;
      const float* ck_iv = ck_tile + static_cast<size_t>(i) * ivf;
      const float* Yend = ck_iv; const float* Ypost = ck_iv + arr;
      const long long rowJ = (static_cast<long long>(tile) * S + i) * RT;     // jump / head stream row
      switch (ph) {
        case BP_HEAD1: {
          // regressor hidden recomputed from the top layer's post-jump state
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < D * RT; e += ncons) {
              const int m = e / D, d = e - m * D;
              const float v = Ypost[static_cast<size_t>(d) * R + (LL - 1) * RT + m];
              c.bufB[d * RT + m] = v;
              p.recA_reg0[(rowJ + m) * D + d] = v;
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Wreg0; op.K = D; op.N = kRegHidden; op.in = c.bufB; op.ode_layout = false;
          op.epi.mode = EPI_STORE; op.epi.bias = p.breg0; op.epi.act = ACT_LEAKY01;
          op.epi.out0 = c.bufA; op.epi.ld0 = RT;
          op.epi.rec = p.recA_reg1; op.epi.rec_row0 = rowJ; op.epi.rec_ld = kRegHidden;
          op.epi.rec_rstride = 1; op.epi.rec_valid = RT;
          do_gemm = true;
          ph = BP_HEAD2;
          break;
        }
        case BP_HEAD2: {
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < kRegHidden * RT; e += ncons) {
              const int k = e / RT, m = e - k * RT;
              const int b = tile * RT + m;
              float ga = 0.f;
              if (b < p.B) {
                const float* gp = p.gpose + (static_cast<size_t>(b) * S + i) * kPoseDim;
#pragma unroll
                for (int o = 0; o < kPoseDim; ++o) ga = fmaf(p.Wreg1[o * kRegHidden + k], gp[o], ga);
              }
              const float a = c.bufA[k * RT + m];
              const float gz = ga * (a > 0.f ? 1.f : 0.1f);
              c.bufA[k * RT + m] = gz;
              p.recG_reg0[(rowJ + m) * kRegHidden + k] = gz;
            }
            if (c.th.ctid < RT * 8) {
              const int m = c.th.ctid / 8, o = c.th.ctid - m * 8;
              const int b = tile * RT + m;
              float v = 0.f;
              if (b < p.B && o < kPoseDim) v = p.gpose[(static_cast<size_t>(b) * S + i) * kPoseDim + o];
              p.recG_reg1[(rowJ + m) * 8 + o] = v;
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Wreg0_raw; op.K = kRegHidden; op.N = D; op.in = c.bufA; op.ode_layout = false;
          op.epi.mode = EPI_ADD; op.epi.act = ACT_NONE;
          op.epi.out0 = c.GY; op.epi.ld0 = R; op.epi.off0 = (LL - 1) * RT;
          do_gemm = true;
          l = LL - 1;
          gg = 0;
          ph = p.rnn_type == 0 ? BP_JUMP_A : BP_GRU_GATES;
          break;
        }
        case BP_JUMP_A: {
          // dpre = gh' (1 - h'^2);  records of the jump Linear ([x ; h^-] -> h')
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < D * RT; e += ncons) {
              const int m = e / D, d = e - m * D;
              const size_t o = static_cast<size_t>(d) * R + l * RT + m;
              const float hp = Ypost[o];
              const float dpre = c.GY[o] * (1.f - hp * hp);
              c.bufA[d * RT + m] = dpre;
              p.recG_rnn[l][(rowJ + m) * D + d] = dpre;
              float x;
              if (l == 0) {
                const int b = tile * RT + m;
                x = 0.f;
                if (b < p.B) {
                  const size_t row = static_cast<size_t>(b) * S + i;
                  x = (d < p.Dv) ? p.fv[row * p.Dv + d] : p.fi[row * (D - p.Dv) + (d - p.Dv)];
                }
              } else {
                x = Ypost[static_cast<size_t>(d) * R + (l - 1) * RT + m];
              }
              float* ra = p.recA_rnn[l] + (rowJ + m) * (2 * static_cast<size_t>(D));
              ra[d] = x;
              ra[D + d] = Yend[o];
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Whh_raw[l]; op.K = D; op.N = D; op.in = c.bufA; op.ode_layout = false;
          op.epi.mode = EPI_STORE; op.epi.act = ACT_NONE;
          op.epi.out0 = c.GY; op.epi.ld0 = R; op.epi.off0 = l * RT;      // gradient of the interval's end state
          do_gemm = true;
          ph = BP_JUMP_B;
          break;
        }
        case BP_JUMP_B: {
          op.W = p.Wih_raw[l]; op.K = D; op.N = D; op.in = c.bufA; op.ode_layout = false;
          op.epi.act = ACT_NONE;
          if (l > 0) {
            op.epi.mode = EPI_ADD; op.epi.out0 = c.GY; op.epi.ld0 = R; op.epi.off0 = (l - 1) * RT;
          } else {
            op.epi.mode = EPI_STORE; op.epi.out0 = nullptr;
            if (p.gfused) {
              op.epi.rec = p.gfused; op.epi.rec_row0 = static_cast<long long>(tile) * RT * S + i;
              op.epi.rec_ld = D; op.epi.rec_rstride = S; op.epi.rec_valid = nvalid;
            }
          }
          do_gemm = (l > 0) || (p.gfused != nullptr);
          if (--l < 0) { it = p.nloops[static_cast<size_t>(tile) * S + i] - 1; ph = BP_ITER; }
          else ph = BP_JUMP_A;
          break;
        }
        // ---------------------------------------------------------------- GRU jump backward
        case BP_GRU_GATES: {
          // re-evaluate the gates of layer l: r, z (K = 2D on [x ; h]), hn = W_hn h + b_hn, n = tanh(W_in x + b_in + r hn)
          float* RG = c.K[1]; float* ZG = c.K[2]; float* HN = c.K[3]; float* NG = c.K[4];     // [D][RT] scratch
          if (gg == 0 && !c.th.producer) {
            for (int e = c.th.ctid; e < D * RT; e += ncons) {
              const int m = e / D, d = e - m * D;
              float x;
              if (l == 0) {
                const int b = tile * RT + m;
                x = 0.f;
                if (b < p.B) {
                  const size_t row = static_cast<size_t>(b) * S + i;
                  x = (d < p.Dv) ? p.fv[row * p.Dv + d] : p.fi[row * (D - p.Dv) + (d - p.Dv)];
                }
              } else {
                x = Ypost[static_cast<size_t>(d) * R + (l - 1) * RT + m];
              }
              const float h = Yend[static_cast<size_t>(d) * R + l * RT + m];
              c.bufA[d * RT + m] = x;
              c.bufA[(D + d) * RT + m] = h;
              float* ra = p.recA_rnn[l] + (rowJ + m) * (2 * static_cast<size_t>(D));
              ra[d] = x; ra[D + d] = h;
            }
            named_bar_sync(1, ncons);
          }
          Epilogue& e = op.epi;
          e.mode = EPI_STORE; e.ld0 = RT;
          op.N = D; op.ode_layout = false; op.in = c.bufA; op.K = 2 * D;
          if (gg == 0) { op.W = p.Wrnn[l][0]; e.bias = p.brnn[l][0]; e.act = ACT_SIGMOID; e.out0 = RG; }
          else if (gg == 1) { op.W = p.Wrnn[l][1]; e.bias = p.brnn[l][1]; e.act = ACT_SIGMOID; e.out0 = ZG; }
          else if (gg == 2) {
            op.W = p.Wrnn[l][3]; e.bias = p.brnn[l][3]; e.act = ACT_NONE; e.out0 = HN;
            op.K = D; op.in = c.bufA + static_cast<size_t>(D) * RT;
          } else {
            op.W = p.Wrnn[l][2]; e.bias = p.brnn[l][2]; e.mode = EPI_GRU_N; e.rg = RG; e.hn = HN; e.out0 = NG;
            op.K = D;
          }
          do_gemm = true;
          if (++gg == 4) { gg = 0; ph = BP_GRU_ELEM; }
          break;
        }
        case BP_GRU_ELEM: {
          // h' = (1 - z) n + z h:  dn = g (1 - z), dz = g (h - n), dh += g z;  n = tanh(a_n + r hn):
          // dn_pre = dn (1 - n^2), dr = dn_pre hn, d(hn) = dn_pre r;  r, z sigmoid
          if (!c.th.producer) {
            const float* RG = c.K[1]; const float* ZG = c.K[2]; const float* HN = c.K[3]; const float* NG = c.K[4];
            const size_t DR = static_cast<size_t>(D) * RT;
            for (int e = c.th.ctid; e < D * RT; e += ncons) {
              const int m = e / D, d = e - m * D;
              const size_t o = static_cast<size_t>(d) * R + l * RT + m;     // tile arrays
              const size_t q = static_cast<size_t>(d) * RT + m;             // [D][RT] scratch
              const float g = c.GY[o], h = Yend[o];
              const float r = RG[q], z = ZG[q], hn = HN[q], n = NG[q];
              const float dn_pre = g * (1.f - z) * (1.f - n * n);
              const float dz_pre = g * (h - n) * z * (1.f - z);
              const float dr_pre = dn_pre * hn * r * (1.f - r);
              const float dhn = dn_pre * r;
              c.GY[o] = g * z;                                              // direct path h -> h'
              c.bufA[q] = dr_pre; c.bufA[DR + q] = dz_pre; c.bufB[q] = dn_pre; c.bufB[DR + q] = dhn;
              float* rg = p.recG_rnn[l] + (rowJ + m) * (6 * static_cast<size_t>(D));
              rg[d] = dr_pre; rg[D + d] = dz_pre; rg[2 * D + d] = dn_pre;               // G_ih
              rg[3 * D + d] = dr_pre; rg[4 * D + d] = dz_pre; rg[5 * D + d] = dhn;       // G_hh
            }
            named_bar_sync(1, ncons);
          }
          gg = 0;
          ph = BP_GRU_BWD;
          break;
        }
        case BP_GRU_BWD: {
          // dh += W_hh^T [dr_pre; dz_pre; dn_pre r],  dx = W_ih^T [dr_pre; dz_pre; dn_pre]; gate g uses rows g*D.. of
          // the PyTorch [3D][D] weight as the K-major operand
          const size_t DR = static_cast<size_t>(D) * RT, DD = static_cast<size_t>(D) * D;
          const int gate = gg % 3;
          const bool hh = gg < 3;
          op.K = D; op.N = D; op.ode_layout = false; op.epi.act = ACT_NONE; op.epi.ld0 = R;
          if (hh) {
            op.W = p.Whh_raw[l] + gate * DD;
            op.in = gate == 0 ? c.bufA : gate == 1 ? c.bufA + DR : c.bufB + DR;
            op.epi.mode = EPI_ADD; op.epi.out0 = c.GY; op.epi.off0 = l * RT;
            do_gemm = true;
          } else {
            op.W = p.Wih_raw[l] + gate * DD;
            op.in = gate == 0 ? c.bufA : gate == 1 ? c.bufA + DR : c.bufB;
            if (l > 0) {
              op.epi.mode = EPI_ADD; op.epi.out0 = c.GY; op.epi.off0 = (l - 1) * RT;
              do_gemm = true;
            } else if (p.gfused) {
              // accumulate the three gate contributions in scratch, emit the rows with the last one
              op.epi.mode = gate == 0 ? EPI_STORE : EPI_ADD; op.epi.out0 = c.K[5]; op.epi.ld0 = RT; op.epi.off0 = 0;
              if (gate == 2) {
                op.epi.rec = p.gfused; op.epi.rec_row0 = static_cast<long long>(tile) * RT * S + i;
                op.epi.rec_ld = D; op.epi.rec_rstride = S; op.epi.rec_valid = nvalid;
              }
              do_gemm = true;
            }
          }
          if (++gg == 6) {
            gg = 0;
            if (--l < 0) { it = p.nloops[static_cast<size_t>(tile) * S + i] - 1; ph = BP_ITER; }
            else ph = BP_GRU_GATES;
          }
          break;
        }
        case BP_ITER: {
          if (it < 0) {
            ph = (--i >= prm.i_lo) ? BP_HEAD1 : BP_TILE_END;
            break;
          }
          const float* slot = ck_iv + 2 * arr + static_cast<size_t>(it) * (arr + 2 * R);
          Y0 = slot;
          __syncthreads();      // previous users of dt / upd are done
          if (!c.th.producer && c.th.ctid < R) {
            c.dt[c.th.ctid] = slot[arr + c.th.ctid];
            c.upd[c.th.ctid] = reinterpret_cast<const int*>(slot + arr + R)[c.th.ctid];
          }
          __syncthreads();
          row_it = p.rec_base[static_cast<size_t>(tile) * S + i] / R + static_cast<long long>(it) * p.ns;
          j = 0;
          ph = BP_RSTAGE;
          break;
        }
        case BP_RSTAGE:
          stage_input_b<RT>(c, Y0, j, row_it + j);
          lam = 0; lin = c.bufA; lout = c.bufB;
          ph = BP_RLAYER;
          break;
        case BP_RLAYER: {
          op.W = p.Wode[lam]; op.K = p.Kode[lam]; op.N = p.Node[lam];
          op.in = lin; op.ode_layout = true;
          op.epi.mode = EPI_STORE; op.epi.bias = p.bode[lam]; op.epi.ld0 = R;
          if (lam == NL - 1) {
            op.epi.act = ACT_TANH; op.epi.out0 = c.K[j];
          } else {
            op.epi.act = p.act; op.epi.out0 = lout;
            op.epi.out1 = c.HS + (static_cast<size_t>(j) * (NL - 1) + lam) * harr; op.epi.ld1 = R;
            op.epi.rec = p.recA_ode[lam + 1]; op.epi.rec_lo = p.recA_ode_lo[lam + 1];
            op.epi.rec_row0 = row_it + j; op.epi.rec_ld = H; op.epi.rec_rstride = p.Rb; op.epi.rec_valid = RT;
          }
          do_gemm = true;
          float* t = lin; lin = lout; lout = t;
          if (++lam == NL) {
            if (++j < p.ns) ph = BP_RSTAGE;
            else { j = p.ns - 1; ph = BP_BSTAGE; }
          }
          break;
        }
        case BP_BSTAGE:
          stage_grad_b<RT>(c, j, row_it + j);
          lam = NL - 1; lin = c.bufA; lout = c.bufB;
          ph = BP_BLAYER;
          break;
        case BP_BLAYER: {
          // g_{lam-1} = (W_lam^T g_lam) * act'(h_{lam-1});  the PyTorch [out][in] weight is the K-major operand
          op.W = p.Wode_raw[lam]; op.K = p.Node[lam]; op.N = p.Kode[lam];
          op.in = lin; op.ode_layout = true; op.epi.ld0 = R;
          if (lam > 0) {
            op.epi.mode = EPI_MUL_DACT; op.epi.act = p.act;
            op.epi.hs = c.HS + (static_cast<size_t>(j) * (NL - 1) + (lam - 1)) * harr; op.epi.ldh = R;
            op.epi.out0 = lout;
            op.epi.rec = p.recG_ode[lam - 1]; op.epi.rec_lo = p.recG_ode_lo[lam - 1];
            op.epi.rec_row0 = row_it + j; op.epi.rec_ld = H; op.epi.rec_rstride = p.Rb; op.epi.rec_valid = RT;
          } else {
            op.epi.mode = EPI_STORE; op.epi.act = ACT_NONE; op.epi.out0 = c.GZ[j];
          }
          do_gemm = true;
          float* t = lin; lin = lout; lout = t;
          if (--lam < 0) ph = (--j >= 0) ? BP_BSTAGE : BP_ITER_END;
          break;
        }
        case BP_ITER_END:
          iter_end_b<RT>(c);
          --it;
          ph = BP_ITER;
          break;
        default:
          ph = BP_TILE_END;
          break;
      }
      if (do_gemm) tile_gemm<RT, LL, true>(c.ring, c.pos, c.th, op.W, op.K, op.N, op.in, op.ode_layout, op.epi);
    }

    // ---- gradient of the initial hidden state (last range), or the carry for the next launch
    __syncthreads();
    if (!c.th.producer && prm.i_lo > 0) {
      for (int e = c.th.ctid; e < D * R; e += ncons) prm.tile_gy[static_cast<size_t>(tile) * D * R + e] = c.GY[e];
    } else if (!c.th.producer && prm.gh0) {
      for (int e = c.th.ctid; e < D * R; e += ncons) {
        const int r = e / D, d = e - r * D;
        const int l2 = r / RT, b = tile * RT + (r % RT);
        if (b < prm.B) prm.gh0[(static_cast<size_t>(l2) * prm.B + b) * D + d] = c.GY[static_cast<size_t>(d) * R + r];
      }
    }
    __syncthreads();
  }
}

// Longest-processing-time order of the tiles for one backward launch: cost = stored solver iterations of the tile over the
// launch's interval range (every one is ns stage recomputations + ns stage pull-backs), rank by counting (ntiles <= 8192).
constexpr int kOrderMaxTiles = 8192;
__global__ void bwd_tile_order_kernel(const int* __restrict__ nloops, int S, int i_lo, int i_hi, int ntiles, int* __restrict__ order,
                                      int* __restrict__ counter) {
  __shared__ int cost[kOrderMaxTiles];
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    int s = 0;
    for (int i = i_lo; i <= i_hi; ++i) s += nloops[static_cast<size_t>(t) * S + i];
    cost[t] = s;
  }
  if (threadIdx.x == 0) *counter = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    const int ct = cost[t];
    int rank = 0;
    for (int u = 0; u < ntiles; ++u) {
      const int cu = cost[u];
      rank += (cu > ct || (cu == ct && u < t)) ? 1 : 0;
    }
    order[rank] = t;
  }
}

template <int RT, int LL>
static cudaError_t launch_one_b(const BwdParams& prm, int grid, size_t smem_bytes, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(odernn_bwd_kernel<RT, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
  if (err != cudaSuccess) return err;
  odernn_bwd_kernel<RT, LL><<<grid, 128 * LL + 32, smem_bytes, stream>>>(prm);
  return cudaGetLastError();
}

cudaError_t launch_odernn_bwd(const BwdParams& prm_in, int rows_per_tile, int grid, size_t smem_bytes,
                              cudaStream_t stream) {
  BwdParams prm = prm_in;
  if (prm.tile_order && prm.tile_counter && prm.ntiles > grid && prm.ntiles <= kOrderMaxTiles) {
    bwd_tile_order_kernel<<<1, 1024, 0, stream>>>(prm.nloops, prm.S, prm.i_lo, prm.i_hi, prm.ntiles,
                                                  const_cast<int*>(prm.tile_order), prm.tile_counter);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  } else {
    prm.tile_order = nullptr; prm.tile_counter = nullptr;          // one wave (or too many tiles to rank): static assignment
  }
  if (rows_per_tile == 4) {
    switch (prm.L) {
      case 1: return launch_one_b<4, 1>(prm, grid, smem_bytes, stream);
      case 2: return launch_one_b<4, 2>(prm, grid, smem_bytes, stream);
      case 3: return launch_one_b<4, 3>(prm, grid, smem_bytes, stream);
      case 4: return launch_one_b<4, 4>(prm, grid, smem_bytes, stream);
    }
  } else if (rows_per_tile == 8) {
    switch (prm.L) {
      case 1: return launch_one_b<8, 1>(prm, grid, smem_bytes, stream);
      case 2: return launch_one_b<8, 2>(prm, grid, smem_bytes, stream);
      case 3: return launch_one_b<8, 3>(prm, grid, smem_bytes, stream);
      case 4: return launch_one_b<8, 4>(prm, grid, smem_bytes, stream);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace odevio
