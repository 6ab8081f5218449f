// Fused Neural-CDE regressor backward (discretise-then-optimise; the accepted step sizes are constants).
//
// Replaces the arithmetic behind `loss.backward()` through (reference file:line)
//   PoseCDE.forward                                  src/models/PoseCDE.py:76-103
//   torchcde.cdeint(..., adjoint=False) -> torchdiffeq odeint dopri5 | rk4   (:98-101)
//   CDEFunc.forward                                  src/models/ODEFunc.py:81-84
//   initial / regressor heads                        src/models/PoseCDE.py:58-61,67-71,96,102
// as driven by scripts/train_model.py:72-78; checked against autograd through oracle/pose_cde.py
// (tests/test_cde_backward_gpu.py).
//
// The forward (cde_fwd.cu, checkpoint mode) left, per accepted step, the tile arrays Z, Y1, K0..K6 and a log entry
// (t_a, t_b, dt, knot landing, the outputs interpolated from the step).  The pullback needs no grid-wide reduction --
// the batch-joint controller only chose the step sizes -- so every tile of R sequences walks the log backwards on its
// own, one persistent CTA per tile list, same T-layout / TMA weight ring as the forward (tile_gemm.cuh):
//   per step:  [knot]  gY1 += J_f(t_b^+, y1)^T gK0next                 else gK_last = gK0next        (FSAL)
//              outputs: gh_i = head^T gpose_i, spread over y0, y1, k_j by the dense output's polynomial weights
//              y1 = y + dt sum b_j k_j:  gK_j += dt b_j gY1, gY += gY1
//              stages i = ns-1 .. :      gz = J_f(t_i, z_i)^T gK_i;  gY += gz;  gK_j += dt a_ij gz  (j < i)
//   J_f^T lam for f(z) = tanh(W_f a_n(z) + b_f) . dX/dt:  recompute the MLP, then per channel group
//      T = tanh(.) (tile_gemm -> staging), g(dX/dt)[c] = sum_h lam[h] T[h,c], G = lam[h] dX[c] (1 - T^2) in place,
//      ga_n += W_f,g^T G (tile_gemm, K = Gc*Hc), and back through the Hc x Hc Linears.
// Weight gradients are NOT accumulated here: each Linear's (input row, pre-activation gradient row) pair goes to
// row-major record streams reduced by wgrad.cu's dense GEMM; the host bounds the record memory by walking the log in
// chunks of steps (one launch per chunk, carried cotangents in per-tile global memory).
#include "cde_params.h"
#include "tile_gemm.cuh"

namespace odevio {

namespace {

__constant__ float kDpCb[7] = {0.0f, 0.2f, 0.3f, 0.8f, static_cast<float>(8.0 / 9.0), 1.0f, 1.0f};

template <int RT>
struct BC {
  const CdeBwdParams* prm;
  TileThread th;
  WeightRing ring;
  RingPos pos;
  float* bufA; float* bufB; float* staging; float* dXs; float* gdX; float* LAM; float* GA;   // shared
  float* GK[kMaxStages]; float* GYN; float* GZ; float* HS;                                    // per-CTA global scratch
  float* GY1; float* GKN;                                                                     // per-tile carried state
  int R, rq4, rq;
};

__device__ __forceinline__ int seg_index(float t, int nk) {
  int cnt = static_cast<int>(ceilf(t));
  cnt = max(0, min(cnt, nk));
  return max(0, min(cnt - 1, nk - 2));
}

__device__ __forceinline__ float obs_val(const CdeBwdParams& p, int b, int o, int ch) {
  if (ch == 0) return p.tobs[static_cast<size_t>(b) * p.So + o];
  const int f = ch - 1;
  const size_t row = static_cast<size_t>(b) * p.So + o;
  return (f < p.Dv) ? p.fv[row * p.Dv + f] : p.fi[row * (p.Hc - p.Dv) + (f - p.Dv)];
}

// dX/dt(t) of the tile's rows -> dXs [Cpad][R]   (same arithmetic as cde_fwd.cu: control_derivative)
template <int RT>
__device__ __forceinline__ void control_derivative_b(BC<RT>& c, int tile, float t) {
  if (c.th.producer) return;
  const CdeBwdParams& p = *c.prm;
  const int R = c.R;
  const int nk = p.interp == CDE_INTERP_LINEAR ? 2 * p.So - 1 : p.So;
  const int seg = seg_index(t, nk);
  const float s = sub_(t, static_cast<float>(seg));
  for (int e = c.th.ctid; e < p.Cpad * R; e += c.th.ncons) {
    const int r = e / p.Cpad, ch = e - r * p.Cpad;
    const int b = tile * R + r;
    float v = 0.f;
    if (b < p.B && ch < p.C) {
      if (p.interp == CDE_INTERP_LINEAR) {
        const int m = seg >> 1;
        if ((seg & 1) == 0) {
          if (ch == 0) v = sub_(obs_val(p, b, m + 1, 0), obs_val(p, b, m, 0));
        } else if (ch > 0) {
          v = sub_(obs_val(p, b, m + 1, ch), obs_val(p, b, m, ch));
        }
      } else {
        const float x0 = obs_val(p, b, seg, ch), x1 = obs_val(p, b, seg + 1, ch);
        const float d = sub_(x1, x0);
        const float m = seg == 0 ? d : sub_(x0, obs_val(p, b, seg - 1, ch));
        v = add_(m, mul_(sub_(d, m), mul_(sub_(4.0f, mul_(3.0f, s)), s)));
      }
    }
    c.dXs[ch * R + r] = v;
    c.gdX[ch * R + r] = 0.f;
  }
}

// pull g(dX/dt) [Cpad][R] back onto the observations: gX[b][o][ch] (+)=      (ch >= 1: the time channel has no gradient)
template <int RT>
__device__ __forceinline__ void control_derivative_pullback(BC<RT>& c, int tile, float t) {
  if (c.th.producer) return;
  const CdeBwdParams& p = *c.prm;
  if (!p.gX) return;
  const int R = c.R;
  const int nk = p.interp == CDE_INTERP_LINEAR ? 2 * p.So - 1 : p.So;
  const int seg = seg_index(t, nk);
  const float s = t - static_cast<float>(seg);
  if (p.interp == CDE_INTERP_LINEAR && (seg & 1) == 0) return;
  for (int e = c.th.ctid; e < p.Cpad * R; e += c.th.ncons) {
    const int r = e / p.Cpad, ch = e - r * p.Cpad;
    const int b = tile * R + r;
    if (b >= p.B || ch >= p.C || ch == 0) continue;
    const float gv = c.gdX[ch * R + r];
    float* gx = p.gX + (static_cast<size_t>(b) * p.So) * p.C + ch;
    if (p.interp == CDE_INTERP_LINEAR) {
      const int m = seg >> 1;
      gx[static_cast<size_t>(m + 1) * p.C] += gv;
      gx[static_cast<size_t>(m) * p.C] -= gv;
    } else if (seg == 0) {
      gx[static_cast<size_t>(1) * p.C] += gv;
      gx[0] -= gv;
    } else {
      const float w = (4.0f - 3.0f * s) * s;
      const float gd = gv * w, gm = gv - gd;
      gx[static_cast<size_t>(seg + 1) * p.C] += gd;
      gx[static_cast<size_t>(seg) * p.C] += gm - gd;
      gx[static_cast<size_t>(seg - 1) * p.C] -= gm;
    }
  }
}

// argument of a vector-field evaluation, rebuilt from the step's checkpoint -> bufA [Hc][R] (+ HS[0], record A_0)
enum { ZK_Z = 0, ZK_Y1, ZK_STAGE, ZK_RK4_1, ZK_RK4_2, ZK_RK4_3 };

template <int RT>
__device__ __forceinline__ void build_arg(BC<RT>& c, const float* ck, int kind, int stage, float dt,
                                          const DevTableau& tab, long long row0) {
  if (c.th.producer) return;
  const CdeBwdParams& p = *c.prm;
  // checkpoint element (array j, feature h, rows 4 r4 ..): ck + j * ck_jstride + h * ck_hstride + 4 r4   (tile layout of
  // cde_fwd.cu: [9][Hc][R]; feature-major layout of cde_tc.cu: [9][Hc][Bpad] with ck pointing at the tile's first row)
  const size_t js = p.ck_jstride, hs = p.ck_hstride;
  const int nvec = p.Hc * c.rq4;
  const float third = static_cast<float>(1.0 / 3.0);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const int hh = e / c.rq4, r4 = e - hh * c.rq4;
    const size_t coff = static_cast<size_t>(hh) * hs + 4 * r4;
    const float4 y4 = ld4(ck + coff);
    const float y[4] = {y4.x, y4.y, y4.z, y4.w};
    float out[4];
    auto ldk = [&](int j, float (&k)[4]) { const float4 v = ld4(ck + (2 + j) * js + coff); k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w; };
    float k0[4], k1[4], k2[4];
    switch (kind) {
      case ZK_Z:
        for (int q = 0; q < 4; ++q) out[q] = y[q];
        break;
      case ZK_Y1: {
        const float4 v = ld4(ck + js + coff);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        break;
      }
      case ZK_STAGE: {       // same operation order as the forward: y + sum_j k_j * fl(a_ij dt)
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        bool any = false;
        for (int j = 0; j < stage; ++j) {
          const float a = tab.a[stage][j];
          if (a == 0.f) continue;
          ldk(j, k0);
          const float w = mul_(a, dt);
          for (int q = 0; q < 4; ++q) acc[q] = any ? add_(acc[q], mul_(k0[q], w)) : mul_(k0[q], w);
          any = true;
        }
        for (int q = 0; q < 4; ++q) out[q] = any ? add_(y[q], acc[q]) : y[q];
        break;
      }
      case ZK_RK4_1:
        ldk(0, k0);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(mul_(dt, k0[q]), third));
        break;
      case ZK_RK4_2:
        ldk(0, k0); ldk(1, k1);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(dt, sub_(k1[q], mul_(k0[q], third))));
        break;
      default:
        ldk(0, k0); ldk(1, k1); ldk(2, k2);
        for (int q = 0; q < 4; ++q) out[q] = add_(y[q], mul_(dt, add_(sub_(k0[q], k1[q]), k2[q])));
        break;
    }
    const float4 o4 = make_float4(out[0], out[1], out[2], out[3]);
    st4(c.bufA + off, o4);
    st4(c.HS + off, o4);
  }
  named_bar_sync(1, c.th.ncons);
  // record A_0 row-major (coalesced along the features)
  for (int e = c.th.ctid; e < p.Hc * c.R; e += c.th.ncons) {
    const int r = e / p.Hc, k = e - r * p.Hc;
    p.recA[0][(row0 + r) * p.Hc + k] = c.bufA[k * c.R + r];
  }
}

// dst[e] (+)= sum_i w_i * src_i[e] over a [Hc][R] array (global scratch); fixed order
template <int RT>
__device__ __forceinline__ void axpy_arr(BC<RT>& c, float* dst, float w, const float* src) {
  if (c.th.producer) return;
  const int nvec = c.prm->Hc * c.rq4;
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const float4 d = ld4(dst + off), s = ld4(src + off);
    st4(dst + off, make_float4(fmaf(w, s.x, d.x), fmaf(w, s.y, d.y), fmaf(w, s.z, d.z), fmaf(w, s.w, d.w)));
  }
}

template <int RT>
__device__ __forceinline__ void fill_arr(BC<RT>& c, float* dst, const float* src) {     // src == nullptr: zero
  if (c.th.producer) return;
  const int nvec = c.prm->Hc * c.rq4;
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    st4(dst + off, src ? ld4(src + off) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// ga[n][r] (+)= sum_k G[k][r] * WT[k][n]   for the pull-back through the final Linear: K = Gc * Hc rows of G (up to 1024), only
// N = Hc outputs.  tile_gemm would leave half of the threads without a column pair (N / 2 = 64 pairs for 128 threads) and stream
// the weight in 4 KB stages; here the K range of every stage is SPLIT over `ks` thread groups (each pair of columns is owned by
// ks threads, their partial sums added in a fixed order through `scratch`), and a stage holds kc = up to 64 k-rows (32 KB).
// Same ring protocol as tile_gemm (one lane of the producer warp streams, every consumer warp releases each stage).
template <int RT, int LL>
__device__ __forceinline__ void gemm_ksplit(const WeightRing& ring, RingPos& pos, const TileThread& th,
                                            const float* __restrict__ WT, int K, int N, const float* inT, float* out,
                                            float* scratch, bool first) {
  constexpr int ncons = 128 * LL;
  constexpr int R = RT * LL;
  const int pairs = N >> 1;
  int ks = 128 / pairs;                       // thread groups per column pair inside a 128-thread row block
  if (ks > 2) ks = 2;                         // one partial set fits the scratch buffer
  if (ks < 1) ks = 1;
  int kc = 64;
  while (kc > 8 && (static_cast<size_t>(kc) * N > ring.stage_floats || K % kc)) kc >>= 1;
  const int nch = K / kc;
  if (th.producer) {
    if (th.lane == 0) {
      const uint32_t bytes = static_cast<uint32_t>(kc) * N * sizeof(float);
      for (int ch = 0; ch < nch; ++ch) {
        mbar_wait(&ring.empty[pos.stage], pos.phase ^ 1u);
        mbar_arrive_expect_tx(&ring.full[pos.stage], bytes);
        tma_load_1d(ring.buf + static_cast<size_t>(pos.stage) * ring.stage_floats, WT + static_cast<size_t>(ch) * kc * N, bytes,
                    &ring.full[pos.stage]);
        pos.advance(ring.nst);
      }
    } else {
      for (int ch = 0; ch < nch; ++ch) pos.advance(ring.nst);
    }
    __syncwarp();
    return;
  }
  const int rb = th.ctid / 128, t = th.ctid - rb * 128;          // row block (RT rows), thread inside it
  const int cp = t % pairs, part = t / pairs;
  const bool active = part < ks && pairs <= 128;
  const int kpp = kc / ks;                                       // k-rows of a stage per thread group
  float acc0[RT], acc1[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
  const float* xrow = inT + rb * RT;
  for (int ch = 0; ch < nch; ++ch) {
    mbar_wait(&ring.full[pos.stage], pos.phase);
    const float* ws = ring.buf + static_cast<size_t>(pos.stage) * ring.stage_floats;
    const uint32_t cur = pos.stage;
    pos.advance(ring.nst);
    if (active) {
      const float* wp = ws + static_cast<size_t>(part) * kpp * N + 2 * cp;
      const float* xp = xrow + (static_cast<size_t>(ch) * kc + part * kpp) * R;
#pragma unroll 4
      for (int kk = 0; kk < kpp; ++kk) {
        const float2 w = *reinterpret_cast<const float2*>(wp + kk * N);
        float x[RT];
#pragma unroll
        for (int q = 0; q < RT / 4; ++q) {
          const float4 v = ld4(xp + kk * R + 4 * q);
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) { acc0[r] = fmaf(x[r], w.x, acc0[r]); acc1[r] = fmaf(x[r], w.y, acc1[r]); }
      }
    }
    __syncwarp();
    if (th.lane == 0) mbar_arrive(&ring.empty[cur]);
  }
  // partial sums of group 1 -> scratch [N][R]; group 0 adds them (fixed order) and writes / accumulates the result
  if (active && part == 1) {
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      st4(scratch + static_cast<size_t>(2 * cp) * R + rb * RT + 4 * q, make_float4(acc0[4 * q], acc0[4 * q + 1], acc0[4 * q + 2], acc0[4 * q + 3]));
      st4(scratch + static_cast<size_t>(2 * cp + 1) * R + rb * RT + 4 * q, make_float4(acc1[4 * q], acc1[4 * q + 1], acc1[4 * q + 2], acc1[4 * q + 3]));
    }
  }
  named_bar_sync(1, ncons);
  if (active && part == 0) {
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      float4 a = make_float4(acc0[4 * q], acc0[4 * q + 1], acc0[4 * q + 2], acc0[4 * q + 3]);
      float4 b = make_float4(acc1[4 * q], acc1[4 * q + 1], acc1[4 * q + 2], acc1[4 * q + 3]);
      float* o0 = out + static_cast<size_t>(2 * cp) * R + rb * RT + 4 * q;
      float* o1 = out + static_cast<size_t>(2 * cp + 1) * R + rb * RT + 4 * q;
      if (ks == 2) {
        const float4 pa = ld4(scratch + static_cast<size_t>(2 * cp) * R + rb * RT + 4 * q);
        const float4 pb = ld4(scratch + static_cast<size_t>(2 * cp + 1) * R + rb * RT + 4 * q);
        a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w; b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
      }
      if (!first) {
        const float4 oa = ld4(o0), ob = ld4(o1);
        a.x += oa.x; a.y += oa.y; a.z += oa.z; a.w += oa.w; b.x += ob.x; b.y += ob.y; b.z += ob.z; b.w += ob.w;
      }
      st4(o0, a); st4(o1, b);
    }
  }
  pos.ready = 0;
  named_bar_sync(1, ncons);
}

struct GemmOpD {
  const float* W; int K; int N;
  const float* in;
  Epilogue epi;
};

enum { D_TILE_BEGIN = 0, D_STEP_BEGIN, D_AFTER_JUMP, D_OUT, D_OUT_APPLY, D_Y1, D_STAGE, D_STAGE_DONE, D_STEP_END,
       D_FINAL, D_F0_DONE, D_OUT0_DONE, D_INIT, D_INIT_GX, D_TILE_END,
       // sub-machines
       V_BEGIN, V_FWD, V_GROUP, V_GROUP_ELEM, V_GROUP_BWD, V_CHAIN0, V_CHAIN, V_END,
       H_1, H_2, H_3 };

}  // namespace

template <int RT, int LL>
__global__ void __launch_bounds__(128 * LL + 32, 1)
cde_bwd_kernel(const __grid_constant__ CdeBwdParams prm, const __grid_constant__ DevTableau tab) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BC<RT> c;
  c.prm = &prm;
  const CdeBwdParams& p = prm;
  const int tid = threadIdx.x;
  constexpr int ncons = 128 * LL;
  constexpr int R = RT * LL;
  c.th.ncons = ncons;
  c.th.lane = tid & 31;
  c.th.producer = tid >= ncons;
  c.th.ctid = c.th.producer ? 0 : tid;
  c.R = R; c.rq4 = R / 4; c.rq = c.th.ctid % (R / 4);

  const int Hc = prm.Hc, S = prm.S, NM = prm.NM, ns = prm.ns;
  float* sm = reinterpret_cast<float*>(smem_raw);
  c.bufA = sm; sm += prm.buf_floats;
  c.bufB = sm; sm += prm.buf_floats;
  c.staging = sm; sm += prm.staging_floats;
  c.dXs = sm; sm += prm.Cpad * R;
  c.gdX = sm; sm += prm.Cpad * R;
  c.LAM = sm; sm += Hc * R;
  c.GA = sm; sm += Hc * R;
  float* stages = sm; sm += static_cast<size_t>(prm.nst) * prm.stage_floats;
  uintptr_t bp = (reinterpret_cast<uintptr_t>(sm) + 15) & ~static_cast<uintptr_t>(15);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bp);
  c.ring.buf = stages;
  c.ring.buf_off = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(stages) - smem_raw);
  c.ring.full = bars;
  c.ring.empty = bars + MAX_STAGES;
  c.ring.stage_floats = prm.stage_floats;
  c.ring.nst = prm.nst;
  c.ring.kc = KC;
  c.pos.stage = 0; c.pos.phase = 0; c.pos.ready = 0;
  if (tid == 0) {
    for (int s = 0; s < prm.nst; ++s) {
      mbar_init(&c.ring.full[s], 1);
      mbar_init(&c.ring.empty[s], ncons / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const size_t arr = static_cast<size_t>(Hc) * R;
  float* sc = prm.scratch + static_cast<size_t>(blockIdx.x) * prm.scratch_floats_per_cta;
  for (int j = 0; j < kMaxStages; ++j) c.GK[j] = sc + j * arr;
  c.GYN = sc + kMaxStages * arr;
  c.GZ = c.GYN + arr;
  c.HS = c.GZ + arr;                         // HS[l] = a_l, l = 0..NM
  const int nk = prm.interp == CDE_INTERP_LINEAR ? 2 * prm.So - 1 : prm.So;
  const int NgTot = prm.ngroups * prm.Ng;
  const int first_stage = prm.fsal ? 1 : 0;
  const int n_stage_vjps = prm.fsal ? ns - 1 : ns;
  const bool adaptive = prm.solver == CDE_SOLVER_DOPRI5;

  for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
    c.GY1 = prm.tile_state + static_cast<size_t>(tile) * 2 * arr;
    c.GKN = c.GY1 + arr;

    int pc = D_TILE_BEGIN, ret = D_TILE_END, hret = D_TILE_END;
    int s = prm.step_hi - 1, st = 0, oi = 0;
    CdeStepRec rec{};
    const float* ck = nullptr;
    // VJP sub-machine registers
    int v_kind = 0, v_stage = 0, v_perturb = 0, v_layer = 0, v_group = 0;
    float v_t = 0.f, v_tt = 0.f, v_dt = 0.f;
    const float* v_cot = nullptr;
    long long v_row0 = 0;
    bool v_first = true, v_time_only = false;
    float* lin = c.bufA; float* lout = c.bufB;
    int h_i = 0;

    auto start_vjp = [&](int kind, int stage, float dts, float t, int perturb, const float* cot, int slot, int back) {
      v_kind = kind; v_stage = stage; v_dt = dts; v_t = t; v_perturb = perturb; v_cot = cot;
      v_row0 = (static_cast<long long>(rec.vjp_base + slot - prm.vjp_lo) * prm.ntiles + tile) * R;
      ret = back; pc = V_BEGIN;
    };
    auto start_head = [&](int i, int back) { h_i = i; hret = back; pc = H_1; };

    while (pc != D_TILE_END) {
      GemmOpD op{};
      bool do_gemm = false;
      switch (pc) {
        case D_TILE_BEGIN:
          if (prm.step_hi == prm.n_acc) {            // first launch of the walk: nothing flows in from the future
            fill_arr<RT>(c, c.GY1, nullptr);
            fill_arr<RT>(c, c.GKN, nullptr);
            if (!c.th.producer) named_bar_sync(1, ncons);
          }
          pc = D_STEP_BEGIN;
          break;
        // ================================================================ one accepted step, backwards
        case D_STEP_BEGIN: {
          if (s < prm.step_lo) { pc = prm.step_lo == 0 ? D_FINAL : D_TILE_END; break; }
          rec = prm.log[1 + s];
          ck = prm.ckpt + static_cast<size_t>(s) * prm.ck_step_stride + static_cast<size_t>(tile) * prm.ck_tile_stride;
          for (int j = 0; j < ns; ++j) fill_arr<RT>(c, c.GK[j], nullptr);
          fill_arr<RT>(c, c.GYN, nullptr);
          if (!c.th.producer) named_bar_sync(1, ncons);
          if (rec.on_jump) {
            // K0 of the next step was f(t_b^+, y1): pull its cotangent back onto y1
            start_vjp(ZK_Y1, 0, rec.dt_s, rec.tb_s, +1, c.GKN, n_stage_vjps, D_AFTER_JUMP);
          } else {
            if (prm.fsal) {
              fill_arr<RT>(c, c.GK[ns - 1], c.GKN);
              if (!c.th.producer) named_bar_sync(1, ncons);
            }
            oi = rec.out_count - 1;
            pc = D_OUT;
          }
          break;
        }
        case D_AFTER_JUMP:
          axpy_arr<RT>(c, c.GY1, 1.0f, c.GZ);
          if (!c.th.producer) named_bar_sync(1, ncons);
          oi = rec.out_count - 1;
          pc = D_OUT;
          break;
        case D_OUT:
          if (oi < 0) { pc = D_Y1; break; }
          start_head(rec.out_first + oi, D_OUT_APPLY);
          break;
        case D_OUT_APPLY: {
          // gh (in GZ) of output h_i, spread by the dense output's weights          (oracle interp_fit / interp_evaluate)
          const double tq = prm.tout[h_i];
          if (adaptive) {
            const float x = static_cast<float>((tq - rec.ta) / (rec.tb - rec.ta));
            const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
            const float wm = 16.f * x2 - 32.f * x3 + 16.f * x4;
            const float cy0 = 1.f - 11.f * x2 + 18.f * x3 - 8.f * x4 + wm;
            const float cy1 = -5.f * x2 + 14.f * x3 - 8.f * x4;
            const float cf0 = rec.dt_s * (x - 4.f * x2 + 5.f * x3 - 2.f * x4);
            const float cf1 = rec.dt_s * (x2 - 3.f * x3 + 2.f * x4);
            if (!c.th.producer) {
              const int nvec = Hc * c.rq4;
              for (int e = c.th.ctid; e < nvec; e += ncons) {
                const size_t off = static_cast<size_t>(e) * 4;
                const float4 g = ld4(c.GZ + off);
                auto acc = [&](float* dst, float w) {
                  const float4 d = ld4(dst + off);
                  st4(dst + off, make_float4(fmaf(w, g.x, d.x), fmaf(w, g.y, d.y), fmaf(w, g.z, d.z), fmaf(w, g.w, d.w)));
                };
                acc(c.GYN, cy0);
                acc(c.GY1, cy1);
                for (int j = 0; j < ns; ++j) {
                  float w = wm * tab.bmid[j] * rec.dt_s;
                  if (j == 0) w += cf0;
                  if (j == ns - 1) w += cf1;
                  if (w != 0.f) acc(c.GK[j], w);
                }
              }
              named_bar_sync(1, ncons);
            }
          } else {
            if (rec.tb == tq) {
              axpy_arr<RT>(c, c.GY1, 1.0f, c.GZ);
            } else {
              const float w = static_cast<float>((tq - rec.ta) / (rec.tb - rec.ta));
              axpy_arr<RT>(c, c.GYN, 1.0f - w, c.GZ);
              axpy_arr<RT>(c, c.GY1, w, c.GZ);
            }
            if (!c.th.producer) named_bar_sync(1, ncons);
          }
          --oi;
          pc = D_OUT;
          break;
        }
        case D_Y1:
          // y1 = y + dt sum_j b_j k_j
          for (int j = 0; j < ns; ++j)
            if (tab.b[j] != 0.f) axpy_arr<RT>(c, c.GK[j], rec.dt_s * tab.b[j], c.GY1);
          axpy_arr<RT>(c, c.GYN, 1.0f, c.GY1);
          if (!c.th.producer) named_bar_sync(1, ncons);
          st = ns - 1;
          pc = D_STAGE;
          break;
        case D_STAGE: {
          if (st < first_stage) { pc = D_STEP_END; break; }
          if (adaptive) {
            const bool last = kDpCb[st] == 1.0f;
            const float ts = last ? rec.tb_s : add_(rec.ta_s, mul_(kDpCb[st], rec.dt_s));
            start_vjp(ZK_STAGE, st, rec.dt_s, ts, last ? -1 : 0, c.GK[st], st - first_stage, D_STAGE_DONE);
          } else {
            const float third = static_cast<float>(1.0 / 3.0);
            if (st == 0) start_vjp(ZK_Z, 0, rec.dt_s, rec.ta_s, 0, c.GK[0], 0, D_STAGE_DONE);
            else if (st == 1) start_vjp(ZK_RK4_1, 0, rec.dt_s, add_(rec.ta_s, mul_(rec.dt_s, third)), 0, c.GK[1], 1, D_STAGE_DONE);
            else if (st == 2) start_vjp(ZK_RK4_2, 0, rec.dt_s, add_(rec.ta_s, mul_(rec.dt_s, mul_(2.0f, third))), 0, c.GK[2], 2, D_STAGE_DONE);
            else start_vjp(ZK_RK4_3, 0, rec.dt_s, rec.tb_s, -1, c.GK[3], 3, D_STAGE_DONE);
          }
          break;
        }
        case D_STAGE_DONE:
          axpy_arr<RT>(c, c.GYN, 1.0f, c.GZ);
          for (int j = 0; j < st; ++j)
            if (tab.a[st][j] != 0.f) axpy_arr<RT>(c, c.GK[j], rec.dt_s * tab.a[st][j], c.GZ);
          if (!c.th.producer) named_bar_sync(1, ncons);
          --st;
          pc = D_STAGE;
          break;
        case D_STEP_END:
          fill_arr<RT>(c, c.GY1, c.GYN);
          fill_arr<RT>(c, c.GKN, prm.fsal ? c.GK[0] : nullptr);
          if (!c.th.producer) named_bar_sync(1, ncons);
          --s;
          pc = D_STEP_BEGIN;
          break;
        // ================================================================ t0: f0, output 0, the initial network
        case D_FINAL:
          if (prm.fsal && prm.n_acc > 0) {
            rec = prm.log[1];
            ck = prm.ckpt + static_cast<size_t>(tile) * prm.ck_tile_stride;
            start_vjp(ZK_Z, 0, 0.f, static_cast<float>(prm.tout[0]), 0, c.GKN, n_stage_vjps + (rec.on_jump ? 1 : 0), D_F0_DONE);
          } else {
            start_head(0, D_OUT0_DONE);
          }
          break;
        case D_F0_DONE:
          axpy_arr<RT>(c, c.GY1, 1.0f, c.GZ);
          if (!c.th.producer) named_bar_sync(1, ncons);
          start_head(0, D_OUT0_DONE);
          break;
        case D_OUT0_DONE:
          axpy_arr<RT>(c, c.GY1, 1.0f, c.GZ);
          if (!c.th.producer) named_bar_sync(1, ncons);
          pc = D_INIT;
          break;
        case D_INIT: {
          // z0 = prev, or tanh(W_init X(knot 0) + b)                                              (PoseCDE.py:96)
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < Hc * R; e += ncons) {
              const int r = e / Hc, h = e - r * Hc;
              const int b = tile * R + r;
              float g = c.GY1[static_cast<size_t>(h) * R + r];
              if (b < p.B) {
                if (p.gz0) g += p.gz0[static_cast<size_t>(b) * Hc + h];
                if (p.has_prev) {
                  p.gprev[static_cast<size_t>(b) * Hc + h] = g;
                } else {
                  const float z = p.z0[static_cast<size_t>(b) * Hc + h];
                  g = g * (1.f - z * z);
                  p.recG_init[static_cast<size_t>(b) * Hc + h] = g;
                }
              } else {
                g = 0.f;
              }
              c.bufA[h * R + r] = g;
            }
            if (!p.has_prev) {
              for (int e = c.th.ctid; e < p.Cpad * R; e += ncons) {
                const int r = e / p.Cpad, ch = e - r * p.Cpad;
                const int b = tile * R + r;
                if (b < p.B) p.recA_init[static_cast<size_t>(b) * p.Cpad + ch] = ch < p.C ? obs_val(p, b, 0, ch) : 0.f;
              }
            }
            named_bar_sync(1, ncons);
          }
          if (!p.has_prev && p.gX) {
            op.W = p.WinitP; op.K = Hc; op.N = p.Cpad; op.in = c.bufA;
            op.epi.mode = EPI_STORE; op.epi.act = ACT_NONE; op.epi.out0 = c.bufB; op.epi.ld0 = R;
            do_gemm = true;
            pc = D_INIT_GX;
          } else {
            pc = D_TILE_END;
          }
          break;
        }
        case D_INIT_GX:
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < p.Cpad * R; e += ncons) {
              const int r = e / p.Cpad, ch = e - r * p.Cpad;
              const int b = tile * R + r;
              if (b < p.B && ch >= 1 && ch < p.C) p.gX[(static_cast<size_t>(b) * p.So) * p.C + ch] += c.bufB[ch * R + r];
            }
            named_bar_sync(1, ncons);
          }
          pc = D_TILE_END;
          break;

        // ================================================================ gz = J_f(t, z)^T lam  -> GZ
        case V_BEGIN: {
          build_arg<RT>(c, ck, v_kind, v_stage, v_dt, tab, v_row0);
          v_tt = v_t;
          if (v_perturb > 0) v_tt = nextafterf(v_tt, v_tt + 1.0f);
          else if (v_perturb < 0) v_tt = nextafterf(v_tt, v_tt - 1.0f);
          control_derivative_b<RT>(c, tile, v_tt);
          if (!c.th.producer) {
            const int nvec = Hc * c.rq4;
            for (int e = c.th.ctid; e < nvec; e += ncons) st4(c.LAM + static_cast<size_t>(e) * 4, ld4(v_cot + static_cast<size_t>(e) * 4));
            named_bar_sync(1, ncons);
          }
          v_time_only = prm.interp == CDE_INTERP_LINEAR && (seg_index(v_tt, nk) & 1) == 0;
          v_layer = 0; lin = c.bufA; lout = c.bufB;
          pc = V_FWD;
          break;
        }
        case V_FWD: {
          op.W = p.Wmlp[v_layer]; op.K = Hc; op.N = Hc; op.in = lin;
          op.epi.mode = EPI_STORE; op.epi.bias = p.bmlp[v_layer]; op.epi.act = p.act;
          op.epi.out0 = lout; op.epi.ld0 = R;
          op.epi.out1 = c.HS + static_cast<size_t>(v_layer + 1) * arr; op.epi.ld1 = R;
          op.epi.rec = p.recA[v_layer + 1]; op.epi.rec_row0 = v_row0; op.epi.rec_ld = Hc;
          op.epi.rec_rstride = 1; op.epi.rec_valid = RT;
          do_gemm = true;
          float* t = lin; lin = lout; lout = t;
          if (++v_layer == NM) { v_group = 0; v_first = true; pc = V_GROUP; }
          break;
        }
        case V_GROUP: {
          if (v_group >= p.ngroups) { pc = V_CHAIN0; break; }
          if (v_time_only && v_group >= 1) {
            // channels that do not move contribute an exact zero: their record columns must read as zero
            if (!c.th.producer) {
              const int ncol4 = (NgTot - p.Ng) / 4;
              for (int e = c.th.ctid; e < R * ncol4; e += ncons) {
                const int r = e / ncol4, q = e - r * ncol4;
                st4(p.recGf + (v_row0 + r) * NgTot + p.Ng + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
              }
            }
            pc = V_CHAIN0;
            break;
          }
          op.W = p.Wfin + static_cast<size_t>(v_group) * Hc * p.Ng; op.K = Hc; op.N = p.Ng; op.in = lin;
          op.epi.mode = EPI_STORE; op.epi.bias = p.bfin + static_cast<size_t>(v_group) * p.Ng; op.epi.act = ACT_TANH;
          op.epi.out0 = c.staging; op.epi.ld0 = R;
          do_gemm = true;
          pc = V_GROUP_ELEM;
          break;
        }
        case V_GROUP_ELEM: {
          if (!c.th.producer) {
            const int c0 = v_group * p.Gc;
            if (p.gX) {
              // g(dX/dt)[c][r] = sum_h lam[h][r] T[(c,h)][r]: tpo threads per output, fixed partial order + butterfly
              const int nout = p.Gc * R;
              int tpo = 1;
              while (tpo < 32 && nout * tpo * 2 <= ncons) tpo *= 2;
              const int per = ncons / tpo;
              const int part = c.th.ctid % tpo;
              const int hspan = (Hc + tpo - 1) / tpo;
              for (int o0 = 0; o0 < nout; o0 += per) {
                const int o = o0 + c.th.ctid / tpo;
                float accv = 0.f;
                if (o < nout) {
                  const int cl = o / R, r = o - cl * R;
                  const int h1 = min(Hc, (part + 1) * hspan);
                  for (int h = part * hspan; h < h1; ++h)
                    accv = fmaf(c.LAM[h * R + r], c.staging[(static_cast<size_t>(cl) * Hc + h) * R + r], accv);
                }
                for (int w = 1; w < tpo; w <<= 1) accv += __shfl_xor_sync(0xffffffffu, accv, w);
                if (o < nout && part == 0) {
                  const int cl = o / R, r = o - cl * R;
                  if (c0 + cl < p.Cpad) c.gdX[(c0 + cl) * R + r] = accv;
                }
              }
              named_bar_sync(1, ncons);
            }
            // G = lam[h] dX[c] (1 - T^2) in place + the final Linear's gradient record
            const int nvec = p.Ng * c.rq4;
            for (int e = c.th.ctid; e < nvec; e += ncons) {
              const int n = e / c.rq4, q4 = 4 * (e - n * c.rq4);
              const int cl = n / Hc, h = n - cl * Hc;
              const int ch = c0 + cl;
              float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ch < p.C) {
                const float4 t4 = ld4(c.staging + static_cast<size_t>(n) * R + q4);
                const float4 l4 = ld4(c.LAM + h * R + q4);
                const float4 d4 = ld4(c.dXs + ch * R + q4);
                g = make_float4(l4.x * d4.x * (1.f - t4.x * t4.x), l4.y * d4.y * (1.f - t4.y * t4.y),
                                l4.z * d4.z * (1.f - t4.z * t4.z), l4.w * d4.w * (1.f - t4.w * t4.w));
              }
              st4(c.staging + static_cast<size_t>(n) * R + q4, g);
              float* rp = p.recGf + (v_row0 + q4) * NgTot + static_cast<size_t>(v_group) * p.Ng + n;
              rp[0] = g.x; rp[NgTot] = g.y; rp[2 * static_cast<size_t>(NgTot)] = g.z; rp[3 * static_cast<size_t>(NgTot)] = g.w;
            }
            named_bar_sync(1, ncons);
          }
          pc = V_GROUP_BWD;
          break;
        }
        case V_GROUP_BWD: {
          if (Hc <= 256) {
            // K = Gc * Hc rows, only Hc outputs: the k-split routine (all threads busy, 32 KB stages); `lout` is free here
            gemm_ksplit<RT, LL>(c.ring, c.pos, c.th, p.WfinT + static_cast<size_t>(v_group) * p.Ng * Hc, p.Ng, Hc, c.staging,
                                c.GA, lout, v_first);
          } else {
            op.W = p.WfinT + static_cast<size_t>(v_group) * p.Ng * Hc; op.K = p.Ng; op.N = Hc; op.in = c.staging;
            op.epi.mode = v_first ? EPI_STORE : EPI_ADD; op.epi.act = ACT_NONE;
            op.epi.out0 = c.GA; op.epi.ld0 = R;
            do_gemm = true;
          }
          v_first = false;
          ++v_group;
          pc = V_GROUP;
          break;
        }
        case V_CHAIN0: {
          // pre-activation gradient of the last Hc -> Hc Linear
          if (!c.th.producer) {
            const float* an = c.HS + static_cast<size_t>(NM) * arr;
            for (int e = c.th.ctid; e < Hc * R; e += ncons) {
              const int r = e / Hc, k = e - r * Hc;
              const float g = c.GA[k * R + r] * dact_from_output(an[static_cast<size_t>(k) * R + r], p.act);
              c.bufA[k * R + r] = g;
              p.recG[NM - 1][(v_row0 + r) * Hc + k] = g;
            }
            named_bar_sync(1, ncons);
          }
          v_layer = NM - 1; lin = c.bufA; lout = c.bufB;
          pc = V_CHAIN;
          break;
        }
        case V_CHAIN: {
          op.W = p.Wmlp_raw[v_layer]; op.K = Hc; op.N = Hc; op.in = lin;
          op.epi.ld0 = R;
          if (v_layer > 0) {
            op.epi.mode = EPI_MUL_DACT; op.epi.act = p.act;
            op.epi.hs = c.HS + static_cast<size_t>(v_layer) * arr; op.epi.ldh = R;
            op.epi.out0 = lout;
            op.epi.rec = p.recG[v_layer - 1]; op.epi.rec_row0 = v_row0; op.epi.rec_ld = Hc;
            op.epi.rec_rstride = 1; op.epi.rec_valid = RT;
          } else {
            op.epi.mode = EPI_STORE; op.epi.act = ACT_NONE; op.epi.out0 = c.GZ;
          }
          do_gemm = true;
          float* t = lin; lin = lout; lout = t;
          if (--v_layer < 0) pc = V_END;
          break;
        }
        case V_END:
          control_derivative_pullback<RT>(c, tile, v_tt);
          if (!c.th.producer) named_bar_sync(1, ncons);
          pc = ret;
          break;

        // ================================================================ gh = head^T gpose_i  -> GZ
        case H_1: {
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < Hc * R; e += ncons) {
              const int r = e / Hc, k = e - r * Hc;
              const int b = tile * R + r;
              float v = 0.f;
              if (b < p.B) {
                v = p.hidden[(static_cast<size_t>(b) * S + h_i) * Hc + k];
                p.recA_reg0[(static_cast<size_t>(b) * S + h_i) * Hc + k] = v;
              }
              c.bufA[k * R + r] = v;
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Wreg0; op.K = Hc; op.N = kRegHidden; op.in = c.bufA;
          op.epi.mode = EPI_STORE; op.epi.bias = p.breg0; op.epi.act = ACT_LEAKY01;
          op.epi.out0 = c.bufB; op.epi.ld0 = R;
          op.epi.rec = p.recA_reg1; op.epi.rec_row0 = static_cast<long long>(tile) * R * S + h_i;
          op.epi.rec_ld = kRegHidden; op.epi.rec_rstride = S; op.epi.rec_valid = RT;
          do_gemm = true;
          pc = H_2;
          break;
        }
        case H_2: {
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < kRegHidden * R; e += ncons) {
              const int r = e / kRegHidden, k = e - r * kRegHidden;
              const int b = tile * R + r;
              float ga = 0.f;
              if (b < p.B) {
                const float* gp = p.gpose + (static_cast<size_t>(b) * S + h_i) * kPoseDim;
#pragma unroll
                for (int o = 0; o < kPoseDim; ++o) ga = fmaf(p.Wreg1[o * kRegHidden + k], gp[o], ga);
              }
              const float a = c.bufB[k * R + r];
              const float gz = ga * (a > 0.f ? 1.f : 0.1f);
              c.bufB[k * R + r] = gz;
              p.recG_reg0[((static_cast<size_t>(tile) * R + r) * S + h_i) * kRegHidden + k] = gz;
            }
            for (int e = c.th.ctid; e < R * 8; e += ncons) {
              const int r = e / 8, o = e - r * 8;
              const int b = tile * R + r;
              float v = 0.f;
              if (b < p.B && o < kPoseDim) v = p.gpose[(static_cast<size_t>(b) * S + h_i) * kPoseDim + o];
              p.recG_reg1[((static_cast<size_t>(tile) * R + r) * S + h_i) * 8 + o] = v;
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Wreg0_raw; op.K = kRegHidden; op.N = Hc; op.in = c.bufB;
          op.epi.mode = EPI_STORE; op.epi.act = ACT_NONE; op.epi.out0 = c.GZ; op.epi.ld0 = R;
          do_gemm = true;
          pc = H_3;
          break;
        }
        case H_3:
          pc = hret;
          break;
        default:
          pc = D_TILE_END;
          break;
      }
      if (do_gemm) tile_gemm<RT, LL>(c.ring, c.pos, c.th, op.W, op.K, op.N, op.in, true, op.epi);
    }
    __syncthreads();
  }
}

template <int RT, int LL>
static cudaError_t launch_cde_b(const CdeBwdParams& prm, const DevTableau& tab, int grid, size_t smem_bytes,
                                cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(cde_bwd_kernel<RT, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
  if (err != cudaSuccess) return err;
  cde_bwd_kernel<RT, LL><<<grid, 128 * LL + 32, smem_bytes, stream>>>(prm, tab);
  return cudaGetLastError();
}

cudaError_t launch_cde_bwd(const CdeBwdParams& prm, const DevTableau& tab, int RT, int LL, int grid,
                           size_t smem_bytes, cudaStream_t stream) {
  if (RT == 8 && LL == 1) return launch_cde_b<8, 1>(prm, tab, grid, smem_bytes, stream);
  if (RT == 8 && LL == 2) return launch_cde_b<8, 2>(prm, tab, grid, smem_bytes, stream);
  return cudaErrorInvalidValue;
}

// ---- the final Linear transposed per channel group: WT[g][n = c_local*Hc + h][k] = W[h*C + c][k]
__global__ void cde_pack_final_t_kernel(const float* __restrict__ W, int Hc, int C, int Gc, int ngroups,
                                        float* __restrict__ WT) {
  const int Ng = Gc * Hc;
  const size_t total = static_cast<size_t>(ngroups) * Ng * Hc;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Hc);
    const size_t gn = i / Hc;
    const int n = static_cast<int>(gn % Ng), g = static_cast<int>(gn / Ng);
    const int cl = n / Hc, h = n - cl * Hc;
    const int ch = g * Gc + cl;
    WT[i] = ch < C ? W[(static_cast<size_t>(h) * C + ch) * Hc + k] : 0.f;
  }
}

cudaError_t cde_pack_final_t(const float* W, int Hc, int C, int Gc, int ngroups, float* WT, cudaStream_t stream) {
  cde_pack_final_t_kernel<<<592, 256, 0, stream>>>(W, Hc, C, Gc, ngroups, WT);
  return cudaGetLastError();
}

// ---- gradient of the final Linear from the packed column order back to nn.Linear's [Hc*C][Hc] / [Hc*C]
__global__ void cde_unpack_final_grad_kernel(const float* __restrict__ dWp, const float* __restrict__ dbp, int Hc, int C,
                                             int Gc, float* __restrict__ dW, float* __restrict__ db) {
  const int Ng = Gc * Hc;
  const size_t total = static_cast<size_t>(Hc) * C * Hc;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Hc);
    const size_t row = i / Hc;                       // h * C + ch
    const int ch = static_cast<int>(row % C), h = static_cast<int>(row / C);
    const int g = ch / Gc, cl = ch - g * Gc;
    const size_t n = static_cast<size_t>(g) * Ng + static_cast<size_t>(cl) * Hc + h;
    dW[i] = dWp[n * Hc + k];
    if (k == 0) db[row] = dbp[n];
  }
}

cudaError_t cde_unpack_final_grad(const float* dWp, const float* dbp, int Hc, int C, int Gc, float* dW, float* db,
                                  cudaStream_t stream) {
  cde_unpack_final_grad_kernel<<<592, 256, 0, stream>>>(dWp, dbp, Hc, C, Gc, dW, db);
  return cudaGetLastError();
}

}  // namespace odevio
