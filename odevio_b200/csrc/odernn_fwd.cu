// Fused ODE-RNN regressor forward: ONE persistent kernel runs, for a tile of sequences, every
// observation interval's ODE solves (all RK stages, Butcher combines, error norm, step-size
// controller, dense-output end point), the RNN/GRU jump and the pose head, without returning
// to the host.
//
// Replaces the arithmetic behind (reference file:line):
//   PoseODERNN.forward / evolve_state      src/models/PoseODERNN.py:88-123, :70-75
//   torchode AutoDiffAdjoint.solve et al.  (call sites src/models/PoseODERNN.py:55-60) -- semantics
//                                          as restated in oracle/torchode_like.py
//   ODEFunc.forward                        src/models/ODEFunc.py:38-39
//   nn.RNN / nn.GRU single step            src/models/PoseODERNN.py:114
//   regressor head                         src/models/PoseODERNN.py:64-68,122
//
// Tile = RT sequences x L rnn layers = R "ODE rows" that share one ODEFunc (the reference's
// jit.fork over layers, PoseODERNN.py:109, becomes extra rows of the same GEMM).  Row r = l*RT+m.
// State/stage vectors of a tile live in per-CTA global scratch (L2 resident) in T-layout
// [d][R]; GEMM operands in shared memory; weights stream through the TMA ring of tile_gemm.cuh.
#include "odernn_params.h"
#include "tile_gemm.cuh"

namespace odevio {

namespace {

struct RowState {      // shared-memory per-row solver state (arrays of R)
  float* t; float* dt; float* tend; float* tmin; float* tmax;
  float* dtstep; float* x;
  int* run; int* noteval; int* upd; int* toeval; int* nsteps; int* nacc; int* status;
};

template <int RT>
struct Ctx {
  const FwdParams* prm;
  TileThread th;
  WeightRing ring;
  RingPos pos;
  float* bufA; float* bufB; float* partial;
  RowState rs;
  float* K[kMaxStages]; float* Y; float* Y1;
  int R;        // rows in the tile
  int rq4;      // R / 4
  int rq;       // this consumer's row quad for elementwise passes
  int tile;     // current tile index
};

// (layer, sequence) of local row (l, b) in the addressing of state / ts / stats / status (see FwdParams::full_B)
__device__ __forceinline__ void global_ids(const FwdParams& p, int l, int b, int& ll, int& bb) {
  ll = l;
  bb = (p.full_B > 0) ? p.seq[p.row_off + b] : b;
}

// ---- elementwise helpers over T-layout [D][R]; one float4 = 4 rows of one feature ---------

__device__ __forceinline__ float4 wsum4(const float* const* K, const float* coef, int n, size_t off,
                                        bool& any) {
  // acc = c0*k0 + c1*k1 + ... left to right, skipping exact zeros (oracle/_weighted_sum).
  // All stage vectors are loaded first (warp-uniform predicates, independent 128-bit loads): one L2 round trip per
  // element instead of one per term -- the per-term load -> use chain was ~12 % of the kernel's warp time (ncu
  // long_scoreboard), see DESIGN.md 4.6 for the same fix in the tensor-core solver.
  static_assert(kMaxStages == 7, "wsum4 is written out for 7 stages");
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool u0 = 0 < n && coef[0] != 0.f, u1 = 1 < n && coef[1] != 0.f, u2 = 2 < n && coef[2] != 0.f,
             u3 = 3 < n && coef[3] != 0.f, u4 = 4 < n && coef[4] != 0.f, u5 = 5 < n && coef[5] != 0.f,
             u6 = 6 < n && coef[6] != 0.f;
  const float4 k0 = u0 ? ld4(K[0] + off) : z4, k1 = u1 ? ld4(K[1] + off) : z4, k2 = u2 ? ld4(K[2] + off) : z4,
               k3 = u3 ? ld4(K[3] + off) : z4, k4 = u4 ? ld4(K[4] + off) : z4, k5 = u5 ? ld4(K[5] + off) : z4,
               k6 = u6 ? ld4(K[6] + off) : z4;
  float4 acc = z4;
  any = false;
  auto term = [&](bool u, const float4& k, float cj) {
    if (!u) return;
    if (!any) {
      acc = make_float4(mul_(k.x, cj), mul_(k.y, cj), mul_(k.z, cj), mul_(k.w, cj));
      any = true;
    } else {
      acc = make_float4(add_(acc.x, mul_(k.x, cj)), add_(acc.y, mul_(k.y, cj)),
                        add_(acc.z, mul_(k.z, cj)), add_(acc.w, mul_(k.w, cj)));
    }
  };
  term(u0, k0, coef[0]); term(u1, k1, coef[1]); term(u2, k2, coef[2]); term(u3, k3, coef[3]);
  term(u4, k4, coef[4]); term(u5, k5, coef[5]); term(u6, k6, coef[6]);
  return acc;
}

__device__ __forceinline__ float4 axpy4(float4 y, float4 dt, float4 acc) {   // y + dt*acc
  return make_float4(add_(y.x, mul_(dt.x, acc.x)), add_(y.y, mul_(dt.y, acc.y)),
                     add_(y.z, mul_(dt.z, acc.z)), add_(y.w, mul_(dt.w, acc.w)));
}

// stage argument i -> bufA (shared T-layout).  i == 0 copies Y.
template <int RT>
__device__ __forceinline__ void stage_input(Ctx<RT>& c, int i) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  const int nvec = p.D * c.rq4;
  const float4 dt = ld4(c.rs.dt + 4 * c.rq);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    float4 y = ld4(c.Y + off);
    if (i > 0) {
      bool any;
      const float4 acc = wsum4(c.K, p.tab.a[i], i, off, any);
      if (any) y = axpy4(y, dt, acc);
    }
    st4(c.bufA + off, y);
  }
  named_bar_sync(1, c.th.ncons);
}

// Error pass: y1 -> Y1, per-row sum of (err / bound)^2 -> partial[].
template <int RT>
__device__ __forceinline__ void error_pass(Ctx<RT>& c) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  const DevTableau& tb = p.tab;
  const int nvec = p.D * c.rq4;
  const int s = tb.n_stages;
  const float4 dt = ld4(c.rs.dt + 4 * c.rq);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const float4 y0 = ld4(c.Y + off);
    bool any;
    float4 acc = tb.ssal ? wsum4(c.K, tb.a[s - 1], s - 1, off, any) : wsum4(c.K, tb.b, s, off, any);
    const float4 y1 = any ? axpy4(y0, dt, acc) : y0;
    st4(c.Y1 + off, y1);
    if (tb.has_err) {
      acc = wsum4(c.K, tb.e, s, off, any);
      const float ex = mul_(dt.x, acc.x), ey = mul_(dt.y, acc.y), ez = mul_(dt.z, acc.z), ew = mul_(dt.w, acc.w);
      const float bx = add_(p.atol, mul_(p.rtol, fmaxf(fabsf(y0.x), fabsf(y1.x))));
      const float by = add_(p.atol, mul_(p.rtol, fmaxf(fabsf(y0.y), fabsf(y1.y))));
      const float bz = add_(p.atol, mul_(p.rtol, fmaxf(fabsf(y0.z), fabsf(y1.z))));
      const float bw = add_(p.atol, mul_(p.rtol, fmaxf(fabsf(y0.w), fabsf(y1.w))));
      const float rx = __fdiv_rn(ex, bx), ry = __fdiv_rn(ey, by), rz = __fdiv_rn(ez, bz), rw = __fdiv_rn(ew, bw);
      sum.x = fmaf(rx, rx, sum.x); sum.y = fmaf(ry, ry, sum.y);
      sum.z = fmaf(rz, rz, sum.z); sum.w = fmaf(rw, rw, sum.w);
    }
  }
  st4(c.partial + 4 * c.th.ctid, sum);
  named_bar_sync(1, c.th.ncons);
}

// Per-row step-size controller + bookkeeping (thread ctid == row).  Returns "row still running".
template <int RT>
__device__ __forceinline__ int controller(Ctx<RT>& c, int loops, int interval) {
  if (c.th.producer || c.th.ctid >= c.R) return 0;
  const FwdParams& p = *c.prm;
  const int r = c.th.ctid;
  RowState& rs = c.rs;
  int run = rs.run[r];
  const float dt = rs.dt[r];
  float t = rs.t[r];
  const float tend = rs.tend[r];
  bool accept = true, finite = true;
  float dt_next = dt;
  float ratio_out = 0.f;
  if (p.tab.has_err) {
    // deterministic reduction of the partial sums of this row's quad, in thread order
    const int rq = r >> 2, j = r & 3;
    float total = 0.f;
    for (int k = rq; k < c.th.ncons; k += c.rq4) total = add_(total, c.partial[4 * k + j]);
    const float ratio = sqrtf(__fdiv_rn(total, static_cast<float>(p.D)));
    finite = isfinite(ratio);
    ratio_out = ratio;
    accept = p.accept_strict ? (ratio < 1.0f) : (ratio <= 1.0f);
    float factor = mul_(p.safety, powf(ratio, p.tab.exponent));
    factor = fminf(fmaxf(factor, p.fmin), p.fmax);
    if (p.floor_factor && accept) factor = fmaxf(factor, 1.0f);
    dt_next = mul_(dt, factor);
  }
  if (p.stats && loops <= p.trace_steps && run) {      // diagnostic trace of the first T steps
    const int b = c.tile * RT + (r % RT);
    if (b < p.B) {
      int* sp = p.stats + ((static_cast<size_t>(interval) * p.L + r / RT) * p.B + b) * (2 + 2 * p.trace_steps);
      sp[2 + 2 * (loops - 1)] = __float_as_int(dt);
      sp[3 + 2 * (loops - 1)] = __float_as_int(ratio_out);
    }
  }
  const int upd = (accept && run) ? 1 : 0;
  rs.nsteps[r] += run;
  rs.nacc[r] += upd;
  const bool lands = p.exact_landing && dt >= sub_(tend, t);
  const float t_new = upd ? (lands ? tend : add_(t, dt)) : t;
  const int toeval = (upd && t_new >= tend && rs.noteval[r]) ? 1 : 0;
  if (toeval) {
    rs.x[r] = __fdiv_rn(sub_(tend, t), dt);
    rs.noteval[r] = 0;
  }
  rs.upd[r] = upd;
  rs.toeval[r] = toeval;
  rs.dtstep[r] = dt;
  t = t_new;
  if (run && !finite) rs.status[r] = max(rs.status[r], 2);
  run = (run && t < tend && finite) ? 1 : 0;
  if (run && loops >= p.max_steps) { rs.status[r] = max(rs.status[r], 1); run = 0; }
  float dtn = run ? dt_next : dt;
  dtn = fminf(fmaxf(dtn, sub_(rs.tmin[r], t)), sub_(rs.tmax[r], t));
  rs.t[r] = t;
  rs.dt[r] = dtn;
  rs.run[r] = run;
  return run;
}

// Commit pass: accepted rows take y1 (or the dense-output value at t_end), FSAL carry.
template <int RT>
__device__ __forceinline__ void commit_pass(Ctx<RT>& c) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  const DevTableau& tb = p.tab;
  const int s = tb.n_stages;
  const int r0 = 4 * c.rq;
  int upd[4], tev[4];
  float dts[4], xs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    upd[j] = c.rs.upd[r0 + j]; tev[j] = c.rs.toeval[r0 + j] && p.endpoint_dense;
    dts[j] = c.rs.dtstep[r0 + j]; xs[j] = c.rs.x[r0 + j];
  }
  if (!(upd[0] | upd[1] | upd[2] | upd[3])) return;
  const bool any_dense = tev[0] | tev[1] | tev[2] | tev[3];
  const int nvec = p.D * c.rq4;
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const float4 y0v = ld4(c.Y + off);
    const float4 y1v = ld4(c.Y1 + off);
    float y0[4] = {y0v.x, y0v.y, y0v.z, y0v.w};
    float y1[4] = {y1v.x, y1v.y, y1v.z, y1v.w};
    float out[4];
    float ymid[4] = {0.f, 0.f, 0.f, 0.f}, f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
    if (any_dense && tb.has_mid) {
      bool any;
      const float4 dtv = make_float4(dts[0], dts[1], dts[2], dts[3]);
      const float4 acc = wsum4(c.K, tb.bmid, s, off, any);
      const float4 ym = axpy4(y0v, dtv, acc);
      const float4 k0 = ld4(c.K[0] + off), kl = ld4(c.K[s - 1] + off);
      ymid[0] = ym.x; ymid[1] = ym.y; ymid[2] = ym.z; ymid[3] = ym.w;
      f0[0] = mul_(dts[0], k0.x); f0[1] = mul_(dts[1], k0.y); f0[2] = mul_(dts[2], k0.z); f0[3] = mul_(dts[3], k0.w);
      f1[0] = mul_(dts[0], kl.x); f1[1] = mul_(dts[1], kl.y); f1[2] = mul_(dts[2], kl.z); f1[3] = mul_(dts[3], kl.w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = y0[j];
      if (upd[j]) {
        v = y1[j];
        if (tev[j]) {
          const float x = xs[j];
          if (tb.has_mid) {
            // quartic through (y0, y1, f0, f1, y_mid); operation order mirrors oracle/dense_eval
            const float a = add_(sub_(mul_(2.0f, sub_(f1[j], f0[j])), mul_(8.0f, add_(y1[j], y0[j]))),
                                 mul_(16.0f, ymid[j]));
            const float b = sub_(add_(add_(sub_(mul_(5.0f, f0[j]), mul_(3.0f, f1[j])), mul_(18.0f, y0[j])),
                                      mul_(14.0f, y1[j])), mul_(32.0f, ymid[j]));
            const float cc = add_(sub_(sub_(sub_(f1[j], mul_(4.0f, f0[j])), mul_(11.0f, y0[j])),
                                       mul_(5.0f, y1[j])), mul_(16.0f, ymid[j]));
            v = add_(mul_(add_(mul_(add_(mul_(add_(mul_(a, x), b), x), cc), x), f0[j]), x), y0[j]);
          } else {
            v = add_(y0[j], mul_(x, sub_(y1[j], y0[j])));
          }
        }
      }
      out[j] = v;
    }
    st4(c.Y + off, make_float4(out[0], out[1], out[2], out[3]));
    if (tb.fsal) {
      const float4 kold = ld4(c.K[0] + off), kl = ld4(c.K[s - 1] + off);
      st4(c.K[0] + off, make_float4(upd[0] ? kl.x : kold.x, upd[1] ? kl.y : kold.y,
                                    upd[2] ? kl.z : kold.z, upd[3] ? kl.w : kold.w));
    }
  }
}

// Fixed-step commit: Y <- y0 + dt * sum b_j k_j  (every row).
template <int RT>
__device__ __forceinline__ void fixed_commit(Ctx<RT>& c) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  const int nvec = p.D * c.rq4;
  const float4 dt = ld4(c.rs.dt + 4 * c.rq);
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) {
    const size_t off = static_cast<size_t>(e) * 4;
    const float4 y0 = ld4(c.Y + off);
    bool any;
    const float4 acc = wsum4(c.K, p.tab.b, p.tab.n_stages, off, any);
    st4(c.Y + off, any ? axpy4(y0, dt, acc) : y0);
  }
}

// Per-row solver state for interval i; returns "row runs" for the CTA-wide OR.
template <int RT>
__device__ __forceinline__ int interval_begin(Ctx<RT>& c, int i) {
  const FwdParams& p = *c.prm;
  RowState& rs = c.rs;
  int run = 0;
  if (!c.th.producer && c.th.ctid < c.R) {
    const int r = c.th.ctid;
    const int b = c.tile * RT + (r % RT);
    float t0 = 0.f, t1 = 0.f;
    if (b < p.B && !p.skip_evolve) {
      int ll, bb;
      global_ids(p, r / RT, b, ll, bb);
      t0 = p.ts[static_cast<size_t>(bb) * p.ts_ld + p.i_off + i];
      t1 = p.ts[static_cast<size_t>(bb) * p.ts_ld + p.i_off + i + 1];
    }
    rs.t[r] = t0; rs.tend[r] = t1;
    rs.tmin[r] = fminf(t0, t1); rs.tmax[r] = fmaxf(t0, t1);
    rs.nsteps[r] = 0; rs.nacc[r] = 0;
    rs.upd[r] = 0; rs.toeval[r] = 0; rs.x[r] = 1.f;
    if (p.adaptive) {
      rs.dt[r] = fminf(fmaxf(p.dt0, sub_(rs.tmin[r], t0)), sub_(rs.tmax[r], t0));
      run = (t0 < t1) ? 1 : 0;
    } else {
      rs.dt[r] = __fdiv_rn(sub_(t1, t0), static_cast<float>(p.substeps));
      run = 1;
    }
    rs.dtstep[r] = rs.dt[r];
    if (p.skip_evolve) run = 0;          // jump + head only: the tensor-core solver kernel evolved the state (odernn_tc.cu)
    rs.run[r] = run; rs.noteval[r] = run;
  }
  return run;
}

// stats[S][L][B][2 + 2T] <- (n_steps, n_accepted) of interval i
template <int RT>
__device__ __forceinline__ void write_stats(Ctx<RT>& c, int i) {
  const FwdParams& p = *c.prm;
  if (c.th.producer || c.th.ctid >= c.R || !p.stats) return;
  const int r = c.th.ctid;
  const int l = r / RT, b = c.tile * RT + (r % RT);
  if (b < p.B) {
    int ll, bb;
    global_ids(p, l, b, ll, bb);
    const int LL_ = p.full_B > 0 ? p.full_L : p.L, BB_ = p.full_B > 0 ? p.full_B : p.B;
    int* sp = p.stats + ((static_cast<size_t>(i + (p.full_B > 0 ? p.i_off : 0)) * LL_ + ll) * BB_ + bb) * (2 + 2 * p.trace_steps);
    sp[0] = c.rs.nsteps[r]; sp[1] = c.rs.nacc[r];
  }
}

// [x ; h] for the jump of layer l -> bufA as T-layout [2D][RT]
template <int RT>
__device__ __forceinline__ void assemble_jump_input(Ctx<RT>& c, int i, int l) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  const int D = p.D;
  const TileThread& th = c.th;
  if (l == 0 && p.Wfuse) {
    // 'soft' fusion: PH_FUSE left cat * Linear(cat) in scratch K[4] ([D][RT])
    for (int e = th.ctid; e < D * RT / 4; e += th.ncons) st4(c.bufA + 4 * e, ld4(c.K[4] + 4 * e));
  } else if (l == 0) {
    for (int e = th.ctid; e < D * RT; e += th.ncons) {
      const int m = e / D, k = e - m * D;          // coalesced along k
      const int b = c.tile * RT + m;
      float v = 0.f;
      if (b < p.B) {
        const size_t row = static_cast<size_t>(b) * p.S_io + p.i_off + i;
        v = (k < p.Dv) ? p.fv[row * p.Dv + k] : p.fi[row * (D - p.Dv) + (k - p.Dv)];
      }
      c.bufA[k * RT + m] = v;
    }
  } else {
    for (int e = th.ctid; e < D * RT / 4; e += th.ncons) st4(c.bufA + 4 * e, ld4(c.bufB + 4 * e));
  }
  for (int e = th.ctid; e < D * RT / 4; e += th.ncons) {
    const int d = e / (RT / 4), q = e - d * (RT / 4);
    st4(c.bufA + static_cast<size_t>(D + d) * RT + 4 * q,
        ld4(c.Y + static_cast<size_t>(d) * c.R + l * RT + 4 * q));
  }
  named_bar_sync(1, th.ncons);
}

// final Linear(128, 6) of the pose head from bufA ([128][RT]) -> pose[b, i, :]
template <int RT>
__device__ __forceinline__ void pose_out(Ctx<RT>& c, int i) {
  if (c.th.producer) return;
  const FwdParams& p = *c.prm;
  if (c.th.ctid < RT * kPoseDim) {
    const int m = c.th.ctid / kPoseDim, o = c.th.ctid - m * kPoseDim;
    const int b = c.tile * RT + m;
    float acc = 0.f;
    for (int k = 0; k < kRegHidden; ++k) acc = fmaf(c.bufA[k * RT + m], p.Wreg1[o * kRegHidden + k], acc);
    if (b < p.B) p.pose[(static_cast<size_t>(b) * p.S_io + p.i_off + i) * kPoseDim + o] = acc + p.breg1[o];
  }
}

// Training checkpoints: copy a T-layout tile array (consumers; same element->thread mapping as the
// elementwise passes, so no barrier is needed against them).
template <int RT>
__device__ __forceinline__ void copy_tile(const Ctx<RT>& c, float* dst, const float* src) {
  if (c.th.producer) return;
  const int nvec = c.prm->D * c.rq4;
  for (int e = c.th.ctid; e < nvec; e += c.th.ncons) st4(dst + static_cast<size_t>(e) * 4, ld4(src + static_cast<size_t>(e) * 4));
}

// One GEMM of the tile program.
struct GemmOp {
  const float* W; int K; int N;
  const float* in; bool ode_layout;
  Epilogue epi;
};

// Phases of the per-tile program.  The whole forward is ONE loop whose body ends in the single
// (inlined) tile_gemm call site: every GEMM of the path -- ODEFunc layers of every RK stage, the
// RNN/GRU jump, the pose head -- is issued from there, so the hot loop is compiled once per
// column-pair count and scheduled by ptxas as straight-line kernel code.
enum { PH_INTERVAL = 0, PH_STEP_BEGIN, PH_STAGE, PH_LAYER, PH_STEP_END, PH_FUSE, PH_JUMP, PH_REG, PH_REG_OUT, PH_TILE_END };

}  // namespace

template <int RT, int LL>
__global__ void __launch_bounds__(128 * LL + 32, 1)
odernn_fwd_kernel(const __grid_constant__ FwdParams prm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ctx<RT> c;
  c.prm = &prm;
  const FwdParams& p = prm;
  const DevTableau& tb = prm.tab;
  const int tid = threadIdx.x;
  constexpr int ncons = 128 * LL;
  c.th.ncons = ncons;
  c.th.lane = tid & 31;
  c.th.producer = tid >= ncons;
  c.th.ctid = c.th.producer ? 0 : tid;
  constexpr int R = RT * LL;
  c.R = R;
  c.rq4 = R / 4;
  c.rq = c.th.ctid % (R / 4);

  // ---- shared memory carve-up (sizes mirrored by plan_odernn in api.cu)
  float* sm = reinterpret_cast<float*>(smem_raw);
  c.bufA = sm; sm += prm.bufA_floats;
  c.bufB = sm; sm += prm.bufB_floats;
  float* stages = sm; sm += static_cast<size_t>(prm.nst) * prm.stage_floats;
  c.partial = sm; sm += 4 * ncons;
  RowState& rs = c.rs;
  rs.t = sm; sm += R; rs.dt = sm; sm += R; rs.tend = sm; sm += R; rs.tmin = sm; sm += R;
  rs.tmax = sm; sm += R; rs.dtstep = sm; sm += R; rs.x = sm; sm += R;
  int* si = reinterpret_cast<int*>(sm);
  rs.run = si; si += R; rs.noteval = si; si += R; rs.upd = si; si += R; rs.toeval = si; si += R;
  rs.nsteps = si; si += R; rs.nacc = si; si += R; rs.status = si; si += R;
  uintptr_t bp = (reinterpret_cast<uintptr_t>(si) + 7) & ~static_cast<uintptr_t>(7);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bp);
  c.ring.buf = stages;
  c.ring.buf_off = static_cast<uint32_t>(reinterpret_cast<unsigned char*>(stages) - smem_raw);
  c.ring.full = bars;
  c.ring.empty = bars + MAX_STAGES;
  c.ring.stage_floats = prm.stage_floats;
  c.ring.nst = prm.nst;
  c.ring.kc = static_cast<uint32_t>(prm.kc);
  c.pos.stage = 0;
  c.pos.phase = 0;
  c.pos.ready = 0;
  if (tid == 0) {
    for (int s = 0; s < prm.nst; ++s) {
      mbar_init(&c.ring.full[s], 1);
      mbar_init(&c.ring.empty[s], ncons / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // ---- per-CTA scratch
  const size_t arr = static_cast<size_t>(prm.D) * R;
  float* sc = prm.scratch + static_cast<size_t>(blockIdx.x) * prm.scratch_floats_per_cta;
  for (int j = 0; j < kMaxStages; ++j) c.K[j] = sc + j * arr;
  c.Y = sc + kMaxStages * arr;
  c.Y1 = c.Y + arr;
  const int D = prm.D;
  const int gemms_per_layer = prm.rnn_type == 0 ? 1 : 4;

  for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
    c.tile = tile;
    // ---- load h0 -> Y (T-layout); rows past B are zero
    if (!c.th.producer) {
      for (int e = c.th.ctid; e < prm.D * R; e += ncons) {
        const int r = e / prm.D, d = e - r * prm.D;      // coalesced along d
        const int l = r / RT, b = tile * RT + (r % RT);
        float v = 0.f;
        if (prm.h0 && b < prm.B) {
          int ll, bb;
          global_ids(prm, l, b, ll, bb);
          v = prm.h0[(static_cast<size_t>(ll) * (prm.full_B > 0 ? prm.full_B : prm.B) + bb) * prm.D + d];
        }
        c.Y[static_cast<size_t>(d) * R + r] = v;
      }
      if (c.th.ctid < R) rs.status[c.th.ctid] = 0;
    }
    __syncthreads();

    int ph = PH_INTERVAL;
    int i = 0, st = 0, j = 0, l = 0, g = 0, loops = 0, any_running = 0, nsaved = 0;
    float* const ck_tile = prm.ckpt ? prm.ckpt + static_cast<size_t>(tile) * prm.ckpt_floats_per_tile : nullptr;
    float* ck_iv = nullptr;
    bool have_k0 = false;
    float* lin = c.bufA;      // ping-pong buffers of the vector-field MLP
    float* lout = c.bufB;
    while (ph != PH_TILE_END) {
      GemmOp op{};
      bool do_gemm = false;
      switch (ph) {
        case PH_INTERVAL: {
          const int run = interval_begin<RT>(c, i);
          any_running = __syncthreads_or(run);
          loops = 0; have_k0 = false;
          nsaved = 0;
          ck_iv = ck_tile ? ck_tile + static_cast<size_t>(i) * ckpt_interval_floats(D, R, p.CK) : nullptr;
          if (any_running) {
            ph = PH_STEP_BEGIN;
          } else {
            write_stats<RT>(c, i);
            l = 0; g = 0; ph = p.evolve_only ? ((++i < p.S) ? PH_INTERVAL : PH_TILE_END) : (p.Wfuse ? PH_FUSE : PH_JUMP);
          }
          break;
        }
        case PH_STEP_BEGIN:
          // first pass evaluates stage 0 (for FSAL methods: torchode's up-front evaluation)
          ++loops;
          st = (tb.fsal && have_k0) ? 1 : 0;
          ph = PH_STAGE;
          break;
        case PH_STAGE:
          stage_input<RT>(c, st);
          j = 0; lin = c.bufA; lout = c.bufB;
          ph = PH_LAYER;
          break;
        case PH_LAYER: {
          // ODEFunc layer j of stage st: lin -> lout (hidden) or -> K[st] (output layer, Tanh)
          op.W = p.Wode[j]; op.K = p.Kode[j]; op.N = p.Node[j];
          op.in = lin; op.ode_layout = true;
          op.epi.mode = EPI_STORE; op.epi.bias = p.bode[j];
          op.epi.out1 = nullptr; op.epi.ld1 = 0; op.epi.off1 = 0;
          op.epi.ld0 = R; op.epi.off0 = 0;
          if (j == p.NL - 1) { op.epi.act = ACT_TANH; op.epi.out0 = c.K[st]; }
          else { op.epi.act = p.act; op.epi.out0 = lout; }
          do_gemm = true;
          float* t = lin; lin = lout; lout = t;
          if (++j == p.NL) { ph = (++st < tb.n_stages) ? PH_STAGE : PH_STEP_END; }
          break;
        }
        case PH_STEP_END: {
          have_k0 = true;
          if (p.adaptive) {
            error_pass<RT>(c);
            const int run = controller<RT>(c, loops, i);
            any_running = __syncthreads_or(run);
            if (ck_iv) {
              const int upd = (!c.th.producer && c.th.ctid < R) ? rs.upd[c.th.ctid] : 0;
              if (__syncthreads_or(upd)) {            // iterations that reject every row leave no trace
                if (nsaved < p.CK) {
                  float* slot = ck_iv + 2 * arr + static_cast<size_t>(nsaved) * (arr + 2 * R);
                  copy_tile<RT>(c, slot, c.Y);
                  if (!c.th.producer && c.th.ctid < R) {
                    slot[arr + c.th.ctid] = rs.dtstep[c.th.ctid];
                    reinterpret_cast<int*>(slot + arr + R)[c.th.ctid] = rs.upd[c.th.ctid];
                  }
                  ++nsaved;
                } else if (!c.th.producer && c.th.ctid < R) {
                  rs.status[c.th.ctid] = max(rs.status[c.th.ctid], 3);     // checkpoint overflow
                }
              }
            }
            commit_pass<RT>(c);
          } else {
            if (ck_iv) {
              if (nsaved < p.CK) {
                float* slot = ck_iv + 2 * arr + static_cast<size_t>(nsaved) * (arr + 2 * R);
                copy_tile<RT>(c, slot, c.Y);
                if (!c.th.producer && c.th.ctid < R) {
                  slot[arr + c.th.ctid] = rs.dt[c.th.ctid];
                  reinterpret_cast<int*>(slot + arr + R)[c.th.ctid] = 1;
                }
                ++nsaved;
              } else if (!c.th.producer && c.th.ctid < R) {
                rs.status[c.th.ctid] = max(rs.status[c.th.ctid], 3);
              }
            }
            fixed_commit<RT>(c);
            if (!c.th.producer && c.th.ctid < R) { rs.nsteps[c.th.ctid] += 1; rs.nacc[c.th.ctid] += 1; }
            any_running = loops < p.substeps;
          }
          if (any_running) {
            ph = PH_STEP_BEGIN;
          } else {
            write_stats<RT>(c, i);
            __syncthreads();
            l = 0; g = 0; ph = p.evolve_only ? ((++i < p.S) ? PH_INTERVAL : PH_TILE_END) : (p.Wfuse ? PH_FUSE : PH_JUMP);
          }
          break;
        }
        case PH_FUSE: {
          // FusionModule 'soft' (src/models/FusionModule.py:20-23): x <- cat(fv, fi) * (W_f cat + b_f), in the kernel
          if (!c.th.producer) {
            for (int e = c.th.ctid; e < D * RT; e += ncons) {
              const int m = e / D, k = e - m * D;
              const int b = tile * RT + m;
              float v = 0.f;
              if (b < p.B) {
                const size_t row = static_cast<size_t>(b) * p.S_io + p.i_off + i;
                v = (k < p.Dv) ? p.fv[row * p.Dv + k] : p.fi[row * (D - p.Dv) + (k - p.Dv)];
              }
              c.bufA[k * RT + m] = v;
            }
            named_bar_sync(1, ncons);
          }
          op.W = p.Wfuse; op.K = D; op.N = D; op.in = c.bufA; op.ode_layout = false;
          op.epi.mode = EPI_MUL_IN; op.epi.bias = p.bfuse; op.epi.act = ACT_NONE;
          op.epi.hs = c.bufA; op.epi.ldh = RT; op.epi.offh = 0;
          op.epi.out0 = c.K[4]; op.epi.ld0 = RT; op.epi.off0 = 0;
          do_gemm = true;
          ph = PH_JUMP;
          break;
        }
        case PH_JUMP: {
          if (ck_iv && l == 0 && g == 0) {           // state at the end of the interval's solves
            copy_tile<RT>(c, ck_iv, c.Y);
            if (tid == 0) p.nloops[static_cast<size_t>(tile) * p.S + i] = nsaved;
          }
          // RNN: one GEMM per layer on [x ; h] (K = 2D).  GRU: r, z (K = 2D), hn (h half), new gate
          // (x half) -- PyTorch gate order (r, z, n).
          if (g == 0) assemble_jump_input<RT>(c, i, l);
          float* RG = c.K[1]; float* ZG = c.K[2]; float* HN = c.K[3];   // GRU scratch, [D][RT]
          Epilogue& e = op.epi;
          e.mode = EPI_STORE; e.ld0 = RT; e.off0 = 0; e.out1 = nullptr; e.ld1 = 0; e.off1 = 0;
          e.rg = e.zg = e.hn = e.hprev = nullptr;
          op.N = D; op.ode_layout = false; op.in = c.bufA; op.K = 2 * D;
          if (p.rnn_type == 0) {
            op.W = p.Wrnn[l][0]; e.bias = p.brnn[l][0]; e.act = ACT_TANH;
            e.out0 = c.bufB; e.out1 = c.Y; e.ld1 = R; e.off1 = l * RT;
          } else if (g == 0) {
            op.W = p.Wrnn[l][0]; e.bias = p.brnn[l][0]; e.act = ACT_SIGMOID; e.out0 = RG;
          } else if (g == 1) {
            op.W = p.Wrnn[l][1]; e.bias = p.brnn[l][1]; e.act = ACT_SIGMOID; e.out0 = ZG;
          } else if (g == 2) {
            op.W = p.Wrnn[l][3]; e.bias = p.brnn[l][3]; e.act = ACT_NONE; e.out0 = HN;
            op.K = D; op.in = c.bufA + static_cast<size_t>(D) * RT;
          } else {
            op.W = p.Wrnn[l][2]; e.bias = p.brnn[l][2]; e.mode = EPI_GRU_NEW;
            e.rg = RG; e.zg = ZG; e.hn = HN; e.hprev = c.bufA + static_cast<size_t>(D) * RT;
            e.out0 = c.bufB; e.out1 = c.Y; e.ld1 = R; e.off1 = l * RT;
            op.K = D;
          }
          do_gemm = true;
          if (++g == gemms_per_layer) { g = 0; if (++l == LL) ph = PH_REG; }
          break;
        }
        case PH_REG: {
          if (ck_iv) copy_tile<RT>(c, ck_iv + arr, c.Y);      // post-jump state
          // pose head on the top layer's output (bufB, [D][RT]): Linear(D,128) + LeakyReLU(0.1)
          op.W = p.Wreg0; op.K = D; op.N = kRegHidden; op.in = c.bufB; op.ode_layout = false;
          op.epi.mode = EPI_STORE; op.epi.bias = p.breg0; op.epi.act = ACT_LEAKY01;
          op.epi.out0 = c.bufA; op.epi.ld0 = RT; op.epi.off0 = 0;
          op.epi.out1 = nullptr; op.epi.ld1 = 0; op.epi.off1 = 0;
          do_gemm = true;
          ph = PH_REG_OUT;
          break;
        }
        case PH_REG_OUT:
          pose_out<RT>(c, i);
          __syncthreads();
          ph = (++i < p.S) ? PH_INTERVAL : PH_TILE_END;
          break;
        default:
          ph = PH_TILE_END;
          break;
      }
      if (do_gemm) tile_gemm<RT, LL, true>(c.ring, c.pos, c.th, op.W, op.K, op.N, op.in, op.ode_layout, op.epi);
    }

    // ---- final hidden state and status
    if (!c.th.producer) {
      for (int e = c.th.ctid; e < prm.D * R; e += ncons) {
        const int r = e / prm.D, d = e - r * prm.D;
        const int l2 = r / RT, b = tile * RT + (r % RT);
        if (b < prm.B) {
          int ll, bb;
          global_ids(prm, l2, b, ll, bb);
          prm.hT[(static_cast<size_t>(ll) * (prm.full_B > 0 ? prm.full_B : prm.B) + bb) * prm.D + d] = c.Y[static_cast<size_t>(d) * R + r];
        }
      }
      if (prm.status && c.th.ctid < RT) {
        const int b = tile * RT + c.th.ctid;
        int stt = 0;
        for (int l2 = 0; l2 < LL; ++l2) stt = max(stt, rs.status[l2 * RT + c.th.ctid]);
        if (b < prm.B) {
          if (prm.full_B > 0) {
            int ll, bb;
            global_ids(prm, 0, b, ll, bb);
            if (stt) atomicMax(prm.status + bb, stt);     // several (layer) rows of a sequence may live in different launches
          } else {
            prm.status[b] = stt;
          }
        }
      }
    }
    __syncthreads();
  }
}

template <int RT, int LL>
static cudaError_t launch_one(const FwdParams& prm, int grid, size_t smem_bytes, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(odernn_fwd_kernel<RT, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
  if (err != cudaSuccess) return err;
  odernn_fwd_kernel<RT, LL><<<grid, 128 * LL + 32, smem_bytes, stream>>>(prm);
  return cudaGetLastError();
}

// host-visible launcher (api.cu).  Supported: 4- and 8-row tiles for L = 1..4, 16-row tiles for L = 1..2.
cudaError_t launch_odernn_fwd(const FwdParams& prm, int rows_per_tile, int grid, size_t smem_bytes,
                              cudaStream_t stream) {
  if (rows_per_tile == 4) {
    switch (prm.L) {
      case 1: return launch_one<4, 1>(prm, grid, smem_bytes, stream);
      case 2: return launch_one<4, 2>(prm, grid, smem_bytes, stream);
      case 3: return launch_one<4, 3>(prm, grid, smem_bytes, stream);
      case 4: return launch_one<4, 4>(prm, grid, smem_bytes, stream);
    }
  } else if (rows_per_tile == 8) {
    switch (prm.L) {
      case 1: return launch_one<8, 1>(prm, grid, smem_bytes, stream);
      case 2: return launch_one<8, 2>(prm, grid, smem_bytes, stream);
      case 3: return launch_one<8, 3>(prm, grid, smem_bytes, stream);
      case 4: return launch_one<8, 4>(prm, grid, smem_bytes, stream);
    }
  } else if (rows_per_tile == 16) {
    switch (prm.L) {
      case 1: return launch_one<16, 1>(prm, grid, smem_bytes, stream);
      case 2: return launch_one<16, 2>(prm, grid, smem_bytes, stream);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace odevio
