// Tensor-core ODE solver: PoseODERNN.evolve_state (reference src/models/PoseODERNN.py:70-75) for every
// (sequence, rnn layer) row of one observation interval, with the ODEFunc GEMMs on tcgen05 (3xTF32,
// fp32-accurate -- ft_layer.cuh) and the whole solver loop of a 128-row tile inside ONE cluster of 8 CTAs:
// stage combines, Butcher tableau, per-row error norm, per-row step-size controller (torchode semantics as
// restated in oracle/torchode_like.py), FSAL -- no host round trip between solver steps.
//
//   row g = l * B + b of the [L, B, D] hidden state; a tile = 128 consecutive rows; the cluster's CTAs split every
//   Linear as an nN x nK grid (ft_layer) and every elementwise pass by FEATURE slice: CTA c owns the `own_nf`
//   features it finalises in the last Linear, for all 128 rows, thread = row.
//   * stage vectors K0..K6, Y, Y1 of the tile: per-cluster L2-resident scratch, feature-major [D][128] so that
//     thread = row accesses are coalesced; every element is only ever touched by its owning CTA;
//   * stage argument y + dt * sum_j a_ij k_j -> written straight into the tensor core's operand image of layer 0
//     (fp32, split hi/lo in shared memory by the splitter warps), fence.proxy.async + cluster barrier;
//   * error norm: thread-local sum over the CTA's features, one 128-float exchange per CTA through L2, summed in
//     CTA order by every CTA -> all 8 CTAs take identical controller decisions (each keeps the full row state in
//     shared memory), so the number of cluster barriers is the same everywhere.
// The jump (nn.RNN / nn.GRU step) and the pose head of the interval run in the FMA kernel with skip_evolve = 1
// (odernn_fwd.cu): M = B rows there, too few for the tensor core to matter (5 % of the FLOPs).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/odevio.h"
#include "common.cuh"
#include "ft_layer.cuh"
#include "odernn_params.h"
#include "odernn_tc.h"

namespace odevio {

namespace {

#ifdef ODEVIO_FT_TIMELINE
#define TC_STAMP(idx) do { if (blockIdx.x == 0 && tile == 0 && threadIdx.x == 0) g_ft_dbg[idx] = clock64(); } while (0)
#else
#define TC_STAMP(idx) do { } while (0)
#endif

// two instantiations (ft_layer.cuh): <8, 8, split> up to 16 tiles (latency: 8 CTAs per tile), <4, 16, pre-split> beyond
// (throughput: 37 co-resident clusters, 128..256-column MMAs, no k-split reduce)
constexpr int TC_MAX_NC = 8;
constexpr int TC_UNIT = 16;          // features per elementwise unit (thread = row, 4 float4 of the operand image)

struct TcParams {
  int M, B, L, D, NL, act;
  int K[FT_MAX_LAYERS], N[FT_MAX_LAYERS], nN[FT_MAX_LAYERS], nK[FT_MAX_LAYERS];
  const float* Wp[FT_MAX_LAYERS];
  const float* Wlo[FT_MAX_LAYERS];   // residual image (pre-split mode only)
  const float* bias[FT_MAX_LAYERS];
  float* part; size_t part_floats;
  float* xa; size_t xa_buf_floats;
  int ntiles, nraw;
  uint32_t raw_stage_bytes, op_stage_bytes;
  // solver
  DevTableau tab;
  int adaptive, substeps;
  float atol, rtol, dt0, safety, fmin, fmax;
  int accept_strict, floor_factor, max_steps, exact_landing;
  float* Y;                 // [L*B][D] row-major hidden state, evolved in place
  const int* seq; int Bsub; // this launch integrates the L * Bsub rows (l, j), j < Bsub, of sequences b = seq[j] (seq == nullptr:
                            // b = j); kernel row g = l * Bsub + j lives at state row l * B + b.  M = L * Bsub.
  const float* ts; int ts_ld, interval;      // row b: ts[b * ts_ld + interval] -> ts[.. + 1]
  int* stats;               // [S][L][B][2] (n_steps, n_accepted) or nullptr
  int* status;              // [B], max over layers
  float* state; size_t state_floats;         // per cluster: (kMaxStages + 2) x [D][128] + [TC_MAX_NC][128] norm partials
};

struct TcRows {      // per-row solver state, replicated in every CTA of the cluster (shared memory)
  float t[FT_ROWS], dt[FT_ROWS], tend[FT_ROWS], tmin[FT_ROWS], tmax[FT_ROWS];
  int run[FT_ROWS], upd[FT_ROWS], nsteps[FT_ROWS], nacc[FT_ROWS], status[FT_ROWS];
  float psum[2][FT_ROWS];
};

// ---- elementwise passes: thread = row, groups of TC_GRP consecutive features of the CTA's slice.
// All loads of a group (TC_GRP x (1 + N) coalesced 128-byte warp requests) are issued before any arithmetic and the
// arithmetic is branch-free: one L2 round trip per group instead of one per element (measured: 55 k -> see DESIGN).
// Weighted sums run left to right over j = 0..N-1 in the oracle's order (oracle/_weighted_sum, odernn_fwd.cu:wsum4);
// the oracle skips exactly-zero coefficients, here they contribute an exact +-0 (identical for finite stage values).
constexpr int TC_GRP = 8;

struct TcSlice {            // what a thread needs to walk its share of the CTA's feature slice
  float* base;              // per-cluster stage vectors: K[j] = base + j * arr, Y = base + 7 * arr, Y1 = base + 8 * arr
  size_t arr;
  int own_f0, ngroups, half, r;
};

template <int N>
__device__ __forceinline__ void tc_load_k(const TcSlice& sl, size_t off, float (&k)[N > 0 ? N : 1][TC_GRP]) {
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) k[j][q] = __ldcg(sl.base + j * sl.arr + off + static_cast<size_t>(q) * FT_ROWS);
}
template <int N>
__device__ __forceinline__ float tc_wsum(const float (&k)[N > 0 ? N : 1][TC_GRP], int q, const float (&cf)[kMaxStages]) {
  float acc = mul_(k[0][q], cf[0]);
#pragma unroll
  for (int j = 1; j < N; ++j) acc = add_(acc, mul_(k[j][q], cf[j]));
  return acc;
}

// stage argument y + dt * sum_{j<N} a_j k_j  ->  fp32 operand image of layer 0 (N = stage index)
template <int N, int KCH, bool SPLIT>
__device__ __forceinline__ void tc_stage_input(const TcSlice& sl, const float* coef, float dt, float* xa, size_t lo_off) {
  float cf[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) cf[j] = j < N ? coef[j] : 0.f;
  const float* Y = sl.base + kMaxStages * sl.arr;
  for (int grp = sl.half; grp < sl.ngroups; grp += 2) {
    const int f0 = sl.own_f0 + TC_GRP * grp;
    const size_t off = static_cast<size_t>(f0) * FT_ROWS + sl.r;
    float y[TC_GRP], k[N > 0 ? N : 1][TC_GRP];
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) y[q] = __ldcg(Y + off + static_cast<size_t>(q) * FT_ROWS);
    tc_load_k<N>(sl, off, k);
    if (N > 0) {
#pragma unroll
      for (int q = 0; q < TC_GRP; ++q) y[q] = add_(y[q], mul_(dt, tc_wsum<N>(k, q, cf)));
    }
#pragma unroll
    for (int q = 0; q < TC_GRP / 4; ++q) {
      const size_t o = xa_offset<KCH>(sl.r, f0 + 4 * q);
      if (SPLIT) {           // one fp32 image, split into hi / lo in shared memory by the splitter warps
        *reinterpret_cast<float4*>(xa + o) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      } else {               // pre-split operand images (TF32-exact high part | exact residual)
        const float4 h = make_float4(ft_hi(y[4 * q]), ft_hi(y[4 * q + 1]), ft_hi(y[4 * q + 2]), ft_hi(y[4 * q + 3]));
        *reinterpret_cast<float4*>(xa + o) = h;
        *reinterpret_cast<float4*>(xa + lo_off + o) =
            make_float4(y[4 * q] - h.x, y[4 * q + 1] - h.y, y[4 * q + 2] - h.z, y[4 * q + 3] - h.w);
      }
    }
  }
}

// y1 -> Y1 and this thread's share of sum_d (err_d / bound_d)^2 (odernn_fwd.cu:error_pass); NS = tableau stages.
// The stage vectors are loaded once and feed both weighted sums.
template <int NS>
__device__ __forceinline__ float tc_error_pass(const TcSlice& sl, const DevTableau& tb, float dt, float atol, float rtol) {
  float cy[kMaxStages], ce[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) {
    cy[j] = j < NS ? (tb.ssal ? (j < NS - 1 ? tb.a[NS - 1][j] : 0.f) : tb.b[j]) : 0.f;
    ce[j] = j < NS ? tb.e[j] : 0.f;
  }
  const float* Y = sl.base + kMaxStages * sl.arr;
  float* Y1 = sl.base + (kMaxStages + 1) * sl.arr;
  float sum = 0.f;
  for (int grp = sl.half; grp < sl.ngroups; grp += 2) {
    const size_t off = static_cast<size_t>(sl.own_f0 + TC_GRP * grp) * FT_ROWS + sl.r;
    float y0[TC_GRP], k[NS][TC_GRP];
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) y0[q] = __ldcg(Y + off + static_cast<size_t>(q) * FT_ROWS);
    tc_load_k<NS>(sl, off, k);
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) {
      const float y1 = add_(y0[q], mul_(dt, tc_wsum<NS>(k, q, cy)));
      __stcg(Y1 + off + static_cast<size_t>(q) * FT_ROWS, y1);
      if (tb.has_err) {
        const float e = mul_(dt, tc_wsum<NS>(k, q, ce));
        const float bound = add_(atol, mul_(rtol, fmaxf(fabsf(y0[q]), fabsf(y1))));
        const float q2 = __fdiv_rn(e, bound);
        sum = fmaf(q2, q2, sum);
      }
    }
  }
  return sum;
}

// fixed step: Y <- y0 + dt * sum b_j k_j (odernn_fwd.cu:fixed_commit)
template <int NS>
__device__ __forceinline__ void tc_fixed_commit(const TcSlice& sl, const DevTableau& tb, float dt) {
  float cb[kMaxStages];
#pragma unroll
  for (int j = 0; j < kMaxStages; ++j) cb[j] = j < NS ? tb.b[j] : 0.f;
  float* Y = sl.base + kMaxStages * sl.arr;
  for (int grp = sl.half; grp < sl.ngroups; grp += 2) {
    const size_t off = static_cast<size_t>(sl.own_f0 + TC_GRP * grp) * FT_ROWS + sl.r;
    float y0[TC_GRP], k[NS][TC_GRP];
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) y0[q] = __ldcg(Y + off + static_cast<size_t>(q) * FT_ROWS);
    tc_load_k<NS>(sl, off, k);
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q)
      __stcg(Y + off + static_cast<size_t>(q) * FT_ROWS, add_(y0[q], mul_(dt, tc_wsum<NS>(k, q, cb))));
  }
}

// accepted rows: Y <- Y1, FSAL carry K0 <- K[ns-1]
__device__ __forceinline__ void tc_commit(const TcSlice& sl, int ns, int fsal) {
  float* Y = sl.base + kMaxStages * sl.arr;
  const float* Y1 = sl.base + (kMaxStages + 1) * sl.arr;
  const float* Kl = sl.base + static_cast<size_t>(ns - 1) * sl.arr;
  for (int grp = sl.half; grp < sl.ngroups; grp += 2) {
    const size_t off = static_cast<size_t>(sl.own_f0 + TC_GRP * grp) * FT_ROWS + sl.r;
    float a[TC_GRP], b[TC_GRP];
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) {
      a[q] = __ldcg(Y1 + off + static_cast<size_t>(q) * FT_ROWS);
      b[q] = fsal ? __ldcg(Kl + off + static_cast<size_t>(q) * FT_ROWS) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < TC_GRP; ++q) {
      __stcg(Y + off + static_cast<size_t>(q) * FT_ROWS, a[q]);
      if (fsal) __stcg(sl.base + off + static_cast<size_t>(q) * FT_ROWS, b[q]);
    }
  }
}

#define TC_DISPATCH_STAGES(n, CALL)                                                                                 \
  switch (n) {                                                                                                      \
    case 1: { constexpr int NSV = 1; CALL; break; }                                                                 \
    case 2: { constexpr int NSV = 2; CALL; break; }                                                                 \
    case 3: { constexpr int NSV = 3; CALL; break; }                                                                 \
    case 4: { constexpr int NSV = 4; CALL; break; }                                                                 \
    case 5: { constexpr int NSV = 5; CALL; break; }                                                                 \
    case 6: { constexpr int NSV = 6; CALL; break; }                                                                 \
    default: { constexpr int NSV = 7; CALL; break; }                                                                \
  }

template <int NC, int KCH, bool SPLIT>
__global__ void __launch_bounds__(FT_THREADS, 1)
odernn_tc_evolve_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t raw_full[FT_RAW_STAGES];
  __shared__ __align__(8) uint64_t raw_empty[FT_RAW_STAGES];
  __shared__ __align__(8) uint64_t op_ready[FT_OP_STAGES];
  __shared__ __align__(8) uint64_t op_empty[FT_OP_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ TcRows rs;

  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x / NC, nclusters = gridDim.x / NC;
  const DevTableau& tb = p.tab;

  if (tid == 0) {
    for (int i = 0; i < FT_RAW_STAGES; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 1); }
    for (int i = 0; i < FT_OP_STAGES; ++i) { mbar_init(&op_ready[i], ft_op_ready_arrivals(SPLIT)); mbar_init(&op_empty[i], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == FT_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  float* xa_cluster = p.xa + static_cast<size_t>(cluster_id) * 4 * p.xa_buf_floats;
  FtCtx c;
  c.smem = smem; c.op_base = smem + static_cast<size_t>(p.nraw) * p.raw_stage_bytes;
  c.raw_full = raw_full; c.raw_empty = raw_empty; c.op_ready = op_ready; c.op_empty = op_empty; c.accum_bar = &accum_bar;
  c.tmem_d = tmem_d; c.crank = crank; c.nraw = static_cast<uint32_t>(p.nraw);
  c.raw_stage_bytes = p.raw_stage_bytes; c.op_stage_bytes = p.op_stage_bytes; c.xa_buf_floats = p.xa_buf_floats;
  c.part = p.part + static_cast<size_t>(cluster_id) * p.part_floats;
  c.g0 = 0; c.accum_phase = 0; c.tile = 0;

  // ---- feature slice owned by this CTA in every elementwise pass = the columns it finalises in the last Linear
  const int D = p.D;
  int own_f0, own_nf;
  {
    const int l = p.NL - 1, nN = p.nN[l], nK = p.nK[l];
    const int cn = static_cast<int>(crank) % nN, ck = static_cast<int>(crank) / nN;
    const int Nc = p.N[l] / nN;
    own_nf = Nc / nK;
    own_f0 = cn * Nc + ck * own_nf;
  }
  const int nunits = own_nf / TC_UNIT;
  const size_t arr = static_cast<size_t>(D) * FT_ROWS;
  float* const st_base = p.state + static_cast<size_t>(cluster_id) * p.state_floats;
  float* const Yc = st_base + kMaxStages * arr;
  float* const Y1 = Yc + arr;
  float* const normpart = Y1 + arr;          // [NC][128]

  const bool epi = warp < FT_EPI_WARPS;
  const int r = tid & (FT_ROWS - 1), half = (tid >> 7) & 1;
  TcSlice sl;
  sl.base = st_base; sl.arr = arr; sl.own_f0 = own_f0; sl.ngroups = own_nf / TC_GRP; sl.half = half; sl.r = r;
  const int ns = tb.n_stages;

  for (int tile = cluster_id; tile < p.ntiles; tile += nclusters) {
    c.tile = tile;
    const int row0 = tile * FT_ROWS;
    const int g = row0 + r;                        // global row of this thread (epilogue warps)
    const bool valid = g < p.M;
    const int lyr = valid ? g / p.Bsub : 0, jseq = valid ? g - lyr * p.Bsub : 0;
    const int b = valid ? (p.seq ? p.seq[jseq] : jseq) : 0;
    const size_t grow = static_cast<size_t>(lyr) * p.B + b;      // state row of this thread's kernel row

    // ---- load the tile's state (row-major [M][D]) into the feature-major scratch, own slice
    if (epi) {
      for (int u = half; u < nunits; u += 2) {
        const int f0 = own_f0 + TC_UNIT * u;
        float v[TC_UNIT];
        if (valid) {
          const float4* src = reinterpret_cast<const float4*>(p.Y + grow * D + f0);
#pragma unroll
          for (int q = 0; q < TC_UNIT / 4; ++q) { const float4 t4 = src[q]; v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w; }
        } else {
#pragma unroll
          for (int q = 0; q < TC_UNIT; ++q) v[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < TC_UNIT; ++q) __stcg(Yc + static_cast<size_t>(f0 + q) * FT_ROWS + r, v[q]);
      }
    }
    // ---- per-row solver state (PoseODERNN.py:70-75; odernn_fwd.cu:interval_begin)
    int run = 0;
    if (tid < FT_ROWS) {
      float t0 = 0.f, t1 = 0.f;
      if (valid) {
        t0 = p.ts[static_cast<size_t>(b) * p.ts_ld + p.interval];
        t1 = p.ts[static_cast<size_t>(b) * p.ts_ld + p.interval + 1];
      }
      rs.t[r] = t0; rs.tend[r] = t1;
      rs.tmin[r] = fminf(t0, t1); rs.tmax[r] = fmaxf(t0, t1);
      rs.nsteps[r] = 0; rs.nacc[r] = 0; rs.upd[r] = 0; rs.status[r] = 0;
      if (p.adaptive) {
        rs.dt[r] = fminf(fmaxf(p.dt0, sub_(rs.tmin[r], t0)), sub_(rs.tmax[r], t0));
        run = (valid && t0 < t1) ? 1 : 0;
      } else {
        rs.dt[r] = __fdiv_rn(sub_(t1, t0), static_cast<float>(p.substeps));
        run = valid ? 1 : 0;
      }
      rs.run[r] = run;
    }
    int any_running = __syncthreads_or(run);
    int loops = 0;
    bool have_k0 = false;

    while (any_running) {
      ++loops;
      TC_STAMP(40);
      for (int st = (tb.fsal && have_k0) ? 1 : 0; st < ns; ++st) {
        // ---- stage argument -> layer-0 operand image (own feature slice, all 128 rows)
        TC_STAMP(41);
        if (epi) {
          const float dt = rs.dt[r];
          if (st == 0) tc_stage_input<0, KCH, SPLIT>(sl, tb.a[0], dt, xa_cluster, p.xa_buf_floats);
          else TC_DISPATCH_STAGES(st, (tc_stage_input<(NSV < kMaxStages ? NSV : kMaxStages - 1), KCH, SPLIT>(sl, tb.a[st], dt, xa_cluster, p.xa_buf_floats)))
          TC_STAMP(42);
          asm volatile("fence.proxy.async.global;" ::: "memory");     // generic-proxy stores -> bulk-copy (async) proxy of all CTAs
          TC_STAMP(43);
        }
        __syncwarp();
        cluster_sync_all();
        TC_STAMP(44);
        // ---- ODEFunc (src/models/ODEFunc.py:38-39) on the tensor cores; last Linear (+ Tanh) -> K[st]
        for (int l = 0; l < p.NL; ++l) {
          FtLayer Ld;
          Ld.K = p.K[l]; Ld.N = p.N[l]; Ld.nN = p.nN[l]; Ld.nK = p.nK[l]; Ld.stamp = l;
          const bool last = l == p.NL - 1;
          Ld.act = last ? ACT_TANH : p.act;
          Ld.out_mode = last ? FT_OUT_FEATURE_MAJOR : FT_OUT_OPERAND;
          Ld.Wp = p.Wp[l]; Ld.Wlo = p.Wlo[l]; Ld.bias = p.bias[l];
          Ld.a_src = xa_cluster + static_cast<size_t>(l & 1) * 2 * p.xa_buf_floats;
          Ld.nx = xa_cluster + static_cast<size_t>((l + 1) & 1) * 2 * p.xa_buf_floats;
          Ld.out = st_base + static_cast<size_t>(st) * arr; Ld.M = p.M; Ld.row0 = row0;
          ft_layer<NC, KCH, SPLIT>(c, Ld);
          TC_STAMP(52 + l);
        }
        TC_STAMP(45);
      }
      have_k0 = true;

      if (p.adaptive) {
        // ---- y1, embedded error, this CTA's share of the per-row error norm (odernn_fwd.cu:error_pass)
        if (epi) {
          const float dt = rs.dt[r];
          float sum = 0.f;
          TC_DISPATCH_STAGES(ns, (sum = tc_error_pass<NSV>(sl, tb, dt, p.atol, p.rtol)))
          rs.psum[half][r] = sum;
          TC_STAMP(46);
          named_bar_sync(1, FT_EPI_WARPS * 32);
          if (tid < FT_ROWS) __stcg(normpart + static_cast<size_t>(crank) * FT_ROWS + r, add_(rs.psum[0][r], rs.psum[1][r]));
        }
        __syncwarp();
        TC_STAMP(47);
        cluster_sync_all();
        TC_STAMP(48);
        // ---- per-row controller, identical in every CTA (odernn_fwd.cu:controller; torchode IntegralController)
        run = 0;
        if (tid < FT_ROWS) {
          run = rs.run[r];
          const float dt = rs.dt[r];
          float t = rs.t[r];
          const float tend = rs.tend[r];
          bool accept = true, finite = true;
          float dt_next = dt;
          if (tb.has_err) {
            float total = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) total = add_(total, __ldcg(normpart + k * FT_ROWS + r));
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
          t = upd ? (lands ? tend : add_(t, dt)) : t;
          rs.upd[r] = upd;
          if (run && !finite) rs.status[r] = max(rs.status[r], 2);
          run = (run && t < tend && finite) ? 1 : 0;
          if (run && loops >= p.max_steps) { rs.status[r] = max(rs.status[r], 1); run = 0; }
          float dtn = run ? dt_next : dt;
          dtn = fminf(fmaxf(dtn, sub_(rs.tmin[r], t)), sub_(rs.tmax[r], t));
          rs.t[r] = t;
          rs.dt[r] = dtn;
          rs.run[r] = run;
        }
        any_running = __syncthreads_or(run);
        TC_STAMP(49);
        // ---- commit: accepted rows take y1; FSAL carry (end point rule "y1": exact landing makes y1 the value at t_end)
        if (epi && rs.upd[r]) tc_commit(sl, ns, tb.fsal);
        TC_STAMP(50);
      } else {
        // ---- fixed step: Y <- y0 + dt * sum b_j k_j for every row (odernn_fwd.cu:fixed_commit)
        if (epi) {
          const float dt = rs.dt[r];
          TC_DISPATCH_STAGES(ns, (tc_fixed_commit<NSV>(sl, tb, dt)))
          if (tid < FT_ROWS && valid) { rs.nsteps[r] += 1; rs.nacc[r] += 1; }
        }
        any_running = loops < p.substeps;
      }
    }

    // ---- evolved state back to [M][D]; stats / status of the interval
    __syncthreads();
    if (epi && valid) {
      for (int u = half; u < nunits; u += 2) {
        const int f0 = own_f0 + TC_UNIT * u;
        float v[TC_UNIT];
#pragma unroll
        for (int q = 0; q < TC_UNIT; ++q) v[q] = __ldcg(Yc + static_cast<size_t>(f0 + q) * FT_ROWS + r);
        float4* dst = reinterpret_cast<float4*>(p.Y + grow * D + f0);
#pragma unroll
        for (int q = 0; q < TC_UNIT / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (crank == 0 && tid < FT_ROWS) {
        if (p.stats) {
          int* sp = p.stats + ((static_cast<size_t>(p.interval) * p.L + lyr) * p.B + b) * 2;
          sp[0] = rs.nsteps[r]; sp[1] = rs.nacc[r];
        }
        if (p.status && rs.status[r]) atomicMax(p.status + b, rs.status[r]);
      }
    }
    __syncthreads();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  cluster_sync_all();
  if (warp == FT_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side

// ---- per-interval sequence order for the row split (api.cu): seq[i][0 .. B - n_side) = the sequences the cluster
// kernel integrates (increasing index), seq[i][B - n_side .. B) = the n_side sequences with the SHORTEST interval i
// (ties by index): they need the fewest solver steps and go to the slower FFMA side launch.  One CTA per interval,
// rank by counting (B <= 8192: ~10 us at B = 1024).
__global__ void tc_select_kernel(const float* __restrict__ ts, int B, int S, int n_side, int* __restrict__ seq) {
  extern __shared__ unsigned char sel_smem[];
  float* dt = reinterpret_cast<float*>(sel_smem);
  int* side = reinterpret_cast<int*>(dt + B);
  const int i = blockIdx.x;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    dt[b] = ts[static_cast<size_t>(b) * (S + 1) + i + 1] - ts[static_cast<size_t>(b) * (S + 1) + i];
  __syncthreads();
  int* out = seq + static_cast<size_t>(i) * B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float d = dt[b];
    int rank = 0;
    for (int o = 0; o < B; ++o) { const float e = dt[o]; rank += (e < d || (e == d && o < b)) ? 1 : 0; }
    side[b] = rank < n_side ? 1 : 0;
    if (rank < n_side) out[B - n_side + rank] = b;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (side[b]) continue;
    int before = 0;
    for (int o = 0; o < b; ++o) before += side[o];
    out[b - before] = b;
  }
}

int odernn_tc_select(const float* ts, int B, int S, int n_side, int* seq, cudaStream_t stream) {
  if (B > 8192 || n_side <= 0 || n_side >= B) return ODEVIO_E_SHAPE;
  const size_t smem = static_cast<size_t>(B) * 8;
  cudaError_t e = cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_select_kernel<<<S, 1024, smem, stream>>>(ts, B, S, n_side, seq);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

static int g_tc_last_clusters = 0, g_tc_last_max_clusters = 0, g_tc_last_rows = 0;      // development: geometry of the last launch

// Measurement hook (bench.py): CUDA events around every solver launch on its own stream, so that the kernel's
// average launch duration is measured live, un-profiled, inside a normal forward (roofline.achieved).
constexpr int kTimingSlots = 512;
static bool g_tc_timing = false;
static int g_tc_timing_n = 0;
static cudaEvent_t g_tc_ev[kTimingSlots][2];
static bool g_tc_ev_made = false;
void odernn_tc_timing_enable(bool on) {
  g_tc_timing = on; g_tc_timing_n = 0;
  if (on && !g_tc_ev_made) {
    for (int i = 0; i < kTimingSlots; ++i) { cudaEventCreate(&g_tc_ev[i][0]); cudaEventCreate(&g_tc_ev[i][1]); }
    g_tc_ev_made = true;
  }
}
int odernn_tc_timing_read(float* total_ms, int* launches) {
  float tot = 0.f;
  for (int i = 0; i < g_tc_timing_n; ++i) {
    cudaError_t e = cudaEventSynchronize(g_tc_ev[i][1]);
    if (e != cudaSuccess) return static_cast<int>(e);
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_tc_ev[i][0], g_tc_ev[i][1]);
    if (e != cudaSuccess) return static_cast<int>(e);
    tot += ms;
  }
  *total_ms = tot; *launches = g_tc_timing_n;
  g_tc_timing_n = 0;
  return 0;
}
int odernn_tc_debug_timeline(long long* host_dst) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_dst, g_ft_dbg, sizeof(long long) * 64));
}
void odernn_tc_last_geometry(int* clusters, int* max_clusters, int* rows) {
  *clusters = g_tc_last_clusters; *max_clusters = g_tc_last_max_clusters; *rows = g_tc_last_rows;
}

struct TcEvolve::Impl {
  FtPlan pl;
  TcParams prm;
  size_t state_floats, off_state, total_bytes;
  int maxc = 0;
};

static int tc_plan(const odevio_odernn_cfg& c, FtPlan& pl, size_t& off_state, size_t& state_floats, size_t& total_bytes) {
  const long long M = static_cast<long long>(c.L) * c.B;
  if (M <= 0 || M > 0x7fffffffLL) return ODEVIO_E_SHAPE;
  // up to 16 tiles (configs[1]: 2048 rows): latency matters, 8 CTAs per tile; beyond: clusters of 4 with wide pre-split MMAs
  const int ntiles = static_cast<int>((M + FT_ROWS - 1) / FT_ROWS);
  // (shapes one instantiation cannot slice -- e.g. H = 128 with 8 CTAs -- take the other one)
  // development override: ODEVIO_TC_MODE=1 (clusters of 8) / 2 (clusters of 4) regardless of the tile count
  const char* force = getenv("ODEVIO_TC_MODE");
  const int first = (force && (force[0] == '1' || force[0] == '2')) ? force[0] - '0' : (ntiles <= 16 ? 1 : 2);
  int rc = ODEVIO_E_SHAPE;
  for (int attempt = 0; attempt < 2 && rc != 0; ++attempt) {
    rc = ft_plan(static_cast<int>(M), c.D, c.H, c.n_hidden, pl, attempt == 0 ? first : 3 - first);
    if (rc == 0) {
      // elementwise groups of 8 features inside the slice every CTA finalises in the last Linear, two thread halves
      const int l = pl.NL - 1;
      const int own_nf = pl.N[l] / pl.nN[l] / pl.nK[l];
      if (own_nf % (2 * TC_GRP) || c.D % 4) rc = ODEVIO_E_SHAPE;
    }
  }
  if (rc != 0) return rc;
  state_floats = (static_cast<size_t>(kMaxStages + 2) * c.D + TC_MAX_NC) * FT_ROWS;
  state_floats = (state_floats + 63) / 64 * 64;
  off_state = (pl.total_bytes / sizeof(float) + 63) / 64 * 64;
  total_bytes = (off_state + state_floats * static_cast<size_t>(pl.nclusters)) * sizeof(float);
  return 0;
}

size_t odernn_tc_workspace_bytes(const odevio_odernn_cfg& c) {
  FtPlan pl;
  size_t off_state, state_floats, total;
  if (tc_plan(c, pl, off_state, state_floats, total) != 0) return 0;
  return total;
}

TcEvolve::TcEvolve() : impl(nullptr) {}
TcEvolve::~TcEvolve() { delete impl; }

int TcEvolve::prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const float* const* ode_w,
                      const float* const* ode_b, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  delete impl;
  impl = new Impl();
  int rc = tc_plan(c, impl->pl, impl->off_state, impl->state_floats, impl->total_bytes);
  if (rc != 0) return rc;
  if (workspace_bytes < impl->total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  FtPlan& pl = impl->pl;
  float* ws = static_cast<float*>(workspace);
  TcParams& p = impl->prm;
  memset(&p, 0, sizeof(p));
  p.M = c.L * c.B; p.B = c.B; p.L = c.L; p.D = c.D; p.NL = pl.NL; p.act = c.activation;
  for (int l = 0; l < pl.NL; ++l) {
    if (!ode_w[l] || !ode_b[l]) return ODEVIO_E_NULL;
    p.K[l] = pl.K[l]; p.N[l] = pl.N[l]; p.nN[l] = pl.nN[l]; p.nK[l] = pl.nK[l];
    ft_pack_weight_kernel<<<296, 256, 0, stream>>>(ode_w[l], pl.N[l], pl.K[l], pl.nN[l], pl.KCH, ws + pl.off_w[l],
                                                   ws + pl.off_wlo[l]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    p.Wp[l] = ws + pl.off_w[l]; p.Wlo[l] = ws + pl.off_wlo[l]; p.bias[l] = ode_b[l];
  }
  p.xa = ws + pl.off_xa; p.xa_buf_floats = pl.xa_buf_floats;
  p.part = ws + pl.off_part; p.part_floats = pl.part_floats;
  p.ntiles = pl.ntiles; p.nraw = pl.nraw; p.raw_stage_bytes = pl.raw_stage_bytes; p.op_stage_bytes = pl.op_stage_bytes;
  p.tab = tab; p.adaptive = adaptive ? 1 : 0; p.substeps = c.substeps;
  p.atol = c.atol; p.rtol = c.rtol; p.dt0 = c.dt0; p.safety = c.safety; p.fmin = c.factor_min; p.fmax = c.factor_max;
  p.accept_strict = c.accept_strict; p.floor_factor = c.floor_factor; p.max_steps = c.max_steps; p.exact_landing = c.exact_landing;
  p.state = ws + impl->off_state; p.state_floats = impl->state_floats;
  return 0;
}

typedef void (*TcKernel)(TcParams);
static TcKernel tc_kernel_of(const FtPlan& pl) {
  return pl.NC == 8 ? static_cast<TcKernel>(odernn_tc_evolve_kernel<8, 8, true>)
                    : static_cast<TcKernel>(odernn_tc_evolve_kernel<4, 16, false>);
}
static cudaError_t tc_launch_config(const FtPlan& pl, cudaLaunchConfig_t& lc, cudaLaunchAttribute& at, int nclusters,
                                    cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(tc_kernel_of(pl), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem_bytes));
  if (e != cudaSuccess) return e;
  memset(&lc, 0, sizeof(lc));
  lc.blockDim = dim3(FT_THREADS); lc.dynamicSmemBytes = pl.smem_bytes; lc.stream = stream;
  at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = pl.NC; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  lc.attrs = &at; lc.numAttrs = 1;
  lc.gridDim = dim3(nclusters * pl.NC);
  return cudaSuccess;
}

int TcEvolve::cluster_size() const { return impl ? impl->pl.NC : 0; }

int TcEvolve::max_clusters() {
  if (!impl) return 0;
  if (impl->maxc > 0) return impl->maxc;
  cudaLaunchConfig_t lc; cudaLaunchAttribute at;
  if (tc_launch_config(impl->pl, lc, at, impl->pl.nclusters, nullptr) != cudaSuccess) { cudaGetLastError(); return impl->pl.nclusters; }
  int maxc = 0;
  if (cudaOccupancyMaxActiveClusters(&maxc, tc_kernel_of(impl->pl), &lc) != cudaSuccess || maxc <= 0) {
    cudaGetLastError();
    maxc = impl->pl.nclusters;
  }
  if (maxc > impl->pl.nclusters) maxc = impl->pl.nclusters;       // scratch is sized for pl.nclusters
  impl->maxc = maxc;
  return maxc;
}

int TcEvolve::evolve(float* Y, int Bsub, const int* seq, const float* ts, int ts_ld, int interval, int* stats, int* status,
                     cudaStream_t stream) {
  if (!impl) return ODEVIO_E_NULL;
  TcParams p = impl->prm;
  FtPlan& pl = impl->pl;
  if (Bsub <= 0 || Bsub > p.B) return ODEVIO_E_SHAPE;
  const int rows = p.L * Bsub;
  p.M = rows; p.ntiles = (rows + FT_ROWS - 1) / FT_ROWS;
  p.Bsub = Bsub; p.seq = seq;
  p.Y = Y; p.ts = ts; p.ts_ld = ts_ld; p.interval = interval; p.stats = stats; p.status = status;
  int nclusters = max_clusters();
  if (nclusters > p.ntiles) nclusters = p.ntiles;
  cudaLaunchConfig_t lc; cudaLaunchAttribute at;
  cudaError_t e = tc_launch_config(pl, lc, at, nclusters, stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  g_tc_last_clusters = nclusters; g_tc_last_max_clusters = impl->maxc; g_tc_last_rows = rows;
  const bool timed = g_tc_timing && g_tc_timing_n < kTimingSlots;
  if (timed) cudaEventRecord(g_tc_ev[g_tc_timing_n][0], stream);
  e = cudaLaunchKernelEx(&lc, tc_kernel_of(pl), p);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (timed) { cudaEventRecord(g_tc_ev[g_tc_timing_n][1], stream); ++g_tc_timing_n; }
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace odevio
