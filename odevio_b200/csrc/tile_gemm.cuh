// Streaming tile GEMM used by every fused kernel on the path.
//
//   out[r][n] = epilogue( sum_k in[r][k] * W[n][k] )     for the rows of ONE sequence tile
//
// * activations live in shared memory in "T-layout": inT[k * ld + r]  (k-major, the tile's rows
//   contiguous), so one broadcast LDS.128 feeds 4 rows of the thread tile;
// * weights are pre-packed K-major (Wt[k][n], see prepack.cu) so that KC consecutive k-rows are
//   ONE contiguous block: a dedicated producer lane streams them L2 -> shared memory with 1-D
//   bulk TMA (cp.async.bulk, SASS UBLKCP) through an mbarrier full/empty ring;
// * each consumer thread owns RT rows x P column-pairs (strided by the threads of its row
//   block) and accumulates in fp32 registers with FFMA, sequentially over k (deterministic).
//
// The producer only touches the ring, so it runs ahead of the consumers across GEMM
// boundaries; control flow is identical for all threads of the CTA.
#pragma once
#include "common.cuh"

namespace odevio {

constexpr int KC = 8;           // k-rows per weight stage; kernels that opt in (tile_gemm<.., ALLOW16>) also run 16-row stages
                                // (ring.kc = 16: half the mbarrier waits / arrives per k and a longer software pipeline:
                                // measured +7 % on the forward, +5 % on the training step)

// tuning switches (A/B-tested on the B200, see profiles/)
#ifndef ODEVIO_PRODUCER_WAIT
#define ODEVIO_PRODUCER_WAIT 0      // 0 spin, 1 nanosleep back-off, 2 hardware-parked try_wait
#endif
#ifndef ODEVIO_EARLY_PROBE
#define ODEVIO_EARLY_PROBE 0        // probe the next stage's full barrier one chunk ahead (no gain at 8-row
                                    // tiles, -28 % at 16-row tiles on B200: profiles/r01_ab_tuning.md)
#endif
constexpr int MAX_STAGES = 4;
constexpr int MAX_P = 4;        // column pairs per thread -> N <= 2 * MAX_P * threads_per_row_block

// Ring geometry (uniform, lives in registers / constant bank).
struct WeightRing {
  float* buf;            // shared: nst * stage_floats
  uint32_t buf_off;      // byte offset of buf from the dynamic shared-memory base
  uint64_t* full;        // [nst] armed by the producer, completed by TMA bytes
  uint64_t* empty;       // [nst] one arrive per consumer warp
  uint32_t stage_floats;
  uint32_t nst;
  uint32_t kc;           // k-rows per stage: a multiple of KC that divides every K streamed through this ring
};

// Running position in the ring; every thread keeps its own copy, all in lock-step.
struct RingPos {
  uint32_t stage;
  uint32_t phase;
  uint32_t ready;        // consumer hint: the full barrier at (stage, phase) was already seen complete
  __device__ __forceinline__ void advance(uint32_t nst) {
    if (++stage == nst) { stage = 0; phase ^= 1u; }
  }
};

// Identity of a thread inside the CTA (consumers first, producer warp last).
struct TileThread {
  int ctid;          // consumer thread id, [0, ncons)
  int ncons;         // consumer threads (multiple of 32)
  int lane;
  bool producer;     // member of the producer warp
};

// Epilogue description: v = act(acc + bias[n]) stored to up to two T-layout destinations, or the
// GRU new-gate combination.
//   EPI_STORE     v = act(acc + bias[n])
//   EPI_GRU_NEW   n = tanh(acc + bias + hn * r);  h' = (hprev - n) * z + n   (ATen gru_cell order)
//   EPI_MUL_DACT  v = acc * act'(hs[n])      backward through a hidden activation; act' is computed
//                                            from the saved activation OUTPUT hs (T-layout)
//   EPI_ADD       v = acc + out0[n]          accumulate into an existing T-layout gradient
// Optionally the value is also appended to a record stream:
//   rec_lo == nullptr: row-major, one row per tile row,
//       rec[(rec_row0 + (rb * RT + r) * rec_rstride) * rec_ld + n]   for r < rec_valid;
//   rec_lo != nullptr: tcgen05 operand blocks (wgrad_tc.cu) -- block rec_row0 holds the tile's R rows of
//       all rec_ld features as K-major core matrices [feature/8][R/4][8][4], split into a
//       TF32-exact high part (rec) and the residual (rec_lo).
//   EPI_GRU_N     v = tanh(acc + bias + hn * r)                 the GRU new gate alone (backward recompute)
//   EPI_MUL_IN    v = (acc + bias) * hs[n]                      FusionModule 'soft': cat * Linear(cat)
enum { EPI_STORE = 0, EPI_GRU_NEW = 1, EPI_MUL_DACT = 2, EPI_ADD = 3, EPI_GRU_N = 4, EPI_MUL_IN = 5 };
struct Epilogue {
  int mode;
  const float* bias;   // [N] or nullptr
  int act;
  float* out0; int ld0; int off0;   // out0[n * ld0 + off0 + rb * RT + r]   (shared or global); may be null
  float* out1; int ld1; int off1;   // optional second destination
  const float* rg; const float* zg; const float* hn; const float* hprev;   // EPI_GRU_NEW: [N][RT]
  const float* hs; int ldh; int offh;                                       // EPI_MUL_DACT
  float* rec; float* rec_lo; long long rec_row0; int rec_ld; int rec_rstride; int rec_valid;
};

// TF32-exact high part of x (low 13 mantissa bits cleared); x - hi is exact in fp32.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// float offset of (feature f, tile row r) inside a record block of R rows (see Epilogue)
__device__ __forceinline__ size_t rec_block_offset(int f, int r, int R) {
  return ((static_cast<size_t>(f >> 3) * (R >> 2) + (r >> 2)) * 8 + (f & 7)) * 4 + (r & 3);
}

// act'(x) expressed through the activation output h = act(x)
__device__ __forceinline__ float dact_from_output(float h, int act) {
  switch (act) {
    case ACT_TANH: return 1.f - h * h;
    case ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case ACT_LEAKY: return h > 0.f ? 1.f : 0.01f;
    case ACT_SOFTPLUS: return 1.f - expf(-h);     // sigmoid(x) with h = log(1 + e^x)
    case ACT_LEAKY01: return h > 0.f ? 1.f : 0.1f;
    default: return 1.f;
  }
}

// One lane of the producer warp: stream Wt[K][N] in KC-row chunks.
__device__ __forceinline__ void pipe_produce(const WeightRing& ring, RingPos& pos,
                                             const float* __restrict__ Wt, int K, int N) {
  const uint32_t bytes = ring.kc * static_cast<uint32_t>(N) * sizeof(float);
  const int nch = K / static_cast<int>(ring.kc);
  for (int ch = 0; ch < nch; ++ch) {
#if ODEVIO_PRODUCER_WAIT == 2
    mbar_wait_parked(&ring.empty[pos.stage], pos.phase ^ 1u);
#elif ODEVIO_PRODUCER_WAIT == 1
    mbar_wait_backoff(&ring.empty[pos.stage], pos.phase ^ 1u);
#else
    mbar_wait(&ring.empty[pos.stage], pos.phase ^ 1u);
#endif
    mbar_arrive_expect_tx(&ring.full[pos.stage], bytes);
    tma_load_1d(ring.buf + static_cast<size_t>(pos.stage) * ring.stage_floats,
                Wt + static_cast<size_t>(ch) * ring.kc * N, bytes, &ring.full[pos.stage]);
    pos.advance(ring.nst);
  }
}

template <int RT>
__device__ __forceinline__ void run_epilogue(const Epilogue& e, int n, int rb, const float (&acc)[RT]) {
  const float b = e.bias ? e.bias[n] : 0.f;
  float v[RT];
  if (e.mode == EPI_STORE) {
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 a4 = apply_act4(make_float4(acc[4 * q] + b, acc[4 * q + 1] + b, acc[4 * q + 2] + b,
                                               acc[4 * q + 3] + b), e.act);
      v[4 * q] = a4.x; v[4 * q + 1] = a4.y; v[4 * q + 2] = a4.z; v[4 * q + 3] = a4.w;
    }
  } else if (e.mode == EPI_GRU_NEW) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const size_t o = static_cast<size_t>(n) * RT + r;
      const float ng = apply_act((acc[r] + b) + e.hn[o] * e.rg[o], ACT_TANH);
      v[r] = (e.hprev[o] - ng) * e.zg[o] + ng;
    }
  } else if (e.mode == EPI_GRU_N) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const size_t o = static_cast<size_t>(n) * RT + r;
      v[r] = apply_act((acc[r] + b) + e.hn[o] * e.rg[o], ACT_TANH);
    }
  } else if (e.mode == EPI_MUL_IN) {
    const float* hp = e.hs + static_cast<size_t>(n) * e.ldh + e.offh + rb * RT;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 h4 = ld4(hp + 4 * q);
      v[4 * q] = (acc[4 * q] + b) * h4.x; v[4 * q + 1] = (acc[4 * q + 1] + b) * h4.y;
      v[4 * q + 2] = (acc[4 * q + 2] + b) * h4.z; v[4 * q + 3] = (acc[4 * q + 3] + b) * h4.w;
    }
  } else if (e.mode == EPI_MUL_DACT) {
    const float* hp = e.hs + static_cast<size_t>(n) * e.ldh + e.offh + rb * RT;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 h4 = ld4(hp + 4 * q);
      v[4 * q] = acc[4 * q] * dact_from_output(h4.x, e.act);
      v[4 * q + 1] = acc[4 * q + 1] * dact_from_output(h4.y, e.act);
      v[4 * q + 2] = acc[4 * q + 2] * dact_from_output(h4.z, e.act);
      v[4 * q + 3] = acc[4 * q + 3] * dact_from_output(h4.w, e.act);
    }
  } else {   // EPI_ADD
    const float* op = e.out0 + static_cast<size_t>(n) * e.ld0 + e.off0 + rb * RT;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 o4 = ld4(op + 4 * q);
      v[4 * q] = acc[4 * q] + o4.x; v[4 * q + 1] = acc[4 * q + 1] + o4.y;
      v[4 * q + 2] = acc[4 * q + 2] + o4.z; v[4 * q + 3] = acc[4 * q + 3] + o4.w;
    }
  }
  if (e.rec && e.rec_lo) {
    // block format: R = rec_rstride rows per block; this thread holds rows rb*RT .. rb*RT + RT - 1 of feature n
    const int R = e.rec_rstride;
    const size_t blk = static_cast<size_t>(e.rec_row0) * e.rec_ld * R;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const size_t o = blk + rec_block_offset(n, rb * RT + 4 * q, R);
      const float4 h4 = make_float4(tf32_hi(v[4 * q]), tf32_hi(v[4 * q + 1]), tf32_hi(v[4 * q + 2]), tf32_hi(v[4 * q + 3]));
      st4(e.rec + o, h4);
      st4(e.rec_lo + o, make_float4(v[4 * q] - h4.x, v[4 * q + 1] - h4.y, v[4 * q + 2] - h4.z, v[4 * q + 3] - h4.w));
    }
  } else if (e.rec) {
    float* rp = e.rec + (e.rec_row0 + static_cast<long long>(rb * RT) * e.rec_rstride) * e.rec_ld + n;
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (r < e.rec_valid) rp[static_cast<long long>(r) * e.rec_rstride * e.rec_ld] = v[r];
  }
  if (e.out0) {
    float* p0 = e.out0 + static_cast<size_t>(n) * e.ld0 + e.off0 + rb * RT;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) st4(p0 + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
  if (e.out1) {
    float* p1 = e.out1 + static_cast<size_t>(n) * e.ld1 + e.off1 + rb * RT;
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) st4(p1 + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
}

// Consumer body.  Inlined into its (single) call site: as an ABI (__noinline__) function ptxas
// serialised every LDS behind the previous step's FFMAs; inlined it double-buffers the operand
// registers.  Returns the advanced ring position packed as stage | phase << 8.
// Operands are addressed as offsets from the dynamic shared-memory base so the compiler emits
// LDS (shared-space) loads in the hot loop instead of generic LD.
template <int RT, int P, int KCT>
__device__ __forceinline__ uint32_t gemm_body(WeightRing ring, uint32_t pos_packed, uint32_t in_off,
                                           int ld, int K, int N, int rb, int cg, int tpb, int lane,
                                           const Epilogue* __restrict__ epi) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const float* __restrict__ inT = reinterpret_cast<const float*>(smem_raw + in_off);
  const float* __restrict__ ring_buf = reinterpret_cast<const float*>(smem_raw + ring.buf_off);
  RingPos pos{pos_packed & 0xffu, (pos_packed >> 8) & 1u, (pos_packed >> 16) & 1u};
  float acc[P][2][RT];
  int col[P];
  const int npairs = N >> 1;
#pragma unroll
  for (int pp = 0; pp < P; ++pp) {
    const int q = cg + pp * tpb;
    col[pp] = 2 * (q < npairs ? q : npairs - 1);   // clamp: surplus slots compute a discarded duplicate
#pragma unroll
    for (int r = 0; r < RT; ++r) { acc[pp][0][r] = 0.f; acc[pp][1][r] = 0.f; }
  }
  const int nch = K / KCT;
  // Software-pipelined over k with two operand register sets: the loads of step kk+1 are issued
  // before the FFMAs of step kk, so the ~30-cycle LDS latency hides behind 16*P FFMAs instead of
  // being exposed twice per step (ncu: short_scoreboard was the top stall without this).
  float xa[RT], xb[RT];
  float2 wa[P], wb[P];
  auto load_operands = [&](const float* __restrict__ xp, const float* __restrict__ ws, int kk,
                           float (&x)[RT], float2 (&w)[P]) {
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 v = ld4(xp + kk * ld + 4 * q);
      x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int pp = 0; pp < P; ++pp) w[pp] = *reinterpret_cast<const float2*>(ws + kk * N + col[pp]);
  };
  auto fma_step = [&](const float (&x)[RT], const float2 (&w)[P]) {
#pragma unroll
    for (int pp = 0; pp < P; ++pp) {
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        acc[pp][0][r] = fmaf(x[r], w[pp].x, acc[pp][0][r]);
        acc[pp][1][r] = fmaf(x[r], w[pp].y, acc[pp][1][r]);
      }
    }
  };
  bool ready = pos.ready != 0;
  for (int ch = 0; ch < nch; ++ch) {
    if (!ready) mbar_wait(&ring.full[pos.stage], pos.phase);
    const float* __restrict__ ws = ring_buf + pos.stage * ring.stage_floats;
    const float* __restrict__ xp = inT + ch * KCT * ld;
    const uint32_t cur = pos.stage;
    pos.advance(ring.nst);
#if ODEVIO_EARLY_PROBE
    // probe the NEXT stage's barrier now; its ~90-cycle predicate latency hides behind this chunk
    ready = mbar_try_wait(&ring.full[pos.stage], pos.phase);
#else
    ready = false;
#endif
    load_operands(xp, ws, 0, xa, wa);
#pragma unroll
    for (int kk = 0; kk < KCT; kk += 2) {
      load_operands(xp, ws, kk + 1, xb, wb);
      fma_step(xa, wa);
      if (kk + 2 < KCT) load_operands(xp, ws, kk + 2, xa, wa);
      fma_step(xb, wb);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[cur]);
  }
  pos.ready = ready ? 1u : 0u;
  const Epilogue e = *epi;
#pragma unroll
  for (int pp = 0; pp < P; ++pp) {
    const int q = cg + pp * tpb;
    if (q < npairs) {
      run_epilogue<RT>(e, 2 * q, rb, acc[pp][0]);
      run_epilogue<RT>(e, 2 * q + 1, rb, acc[pp][1]);
    }
  }
  return pos.stage | (pos.phase << 8) | (pos.ready << 16);
}

// All threads of the CTA call this with identical arguments.
//   Wt: packed [K][N] (K % KC == 0, N even, N <= 2*MAX_P*tpb), inT: shared T-layout.
//   ode_layout: vector-field geometry (LL row blocks of RT rows, row block rb starts at inT + rb*RT,
//   row stride RT*LL, 128 threads per row block) vs jump geometry (one row block, stride RT).
// Ends with a consumer-wide named barrier so the epilogue's stores are visible to the next phase.
template <int RT, int LL, bool ALLOW16 = false>
__device__ __forceinline__ void tile_gemm(const WeightRing& ring, RingPos& pos, const TileThread& th,
                                          const float* __restrict__ Wt, int K, int N,
                                          const float* inT, bool ode_layout, const Epilogue& epi) {
  if (th.producer) {
    if (th.lane == 0) pipe_produce(ring, pos, Wt, K, N);
    __syncwarp();
    return;
  }
  constexpr int ncons = 128 * LL;
  const int nrb = ode_layout ? LL : 1;
  const int ld = ode_layout ? RT * LL : RT;
  const int tpb = ncons / nrb;
  const int rb = th.ctid / tpb;
  const int cg = th.ctid - rb * tpb;
  const int P = ((N >> 1) + tpb - 1) / tpb;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t in_rb = static_cast<uint32_t>(reinterpret_cast<const unsigned char*>(inT + rb * RT) - smem_raw);
  const uint32_t pk = pos.stage | (pos.phase << 8) | (pos.ready << 16);
  uint32_t nk;
  if (ALLOW16 && ring.kc == 16) {      // 16-row stages (K % 16 == 0 guaranteed by the host plan)
    switch (P) {
      case 1: nk = gemm_body<RT, 1, 16>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      case 2: nk = gemm_body<RT, 2, 16>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      case 3: nk = gemm_body<RT, 3, 16>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      default: nk = gemm_body<RT, 4, 16>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
    }
  } else {
    switch (P) {
      case 1: nk = gemm_body<RT, 1, KC>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      case 2: nk = gemm_body<RT, 2, KC>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      case 3: nk = gemm_body<RT, 3, KC>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
      default: nk = gemm_body<RT, 4, KC>(ring, pk, in_rb, ld, K, N, rb, cg, tpb, th.lane, &epi); break;
    }
  }
  pos.stage = nk & 0xffu;
  pos.phase = (nk >> 8) & 1u;
  pos.ready = (nk >> 16) & 1u;
  named_bar_sync(1, ncons);
}

}  // namespace odevio
