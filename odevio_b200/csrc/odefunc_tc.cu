// ODEFunc.forward on the 5th-generation tensor cores: batched evaluation of the vector field
//   f(x) = tanh(W_n a(... a(W_0 x + b_0) ...) + b_n)            reference src/models/ODEFunc.py:38-39
// for M rows, fp32-accurate through the 3xTF32 split (tcgen05.mma.kind::tf32, fp32 accumulators
// in TMEM).  This is the tensor-core building block of the solver kernels: a 128-row tile is
// owned by a CLUSTER of 8 CTAs, CTA c computing output columns [c N/8, (c+1) N/8) of every Linear,
// so that 16 tiles (B = 1024 sequences x 2 rnn layers) already occupy 128 SMs.
//
// Per layer and CTA:
//   * the layer input (128 x K, written by the previous layer's epilogues of all 8 CTAs) and the
//     CTA's weight slice stream L2 -> shared memory in K-chunks of 32 through a bulk-TMA / mbarrier
//     ring; both are stored in global memory directly in the tensor core's canonical K-major
//     no-swizzle image (8 x 16 B core matrices) as a TF32-exact high part plus the exact residual;
//   * one elected thread issues, per 8-wide k-step, D += A_hi W_hi + A_lo W_hi + A_hi W_lo;
//   * four epilogue warps (thread = row) read the accumulators with tcgen05.ld, add the bias, apply
//     the activation and write their 128 x N/8 slice of the next layer's operand (hi / lo) -- or
//     the fp32 result of the last layer -- followed by a cluster barrier.
// Activations therefore cross CTAs through L2 (8 KB .. 48 KB per CTA and layer), not DSMEM, whose
// ~20 B/clk/SM would cost more than the MMAs (B300_MICROARCH.md, CGA/DSMEM table).  Multicasting the
// shared A chunks (cp.async.bulk ... .multicast::cluster with cluster-wide slot release) was measured
// and removed: it cut L2 reads 2.3x and changed nothing -- the bound is the per-MMA A-operand read
// from shared memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/odevio.h"
#include "../../include/odevio_debug.h"
#include "common.cuh"
#include "ft_layer.cuh"

namespace odevio {

namespace {

#undef FT_STAMP
#ifdef ODEVIO_FT_TIMELINE
#define FT_STAMP(idx) do { if (blockIdx.x == 0 && tile == 0) g_ft_dbg[idx] = clock64(); } while (0)
#else
#define FT_STAMP(idx) do { } while (0)
#endif

struct FtParams {
  int M, NL, act;
  int K[FT_MAX_LAYERS], N[FT_MAX_LAYERS];      // layer shapes
  int nN[FT_MAX_LAYERS], nK[FT_MAX_LAYERS];    // the cluster's CTAs form an nN x nK grid per layer: CTA (cn, ck) computes
                                               // output columns slice cn over the k range ck; nK > 1: partials reduced through L2
  float* part;                                 // per cluster: [nK*nN][Nsl][128] fp32 partial sums, column-major (nK > 1)
  size_t part_floats;                          // per cluster
  const float* Wp[FT_MAX_LAYERS];              // packed [c][K/KCH][Nc/8][KCH/4][8][4]: fp32 (FT_SPLIT) or TF32-exact high part
  const float* Wlo[FT_MAX_LAYERS];             // residual (only !FT_SPLIT)
  const float* bias[FT_MAX_LAYERS];
  const float* x;                              // [M][K[0]]
  float* out;                                  // [M][N[NL-1]]
  float* xa;                                   // per cluster: 2 buffers x 128 x Kmax floats (fp32 operand image)
  size_t xa_buf_floats;                        // 128 * Kmax
  int ntiles, nraw;
  uint32_t raw_stage_bytes;                    // (128 + Ncmax) * 16 * 4: one fp32 A chunk + one fp32 W chunk
  uint32_t op_stage_bytes;                     // 2 * raw: hi and lo images
};


template <int FT_NC, int FT_KCH, bool FT_SPLIT>
__global__ void __launch_bounds__(FT_THREADS, 1)
odefunc_tc_kernel(const __grid_constant__ FtParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t raw_full[FT_RAW_STAGES];      // bulk TMA -> splitter
  __shared__ __align__(8) uint64_t raw_empty[FT_RAW_STAGES];     // splitter -> producer
  __shared__ __align__(8) uint64_t op_ready[FT_OP_STAGES];       // splitter -> MMA issuer
  __shared__ __align__(8) uint64_t op_empty[FT_OP_STAGES];       // tcgen05.commit -> splitter
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x / FT_NC, nclusters = gridDim.x / FT_NC;

  if (tid == 0) {
    for (int i = 0; i < FT_RAW_STAGES; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 1); }
    for (int i = 0; i < FT_OP_STAGES; ++i) { mbar_init(&op_ready[i], ft_op_ready_arrivals(FT_SPLIT)); mbar_init(&op_empty[i], 1); }
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

  float* xa_cluster = p.xa + static_cast<size_t>(cluster_id) * 4 * p.xa_buf_floats;   // [ping-pong][hi|lo]
  FtCtx c;
  c.smem = smem; c.op_base = smem + static_cast<size_t>(p.nraw) * p.raw_stage_bytes;
  c.raw_full = raw_full; c.raw_empty = raw_empty; c.op_ready = op_ready; c.op_empty = op_empty; c.accum_bar = &accum_bar;
  c.tmem_d = tmem_d; c.crank = crank; c.nraw = static_cast<uint32_t>(p.nraw);
  c.raw_stage_bytes = p.raw_stage_bytes; c.op_stage_bytes = p.op_stage_bytes; c.xa_buf_floats = p.xa_buf_floats;
  c.part = p.part + static_cast<size_t>(cluster_id) * p.part_floats;
  c.g0 = 0;                       // chunks issued so far (all roles count identically): chunk g uses raw stage
                                  // g % RAW, operand stage g % OP, barrier parity (g / stages) & 1
  c.accum_phase = 0; c.tile = 0;

  for (int tile = cluster_id; tile < p.ntiles; tile += nclusters) {
    const int row0 = tile * FT_ROWS;
    // ---- layer-0 operand: this CTA converts its K/8 feature slice of the tile's rows
    if (warp < 4) {
      const int r = tid, K0 = p.K[0], ks = K0 / FT_NC;
      float* dst0 = xa_cluster;
      for (int kc = 0; kc < ks; kc += 32) {
        const int k0 = static_cast<int>(crank) * ks + kc;
        float v[32];
        if (row0 + r < p.M) {
          const float4* src = reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row0 + r) * K0 + k0);
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float4 t = src[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0.f;
        }
        if (FT_SPLIT) store_chunk<FT_KCH>(dst0, r, k0, v);
        else store_chunk_hilo<FT_KCH>(dst0, dst0 + p.xa_buf_floats, r, k0, v);
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");      // generic-proxy stores -> visible to the bulk-copy (async) proxy
    }
    __syncwarp();
    if (tid == 0) FT_STAMP(0);
    cluster_sync_all();
    if (tid == 0) FT_STAMP(1);

    for (int l = 0; l < p.NL; ++l) {
      FtLayer L;
      L.K = p.K[l]; L.N = p.N[l]; L.nN = p.nN[l]; L.nK = p.nK[l]; L.stamp = l;
      const bool last = l == p.NL - 1;
      L.act = last ? ACT_TANH : p.act;
      L.out_mode = last ? FT_OUT_ROWS : FT_OUT_OPERAND;
      L.Wp = p.Wp[l]; L.Wlo = p.Wlo[l]; L.bias = p.bias[l];
      L.a_src = xa_cluster + static_cast<size_t>(l & 1) * 2 * p.xa_buf_floats;
      L.nx = xa_cluster + static_cast<size_t>((l + 1) & 1) * 2 * p.xa_buf_floats;
      L.out = p.out; L.M = p.M; L.row0 = row0;
      c.tile = tile;
      ft_layer<FT_NC, FT_KCH, FT_SPLIT>(c, L);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  cluster_sync_all();          // no CTA may leave while peers can still multicast into it / arrive on its barriers
  if (warp == FT_WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
  }
}

template <int NC, int KCH, bool SPLIT>
cudaError_t ft_launch(const FtParams& p, FtPlan& pl, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(odefunc_tc_kernel<NC, KCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(pl.smem_bytes));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.blockDim = dim3(FT_THREADS); lc.dynamicSmemBytes = pl.smem_bytes; lc.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = NC; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  lc.attrs = &at; lc.numAttrs = 1;
  // clusters must sit inside one GPC: launch no more clusters than can be co-resident
  lc.gridDim = dim3(pl.nclusters * NC);
  int maxc = 0;
  if (cudaOccupancyMaxActiveClusters(&maxc, odefunc_tc_kernel<NC, KCH, SPLIT>, &lc) == cudaSuccess && maxc > 0) {
    if (pl.nclusters > maxc) pl.nclusters = maxc;
  } else {
    cudaGetLastError();
  }
  lc.gridDim = dim3(pl.nclusters * NC);
  return cudaLaunchKernelEx(&lc, odefunc_tc_kernel<NC, KCH, SPLIT>, p);
}

}  // namespace
}  // namespace odevio

using namespace odevio;

extern "C" {

// development only: copy the kernel's timeline stamps (64 x int64) to the host
ODEVIO_API int32_t odevio_debug_odefunc_timeline(long long* host_dst) {
  return static_cast<int32_t>(cudaMemcpyFromSymbol(host_dst, g_ft_dbg, sizeof(long long) * 64));
}

size_t odevio_odefunc_workspace_bytes(int32_t M, int32_t D, int32_t H, int32_t n_hidden) {
  FtPlan pl;
  if (ft_plan(M, D, H, n_hidden, pl) != 0) return 0;
  return pl.total_bytes;
}

int32_t odevio_odefunc_forward(int32_t M, int32_t D, int32_t H, int32_t n_hidden, int32_t activation,
                               const float* const* weights, const float* const* biases,
                               const float* x, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!weights || !biases || !x || !out || !workspace) return ODEVIO_E_NULL;
  if (activation < 0 || activation > ODEVIO_ACT_SOFTPLUS) return ODEVIO_E_ENUM;
  FtPlan pl;
  const int rc = ft_plan(M, D, H, n_hidden, pl);
  if (rc != 0) return rc;
  if (workspace_bytes < pl.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  FtParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.NL = pl.NL; p.act = activation;
  for (int l = 0; l < pl.NL; ++l) {
    if (!weights[l] || !biases[l]) return ODEVIO_E_NULL;
    p.K[l] = pl.K[l]; p.N[l] = pl.N[l]; p.nN[l] = pl.nN[l]; p.nK[l] = pl.nK[l];
    ft_pack_weight_kernel<<<296, 256, 0, stream>>>(weights[l], pl.N[l], pl.K[l], pl.nN[l], pl.KCH, ws + pl.off_w[l],
                                                   ws + pl.off_wlo[l]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int32_t>(e);
    p.Wp[l] = ws + pl.off_w[l]; p.Wlo[l] = ws + pl.off_wlo[l]; p.bias[l] = biases[l];
  }
  p.x = x; p.out = out; p.xa = ws + pl.off_xa; p.xa_buf_floats = pl.xa_buf_floats;
  p.part = ws + pl.off_part; p.part_floats = pl.part_floats;
  p.ntiles = pl.ntiles; p.nraw = pl.nraw; p.raw_stage_bytes = pl.raw_stage_bytes; p.op_stage_bytes = pl.op_stage_bytes;
  cudaError_t e = pl.NC == 8 ? ft_launch<8, 8, true>(p, pl, stream) : ft_launch<4, 16, false>(p, pl, stream);
  if (e != cudaSuccess) return static_cast<int32_t>(e);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}

}  // extern "C"
