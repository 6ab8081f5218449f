// extern "C" surface of libodevio_b200.so (declared in include/odevio.h).
// Host side only: argument validation, workspace planning, weight pre-pack launches and the
// fused-kernel launch, all asynchronous on the caller's stream.
#include <cuda_runtime.h>
#include <string.h>

#include <mutex>

#include "../../include/odevio.h"
#include "../../include/odevio_debug.h"
#include "odernn_params.h"
#include "odernn_tc.h"
#include "odernn_h3.h"
#include <stdlib.h>

#include "cde_params.h"

namespace odevio {

cudaError_t transpose_pack(const float* src, int N, int K, float* dst, int ldN, int k_off, int n_off,
                           cudaStream_t stream);
cudaError_t bias_sum(const float* a, const float* b, float* dst, int n, cudaStream_t stream);
cudaError_t launch_odernn_fwd(const FwdParams& prm, int rows_per_tile, int grid, size_t smem_bytes,
                              cudaStream_t stream);
cudaError_t launch_odernn_bwd(const BwdParams& prm, int rows_per_tile, int grid, size_t smem_bytes,
                              cudaStream_t stream);
cudaError_t launch_cde_fwd(const CdeParams& prm, const DevTableau& tab, int RT, int LL, int grid,
                           size_t smem_bytes, cudaStream_t stream);
cudaError_t cde_pack_final(const float* W, const float* b, int Hc, int C, int Gc, int ngroups, float* Wp,
                           float* bp, cudaStream_t stream);
cudaError_t cde_tc_debug_timeline(long long* host_dst);
cudaError_t launch_cde_tc(const CdeParams& prm, const DevTableau& tab, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t cde_tc_pack(const float* W, const float* b, int Hc, int C, unsigned char* Wimg, float* bval, float* W0t, float* b0,
                        cudaStream_t stream);
cudaError_t launch_cde_bwd(const CdeBwdParams& prm, const DevTableau& tab, int RT, int LL, int grid,
                           size_t smem_bytes, cudaStream_t stream);
cudaError_t cde_pack_final_t(const float* W, int Hc, int C, int Gc, int ngroups, float* WT, cudaStream_t stream);
cudaError_t cde_unpack_final_grad(const float* dWp, const float* dbp, int Hc, int C, int Gc, float* dW, float* db,
                                  cudaStream_t stream);
cudaError_t wgrad_linear_ex(const float* G, int ldg, const float* A, int lda, long long M, int N, int K,
                            float* dW0, float* dW1, int k_split, float* db0, float* db1, float* part, int nsm,
                            int accumulate, cudaStream_t stream);
int wgrad_splits(long long M, int N, int K, int nsm);
int wgrad_tc_bn(int K);
int wgrad_tc_splits(long long nblocks, int N, int K, int nsm);
cudaError_t wgrad_linear_tc(const float* Ghi, const float* Glo, const float* Ahi, const float* Alo,
                            long long nblocks, int N, int K, int R, float* dW, float* db, float* part, int nsm,
                            cudaStream_t stream);
cudaError_t wgrad_linear_tc_ex(const float* Ghi, const float* Glo, const float* Ahi, const float* Alo,
                               long long nblocks, int N, int K, int R, float* dW, float* db, float* part, int nsm,
                               int accumulate, cudaStream_t stream);
cudaError_t wgrad_linear(const float* G, int ldg, const float* A, int lda, long long M, int N, int K,
                         float* dW0, float* dW1, int k_split, float* db0, float* db1, float* part, int nsm,
                         cudaStream_t stream);

namespace {

constexpr size_t kSmemLimit = 232448;   // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int kStageK = 8;              // == KC in tile_gemm.cuh: rows per weight stage of the backward / CDE kernels;
                                        // the forward kernel takes 16-row stages when D, H allow it (OdePlan::kc)
constexpr int kMaxStagesRing = 4;       // == MAX_STAGES

// ------------------------------------------------------------------ tableaus (oracle/tableaus.py)
void zero_tab(DevTableau& t) { memset(&t, 0, sizeof(t)); }

void set_row(DevTableau& t, int i, const double* row, int n) {
  for (int j = 0; j < n; ++j) t.a[i][j] = static_cast<float>(row[j]);
}

bool make_tableau(int solver, DevTableau& t) {
  zero_tab(t);
  switch (solver) {
    case ODEVIO_SOLVER_DOPRI5: {
      t.n_stages = 7; t.fsal = 1; t.ssal = 1; t.has_err = 1; t.has_mid = 1; t.exponent = -1.0f / 5.0f;
      const double a1[] = {1.0 / 5};
      const double a2[] = {3.0 / 40, 9.0 / 40};
      const double a3[] = {44.0 / 45, -56.0 / 15, 32.0 / 9};
      const double a4[] = {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729};
      const double a5[] = {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656};
      const double a6[] = {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
      set_row(t, 1, a1, 1); set_row(t, 2, a2, 2); set_row(t, 3, a3, 3);
      set_row(t, 4, a4, 4); set_row(t, 5, a5, 5); set_row(t, 6, a6, 6);
      const double b[] = {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0.0};
      const double e[] = {0.0012326388888888873, 0.0, -0.004252770290506136, 0.036979166666666674,
                          -0.05086379716981132, 0.04190476190476192, -0.025};
      const double m[] = {0.10013431883002395, 0.0, 0.3918321794184259, -0.02982460176594817,
                          0.05893268337240795, -0.04497888809104361, 0.023904308236133973};
      for (int j = 0; j < 7; ++j) { t.b[j] = (float)b[j]; t.e[j] = (float)e[j]; t.bmid[j] = (float)m[j]; }
      return true;
    }
    case ODEVIO_SOLVER_TSIT5: {
      t.n_stages = 7; t.fsal = 1; t.ssal = 1; t.has_err = 1; t.has_mid = 1; t.exponent = -1.0f / 5.0f;
      const double a1[] = {0.161};
      const double a2[] = {-0.008480655492356989, 0.335480655492357};
      const double a3[] = {2.8971530571054935, -6.359448489975075, 4.3622954328695815};
      const double a4[] = {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525};
      const double a5[] = {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
                           -0.028269050394068383};
      const double a6[] = {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
                           -3.290069515436081, 2.324710524099774};
      set_row(t, 1, a1, 1); set_row(t, 2, a2, 2); set_row(t, 3, a3, 3);
      set_row(t, 4, a4, 4); set_row(t, 5, a5, 5); set_row(t, 6, a6, 6);
      const double e[] = {-0.001780011052225777, -0.0008164344596567469, 0.007880878010261995,
                          -0.1447110071732629, 0.5823571654525552, -0.45808210592918697,
                          0.015151515151515152};
      const double m[] = {0.10741235230096871, 0.01135625, 0.39560903056045305, -0.34475214352593553,
                          1.3161853649581645, -1.0170608542936508, 0.031249999999999993};
      for (int j = 0; j < 6; ++j) t.b[j] = (float)a6[j];
      t.b[6] = 0.f;
      for (int j = 0; j < 7; ++j) { t.e[j] = (float)e[j]; t.bmid[j] = (float)m[j]; }
      return true;
    }
    case ODEVIO_SOLVER_HEUN:
      t.n_stages = 2; t.has_err = 1; t.exponent = -1.0f / 2.0f;
      t.a[1][0] = 1.f; t.b[0] = 0.5f; t.b[1] = 0.5f; t.e[0] = -0.5f; t.e[1] = 0.5f;
      return true;
    case ODEVIO_SOLVER_EULER:
      t.n_stages = 1; t.exponent = -1.0f; t.b[0] = 1.f;
      return true;
    case ODEVIO_SOLVER_RK4:
      t.n_stages = 4; t.exponent = -1.0f / 4.0f;
      t.a[1][0] = 0.5f; t.a[2][1] = 0.5f; t.a[3][2] = 1.f;
      t.b[0] = (float)(1.0 / 6); t.b[1] = (float)(1.0 / 3); t.b[2] = (float)(1.0 / 3); t.b[3] = (float)(1.0 / 6);
      return true;
    case ODEVIO_SOLVER_RK4_38:
      t.n_stages = 4; t.exponent = -1.0f / 4.0f;
      t.a[1][0] = (float)(1.0 / 3); t.a[2][0] = (float)(-1.0 / 3); t.a[2][1] = 1.f;
      t.a[3][0] = 1.f; t.a[3][1] = -1.f; t.a[3][2] = 1.f;
      t.b[0] = (float)(1.0 / 8); t.b[1] = (float)(3.0 / 8); t.b[2] = (float)(3.0 / 8); t.b[3] = (float)(1.0 / 8);
      return true;
    default:
      return false;
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;
  }
  return n;
}

// Everything derived from cfg: launch geometry, shared-memory carve-up, workspace offsets.
struct OdePlan {
  int RT, R, ncons, threads, ntiles, grid, nst, G, nsm, CK, kc;
  size_t ckpt_head_bytes, ckpt_floats_per_tile;
  size_t bufA_floats, bufB_floats, stage_floats, smem_bytes;
  size_t off_Wode[kMaxLinears];                 // float offsets into the workspace
  size_t off_Wrnn[kMaxRnnLayers][4], off_brnn[kMaxRnnLayers][4];
  size_t off_Wreg0, off_Wfuse;
  size_t off_scratch, scratch_floats_per_cta;
  size_t total_bytes;
  int Kode[kMaxLinears], Node[kMaxLinears];
};

int plan_odernn(const odevio_odernn_cfg& c, OdePlan& pl) {
  if (c.B <= 0 || c.S <= 0 || c.D <= 0 || c.H <= 0) return ODEVIO_E_SHAPE;
  if (c.n_hidden < 1 || c.n_hidden + 1 > ODEVIO_MAX_ODE_LINEARS) return ODEVIO_E_SHAPE;
  if (c.L < 1 || c.L > ODEVIO_MAX_RNN_LAYERS) return ODEVIO_E_SHAPE;
  if (c.D % 8 || c.H % 8) return ODEVIO_E_SHAPE;                 // K chunks of 8, float4 rows
  if (c.activation < 0 || c.activation > ODEVIO_ACT_SOFTPLUS) return ODEVIO_E_ENUM;
  if (c.rnn_type != ODEVIO_RNN_TANH && c.rnn_type != ODEVIO_RNN_GRU) return ODEVIO_E_ENUM;
  if (c.solver < 0 || c.solver > ODEVIO_SOLVER_RK4_38) return ODEVIO_E_ENUM;
  if (c.precision != ODEVIO_PRECISION_FP32 && c.precision != ODEVIO_PRECISION_TF32X3 && c.precision != ODEVIO_PRECISION_FP16X3)
    return ODEVIO_E_ENUM;
  // tensor-core solvers: no step trace; the literal dense end point (endpoint_dense) is in odernn_h3.cu (FP16X3), not in
  // the round-1 3xTF32 kernel.  Training (save_checkpoints): the one-launch FP16X3 forward writes the checkpoints itself;
  // every other combination runs the FMA forward.
  if (c.precision != ODEVIO_PRECISION_FP32 && !c.save_checkpoints &&
      (c.trace_steps || (c.endpoint_dense && c.precision != ODEVIO_PRECISION_FP16X3))) return ODEVIO_E_ENUM;
  if (c.rows_per_tile != 0 && c.rows_per_tile != 4 && c.rows_per_tile != 8 && c.rows_per_tile != 16) return ODEVIO_E_SHAPE;
  const bool fixed = c.solver == ODEVIO_SOLVER_RK4 || c.solver == ODEVIO_SOLVER_RK4_38;
  if (fixed && c.substeps < 1) return ODEVIO_E_SHAPE;
  if (!fixed && c.max_steps < 1) return ODEVIO_E_SHAPE;
  if (c.trace_steps < 0 || c.trace_steps > 64) return ODEVIO_E_SHAPE;

  const int nsm = sm_count();
  const int NL = c.n_hidden + 1;
  int nmax = kRegHidden;
  for (int j = 0; j < NL; ++j) {
    pl.Kode[j] = j == 0 ? c.D : c.H;
    pl.Node[j] = j == NL - 1 ? c.D : c.H;
    if (pl.Node[j] > nmax) nmax = pl.Node[j];
  }
  if (c.D > nmax) nmax = c.D;
  // column pairs per thread <= 4: ODE GEMMs use 128 threads per row block, the jump all consumers
  if (nmax > 2 * 4 * 128) return ODEVIO_E_SHAPE;
  pl.ncons = 128 * c.L;
  pl.threads = pl.ncons + 32;
  pl.G = c.rnn_type == ODEVIO_RNN_GRU ? 3 : 1;
  // 16-row weight stages halve the per-chunk mbarrier overhead of the consumers (measured: forward 63.0 -> 58.6 ms);
  // they need every streamed K (D, H, 2D) to be a multiple of 16 and at least two stages in shared memory
  const bool kc16_ok = c.D % 16 == 0 && c.H % 16 == 0;

  // shared-memory carve-up for a tile of rt sequences (mirrors odernn_fwd_kernel); false if it
  // cannot hold at least two weight stages
  auto fit_kc = [&](int rt, int kc) -> bool {
    pl.kc = kc;
    pl.stage_floats = static_cast<size_t>(kc) * nmax;
    const int R = rt * c.L;
    const size_t maxdh = static_cast<size_t>(c.D > c.H ? c.D : c.H);
    size_t a = maxdh * R, a2 = static_cast<size_t>(2) * c.D * rt;
    pl.bufA_floats = a > a2 ? a : a2;
    size_t b = static_cast<size_t>(c.H) * R, b2 = static_cast<size_t>(c.D) * rt;
    pl.bufB_floats = b > b2 ? b : b2;
    const size_t fixed_bytes = (pl.bufA_floats + pl.bufB_floats + 4 * static_cast<size_t>(pl.ncons) +
                                14 * static_cast<size_t>(R)) * sizeof(float) + 8 + 2 * kMaxStagesRing * 8 + 128;
    if (fixed_bytes + 2 * pl.stage_floats * sizeof(float) > kSmemLimit) return false;
    size_t nst = (kSmemLimit - fixed_bytes) / (pl.stage_floats * sizeof(float));
    if (nst > kMaxStagesRing) nst = kMaxStagesRing;
    pl.nst = static_cast<int>(nst);
    pl.smem_bytes = fixed_bytes + nst * pl.stage_floats * sizeof(float);
    return true;
  };
  auto fit = [&](int rt) -> bool { return (kc16_ok && fit_kc(rt, 16)) || fit_kc(rt, kStageK); };
  int rt = c.rows_per_tile;
  if (c.save_checkpoints) {
    // training: the backward kernel exists for 4- and 8-row tiles and the y1 end-point rule
    if (c.endpoint_dense) return ODEVIO_E_ENUM;
    if (c.ckpt_loops < 0 || c.ckpt_loops > 4096) return ODEVIO_E_SHAPE;
    if (rt == 16) return ODEVIO_E_SHAPE;
    if (rt == 0) {
      // 4-sequence tiles when 8-sequence ones would leave half of the SMs idle (a rank of the 8-GPU training step holds
      // 512 sequences = 64 tiles of 8): the step time of a single wave is the per-tile latency, which grows with the rows
      rt = fit(8) ? 8 : 4;
      if (rt == 8 && 2 * ((c.B + 7) / 8) <= nsm && (4 * c.L) % 8 == 0 && fit(4)) rt = 4;
    }
  }
  if (rt == 0) {
    // Tile height: 16-row tiles amortise the weight stream better (~1.6x the time of an 8-row
    // tile for 2x the rows) but halve the CTA count and need <= 3 column pairs per thread to
    // stay in registers; fall back to smaller tiles when shared memory does not fit.
    const long w8 = ((c.B + 7) / 8 + nsm - 1) / nsm, w16 = ((c.B + 15) / 16 + nsm - 1) / nsm;
    if (c.L <= 2 && nmax <= 768 && w16 * 16 < w8 * 10 && fit(16)) rt = 16;
    else if (fit(8)) rt = 8;
    else rt = 4;
  }
  if (rt == 16 && c.L > 2) return ODEVIO_E_SHAPE;
  if (!fit(rt)) return ODEVIO_E_SHAPE;
  pl.RT = rt;
  pl.R = rt * c.L;
  pl.ntiles = (c.B + rt - 1) / rt;
  pl.grid = pl.ntiles < nsm ? pl.ntiles : nsm;

  // ---- workspace layout (floats), every block 64-float (256 B) aligned
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
  for (int j = 0; j < NL; ++j) pl.off_Wode[j] = take(static_cast<size_t>(pl.Kode[j]) * pl.Node[j]);
  const size_t DD = static_cast<size_t>(c.D) * c.D;
  for (int l = 0; l < c.L; ++l) {
    if (pl.G == 1) {
      pl.off_Wrnn[l][0] = take(2 * DD); pl.off_brnn[l][0] = take(c.D);
    } else {
      pl.off_Wrnn[l][0] = take(2 * DD); pl.off_Wrnn[l][1] = take(2 * DD);
      pl.off_Wrnn[l][2] = take(DD); pl.off_Wrnn[l][3] = take(DD);
      for (int g = 0; g < 4; ++g) pl.off_brnn[l][g] = take(c.D);
    }
  }
  pl.off_Wreg0 = take(static_cast<size_t>(c.D) * kRegHidden);
  pl.off_Wfuse = take(DD);
  pl.scratch_floats_per_cta = align_up(static_cast<size_t>(kMaxStages + 2) * c.D * pl.R, 64);
  pl.off_scratch = take(pl.scratch_floats_per_cta * pl.grid);
  pl.total_bytes = off * sizeof(float);
  pl.nsm = nsm;
  pl.CK = c.ckpt_loops > 0 ? c.ckpt_loops : (fixed ? c.substeps : 16);
  pl.ckpt_head_bytes = align_up(static_cast<size_t>(pl.ntiles) * c.S * sizeof(int32_t), 256);
  pl.ckpt_floats_per_tile = align_up(static_cast<size_t>(c.S) * ckpt_interval_floats(c.D, pl.R, pl.CK), 64);
  return 0;
}

// Backward: stages entering y1, shared-memory carve-up, workspace layout for `ode_rows` record rows.
struct BwdPlan {
  int ns, nst, kc;
  size_t buf_floats, stage_floats, smem_bytes;
  size_t off_Wode[kMaxLinears], off_Wreg0;
  size_t off_scratch, scratch_floats_per_cta;
  size_t off_recA_ode[kMaxLinears], off_recG_ode[kMaxLinears], off_recA_ode_lo[kMaxLinears], off_recG_ode_lo[kMaxLinears];
  size_t off_recA_rnn[kMaxRnnLayers], off_recG_rnn[kMaxRnnLayers];
  size_t off_recA_reg0, off_recG_reg0, off_recA_reg1, off_recG_reg1;
  size_t off_Wrnn[kMaxRnnLayers][4], off_brnn[kMaxRnnLayers][4];      // GRU: gate re-evaluation weights
  int Rb;
  int GW;                                                              // G-record width per jump row: D (rnn) | 6D (gru)
  size_t off_part, part_floats, off_tile_gy, off_tile_order;
  long long jump_rows;
  size_t total_bytes;
};

int plan_odernn_bwd(const odevio_odernn_cfg& c, const OdePlan& pl, long long ode_rows, BwdPlan& bp) {
  if (ode_rows < 0 || ode_rows % pl.R) return ODEVIO_E_SHAPE;
  // the ODEFunc weight gradients run on tcgen05 (wgrad_tc.cu): 128-row output tiles, 8-row k-steps
  if (pl.R % 4 || pl.R > 32 || c.D % 128 || c.H % 128) return ODEVIO_E_SHAPE;
  bp.Rb = (pl.R + 7) / 8 * 8;                     // record block rows (8-row k-steps); padding rows are zero
  DevTableau tb;
  if (!make_tableau(c.solver, tb)) return ODEVIO_E_ENUM;
  bp.ns = tb.ssal ? tb.n_stages - 1 : tb.n_stages;
  const int NL = c.n_hidden + 1;
  const size_t maxdh = static_cast<size_t>(c.D > c.H ? c.D : c.H);
  bp.buf_floats = maxdh * pl.R;
  if (bp.buf_floats < static_cast<size_t>(2) * c.D * pl.RT) bp.buf_floats = static_cast<size_t>(2) * c.D * pl.RT;
  const size_t nmax_ = pl.stage_floats / pl.kc;               // widest streamed weight row
  const size_t fixed_bytes = (2 * bp.buf_floats + 2 * static_cast<size_t>(pl.R)) * sizeof(float) + 8 +
                             2 * kMaxStagesRing * 8 + 128;
  // 16-row weight stages when two of them fit (every K of the backward is D, 2D, H or 128: multiples of 16 here)
  bp.kc = (fixed_bytes + 2 * 16 * nmax_ * sizeof(float) <= kSmemLimit) ? 16 : kStageK;
  bp.stage_floats = static_cast<size_t>(bp.kc) * nmax_;
  if (fixed_bytes + 2 * bp.stage_floats * sizeof(float) > kSmemLimit) return ODEVIO_E_SHAPE;
  size_t nst = (kSmemLimit - fixed_bytes) / (bp.stage_floats * sizeof(float));
  if (nst > kMaxStagesRing) nst = kMaxStagesRing;
  bp.nst = static_cast<int>(nst);
  bp.smem_bytes = fixed_bytes + nst * bp.stage_floats * sizeof(float);

  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
  for (int j = 0; j < NL; ++j) bp.off_Wode[j] = take(static_cast<size_t>(pl.Kode[j]) * pl.Node[j]);
  bp.off_Wreg0 = take(static_cast<size_t>(c.D) * kRegHidden);
  const bool gru = c.rnn_type == ODEVIO_RNN_GRU;
  bp.GW = gru ? 6 * c.D : c.D;
  if (gru) {
    const size_t DDg = static_cast<size_t>(c.D) * c.D;
    for (int l = 0; l < c.L; ++l) {
      bp.off_Wrnn[l][0] = take(2 * DDg); bp.off_Wrnn[l][1] = take(2 * DDg);
      bp.off_Wrnn[l][2] = take(DDg); bp.off_Wrnn[l][3] = take(DDg);
      for (int g = 0; g < 4; ++g) bp.off_brnn[l][g] = take(c.D);
    }
  }
  const size_t arr = static_cast<size_t>(c.D) * pl.R, harr = static_cast<size_t>(c.H) * pl.R;
  bp.scratch_floats_per_cta = align_up((2 * kMaxStages + 1) * arr + static_cast<size_t>(kMaxStages) * (NL - 1) * harr, 64);
  bp.off_scratch = take(bp.scratch_floats_per_cta * pl.grid);
  const size_t M = static_cast<size_t>(ode_rows / pl.R) * bp.Rb;
  for (int j = 0; j < NL; ++j) {
    bp.off_recA_ode[j] = take(M * pl.Kode[j]);
    bp.off_recA_ode_lo[j] = take(M * pl.Kode[j]);
    bp.off_recG_ode[j] = take(M * pl.Node[j]);
    bp.off_recG_ode_lo[j] = take(M * pl.Node[j]);
  }
  bp.jump_rows = static_cast<long long>(pl.ntiles) * c.S * pl.RT;
  const size_t MJ = static_cast<size_t>(bp.jump_rows);
  for (int l = 0; l < c.L; ++l) {
    bp.off_recA_rnn[l] = take(MJ * 2 * c.D);
    bp.off_recG_rnn[l] = take(MJ * bp.GW);
  }
  bp.off_recA_reg0 = take(MJ * c.D);
  bp.off_recG_reg0 = take(MJ * kRegHidden);
  bp.off_recA_reg1 = take(MJ * kRegHidden);
  bp.off_recG_reg1 = take(MJ * 8);
  // split-M partial buffer: the largest Linear's partials
  size_t part = 0;
  auto need = [&](long long m, int n, int k) {
    size_t v = static_cast<size_t>(wgrad_splits(m, n, k, pl.nsm)) * n * k;
    if (v < static_cast<size_t>(256) * n) v = static_cast<size_t>(256) * n;
    if (v > part) part = v;
  };
  for (int j = 0; j < NL; ++j) {
    size_t v = static_cast<size_t>(wgrad_tc_splits(ode_rows / pl.R, pl.Node[j], pl.Kode[j], pl.nsm)) * pl.Node[j] * pl.Kode[j];
    if (v < static_cast<size_t>(256) * pl.Node[j]) v = static_cast<size_t>(256) * pl.Node[j];
    if (v > part) part = v;
  }
  if (gru) need(bp.jump_rows, 3 * c.D, c.D); else need(bp.jump_rows, c.D, 2 * c.D);
  need(bp.jump_rows, kRegHidden, c.D);
  need(bp.jump_rows, kPoseDim, kRegHidden);
  bp.part_floats = part;
  bp.off_part = take(part);
  bp.off_tile_gy = take(static_cast<size_t>(pl.ntiles) * c.D * pl.R);      // hidden-state gradient carried between interval ranges
  bp.off_tile_order = take(static_cast<size_t>(pl.ntiles) + 64);          // int [ntiles] work-queue order + the queue head
  bp.total_bytes = off * sizeof(float);
  return 0;
}

size_t ckpt_total_bytes(const odevio_odernn_cfg& c, const OdePlan& pl) {
  (void)c;
  return pl.ckpt_head_bytes + static_cast<size_t>(pl.ntiles) * pl.ckpt_floats_per_tile * sizeof(float);
}

// ------------------------------------------------------------------ CDE planning
struct CdePlan {
  int tc, Bpad, nrt, RP, dx_cache;          // tensor-core kernel (cde_tc.cu): rows padded to 128, row tiles, rows per CTA in the row phase
  size_t off_Wimg, off_bval, off_W0t, off_b0, off_state, off_Ximg, off_dXg;
  int RT, LL, R, ntiles, grid, nst, C, Cpad, Gc, ngroups, Ng, nsm;
  size_t buf_floats, staging_floats, stage_floats, smem_bytes;
  size_t off_Wmlp[kMaxLinears], off_Wfin, off_bfin, off_Winit, off_Wreg0;
  size_t off_scratch, scratch_floats_per_tile, off_red, off_bar;
  size_t total_bytes;
};

int plan_cde(const odevio_cde_cfg& c, CdePlan& pl) {
  memset(&pl, 0, sizeof(pl));
  if (c.B <= 0 || c.S < 1 || c.S > kCdeMaxOut || c.So < 2 || c.So < c.S) return ODEVIO_E_SHAPE;
  if (c.Hc < 8 || c.Hc % 8 || c.Hc > 1024) return ODEVIO_E_SHAPE;
  if (c.n_layers < 1 || c.n_layers + 1 > ODEVIO_MAX_ODE_LINEARS) return ODEVIO_E_SHAPE;
  if (c.activation < 0 || c.activation > ODEVIO_ACT_SOFTPLUS) return ODEVIO_E_ENUM;
  if (c.solver != ODEVIO_CDE_SOLVER_DOPRI5 && c.solver != ODEVIO_CDE_SOLVER_RK4) return ODEVIO_E_ENUM;
  if (c.interp != ODEVIO_CDE_INTERP_LINEAR && c.interp != ODEVIO_CDE_INTERP_CUBIC) return ODEVIO_E_ENUM;
  if (c.rows_per_tile != 0 && c.rows_per_tile != 8 && c.rows_per_tile != 16) return ODEVIO_E_SHAPE;
  if (c.solver == ODEVIO_CDE_SOLVER_DOPRI5 && c.max_steps < 1) return ODEVIO_E_SHAPE;
  if (c.step_size < 0.0) return ODEVIO_E_SHAPE;
  const int nsm = sm_count();
  pl.nsm = nsm;
  pl.C = c.Hc + 1;
  pl.Cpad = (pl.C + 7) / 8 * 8;
  pl.tc = 0;
  if (c.precision == ODEVIO_PRECISION_FP16X3) {
    // one CTA per hidden unit (cooperative grid), rows padded to the 128-row MMA tiles
    if (c.Hc != 32 && c.Hc != 64 && c.Hc != 128) return ODEVIO_E_SHAPE;
    if (c.Hc > nsm) return ODEVIO_E_SHAPE;
    pl.tc = 1;
    pl.Bpad = (c.B + 127) / 128 * 128;
    pl.nrt = pl.Bpad / 128;
    pl.RP = 4;
    while (pl.RP < 32 && pl.RP * c.Hc < pl.Bpad) pl.RP *= 2;
    if (pl.RP * c.Hc < pl.Bpad) return ODEVIO_E_SHAPE;
    pl.grid = c.Hc;
    size_t rows = static_cast<size_t>(c.Hc > pl.Cpad ? c.Hc : pl.Cpad);
    if (rows < static_cast<size_t>(kRegHidden)) rows = kRegHidden;
    pl.smem_bytes = 4u * c.Hc * c.Hc + 2u * 512u * c.Hc + (2 * rows * pl.RP + pl.RP + c.Hc + 512) * sizeof(float) + 64;
    if (pl.smem_bytes > kSmemLimit) return ODEVIO_E_SHAPE;
    // per-segment (m, d) cache of the control path's differences, when it fits next to the rest (1 KB static + alignment)
    const size_t dxc = static_cast<size_t>(2) * pl.RP * pl.C * sizeof(float);
    pl.dx_cache = pl.smem_bytes + dxc + 2048 <= kSmemLimit ? 1 : 0;
    if (pl.dx_cache) pl.smem_bytes += dxc;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
    for (int j = 0; j < c.n_layers; ++j) pl.off_Wmlp[j] = take(static_cast<size_t>(c.Hc) * c.Hc);
    pl.off_Winit = take(static_cast<size_t>(pl.Cpad) * c.Hc);
    pl.off_Wreg0 = take(static_cast<size_t>(c.Hc) * kRegHidden);
    pl.off_Wimg = take(static_cast<size_t>(c.Hc) * c.Hc * c.Hc);            // 4 Hc^3 bytes
    pl.off_bval = take(static_cast<size_t>(c.Hc) * c.Hc);
    pl.off_W0t = take(static_cast<size_t>(c.Hc) * c.Hc);
    pl.off_b0 = take(static_cast<size_t>(c.Hc));
    pl.off_state = take(static_cast<size_t>(2 + kMaxStages) * c.Hc * pl.Bpad);
    pl.off_Ximg = take(static_cast<size_t>(pl.nrt) * 128 * c.Hc);          // nrt * 512 Hc bytes
    pl.off_dXg = take(static_cast<size_t>(pl.Bpad) * c.Hc);
    pl.off_red = take(static_cast<size_t>(2) * pl.grid * 2 * 2);
    pl.off_bar = take(64);
    pl.total_bytes = off * sizeof(float);
    return 0;
  }
  if (c.precision != ODEVIO_PRECISION_FP32) return ODEVIO_E_ENUM;
  pl.Gc = 1024 / c.Hc; if (pl.Gc < 1) pl.Gc = 1; if (pl.Gc > pl.C) pl.Gc = pl.C;
  pl.ngroups = (pl.C + pl.Gc - 1) / pl.Gc;
  pl.Ng = pl.Gc * c.Hc;
  int nmax = pl.Ng > kRegHidden ? pl.Ng : kRegHidden;
  pl.stage_floats = static_cast<size_t>(kStageK) * nmax;
  auto fit = [&](int R) -> bool {
    size_t rows = static_cast<size_t>(c.Hc > pl.Cpad ? c.Hc : pl.Cpad);
    if (rows < static_cast<size_t>(kRegHidden)) rows = kRegHidden;
    pl.buf_floats = rows * R;
    pl.staging_floats = static_cast<size_t>(pl.Ng) * R;
    const size_t fixed_bytes = (2 * pl.buf_floats + pl.staging_floats + static_cast<size_t>(pl.Cpad) * R +
                                static_cast<size_t>(c.Hc) * R) * sizeof(float) + 16 + 72 * sizeof(double) +
                               2 * kMaxStagesRing * 8 + 128;
    if (fixed_bytes + 2 * pl.stage_floats * sizeof(float) > kSmemLimit) return false;
    size_t nst = (kSmemLimit - fixed_bytes) / (pl.stage_floats * sizeof(float));
    if (nst > kMaxStagesRing) nst = kMaxStagesRing;
    pl.nst = static_cast<int>(nst);
    pl.smem_bytes = fixed_bytes + nst * pl.stage_floats * sizeof(float);
    return true;
  };
  int R = c.rows_per_tile;
  if (R == 0) R = (c.B <= 8 * nsm) ? 8 : 16;         // more CTAs while the batch fits one wave
  if (!fit(R)) { if (R == 16 && fit(8)) R = 8; else return ODEVIO_E_SHAPE; }
  pl.R = R; pl.RT = 8; pl.LL = R / 8;
  pl.ntiles = (c.B + R - 1) / R;
  pl.grid = pl.ntiles < nsm ? pl.ntiles : nsm;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
  for (int j = 0; j < c.n_layers; ++j) pl.off_Wmlp[j] = take(static_cast<size_t>(c.Hc) * c.Hc);
  pl.off_Wfin = take(static_cast<size_t>(pl.ngroups) * c.Hc * pl.Ng);
  pl.off_bfin = take(static_cast<size_t>(pl.ngroups) * pl.Ng);
  pl.off_Winit = take(static_cast<size_t>(pl.Cpad) * c.Hc);
  pl.off_Wreg0 = take(static_cast<size_t>(c.Hc) * kRegHidden);
  pl.scratch_floats_per_tile = align_up(static_cast<size_t>(2 + kMaxStages) * c.Hc * R, 64);
  pl.off_scratch = take(pl.scratch_floats_per_tile * pl.ntiles);
  pl.off_red = take(static_cast<size_t>(2) * pl.grid * 2 * 2);      // doubles as float pairs
  pl.off_bar = take(64);
  pl.total_bytes = off * sizeof(float);
  return 0;
}

// checkpoint buffer of the CDE training forward: [log: 1 + steps entries][steps][ntiles][Z, Y1, K0..K6][Hc][R]
size_t cde_ckpt_log_bytes(int steps) { return align_up(sizeof(CdeStepRec) * static_cast<size_t>(1 + steps), 256); }
size_t cde_ckpt_step_floats(const odevio_cde_cfg& c, const CdePlan& pl) {
  if (pl.tc) return static_cast<size_t>(2 + kMaxStages) * c.Hc * pl.Bpad;          // cde_tc.cu: [9][Hc][Bpad]
  return static_cast<size_t>(pl.ntiles) * (2 + kMaxStages) * c.Hc * pl.R;           // cde_fwd.cu: [tile][9][Hc][R]
}

struct CdeBwdPlan {
  int nst, NgTot;
  long long Mc;            // record rows per chunk
  size_t buf_floats, staging_floats, stage_floats, smem_bytes;
  size_t off_Wmlp[kMaxLinears], off_Wfin, off_bfin, off_WfinT, off_WinitP, off_Wreg0;
  size_t off_recA[kMaxLinears + 1], off_recG[kMaxLinears], off_recGf;
  size_t off_recA_reg0, off_recG_reg0, off_recA_reg1, off_recG_reg1, off_recA_init, off_recG_init;
  size_t off_tile_state, off_scratch, scratch_floats_per_cta, off_dWp, off_dbp, off_dWinit, off_part;
  size_t total_bytes;
};

int plan_cde_bwd(const odevio_cde_cfg& c, const CdePlan& pl, int chunk_vjps, CdeBwdPlan& bp) {
  if (chunk_vjps < 8) return ODEVIO_E_SHAPE;
  if (pl.R != 8 && pl.R != 16) return ODEVIO_E_SHAPE;
  const int R = pl.R, Hc = c.Hc;
  bp.NgTot = pl.ngroups * pl.Ng;
  int nmax = pl.Ng > kRegHidden ? pl.Ng : kRegHidden;
  if (pl.Cpad > nmax) nmax = pl.Cpad;
  if (pl.Cpad > 1024) return ODEVIO_E_SHAPE;
  bp.stage_floats = static_cast<size_t>(kStageK) * nmax;
  size_t rows = static_cast<size_t>(Hc > pl.Cpad ? Hc : pl.Cpad);
  if (rows < static_cast<size_t>(kRegHidden)) rows = kRegHidden;
  bp.buf_floats = rows * R;
  bp.staging_floats = static_cast<size_t>(pl.Ng) * R;
  const size_t fixed_bytes = (2 * bp.buf_floats + bp.staging_floats + 2 * static_cast<size_t>(pl.Cpad) * R +
                              2 * static_cast<size_t>(Hc) * R) * sizeof(float) + 16 + 2 * kMaxStagesRing * 8 + 128;
  if (fixed_bytes + 2 * bp.stage_floats * sizeof(float) > kSmemLimit) return ODEVIO_E_SHAPE;
  size_t nst = (kSmemLimit - fixed_bytes) / (bp.stage_floats * sizeof(float));
  if (nst > kMaxStagesRing) nst = kMaxStagesRing;
  bp.nst = static_cast<int>(nst);
  bp.smem_bytes = fixed_bytes + nst * bp.stage_floats * sizeof(float);
  bp.Mc = static_cast<long long>(chunk_vjps) * pl.ntiles * R;
  const size_t Mc = static_cast<size_t>(bp.Mc);
  const size_t rowsBS = static_cast<size_t>(pl.ntiles) * R * c.S;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
  for (int j = 0; j < c.n_layers; ++j) bp.off_Wmlp[j] = take(static_cast<size_t>(Hc) * Hc);
  bp.off_Wfin = take(static_cast<size_t>(pl.ngroups) * Hc * pl.Ng);
  bp.off_bfin = take(static_cast<size_t>(bp.NgTot));
  bp.off_WfinT = take(static_cast<size_t>(pl.ngroups) * pl.Ng * Hc);
  bp.off_WinitP = take(static_cast<size_t>(Hc) * pl.Cpad);
  bp.off_Wreg0 = take(static_cast<size_t>(Hc) * kRegHidden);
  for (int j = 0; j <= c.n_layers; ++j) bp.off_recA[j] = take(Mc * Hc);
  for (int j = 0; j < c.n_layers; ++j) bp.off_recG[j] = take(Mc * Hc);
  bp.off_recGf = take(Mc * bp.NgTot);
  bp.off_recA_reg0 = take(rowsBS * Hc);
  bp.off_recG_reg0 = take(rowsBS * kRegHidden);
  bp.off_recA_reg1 = take(rowsBS * kRegHidden);
  bp.off_recG_reg1 = take(rowsBS * 8);
  bp.off_recA_init = take(static_cast<size_t>(pl.ntiles) * R * pl.Cpad);
  bp.off_recG_init = take(static_cast<size_t>(pl.ntiles) * R * Hc);
  bp.off_tile_state = take(static_cast<size_t>(pl.ntiles) * 2 * Hc * R);
  bp.scratch_floats_per_cta = align_up(static_cast<size_t>(kMaxStages + 2 + c.n_layers + 1) * Hc * R, 64);
  bp.off_scratch = take(bp.scratch_floats_per_cta * pl.grid);
  bp.off_dWp = take(static_cast<size_t>(bp.NgTot) * Hc);
  bp.off_dbp = take(static_cast<size_t>(bp.NgTot));
  bp.off_dWinit = take(static_cast<size_t>(Hc) * pl.Cpad);
  // partial sums of the weight-gradient GEMMs
  size_t part = 0;
  auto need = [&](long long M, int N, int K) {
    const size_t a = static_cast<size_t>(wgrad_splits(M, N, K, pl.nsm)) * N * K;
    const size_t b = static_cast<size_t>(256) * N;
    if (a > part) part = a;
    if (b > part) part = b;
  };
  need(bp.Mc, Hc, Hc);
  need(bp.Mc, bp.NgTot, Hc);
  need(static_cast<long long>(c.B) * c.S, kRegHidden, Hc);
  need(static_cast<long long>(c.B) * c.S, kPoseDim, kRegHidden);
  need(c.B, Hc, pl.Cpad);
  bp.off_part = take(part);
  bp.total_bytes = off * sizeof(float);
  return 0;
}

#define ODEVIO_CUDA_TRY(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return static_cast<int32_t>(_e);     \
  } while (0)

// ---- tensor-core mode: rows the cluster kernel cannot take in full rounds run concurrently in the FMA kernel
int g_last_precision = 0;       // development hooks: which solver the last forward used
constexpr int kSideRT = 8;      // sequences (= rows, L = 1) per CTA of the side launch
size_t tc_side_scratch_bytes(const odevio_odernn_cfg& c, int nsm) {      // <= 8 rows per CTA, <= nsm CTAs
  return align_up(align_up(static_cast<size_t>(kMaxStages + 2) * c.D * 8, 64) * sizeof(float) * static_cast<size_t>(nsm), 256);
}

size_t tc_seq_table_bytes(const odevio_odernn_cfg& c) {
  return align_up(static_cast<size_t>(c.S) * c.B * sizeof(int32_t), 256);
}

// Helper stream + fork/join events per device, created on first use (the only persistent objects of the library).
struct SideStream { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
int side_stream(SideStream** out) {
  static std::mutex mu;
  static SideStream table[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev < 0 || dev >= 64) return ODEVIO_E_DEVICE;
  std::lock_guard<std::mutex> lock(mu);
  SideStream& t = table[dev];
  if (!t.stream) {
    if ((e = cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming)) != cudaSuccess) return static_cast<int>(e);
  }
  *out = &t;
  return 0;
}

}  // namespace
}  // namespace odevio

using namespace odevio;

extern "C" {

int32_t odevio_version(void) { return ODEVIO_ABI_VERSION; }

const char* odevio_error_string(int32_t code) {
  switch (code) {
    case 0: return "ok";
    case ODEVIO_E_NULL: return "required pointer is NULL";
    case ODEVIO_E_SHAPE: return "unsupported or inconsistent dimension";
    case ODEVIO_E_ENUM: return "unknown activation / rnn / solver / precision id";
    case ODEVIO_E_WORKSPACE: return "workspace too small or misaligned";
    case ODEVIO_E_DEVICE: return "no usable sm_100 device";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown odevio error";
  }
}

void odevio_odernn_default_cfg(odevio_odernn_cfg* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->B = 1; cfg->S = 10; cfg->D = 768; cfg->H = 512; cfg->n_hidden = 3; cfg->L = 2;
  cfg->activation = ODEVIO_ACT_TANH; cfg->rnn_type = ODEVIO_RNN_TANH; cfg->solver = ODEVIO_SOLVER_DOPRI5;
  cfg->substeps = 1;
  cfg->atol = 1e-6f; cfg->rtol = 1e-2f; cfg->dt0 = 1e-4f;
  cfg->safety = 0.9f; cfg->factor_min = 0.2f; cfg->factor_max = 10.0f;
  cfg->accept_strict = 1; cfg->floor_factor = 0; cfg->endpoint_dense = 0; cfg->exact_landing = 1;
  cfg->max_steps = 100000;
  cfg->precision = ODEVIO_PRECISION_FP32;
  cfg->ckpt_loops = 0;
}

size_t odevio_odernn_workspace_bytes(const odevio_odernn_cfg* cfg) {
  if (!cfg) return 0;
  OdePlan pl;
  if (plan_odernn(*cfg, pl) != 0) return 0;
  if (cfg->precision == ODEVIO_PRECISION_TF32X3) {
    const size_t tcb = odernn_tc_workspace_bytes(*cfg);
    return tcb ? align_up(pl.total_bytes, 256) + align_up(tcb, 256) + tc_side_scratch_bytes(*cfg, pl.nsm) +
                     tc_seq_table_bytes(*cfg) : 0;
  }
  if (cfg->precision == ODEVIO_PRECISION_FP16X3) {
    const size_t hb = odernn_h3_workspace_bytes(*cfg);
    return hb ? align_up(pl.total_bytes, 1024) + align_up(hb, 1024) : 0;
  }
  return pl.total_bytes;
}

int32_t odevio_odernn_forward(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                              const float* fv, const float* fi, int32_t Dv,
                              const float* ts, const float* h0,
                              float* pose, float* hT, int32_t* stats, int32_t* status,
                              void* ckpt, size_t ckpt_bytes,
                              void* workspace, size_t workspace_bytes, void* stream_) {
  if (!cfg || !w || !ts || !hT || !workspace) return ODEVIO_E_NULL;
  if (!cfg->evolve_only && (!fv || !pose)) return ODEVIO_E_NULL;
  if (cfg->save_checkpoints && (!ckpt || cfg->evolve_only)) return ODEVIO_E_NULL;
  odevio_odernn_cfg c = *cfg;
  // training forward: on tcgen05 when the one-launch kernel covers the configuration (tanh rnn, L <= 2, fusion applied by
  // the host), else the FMA kernel -- the checkpoint layout is the same, the backward does not care who wrote it
  if (c.save_checkpoints && c.precision != ODEVIO_PRECISION_FP32 &&
      !(c.precision == ODEVIO_PRECISION_FP16X3 && odernn_h3_can_fuse_jump(c) && !w->fuse_w && !w->fuse_b && !c.trace_steps &&
        !c.endpoint_dense))
    c.precision = ODEVIO_PRECISION_FP32;
  OdePlan pl;
  const int rc = plan_odernn(c, pl);
  if (rc != 0) return rc;
  if (!c.evolve_only && (Dv <= 0 || Dv > c.D || (Dv < c.D && !fi) || (Dv == c.D && fi))) return ODEVIO_E_SHAPE;
  if (workspace_bytes < pl.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  if (c.save_checkpoints && (ckpt_bytes < ckpt_total_bytes(c, pl) || (reinterpret_cast<uintptr_t>(ckpt) & 255)))
    return ODEVIO_E_WORKSPACE;
  const int NL = c.n_hidden + 1;
  for (int j = 0; j < NL; ++j) if (!w->ode_w[j] || !w->ode_b[j]) return ODEVIO_E_NULL;
  if (!c.evolve_only) {
    for (int l = 0; l < c.L; ++l)
      if (!w->rnn_w_ih[l] || !w->rnn_w_hh[l] || !w->rnn_b_ih[l] || !w->rnn_b_hh[l]) return ODEVIO_E_NULL;
    if (!w->reg_w0 || !w->reg_b0 || !w->reg_w1 || !w->reg_b1) return ODEVIO_E_NULL;
  }

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  const int D = c.D;

  if (c.precision == ODEVIO_PRECISION_FP16X3 && (c.evolve_only || (odernn_h3_can_fuse_jump(c) && !w->fuse_w && !w->fuse_b))) {
    // ONE launch per forward (odernn_h3.cu): every cluster of 4 CTAs walks its 64-row tile through all S intervals --
    // solver loops, rnn jump and pose head on tcgen05 -- without returning to the host (PoseODERNN.py:108-122).  None of
    // the FMA kernel's packed weights is needed.
    DevTableau tab;
    if (!make_tableau(c.solver, tab)) return ODEVIO_E_ENUM;
    const bool adaptive = !(c.solver == ODEVIO_SOLVER_RK4 || c.solver == ODEVIO_SOLVER_RK4_38);
    const size_t h_off = align_up(pl.total_bytes, 1024);
    const size_t h_bytes = align_up(odernn_h3_workspace_bytes(c), 1024);
    if (h_bytes == 0) return ODEVIO_E_SHAPE;
    if (workspace_bytes < h_off + h_bytes) return ODEVIO_E_WORKSPACE;
    g_last_precision = c.precision;
    H3Evolve h3;
    const int prc = h3.prepare(c, tab, adaptive, w, !c.evolve_only, !c.weights_prepacked, static_cast<unsigned char*>(workspace) + h_off,
                               h_bytes, stream);
    if (prc != 0) return prc;
    if (status) ODEVIO_CUDA_TRY(cudaMemsetAsync(status, 0, static_cast<size_t>(c.B) * sizeof(int32_t), stream));
    if (c.save_checkpoints)
      h3.set_checkpoints(reinterpret_cast<float*>(static_cast<unsigned char*>(ckpt) + pl.ckpt_head_bytes), static_cast<int*>(ckpt),
                         pl.ckpt_floats_per_tile, pl.CK, pl.RT, pl.ntiles, c.S);
    return h3.run(h0, hT, ts, c.S + 1, 0, c.S, fv, fi, Dv, c.S, pose, stats, status, stream);
  }

  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.S = c.S; p.D = D; p.H = c.H; p.NL = NL; p.L = c.L;
  p.act = c.activation; p.rnn_type = c.rnn_type;
  p.adaptive = !(c.solver == ODEVIO_SOLVER_RK4 || c.solver == ODEVIO_SOLVER_RK4_38);
  p.substeps = c.substeps;
  p.atol = c.atol; p.rtol = c.rtol; p.dt0 = c.dt0; p.safety = c.safety; p.fmin = c.factor_min; p.fmax = c.factor_max;
  p.accept_strict = c.accept_strict; p.floor_factor = c.floor_factor; p.endpoint_dense = c.endpoint_dense;
  p.max_steps = c.max_steps; p.exact_landing = c.exact_landing; p.trace_steps = c.trace_steps;
  p.evolve_only = c.evolve_only ? 1 : 0;
  p.skip_evolve = 0; p.S_io = c.S; p.i_off = 0;
  p.full_B = 0; p.full_L = 0; p.row_off = 0; p.ts_ld = c.S + 1; p.seq = nullptr;
  if (!make_tableau(c.solver, p.tab)) return ODEVIO_E_ENUM;

  // ---- pre-pack weights into the workspace
  for (int j = 0; j < NL; ++j) {
    float* dst = ws + pl.off_Wode[j];
    ODEVIO_CUDA_TRY(transpose_pack(w->ode_w[j], pl.Node[j], pl.Kode[j], dst, pl.Node[j], 0, 0, stream));
    p.Wode[j] = dst; p.bode[j] = w->ode_b[j]; p.Kode[j] = pl.Kode[j]; p.Node[j] = pl.Node[j];
  }
  const size_t DD = static_cast<size_t>(D) * D;
  for (int l = 0; l < (c.evolve_only ? 0 : c.L); ++l) {
    if (pl.G == 1) {
      float* dst = ws + pl.off_Wrnn[l][0];
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l], D, D, dst, D, 0, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l], D, D, dst, D, D, 0, stream));
      float* bd = ws + pl.off_brnn[l][0];
      ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l], w->rnn_b_hh[l], bd, D, stream));
      p.Wrnn[l][0] = dst; p.brnn[l][0] = bd;
    } else {
      // PyTorch GRU gate order (r, z, n) along the rows of weight_ih / weight_hh
      float* wr = ws + pl.off_Wrnn[l][0]; float* wz = ws + pl.off_Wrnn[l][1];
      float* wi = ws + pl.off_Wrnn[l][2]; float* wh = ws + pl.off_Wrnn[l][3];
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l], D, D, wr, D, 0, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l], D, D, wr, D, D, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l] + DD, D, D, wz, D, 0, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l] + DD, D, D, wz, D, D, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l] + 2 * DD, D, D, wi, D, 0, 0, stream));
      ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l] + 2 * DD, D, D, wh, D, 0, 0, stream));
      float* br = ws + pl.off_brnn[l][0]; float* bz = ws + pl.off_brnn[l][1];
      float* bi = ws + pl.off_brnn[l][2]; float* bh = ws + pl.off_brnn[l][3];
      ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l], w->rnn_b_hh[l], br, D, stream));
      ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l] + D, w->rnn_b_hh[l] + D, bz, D, stream));
      ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l] + 2 * D, nullptr, bi, D, stream));
      ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_hh[l] + 2 * D, nullptr, bh, D, stream));
      p.Wrnn[l][0] = wr; p.Wrnn[l][1] = wz; p.Wrnn[l][2] = wi; p.Wrnn[l][3] = wh;
      p.brnn[l][0] = br; p.brnn[l][1] = bz; p.brnn[l][2] = bi; p.brnn[l][3] = bh;
    }
  }
  if (!c.evolve_only) {
    float* dst = ws + pl.off_Wreg0;
    ODEVIO_CUDA_TRY(transpose_pack(w->reg_w0, kRegHidden, D, dst, kRegHidden, 0, 0, stream));
    p.Wreg0 = dst; p.breg0 = w->reg_b0; p.Wreg1 = w->reg_w1; p.breg1 = w->reg_b1;
  }
  if (c.evolve_only) {
    // no fusion / jump / head weights are read
  } else if (w->fuse_w && w->fuse_b) {
    float* dst = ws + pl.off_Wfuse;
    ODEVIO_CUDA_TRY(transpose_pack(w->fuse_w, D, D, dst, D, 0, 0, stream));
    p.Wfuse = dst; p.bfuse = w->fuse_b;
  } else if (w->fuse_w || w->fuse_b) {
    return ODEVIO_E_NULL;
  }
  p.fv = fv; p.fi = fi; p.Dv = Dv; p.ts = ts; p.h0 = h0;
  p.pose = pose; p.hT = hT; p.stats = stats; p.status = status;
  p.scratch = ws + pl.off_scratch; p.scratch_floats_per_cta = pl.scratch_floats_per_cta;
  p.ntiles = pl.ntiles; p.nst = pl.nst; p.kc = pl.kc;
  p.bufA_floats = static_cast<int>(pl.bufA_floats); p.bufB_floats = static_cast<int>(pl.bufB_floats);
  p.stage_floats = static_cast<int>(pl.stage_floats);

  if (c.save_checkpoints) {
    p.nloops = static_cast<int*>(ckpt);
    p.ckpt = reinterpret_cast<float*>(static_cast<unsigned char*>(ckpt) + pl.ckpt_head_bytes);
    p.ckpt_floats_per_tile = pl.ckpt_floats_per_tile;
    p.CK = pl.CK;
  }

  g_last_precision = c.precision;
  if (c.precision == ODEVIO_PRECISION_FP16X3) {
    // Per interval: the 3xFP16 cluster kernel (odernn_h3.cu) evolves ALL L*B rows of the state in place (32 clusters of 4
    // take the 2048 rows of configs[1] in one round: no side launch), then the FMA kernel runs the jump + head.  (GRU,
    // "soft" fusion in the kernel, L > 2; the other configurations returned above with ONE launch per forward.)
    const size_t h_off = align_up(pl.total_bytes, 1024);
    const size_t h_bytes = align_up(odernn_h3_workspace_bytes(c), 1024);
    if (h_bytes == 0) return ODEVIO_E_SHAPE;
    if (workspace_bytes < h_off + h_bytes) return ODEVIO_E_WORKSPACE;
    H3Evolve h3;
    const int prc = h3.prepare(c, p.tab, p.adaptive != 0, w, false, !c.weights_prepacked, static_cast<unsigned char*>(workspace) + h_off,
                               h_bytes, stream);
    if (prc != 0) return prc;
    if (status) ODEVIO_CUDA_TRY(cudaMemsetAsync(status, 0, static_cast<size_t>(c.B) * sizeof(int32_t), stream));
    p.S = 1; p.skip_evolve = 1; p.S_io = c.S; p.stats = nullptr; p.status = nullptr; p.h0 = hT; p.hT = hT; p.ts = nullptr;
    for (int i = 0; i < c.S; ++i) {
      const int erc = h3.run(i == 0 ? h0 : hT, hT, ts, c.S + 1, i, 1, nullptr, nullptr, 0, c.S, nullptr, stats, status, stream);
      if (erc != 0) return erc;
      p.i_off = i;
      ODEVIO_CUDA_TRY(launch_odernn_fwd(p, pl.RT, pl.grid, pl.smem_bytes, stream));
    }
    return 0;
  }
  if (c.precision == ODEVIO_PRECISION_TF32X3) {
    // Per interval: the cluster kernel evolves all L*B rows of the state in place on the tensor cores, then the FMA
    // kernel runs the interval's jump + head (skip_evolve).  Same stream, no host synchronisation.
    const size_t tc_off = align_up(pl.total_bytes, 256);
    const size_t tc_bytes = align_up(odernn_tc_workspace_bytes(c), 256);
    const size_t side_bytes = tc_side_scratch_bytes(c, pl.nsm);
    if (tc_bytes == 0) return ODEVIO_E_SHAPE;
    if (workspace_bytes < tc_off + tc_bytes + side_bytes + tc_seq_table_bytes(c)) return ODEVIO_E_WORKSPACE;
    TcEvolve tc;
    const int prc = tc.prepare(c, p.tab, p.adaptive != 0, w->ode_w, w->ode_b, static_cast<unsigned char*>(workspace) + tc_off,
                               tc_bytes, stream);
    if (prc != 0) return prc;
    const size_t state_bytes = static_cast<size_t>(c.L) * c.B * D * sizeof(float);
    if (h0) { if (h0 != hT) ODEVIO_CUDA_TRY(cudaMemcpyAsync(hT, h0, state_bytes, cudaMemcpyDeviceToDevice, stream)); }
    else ODEVIO_CUDA_TRY(cudaMemsetAsync(hT, 0, state_bytes, stream));
    if (status) ODEVIO_CUDA_TRY(cudaMemsetAsync(status, 0, static_cast<size_t>(c.B) * sizeof(int32_t), stream));

    // Row split.  A B200 holds 15-16 clusters of 8 CTAs at once; every further 128-row tile would cost a whole extra
    // round of the latency-bound cluster kernel.  The cluster kernel therefore takes full rounds only (all L layer rows
    // of Bsub sequences), and when the other sequences fit one wave of the FFMA kernel on the SMs the clusters leave idle
    // they run there, concurrently (helper stream).  The FFMA tiles are slower per evaluation, so per interval they get
    // the sequences with the SHORTEST interval -- the fewest solver steps (tc_select_kernel).
    const int M = c.L * c.B;
    const int ntiles_tc = (M + 127) / 128, maxc = tc.max_clusters();
    int Bsub = c.B, n_side = 0;
    const int side_rt = c.L >= 2 ? 4 : kSideRT;
    if (maxc > 0 && ntiles_tc > maxc && ntiles_tc % maxc != 0 && c.B <= 8192) {
      const int full_rows = (ntiles_tc / maxc) * maxc * 128;
      const int idle_sms = pl.nsm - maxc * tc.cluster_size();
      if (full_rows % c.L == 0) {
        const int bs = full_rows / c.L, ns_ = c.B - bs;
        if (ns_ > 0 && (ns_ + side_rt - 1) / side_rt <= idle_sms && side_rt * c.L <= 8) { Bsub = bs; n_side = ns_; }
      }
    }
    FwdParams ps = p;          // side launch: evolve_only on the last n_side sequences of the per-interval order
    OdePlan pls;
    SideStream* side = nullptr;
    int* seq = nullptr;
    if (n_side > 0) {
      odevio_odernn_cfg cs = c;
      cs.B = n_side; cs.S = 1; cs.evolve_only = 1; cs.rows_per_tile = side_rt; cs.precision = ODEVIO_PRECISION_FP32;
      const int src = plan_odernn(cs, pls);
      if (src != 0) return src;
      const int ssrc = side_stream(&side);
      if (ssrc != 0) return ssrc;
      seq = reinterpret_cast<int*>(static_cast<unsigned char*>(workspace) + tc_off + tc_bytes + side_bytes);
      const int selrc = odernn_tc_select(ts, c.B, c.S, n_side, seq, stream);
      if (selrc != 0) return selrc;
      ps.B = n_side; ps.S = 1; ps.evolve_only = 1; ps.skip_evolve = 0;
      ps.full_B = c.B; ps.full_L = c.L; ps.row_off = Bsub; ps.ts_ld = c.S + 1; ps.ts = ts;
      ps.h0 = hT; ps.hT = hT;
      ps.stats = stats; ps.status = status; ps.pose = nullptr; ps.fv = nullptr; ps.fi = nullptr; ps.Wfuse = nullptr;
      ps.scratch = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + tc_off + tc_bytes);
      ps.scratch_floats_per_cta = pls.scratch_floats_per_cta;
      ps.ntiles = pls.ntiles; ps.nst = pls.nst; ps.kc = pls.kc;
      ps.bufA_floats = static_cast<int>(pls.bufA_floats); ps.bufB_floats = static_cast<int>(pls.bufB_floats);
      ps.stage_floats = static_cast<int>(pls.stage_floats);
    }
    p.S = 1; p.skip_evolve = 1; p.S_io = c.S; p.stats = nullptr; p.status = nullptr; p.h0 = hT; p.hT = hT; p.ts = nullptr;
    for (int i = 0; i < c.S; ++i) {
      if (side) {
        ODEVIO_CUDA_TRY(cudaEventRecord(side->fork, stream));
        ODEVIO_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
      }
      const int* seq_i = seq ? seq + static_cast<size_t>(i) * c.B : nullptr;
      const int erc = tc.evolve(hT, Bsub, seq_i, ts, c.S + 1, i, stats, status, stream);   // clusters first: they need whole GPCs
      if (erc != 0) return erc;
      if (side) {
        ps.i_off = i; ps.seq = seq_i;
        ODEVIO_CUDA_TRY(launch_odernn_fwd(ps, pls.RT, pls.grid, pls.smem_bytes, side->stream));
        ODEVIO_CUDA_TRY(cudaEventRecord(side->join, side->stream));
        ODEVIO_CUDA_TRY(cudaStreamWaitEvent(stream, side->join, 0));
      }
      if (!c.evolve_only) {
        p.i_off = i;
        ODEVIO_CUDA_TRY(launch_odernn_fwd(p, pl.RT, pl.grid, pl.smem_bytes, stream));
      }
    }
    return 0;
  }
  ODEVIO_CUDA_TRY(launch_odernn_fwd(p, pl.RT, pl.grid, pl.smem_bytes, stream));
  return 0;
}

// development only: out[0] = clusters launched, out[1] = co-resident maximum, out[2] = rows of the last tensor-core solver launch
int32_t odevio_debug_tc_geometry(int32_t* out) {
  if (!out) return ODEVIO_E_NULL;
  int a = 0, b = 0, r = 0;
  odernn_tc_last_geometry(&a, &b, &r);
  if (g_last_precision == ODEVIO_PRECISION_FP16X3) odernn_h3_last_geometry(&a, &b, &r);
  out[0] = a; out[1] = b; out[2] = r;
  return 0;
}

int32_t odevio_debug_tc_timeline(long long* host_dst) {
  return host_dst ? odernn_tc_debug_timeline(host_dst) : ODEVIO_E_NULL;
}

int32_t odevio_debug_cde_tc_timeline(long long* host_dst) {
  return host_dst ? static_cast<int32_t>(cde_tc_debug_timeline(host_dst)) : ODEVIO_E_NULL;
}

int32_t odevio_debug_h3_timeline(long long* host_dst) {
  return host_dst ? odernn_h3_debug_timeline(host_dst) : ODEVIO_E_NULL;
}

int32_t odevio_debug_tc_timing(int32_t enable, float* total_ms, int32_t* launches) {
  if (enable >= 0) { odernn_tc_timing_enable(enable != 0); odernn_h3_timing_enable(enable != 0); return 0; }
  if (!total_ms || !launches) return ODEVIO_E_NULL;
  int n = 0;
  const int rc = g_last_precision == ODEVIO_PRECISION_FP16X3 ? odernn_h3_timing_read(total_ms, &n) : odernn_tc_timing_read(total_ms, &n);
  *launches = n;
  return rc;
}

int32_t odevio_odernn_geometry(const odevio_odernn_cfg* cfg, int32_t* out) {
  if (!cfg || !out) return ODEVIO_E_NULL;
  OdePlan pl;
  const int rc = plan_odernn(*cfg, pl);
  if (rc != 0) return rc;
  DevTableau tb;
  if (!make_tableau(cfg->solver, tb)) return ODEVIO_E_ENUM;
  out[0] = pl.RT; out[1] = pl.R; out[2] = pl.ntiles;
  out[3] = tb.ssal ? tb.n_stages - 1 : tb.n_stages;
  out[4] = pl.CK; out[5] = pl.grid; out[6] = 0; out[7] = 0;
  return 0;
}

size_t odevio_odernn_ckpt_bytes(const odevio_odernn_cfg* cfg) {
  if (!cfg) return 0;
  OdePlan pl;
  if (plan_odernn(*cfg, pl) != 0) return 0;
  return ckpt_total_bytes(*cfg, pl);
}

size_t odevio_odernn_backward_workspace_bytes(const odevio_odernn_cfg* cfg, int64_t ode_rows) {
  if (!cfg || !cfg->save_checkpoints) return 0;
  OdePlan pl;
  if (plan_odernn(*cfg, pl) != 0) return 0;
  BwdPlan bp;
  if (plan_odernn_bwd(*cfg, pl, ode_rows, bp) != 0) return 0;
  return bp.total_bytes;
}

static int32_t odernn_backward_impl(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                                    const float* fv, const float* fi, int32_t Dv,
                                    const void* ckpt, size_t ckpt_bytes,
                                    const int64_t* rec_base, int64_t ode_rows, int64_t ode_rows_plan, int32_t i_lo, int32_t i_hi,
                                    const float* grad_pose, const float* grad_hT,
                                    const odevio_odernn_grads* g, float* grad_fused, float* grad_h0,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  if (!cfg || !w || !fv || !ckpt || !rec_base || !grad_pose || !g || !workspace) return ODEVIO_E_NULL;
  const odevio_odernn_cfg& c = *cfg;
  if (!c.save_checkpoints) return ODEVIO_E_ENUM;
  if (i_lo < 0 || i_hi >= c.S || i_lo > i_hi || ode_rows > ode_rows_plan) return ODEVIO_E_SHAPE;
  const bool first_range = i_hi == c.S - 1, last_range = i_lo == 0;
  OdePlan pl;
  int rc = plan_odernn(c, pl);
  if (rc != 0) return rc;
  BwdPlan bp;
  rc = plan_odernn_bwd(c, pl, ode_rows_plan, bp);
  if (rc != 0) return rc;
  if (ode_rows % pl.R) return ODEVIO_E_SHAPE;
  if (Dv <= 0 || Dv > c.D || (Dv < c.D && !fi) || (Dv == c.D && fi)) return ODEVIO_E_SHAPE;
  if (workspace_bytes < bp.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  if (ckpt_bytes < ckpt_total_bytes(c, pl) || (reinterpret_cast<uintptr_t>(ckpt) & 255)) return ODEVIO_E_WORKSPACE;
  const int NL = c.n_hidden + 1;
  for (int j = 0; j < NL; ++j) if (!w->ode_w[j] || !w->ode_b[j] || !g->ode_w[j] || !g->ode_b[j]) return ODEVIO_E_NULL;
  for (int l = 0; l < c.L; ++l)
    if (!w->rnn_w_ih[l] || !w->rnn_w_hh[l] || !g->rnn_w_ih[l] || !g->rnn_w_hh[l] || !g->rnn_b_ih[l] || !g->rnn_b_hh[l])
      return ODEVIO_E_NULL;
  if (!w->reg_w0 || !w->reg_b0 || !w->reg_w1 || !g->reg_w0 || !g->reg_b0 || !g->reg_w1 || !g->reg_b1) return ODEVIO_E_NULL;

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  const int D = c.D;
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.S = c.S; p.D = D; p.H = c.H; p.NL = NL; p.L = c.L;
  p.act = c.activation; p.rnn_type = c.rnn_type;
  if (!make_tableau(c.solver, p.tab)) return ODEVIO_E_ENUM;
  p.ns = bp.ns;
  for (int j = 0; j < NL; ++j) {
    float* dst = ws + bp.off_Wode[j];
    if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->ode_w[j], pl.Node[j], pl.Kode[j], dst, pl.Node[j], 0, 0, stream));
    p.Wode[j] = dst; p.bode[j] = w->ode_b[j]; p.Kode[j] = pl.Kode[j]; p.Node[j] = pl.Node[j];
    p.Wode_raw[j] = w->ode_w[j];
    p.recA_ode[j] = ws + bp.off_recA_ode[j]; p.recG_ode[j] = ws + bp.off_recG_ode[j];
    p.recA_ode_lo[j] = ws + bp.off_recA_ode_lo[j]; p.recG_ode_lo[j] = ws + bp.off_recG_ode_lo[j];
  }
  const size_t DDb = static_cast<size_t>(D) * D;
  for (int l = 0; l < c.L; ++l) {
    p.Wih_raw[l] = w->rnn_w_ih[l]; p.Whh_raw[l] = w->rnn_w_hh[l];
    p.recA_rnn[l] = ws + bp.off_recA_rnn[l]; p.recG_rnn[l] = ws + bp.off_recG_rnn[l];
    if (c.rnn_type == ODEVIO_RNN_GRU) {
      if (!w->rnn_b_ih[l] || !w->rnn_b_hh[l]) return ODEVIO_E_NULL;
      float* wr = ws + bp.off_Wrnn[l][0]; float* wz = ws + bp.off_Wrnn[l][1];
      float* wi = ws + bp.off_Wrnn[l][2]; float* wh = ws + bp.off_Wrnn[l][3];
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l], D, D, wr, D, 0, 0, stream));
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l], D, D, wr, D, D, 0, stream));
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l] + DDb, D, D, wz, D, 0, 0, stream));
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l] + DDb, D, D, wz, D, D, 0, stream));
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_ih[l] + 2 * DDb, D, D, wi, D, 0, 0, stream));
      if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->rnn_w_hh[l] + 2 * DDb, D, D, wh, D, 0, 0, stream));
      float* br = ws + bp.off_brnn[l][0]; float* bz = ws + bp.off_brnn[l][1];
      float* bi = ws + bp.off_brnn[l][2]; float* bh = ws + bp.off_brnn[l][3];
      if (first_range) ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l], w->rnn_b_hh[l], br, D, stream));
      if (first_range) ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l] + D, w->rnn_b_hh[l] + D, bz, D, stream));
      if (first_range) ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_ih[l] + 2 * D, nullptr, bi, D, stream));
      if (first_range) ODEVIO_CUDA_TRY(bias_sum(w->rnn_b_hh[l] + 2 * D, nullptr, bh, D, stream));
      p.Wrnn[l][0] = wr; p.Wrnn[l][1] = wz; p.Wrnn[l][2] = wi; p.Wrnn[l][3] = wh;
      p.brnn[l][0] = br; p.brnn[l][1] = bz; p.brnn[l][2] = bi; p.brnn[l][3] = bh;
    }
  }
  {
    float* dst = ws + bp.off_Wreg0;
    if (first_range) ODEVIO_CUDA_TRY(transpose_pack(w->reg_w0, kRegHidden, D, dst, kRegHidden, 0, 0, stream));
    p.Wreg0 = dst; p.breg0 = w->reg_b0; p.Wreg0_raw = w->reg_w0; p.Wreg1 = w->reg_w1;
  }
  p.recA_reg0 = ws + bp.off_recA_reg0; p.recG_reg0 = ws + bp.off_recG_reg0;
  p.recA_reg1 = ws + bp.off_recA_reg1; p.recG_reg1 = ws + bp.off_recG_reg1;
  p.fv = fv; p.fi = fi; p.Dv = Dv;
  p.gpose = grad_pose; p.ghT = grad_hT; p.gh0 = grad_h0; p.gfused = grad_fused;
  p.nloops = static_cast<const int*>(ckpt);
  p.ckpt = reinterpret_cast<const float*>(static_cast<const unsigned char*>(ckpt) + pl.ckpt_head_bytes);
  p.ckpt_floats_per_tile = pl.ckpt_floats_per_tile; p.CK = pl.CK;
  p.rec_base = reinterpret_cast<const long long*>(rec_base);
  p.Rb = bp.Rb;
  if (bp.Rb != pl.R) {
    // padding rows of every record block must read as zero in the weight-gradient GEMMs
    const size_t Mb = static_cast<size_t>(ode_rows / pl.R) * bp.Rb;
    for (int j = 0; j < NL; ++j) {
      ODEVIO_CUDA_TRY(cudaMemsetAsync(p.recA_ode[j], 0, sizeof(float) * Mb * pl.Kode[j], stream));
      ODEVIO_CUDA_TRY(cudaMemsetAsync(p.recA_ode_lo[j], 0, sizeof(float) * Mb * pl.Kode[j], stream));
      ODEVIO_CUDA_TRY(cudaMemsetAsync(p.recG_ode[j], 0, sizeof(float) * Mb * pl.Node[j], stream));
      ODEVIO_CUDA_TRY(cudaMemsetAsync(p.recG_ode_lo[j], 0, sizeof(float) * Mb * pl.Node[j], stream));
    }
  }
  p.scratch = ws + bp.off_scratch; p.scratch_floats_per_cta = bp.scratch_floats_per_cta;
  p.i_lo = i_lo; p.i_hi = i_hi; p.tile_gy = ws + bp.off_tile_gy;
  p.tile_order = reinterpret_cast<const int*>(ws + bp.off_tile_order);
  p.tile_counter = reinterpret_cast<int*>(ws + bp.off_tile_order) + pl.ntiles + 32;
  p.ntiles = pl.ntiles; p.nst = bp.nst; p.kc = bp.kc;
  p.buf_floats = static_cast<int>(bp.buf_floats); p.stage_floats = static_cast<int>(bp.stage_floats);

  ODEVIO_CUDA_TRY(launch_odernn_bwd(p, pl.RT, pl.grid, bp.smem_bytes, stream));

  // ---- deferred weight gradients: one dense GEMM per Linear over its record stream
  float* part = ws + bp.off_part;
  for (int j = 0; j < NL; ++j)       // ODEFunc Linears: tcgen05 3xTF32 GEMMs over the block-format streams
    ODEVIO_CUDA_TRY(wgrad_linear_tc_ex(p.recG_ode[j], p.recG_ode_lo[j], p.recA_ode[j], p.recA_ode_lo[j], ode_rows / pl.R,
                                       pl.Node[j], pl.Kode[j], bp.Rb, g->ode_w[j], g->ode_b[j], part, pl.nsm,
                                       first_range ? 0 : 1, stream));
  if (!last_range) return 0;          // the jump / head records are complete after the range that holds interval 0
  for (int l = 0; l < c.L; ++l) {
    if (c.rnn_type == ODEVIO_RNN_GRU) {
      // records: A = [x | h] (ld 2D), G = [G_ih | G_hh] (ld 6D); weight_ih / weight_hh are [3D][D]
      ODEVIO_CUDA_TRY(wgrad_linear(p.recG_rnn[l], 6 * D, p.recA_rnn[l], 2 * D, bp.jump_rows, 3 * D, D,
                                   g->rnn_w_ih[l], nullptr, 0, g->rnn_b_ih[l], nullptr, part, pl.nsm, stream));
      ODEVIO_CUDA_TRY(wgrad_linear(p.recG_rnn[l] + 3 * D, 6 * D, p.recA_rnn[l] + D, 2 * D, bp.jump_rows, 3 * D, D,
                                   g->rnn_w_hh[l], nullptr, 0, g->rnn_b_hh[l], nullptr, part, pl.nsm, stream));
    } else {
      ODEVIO_CUDA_TRY(wgrad_linear(p.recG_rnn[l], D, p.recA_rnn[l], 2 * D, bp.jump_rows, D, 2 * D,
                                   g->rnn_w_ih[l], g->rnn_w_hh[l], D, g->rnn_b_ih[l], g->rnn_b_hh[l], part, pl.nsm, stream));
    }
  }
  ODEVIO_CUDA_TRY(wgrad_linear(p.recG_reg0, kRegHidden, p.recA_reg0, D, bp.jump_rows, kRegHidden, D,
                               g->reg_w0, nullptr, 0, g->reg_b0, nullptr, part, pl.nsm, stream));
  ODEVIO_CUDA_TRY(wgrad_linear(p.recG_reg1, 8, p.recA_reg1, kRegHidden, bp.jump_rows, kPoseDim, kRegHidden,
                               g->reg_w1, nullptr, 0, g->reg_b1, nullptr, part, pl.nsm, stream));
  return 0;
}

int32_t odevio_odernn_backward(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                               const float* fv, const float* fi, int32_t Dv,
                               const void* ckpt, size_t ckpt_bytes,
                               const int64_t* rec_base, int64_t ode_rows,
                               const float* grad_pose, const float* grad_hT,
                               const odevio_odernn_grads* g, float* grad_fused, float* grad_h0,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (!cfg) return ODEVIO_E_NULL;
  return odernn_backward_impl(cfg, w, fv, fi, Dv, ckpt, ckpt_bytes, rec_base, ode_rows, ode_rows, 0, cfg->S - 1, grad_pose,
                              grad_hT, g, grad_fused, grad_h0, workspace, workspace_bytes, stream_);
}

int32_t odevio_odernn_backward_range(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                                     const float* fv, const float* fi, int32_t Dv,
                                     const void* ckpt, size_t ckpt_bytes,
                                     const int64_t* rec_base, int64_t ode_rows, int64_t ode_rows_plan,
                                     int32_t i_lo, int32_t i_hi,
                                     const float* grad_pose, const float* grad_hT,
                                     const odevio_odernn_grads* g, float* grad_fused, float* grad_h0,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  return odernn_backward_impl(cfg, w, fv, fi, Dv, ckpt, ckpt_bytes, rec_base, ode_rows, ode_rows_plan, i_lo, i_hi, grad_pose,
                              grad_hT, g, grad_fused, grad_h0, workspace, workspace_bytes, stream_);
}

void odevio_cde_default_cfg(odevio_cde_cfg* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->B = 1; cfg->S = 10; cfg->So = 10; cfg->Hc = 128; cfg->n_layers = 3;     // scripts/config.py:74-78
  cfg->activation = ODEVIO_ACT_TANH; cfg->solver = ODEVIO_CDE_SOLVER_DOPRI5; cfg->interp = ODEVIO_CDE_INTERP_LINEAR;
  cfg->atol = 1e-6f; cfg->rtol = 1e-4f;                                         // PoseCDE.py:101
  cfg->step_size = 0.0; cfg->max_steps = 100000; cfg->rows_per_tile = 0;
}

size_t odevio_cde_workspace_bytes(const odevio_cde_cfg* cfg) {
  if (!cfg) return 0;
  CdePlan pl;
  if (plan_cde(*cfg, pl) != 0) return 0;
  return pl.total_bytes;
}

static int32_t cde_forward_impl(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                                const float* tobs, const float* fv, const float* fi, int32_t Dv,
                                const double* tout, const float* z0_in,
                                float* pose, float* z0_out, float* hidden, int32_t* stats,
                                void* ckpt, size_t ckpt_bytes, int32_t ckpt_steps,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  if (!cfg || !w || !tobs || !fv || !tout || !pose || !z0_out || !workspace) return ODEVIO_E_NULL;
  const odevio_cde_cfg& c = *cfg;
  CdePlan pl;
  const int rc = plan_cde(c, pl);
  if (rc != 0) return rc;
  if (ckpt) {
    if (ckpt_steps < 1 || !hidden) return ODEVIO_E_SHAPE;
    const size_t need = cde_ckpt_log_bytes(ckpt_steps) + cde_ckpt_step_floats(c, pl) * ckpt_steps * sizeof(float);
    if (ckpt_bytes < need || (reinterpret_cast<uintptr_t>(ckpt) & 255)) return ODEVIO_E_WORKSPACE;
  }
  if (Dv <= 0 || Dv > c.Hc || (Dv < c.Hc && !fi) || (Dv == c.Hc && fi)) return ODEVIO_E_SHAPE;
  if (workspace_bytes < pl.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  for (int j = 0; j <= c.n_layers; ++j) if (!w->cde_w[j] || !w->cde_b[j]) return ODEVIO_E_NULL;
  if (!w->init_w || !w->init_b || !w->reg_w0 || !w->reg_b0 || !w->reg_w1 || !w->reg_b1) return ODEVIO_E_NULL;

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  CdeParams p;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.S = c.S; p.So = c.So; p.Hc = c.Hc; p.C = pl.C; p.Cpad = pl.Cpad; p.NM = c.n_layers;
  p.act = c.activation; p.solver = c.solver; p.interp = c.interp;
  p.atol = c.atol; p.rtol = c.rtol; p.step_size = c.step_size; p.max_steps = c.max_steps;
  for (int j = 0; j < c.n_layers; ++j) {
    float* dst = ws + pl.off_Wmlp[j];
    ODEVIO_CUDA_TRY(transpose_pack(w->cde_w[j], c.Hc, c.Hc, dst, c.Hc, 0, 0, stream));
    p.Wmlp[j] = dst; p.bmlp[j] = w->cde_b[j];
  }
  if (pl.tc) {
    unsigned char* wimg = reinterpret_cast<unsigned char*>(ws + pl.off_Wimg);
    ODEVIO_CUDA_TRY(cde_tc_pack(w->cde_w[c.n_layers], w->cde_b[c.n_layers], c.Hc, pl.C, wimg, ws + pl.off_bval,
                                ws + pl.off_W0t, ws + pl.off_b0, stream));
    p.Wimg = wimg; p.bval = ws + pl.off_bval; p.W0t = ws + pl.off_W0t; p.b0 = ws + pl.off_b0;
    p.Bpad = pl.Bpad; p.nrt = pl.nrt; p.RP = pl.RP; p.dx_cache = pl.dx_cache;
    { const char* ft = getenv("ODEVIO_CDE_FAST_TANH"); p.fast_tanh = (ft && ft[0] == '1') ? 1 : 0; }
    p.state = ws + pl.off_state; p.Ximg = reinterpret_cast<unsigned char*>(ws + pl.off_Ximg); p.dXg = ws + pl.off_dXg;
  } else {
    ODEVIO_CUDA_TRY(cde_pack_final(w->cde_w[c.n_layers], w->cde_b[c.n_layers], c.Hc, pl.C, pl.Gc, pl.ngroups,
                                   ws + pl.off_Wfin, ws + pl.off_bfin, stream));
    p.Wfin = ws + pl.off_Wfin; p.bfin = ws + pl.off_bfin; p.Gc = pl.Gc; p.ngroups = pl.ngroups; p.Ng = pl.Ng;
  }
  ODEVIO_CUDA_TRY(cudaMemsetAsync(ws + pl.off_Winit, 0, sizeof(float) * pl.Cpad * c.Hc, stream));
  ODEVIO_CUDA_TRY(transpose_pack(w->init_w, c.Hc, pl.C, ws + pl.off_Winit, c.Hc, 0, 0, stream));
  p.Winit = ws + pl.off_Winit; p.binit = w->init_b;
  ODEVIO_CUDA_TRY(transpose_pack(w->reg_w0, kRegHidden, c.Hc, ws + pl.off_Wreg0, kRegHidden, 0, 0, stream));
  p.Wreg0 = ws + pl.off_Wreg0; p.breg0 = w->reg_b0; p.Wreg1 = w->reg_w1; p.breg1 = w->reg_b1;
  p.tobs = tobs; p.fv = fv; p.fi = fi; p.Dv = Dv; p.tout = tout; p.z0_in = z0_in;
  p.pose = pose; p.z0_out = z0_out; p.hout = hidden; p.stats = stats;
  if (!pl.tc) { p.scratch = ws + pl.off_scratch; p.scratch_floats_per_tile = pl.scratch_floats_per_tile; }
  p.red = reinterpret_cast<double*>(ws + pl.off_red);
  p.bar = reinterpret_cast<unsigned int*>(ws + pl.off_bar);
  ODEVIO_CUDA_TRY(cudaMemsetAsync(p.bar, 0, 256, stream));
  p.ntiles = pl.ntiles; p.nst = pl.nst;
  p.buf_floats = static_cast<int>(pl.buf_floats); p.stage_floats = static_cast<int>(pl.stage_floats);
  p.staging_floats = static_cast<int>(pl.staging_floats);
  if (ckpt) {
    p.log = static_cast<CdeStepRec*>(ckpt);
    p.ckpt = reinterpret_cast<float*>(static_cast<unsigned char*>(ckpt) + cde_ckpt_log_bytes(ckpt_steps));
    p.ckpt_cap = ckpt_steps;
    ODEVIO_CUDA_TRY(cudaMemsetAsync(ckpt, 0, sizeof(CdeStepRec), stream));
  }
  DevTableau tab;
  if (!make_tableau(ODEVIO_SOLVER_DOPRI5, tab)) return ODEVIO_E_ENUM;
  if (pl.tc) ODEVIO_CUDA_TRY(launch_cde_tc(p, tab, pl.grid, pl.smem_bytes, stream));
  else ODEVIO_CUDA_TRY(launch_cde_fwd(p, tab, pl.RT, pl.LL, pl.grid, pl.smem_bytes, stream));
  return 0;
}

int32_t odevio_cde_forward(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                           const float* tobs, const float* fv, const float* fi, int32_t Dv,
                           const double* tout, const float* z0_in,
                           float* pose, float* z0_out, float* hidden, int32_t* stats,
                           void* workspace, size_t workspace_bytes, void* stream_) {
  return cde_forward_impl(cfg, w, tobs, fv, fi, Dv, tout, z0_in, pose, z0_out, hidden, stats, nullptr, 0, 0,
                          workspace, workspace_bytes, stream_);
}

size_t odevio_cde_ckpt_bytes(const odevio_cde_cfg* cfg, int32_t ckpt_steps) {
  if (!cfg || ckpt_steps < 1) return 0;
  CdePlan pl;
  if (plan_cde(*cfg, pl) != 0) return 0;
  return cde_ckpt_log_bytes(ckpt_steps) + cde_ckpt_step_floats(*cfg, pl) * ckpt_steps * sizeof(float);
}

int32_t odevio_cde_forward_ckpt(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                                const float* tobs, const float* fv, const float* fi, int32_t Dv,
                                const double* tout, const float* z0_in,
                                float* pose, float* z0_out, float* hidden, int32_t* stats,
                                void* ckpt, size_t ckpt_bytes, int32_t ckpt_steps,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  if (!ckpt) return ODEVIO_E_NULL;
  return cde_forward_impl(cfg, w, tobs, fv, fi, Dv, tout, z0_in, pose, z0_out, hidden, stats, ckpt, ckpt_bytes,
                          ckpt_steps, workspace, workspace_bytes, stream_);
}

size_t odevio_cde_backward_workspace_bytes(const odevio_cde_cfg* cfg, int32_t chunk_vjps) {
  if (!cfg) return 0;
  odevio_cde_cfg cb = *cfg;
  cb.precision = ODEVIO_PRECISION_FP32;            // the backward kernel's geometry is the CUDA-core one whoever ran the forward
  CdePlan pl;
  if (plan_cde(cb, pl) != 0) return 0;
  CdeBwdPlan bp;
  if (plan_cde_bwd(*cfg, pl, chunk_vjps, bp) != 0) return 0;
  return bp.total_bytes;
}

int32_t odevio_cde_backward(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                            const float* tobs, const float* fv, const float* fi, int32_t Dv,
                            const double* tout, int32_t has_prev, const float* hidden, const float* z0,
                            const void* ckpt, size_t ckpt_bytes, int32_t ckpt_steps,
                            const int32_t* vjp_base, int32_t n_accepted, int32_t chunk_vjps,
                            const float* grad_pose, const float* grad_z0,
                            const odevio_cde_grads* g, float* grad_x, float* grad_prev,
                            void* workspace, size_t workspace_bytes, void* stream_) {
  if (!cfg || !w || !tobs || !fv || !tout || !hidden || !z0 || !ckpt || !vjp_base || !grad_pose || !g || !workspace)
    return ODEVIO_E_NULL;
  odevio_cde_cfg c = *cfg;
  CdePlan plf;                                     // the forward's plan: its checkpoint layout
  int rc = plan_cde(c, plf);
  if (rc != 0) return rc;
  c.precision = ODEVIO_PRECISION_FP32;             // the backward kernel's geometry is the CUDA-core one whoever ran the forward
  CdePlan pl;
  rc = plan_cde(c, pl);
  if (rc != 0) return rc;
  CdeBwdPlan bp;
  rc = plan_cde_bwd(c, pl, chunk_vjps, bp);
  if (rc != 0) return rc;
  if (Dv <= 0 || Dv > c.Hc || (Dv < c.Hc && !fi) || (Dv == c.Hc && fi)) return ODEVIO_E_SHAPE;
  if (n_accepted < 0 || n_accepted > ckpt_steps) return ODEVIO_E_SHAPE;
  if (has_prev && !grad_prev) return ODEVIO_E_NULL;
  if (workspace_bytes < bp.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  {
    const size_t need = cde_ckpt_log_bytes(ckpt_steps) + cde_ckpt_step_floats(c, plf) * ckpt_steps * sizeof(float);
    if (ckpt_bytes < need || (reinterpret_cast<uintptr_t>(ckpt) & 255)) return ODEVIO_E_WORKSPACE;
  }
  const int NM = c.n_layers, Hc = c.Hc;
  for (int j = 0; j <= NM; ++j) if (!w->cde_w[j] || !w->cde_b[j] || !g->cde_w[j] || !g->cde_b[j]) return ODEVIO_E_NULL;
  if (!w->init_w || !w->init_b || !w->reg_w0 || !w->reg_b0 || !w->reg_w1) return ODEVIO_E_NULL;
  if (!g->reg_w0 || !g->reg_b0 || !g->reg_w1 || !g->reg_b1) return ODEVIO_E_NULL;
  if (!has_prev && (!g->init_w || !g->init_b)) return ODEVIO_E_NULL;

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  CdeBwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.S = c.S; p.So = c.So; p.Hc = Hc; p.C = pl.C; p.Cpad = pl.Cpad; p.NM = NM;
  p.act = c.activation; p.solver = c.solver; p.interp = c.interp;
  DevTableau tab;
  if (!make_tableau(c.solver == ODEVIO_CDE_SOLVER_DOPRI5 ? ODEVIO_SOLVER_DOPRI5 : ODEVIO_SOLVER_RK4_38, tab)) return ODEVIO_E_ENUM;
  p.ns = tab.n_stages; p.fsal = tab.fsal;
  for (int j = 0; j < NM; ++j) {
    float* dst = ws + bp.off_Wmlp[j];
    ODEVIO_CUDA_TRY(transpose_pack(w->cde_w[j], Hc, Hc, dst, Hc, 0, 0, stream));
    p.Wmlp[j] = dst; p.bmlp[j] = w->cde_b[j]; p.Wmlp_raw[j] = w->cde_w[j];
  }
  ODEVIO_CUDA_TRY(cde_pack_final(w->cde_w[NM], w->cde_b[NM], Hc, pl.C, pl.Gc, pl.ngroups, ws + bp.off_Wfin,
                                 ws + bp.off_bfin, stream));
  ODEVIO_CUDA_TRY(cde_pack_final_t(w->cde_w[NM], Hc, pl.C, pl.Gc, pl.ngroups, ws + bp.off_WfinT, stream));
  p.Wfin = ws + bp.off_Wfin; p.bfin = ws + bp.off_bfin; p.WfinT = ws + bp.off_WfinT;
  p.Gc = pl.Gc; p.ngroups = pl.ngroups; p.Ng = pl.Ng;
  ODEVIO_CUDA_TRY(cudaMemsetAsync(ws + bp.off_WinitP, 0, sizeof(float) * Hc * pl.Cpad, stream));
  ODEVIO_CUDA_TRY(cudaMemcpy2DAsync(ws + bp.off_WinitP, sizeof(float) * pl.Cpad, w->init_w, sizeof(float) * pl.C,
                                    sizeof(float) * pl.C, Hc, cudaMemcpyDeviceToDevice, stream));
  p.WinitP = ws + bp.off_WinitP;
  ODEVIO_CUDA_TRY(transpose_pack(w->reg_w0, kRegHidden, Hc, ws + bp.off_Wreg0, kRegHidden, 0, 0, stream));
  p.Wreg0 = ws + bp.off_Wreg0; p.breg0 = w->reg_b0; p.Wreg0_raw = w->reg_w0; p.Wreg1 = w->reg_w1;
  p.tobs = tobs; p.fv = fv; p.fi = fi; p.Dv = Dv; p.tout = tout;
  p.hidden = hidden; p.z0 = z0; p.has_prev = has_prev ? 1 : 0;
  p.gpose = grad_pose; p.gz0 = grad_z0; p.gX = grad_x; p.gprev = grad_prev;
  p.log = static_cast<const CdeStepRec*>(ckpt);
  p.ckpt = reinterpret_cast<const float*>(static_cast<const unsigned char*>(ckpt) + cde_ckpt_log_bytes(ckpt_steps));
  p.n_acc = n_accepted;
  {
    const size_t arr = static_cast<size_t>(Hc) * pl.R;
    if (plf.tc) {          // cde_tc.cu: [step][9][Hc][Bpad]; a tile's rows start at column tile * R
      p.ck_step_stride = static_cast<size_t>(2 + kMaxStages) * Hc * plf.Bpad; p.ck_tile_stride = pl.R;
      p.ck_jstride = static_cast<size_t>(Hc) * plf.Bpad; p.ck_hstride = plf.Bpad;
    } else {               // cde_fwd.cu: [step][tile][9][Hc][R]
      p.ck_step_stride = static_cast<size_t>(pl.ntiles) * (2 + kMaxStages) * arr; p.ck_tile_stride = (2 + kMaxStages) * arr;
      p.ck_jstride = arr; p.ck_hstride = pl.R;
    }
  }
  for (int j = 0; j <= NM; ++j) p.recA[j] = ws + bp.off_recA[j];
  for (int j = 0; j < NM; ++j) p.recG[j] = ws + bp.off_recG[j];
  p.recGf = ws + bp.off_recGf;
  p.recA_reg0 = ws + bp.off_recA_reg0; p.recG_reg0 = ws + bp.off_recG_reg0;
  p.recA_reg1 = ws + bp.off_recA_reg1; p.recG_reg1 = ws + bp.off_recG_reg1;
  p.recA_init = ws + bp.off_recA_init; p.recG_init = ws + bp.off_recG_init;
  p.tile_state = ws + bp.off_tile_state;
  p.scratch = ws + bp.off_scratch; p.scratch_floats_per_cta = bp.scratch_floats_per_cta;
  p.ntiles = pl.ntiles; p.nst = bp.nst;
  p.buf_floats = static_cast<int>(bp.buf_floats); p.stage_floats = static_cast<int>(bp.stage_floats);
  p.staging_floats = static_cast<int>(bp.staging_floats);

  float* part = ws + bp.off_part;
  float* dWp = ws + bp.off_dWp; float* dbp = ws + bp.off_dbp;
  const long long rows_per_vjp = static_cast<long long>(pl.ntiles) * pl.R;
  // walk the log backwards, at most chunk_vjps pullbacks (record rows) per launch
  int hi = n_accepted;
  bool first = true;
  do {
    int lo = hi;
    while (lo > 0 && vjp_base[hi] - vjp_base[lo - 1] <= chunk_vjps) --lo;
    if (lo == hi && hi > 0) return ODEVIO_E_SHAPE;          // one step alone exceeds the chunk
    p.step_lo = lo; p.step_hi = hi; p.vjp_lo = n_accepted > 0 ? vjp_base[lo] : 0;
    ODEVIO_CUDA_TRY(launch_cde_bwd(p, tab, pl.RT, pl.LL, pl.grid, bp.smem_bytes, stream));
    const long long M = n_accepted > 0 ? static_cast<long long>(vjp_base[hi] - vjp_base[lo]) * rows_per_vjp : 0;
    const int acc = first ? 0 : 1;
    for (int j = 0; j < NM; ++j)
      ODEVIO_CUDA_TRY(wgrad_linear_ex(p.recG[j], Hc, p.recA[j], Hc, M, Hc, Hc, g->cde_w[j], nullptr, 0, g->cde_b[j], nullptr,
                                      part, pl.nsm, acc, stream));
    ODEVIO_CUDA_TRY(wgrad_linear_ex(p.recGf, bp.NgTot, p.recA[NM], Hc, M, bp.NgTot, Hc, dWp, nullptr, 0, dbp, nullptr,
                                    part, pl.nsm, acc, stream));
    first = false;
    hi = lo;
  } while (hi > 0);
  ODEVIO_CUDA_TRY(cde_unpack_final_grad(dWp, dbp, Hc, pl.C, pl.Gc, g->cde_w[NM], g->cde_b[NM], stream));
  const long long BS = static_cast<long long>(c.B) * c.S;
  ODEVIO_CUDA_TRY(wgrad_linear(p.recG_reg0, kRegHidden, p.recA_reg0, Hc, BS, kRegHidden, Hc, g->reg_w0, nullptr, 0,
                               g->reg_b0, nullptr, part, pl.nsm, stream));
  ODEVIO_CUDA_TRY(wgrad_linear(p.recG_reg1, 8, p.recA_reg1, kRegHidden, BS, kPoseDim, kRegHidden, g->reg_w1, nullptr, 0,
                               g->reg_b1, nullptr, part, pl.nsm, stream));
  if (!has_prev) {
    // K = Cpad (the GEMM's float4 stores need K % 4 == 0; the padded input columns are zero), then drop the padding
    float* dWi = ws + bp.off_dWinit;
    ODEVIO_CUDA_TRY(wgrad_linear(p.recG_init, Hc, p.recA_init, pl.Cpad, c.B, Hc, pl.Cpad, dWi, nullptr, 0, g->init_b,
                                 nullptr, part, pl.nsm, stream));
    ODEVIO_CUDA_TRY(cudaMemcpy2DAsync(g->init_w, sizeof(float) * pl.C, dWi, sizeof(float) * pl.Cpad, sizeof(float) * pl.C,
                                      Hc, cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

}  // extern "C"
