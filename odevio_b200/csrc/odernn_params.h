// Kernel parameter block of the fused ODE-RNN kernels (host <-> device, internal).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace odevio {

constexpr int kMaxStages = 7;
constexpr int kMaxLinears = 6;
constexpr int kMaxRnnLayers = 4;
constexpr int kRegHidden = 128;   // regressor: Linear(D,128) -> LeakyReLU(0.1) -> Linear(128,6)
constexpr int kPoseDim = 6;

// Explicit Runge-Kutta tableau in fp32 (same numbers as oracle/tableaus.py).
struct DevTableau {
  int n_stages;
  int fsal;       // last stage's vector field is next step's first
  int ssal;       // y1 is the last stage's argument
  int has_err;    // embedded error estimate present
  int has_mid;    // quartic dense output available (else linear)
  float exponent; // -1 / order
  float a[kMaxStages][kMaxStages];
  float b[kMaxStages];
  float e[kMaxStages];
  float bmid[kMaxStages];
};

struct FwdParams {
  // dims / enums
  int B, S, D, H, NL, L;     // NL = number of Linear layers of ODEFunc (= n_hidden + 1)
  int act, rnn_type, adaptive, substeps;
  // controller
  float atol, rtol, dt0, safety, fmin, fmax;
  int accept_strict, floor_factor, endpoint_dense, max_steps, exact_landing, trace_steps;
  int evolve_only;           // PoseODERNN.evolve_state: ODE solves only (no jump / head)
  int skip_evolve;           // jump + head only (the state was evolved by the tensor-core solver kernel); ts is not read
  int S_io, i_off;           // features / poses of interval i live at row b * S_io + i_off + i (S_io = S, i_off = 0 normally)
  // sub-launch on part of the sequences of a larger [full_L = L, full_B] problem (tensor-core mode: the sequences the
  // cluster kernel leaves to the FFMA kernel): local sequence b is sequence bb = seq[row_off + b] of the full problem;
  // hidden state (h0 / hT are the FULL [L, full_B, D] arrays), timestamps, stats and status are addressed with bb.
  // full_B = 0: plain launch.
  int full_B, full_L, row_off, ts_ld;
  const int* seq;
  DevTableau tab;
  // packed weights (K-major [K][N]) and biases
  const float* Wode[kMaxLinears];
  const float* bode[kMaxLinears];
  int Kode[kMaxLinears], Node[kMaxLinears];
  // rnn (tanh): [0] = [W_ih^T ; W_hh^T] ([2D][D]), bias[0] = b_ih + b_hh
  // gru:        [0] = r-gate cat ([2D][D]), [1] = z-gate cat, [2] = W_in^T ([D][D]), [3] = W_hn^T
  const float* Wrnn[kMaxRnnLayers][4];
  const float* brnn[kMaxRnnLayers][4];
  const float* Wreg0;  // [D][128] packed
  const float* breg0;  // [128]
  const float* Wreg1;  // [6][128]  (PyTorch layout, used directly)
  const float* breg1;  // [6]
  const float* Wfuse;  // FusionModule 'soft': packed [D][D] or nullptr ('cat')
  const float* bfuse;  // [D]
  // io
  const float* fv; const float* fi; int Dv;
  const float* ts; const float* h0;
  float* pose; float* hT; int* stats; int* status;
  // per-CTA global scratch (L2 resident), T-layout [D][R] arrays: K[0..6], Y, Y1
  float* scratch; size_t scratch_floats_per_cta;
  // launch geometry
  int ntiles, nst;
  int bufA_floats, bufB_floats, stage_floats;
  int kc;                    // k-rows per weight stage (8 or 16; divides D, H)
  // training checkpoints (nullptr: inference).  Per tile, per interval i (T-layout arrays of D*R floats):
  //   [Yend | Ypost | CK x (Y0 | dt[R] | upd[R])]; nloops[tile * S + i] = stored solver iterations
  float* ckpt; int* nloops; size_t ckpt_floats_per_tile; int CK;
};

inline __host__ __device__ size_t ckpt_interval_floats(int D, int R, int CK) {
  return static_cast<size_t>(2) * D * R + static_cast<size_t>(CK) * (static_cast<size_t>(D) * R + 2 * R);
}

// Backward kernel parameters (odernn_bwd.cu).
struct BwdParams {
  int B, S, D, H, NL, L;
  int act, rnn_type;
  DevTableau tab;
  int ns;                                    // stages that enter y1 (n_stages - 1 for FSAL/SSAL tableaus)
  // forward-packed (K-major) weights for the stage re-evaluation
  const float* Wode[kMaxLinears];
  const float* bode[kMaxLinears];
  int Kode[kMaxLinears], Node[kMaxLinears];
  // PyTorch-layout [out][in] weights: the K-major operand of the transposed products W^T g
  const float* Wode_raw[kMaxLinears];
  const float* Wih_raw[kMaxRnnLayers];       // rnn.weight_ih_l{k}: [G*D][D], G = 1 | 3 (gates r, z, n)
  const float* Whh_raw[kMaxRnnLayers];
  // GRU only: forward-packed gate weights / biases for the gate re-evaluation (same packing as FwdParams)
  const float* Wrnn[kMaxRnnLayers][4];
  const float* brnn[kMaxRnnLayers][4];
  const float* Wreg0;      // packed [D][128]
  const float* breg0;
  const float* Wreg0_raw;  // [128][D]
  const float* Wreg1;      // [6][128]
  // io
  const float* fv; const float* fi; int Dv;
  const float* gpose;      // [B,S,6]
  const float* ghT;        // [L,B,D] or nullptr
  int i_lo, i_hi;          // this launch walks the intervals i_hi .. i_lo (a training step may be split into interval ranges)
  float* tile_gy;          // [ntiles][D*R]: gradient of the hidden state carried from one range to the next
  const int* tile_order;   // [ntiles] tiles by decreasing cost of this launch (stored solver iterations of the range), or nullptr
  int* tile_counter;       // work queue head: a CTA takes tile_order[atomicAdd(tile_counter, 1)] (zeroed by the ordering kernel)
  float* gh0;              // [L,B,D]
  float* gfused;           // [B,S,D] or nullptr
  // checkpoints written by the forward
  const float* ckpt; const int* nloops; size_t ckpt_floats_per_tile; int CK;
  // record streams (row-major) for the deferred weight-gradient GEMMs
  // ODE-layer streams: tcgen05 operand blocks of R rows (tile_gemm.cuh: rec_block_offset), hi / lo parts
  const long long* rec_base;                 // [ntiles * S] first ODE-stream row of (tile, interval); block = row / R
  int Rb;                                    // rows per record block = R rounded up to 8 (padding rows stay zero)
  float* recA_ode[kMaxLinears]; float* recG_ode[kMaxLinears];     // hi: [blocks][K_j * R], [blocks][N_j * R]
  float* recA_ode_lo[kMaxLinears]; float* recG_ode_lo[kMaxLinears];
  float* recA_rnn[kMaxRnnLayers]; float* recG_rnn[kMaxRnnLayers]; // [ntiles*S*RT][2D], [..][D] (GRU: [..][6D] = [G_ih | G_hh])
  float* recA_reg0; float* recG_reg0;        // [ntiles*S*RT][D], [..][128]
  float* recA_reg1; float* recG_reg1;        // [ntiles*S*RT][128], [..][8] (6 used)
  // per-CTA scratch: K[7], GZ[7], GY (D*R each), HS[7][NL-1] (H*R each)
  float* scratch; size_t scratch_floats_per_cta;
  int ntiles, nst;
  int buf_floats, stage_floats;
  int kc;                                    // k-rows per weight stage (8 or 16)
};

}  // namespace odevio
