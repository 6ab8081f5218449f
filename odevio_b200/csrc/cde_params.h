// Kernel parameter block of the fused Neural-CDE forward (host <-> device, internal).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "odernn_params.h"

namespace odevio {

constexpr int kCdeMaxOut = 64;        // output times per call (S)

enum { CDE_SOLVER_DOPRI5 = 0, CDE_SOLVER_RK4 = 1 };
enum { CDE_INTERP_LINEAR = 0, CDE_INTERP_CUBIC = 1 };

struct CdeStepRec;

struct CdeParams {
  int B, S, So;            // sequences, output times, observations in the control path (So >= 2)
  int Hc, C, Cpad;         // hidden, channels (Hc + 1), channels padded to a multiple of 8
  int NM;                  // Hc -> Hc Linear layers of CDEFunc (= cde_fn_num_layers)
  int act, solver, interp;
  float atol, rtol;
  double step_size;        // rk4: > 0 -> fixed grid of this spacing with linear output interpolation
  int max_steps;
  // CDEFunc MLP, packed K-major [Hc][Hc]
  const float* Wmlp[kMaxLinears];
  const float* bmlp[kMaxLinears];
  // final Linear Hc -> Hc*C, packed per channel group g: [Hc][Gc*Hc] with column c_local*Hc + h
  const float* Wfin; const float* bfin;
  int Gc, ngroups, Ng;     // channels per group, groups, Ng = Gc * Hc
  const float* Winit;      // packed [Cpad][Hc]
  const float* binit;
  const float* Wreg0;      // packed [Hc][128]
  const float* breg0;
  const float* Wreg1;      // [6][128]
  const float* breg1;
  // io
  const float* tobs;       // [B, So]   channel 0 of the observations
  const float* fv; const float* fi; int Dv;   // [B, So, Dv], [B, So, Hc - Dv]
  const double* tout;      // [S] output times (integration variable)
  const float* z0_in;      // [B, Hc] or nullptr
  float* pose;             // [B, S, 6]
  float* z0_out;           // [B, Hc]
  float* hout;             // [B, S, Hc] or nullptr
  int* stats;              // [4]: n_steps, n_accepted, n_f_evals, status
  // per-TILE global scratch (T-layout [Hc][R]): Z, Y1, K[7]
  float* scratch; size_t scratch_floats_per_tile;
  // grid-wide reduction / barrier
  double* red;             // [2][grid][2]
  unsigned int* bar;       // monotonic arrival counter (zeroed by the host before launch)
  int ntiles, nst;
  int buf_floats, stage_floats, staging_floats;
  // training: checkpoints of every accepted step (cde_bwd.cu): ckpt[step][tile][Z, Y1, K0..K6][Hc][R] + the step log
  float* ckpt; int ckpt_cap;
  struct CdeStepRec* log;  // [1 + ckpt_cap]: slot 0 is the header (CdeLogHead)
  // tensor-core kernel (cde_tc.cu): CTA h owns hidden unit h of the final Linear for ALL rows
  int Bpad, nrt, RP;            // rows padded to 128, row tiles, rows per CTA in the row phase
  float* state;                 // [9][Hc][Bpad] fp32 feature-major: Z, Y1, K0..K6
  unsigned char* Ximg;          // [nrt][hi|lo][128 x Hc] fp16, K-major canonical image of the final Linear's input
  float* dXg;                   // [nrt][Hc / 4][128 rows][4]  dX/dt of the value channels
  const unsigned char* Wimg;    // [Hc (h)][hi|lo][Hc (c) x Hc (k)] fp16 K-major canonical image: W[h*C + 1 + c][k]
  const float* bval;            // [Hc (h)][Hc (c)]  bias of the value channels
  const float* W0t;             // [Hc (k)][Hc (h)]  time channel W[h*C + 0][k], K-major
  const float* b0;              // [Hc]
  int dx_cache;                 // (m, d) of the current knot segment cached in shared memory ([2][RP][C] floats fit)
  int fast_tanh;                // epilogue tanh on the SFU approximations (absolute error ~2e-7) instead of tanhf
};

// One accepted solver step as the backward needs it (the step sizes are constants of the pullback).
struct CdeStepRec {
  double ta, tb;           // the step [ta, tb] in the integration variable
  float dt_s, ta_s, tb_s;  // the same in the state dtype, as the stages used them
  int on_jump;             // the step was shortened onto a knot: K0 of the next step is f(tb^+, y1), not k6
  int out_first, out_count;// output times interpolated from this step
  int vjp_base;            // vector-field pullbacks of all earlier steps (record-stream row block)
  int pad;
};
struct CdeLogHead { int n_acc, n_vjp, status, pad[9]; };
static_assert(sizeof(CdeStepRec) == 48 && sizeof(CdeLogHead) == 48, "step log layout");

// vector-field pullbacks of one accepted step: every stage that is evaluated inside the step, the re-evaluation after a
// knot, and (step 0 of an FSAL scheme) the initial f(t0, z0)
__host__ __device__ inline int cde_step_vjps(int ns, int fsal, int on_jump, int step) {
  return (fsal ? ns - 1 : ns) + (on_jump ? 1 : 0) + ((fsal && step == 0) ? 1 : 0);
}

// Parameter block of the fused CDE backward (cde_bwd.cu).
struct CdeBwdParams {
  int B, S, So, Hc, C, Cpad, NM, act, solver, interp;
  int ns, fsal;
  const float* Wmlp[kMaxLinears];      // packed K-major [Hc][Hc] (recompute)
  const float* bmlp[kMaxLinears];
  const float* Wmlp_raw[kMaxLinears];  // PyTorch [out][in]: the K-major operand of W^T g
  const float* Wfin; const float* bfin;// packed groups [g][Hc][Ng]
  const float* WfinT;                  // [g][Ng][Hc]: row n = c_local*Hc + h holds W[h*C + c][:]
  int Gc, ngroups, Ng;
  const float* WinitP;                 // [Hc][Cpad]: initial.0.weight rows padded (K-major operand of W^T g)
  const float* Wreg0; const float* breg0; const float* Wreg0_raw; const float* Wreg1;
  const float* tobs; const float* fv; const float* fi; int Dv;
  const double* tout;
  const float* hidden;                 // [B,S,Hc] forward output
  const float* z0;                     // [B,Hc]   forward output
  int has_prev;
  const float* gpose;                  // [B,S,6]
  const float* gz0;                    // [B,Hc] or nullptr: gradient of the returned z0
  float* gX;                           // [B,So,C] accumulated (zeroed by the caller) or nullptr
  float* gprev;                        // [B,Hc] (has_prev)
  const float* ckpt; const CdeStepRec* log;
  size_t ck_step_stride, ck_tile_stride, ck_jstride, ck_hstride;   // checkpoint layout (floats): cde_fwd.cu tiles or cde_tc.cu feature-major
  int step_lo, step_hi, n_acc, vjp_lo; // this launch walks steps step_hi-1 .. step_lo; records are relative to vjp_lo
  float* recA[kMaxLinears + 1];        // inputs of the CDEFunc Linears, row-major [M][Hc]
  float* recG[kMaxLinears];            // pre-activation gradients of the Hc->Hc Linears [M][Hc]
  float* recGf;                        // final Linear [M][ngroups*Ng] in the packed column order
  float* recA_reg0; float* recG_reg0; float* recA_reg1; float* recG_reg1;   // head, rows b*S + i
  float* recA_init; float* recG_init;  // initial Linear, rows b ([B][Cpad], [B][Hc])
  float* tile_state;                   // per tile: gY1, gK0next ([2][Hc][R]) carried across launches
  float* scratch; size_t scratch_floats_per_cta;
  int ntiles, nst;
  int buf_floats, stage_floats, staging_floats;
};

}  // namespace odevio
