// Kernel parameter block of the fused Neural-CDE forward (host <-> device, internal).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "odernn_params.h"

namespace odevio {

constexpr int kCdeMaxOut = 64;        // output times per call (S)

enum { CDE_SOLVER_DOPRI5 = 0, CDE_SOLVER_RK4 = 1 };
enum { CDE_INTERP_LINEAR = 0, CDE_INTERP_CUBIC = 1 };

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
};

}  // namespace odevio
