/*
 * odevio.h -- C ABI of the B200-native ODE-VIO latent-dynamics integration path.
 *
 * The reference (mc1017/ODE-VIO) has no FFI layer: its seam for this path is the
 * nn.Module contract of the pose regressors.  Each entry point below replaces the
 * arithmetic behind one reference interface (file:line into the reference tree):
 *
 *   odevio_odernn_forward      PoseODERNN.forward            src/models/PoseODERNN.py:88-123
 *                              + PoseODERNN.evolve_state     src/models/PoseODERNN.py:70-75
 *                              + torchode AutoDiffAdjoint.solve / Dopri5|Tsit5|Heun|Euler.step /
 *                                IntegralController (call sites src/models/PoseODERNN.py:55-60,125-137)
 *                              + ODEFunc.forward             src/models/ODEFunc.py:38-39
 *                              + nn.RNN / nn.GRU one-step    src/models/PoseODERNN.py:114,139-148
 *                              + regressor head              src/models/PoseODERNN.py:64-68,122
 *   odevio_odernn_backward     loss.backward() through the above (to.AutoDiffAdjoint is plain
 *                              autograd = discretise-then-optimise)  scripts/train_model.py:78
 *   odevio_cde_backward        loss.backward() through PoseCDE.forward (cdeint adjoint=False = plain autograd through
 *                              torchdiffeq's loop)           src/models/PoseCDE.py:98-101, scripts/train_model.py:78
 *   odevio_cde_forward         PoseCDE.forward               src/models/PoseCDE.py:76-103
 *                              + torchcde linear_interpolation_coeffs / LinearInterpolation / cdeint
 *                                (call sites src/models/PoseCDE.py:94-101)
 *                              + CDEFunc.forward             src/models/ODEFunc.py:81-84
 *   odevio_*_workspace_bytes   (torch allocator; the reference allocates implicitly)
 *
 * Conventions: plain pointers and sizes only; every pointer is DEVICE memory owned by the
 * caller (PyTorch in the shipped host layer) unless marked HOST; all tensors are contiguous
 * row-major fp32 with the PyTorch shapes quoted; `stream` is a cudaStream_t passed as void*.
 * Calls are asynchronous on `stream`, allocate nothing, never synchronise the device, and are
 * re-entrant across devices (the current device of the calling thread is used).
 * Thread safety: calls on DIFFERENT devices may run concurrently; calls on the same device must be serialised by the
 * caller (odevio_odernn_forward with ODEVIO_PRECISION_TF32X3 shares one side stream + fork / join events per device, and
 * the measurement hooks of odevio_debug.h keep process-wide state).
 * Return value: 0 ok; <0 invalid argument (ODEVIO_E_*); >0 a cudaError_t from launch.
 * Per-row solver failures (non-finite error norm, max_steps hit) are reported through the
 * `status` output, never by hanging.
 */
#ifndef ODEVIO_H_
#define ODEVIO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODEVIO_ABI_VERSION 4

#if defined(__GNUC__)
#define ODEVIO_API __attribute__((visibility("default")))
#else
#define ODEVIO_API
#endif

#define ODEVIO_MAX_ODE_LINEARS 6   /* ode_fn_num_layers + 1 <= 6 */
#define ODEVIO_MAX_RNN_LAYERS 4
#define ODEVIO_MAX_STAGES 7

/* error codes (negative returns) */
#define ODEVIO_E_NULL        -1   /* required pointer is NULL */
#define ODEVIO_E_SHAPE       -2   /* unsupported / inconsistent dimension */
#define ODEVIO_E_ENUM        -3   /* unknown activation / rnn / solver id */
#define ODEVIO_E_WORKSPACE   -4   /* workspace too small or misaligned */
#define ODEVIO_E_DEVICE      -5   /* not an sm_100 device / no device */

/* activations: reference src/models/ODEFunc.py:23-36 */
enum { ODEVIO_ACT_TANH = 0, ODEVIO_ACT_RELU = 1, ODEVIO_ACT_LEAKY_RELU = 2, ODEVIO_ACT_SOFTPLUS = 3 };
/* recurrent jump: reference src/models/PoseODERNN.py:139-148 */
enum { ODEVIO_RNN_TANH = 0, ODEVIO_RNN_GRU = 1 };
/* solver menu: reference src/models/PoseODERNN.py:125-137 (+ north_star fixed-step rk4, and
 * torchdiffeq's 3/8-rule "rk4" reachable through src/models/PoseCDE.py:72,101) */
enum {
  ODEVIO_SOLVER_DOPRI5 = 0, ODEVIO_SOLVER_TSIT5 = 1, ODEVIO_SOLVER_HEUN = 2,
  ODEVIO_SOLVER_EULER = 3, ODEVIO_SOLVER_RK4 = 4, ODEVIO_SOLVER_RK4_38 = 5
};
/* per-row status codes written to `status` */
enum { ODEVIO_STATUS_OK = 0, ODEVIO_STATUS_MAX_STEPS = 1, ODEVIO_STATUS_INFINITE_NORM = 2,
       ODEVIO_STATUS_CKPT_OVERFLOW = 3 /* training: more solver iterations than cfg.ckpt_loops */ };
/* arithmetic mode of the vector-field GEMMs */
enum { ODEVIO_PRECISION_FP32 = 0,    /* CUDA-core FFMA, one persistent kernel per forward */
       ODEVIO_PRECISION_TF32X3 = 1,  /* ODEFunc GEMMs on tcgen05 as 3xTF32 (hi/lo split, fp32-accurate: <= 1e-5 on poses);
                                        per interval one cluster kernel runs the whole solver loop of every 128-row
                                        tile (odernn_tc.cu), then the FMA kernel runs the jump + head.  Inference only
                                        (save_checkpoints = 0), endpoint_dense = 0, trace_steps = 0; D, H multiples of 64 */
       ODEVIO_PRECISION_FP16X3 = 2   /* second-generation tensor-core solver (odernn_h3.cu): ODEFunc GEMMs on tcgen05 as
                                        3xFP16 (x = hi + lo * 2^-11, products hi*hi + lo*hi + hi*lo, fp32 accumulate: the
                                        same 2^-22 product accuracy as 3xTF32 at twice the MMA rate and half the operand
                                        bytes), weights on the M side, clusters of 4 CTAs around 64-row tiles.  Same
                                        restrictions as TF32X3; D / 4 and H / 4 multiples of 32 in [128, 256]
                                        (D = 768, H in {512, 768, 1024}); |state|, |activation| < 65504 (fp16 range) */ };

typedef struct odevio_odernn_cfg {
  int32_t B;            /* sequences */
  int32_t S;            /* observation intervals (= seq_len - 1) */
  int32_t D;            /* state dim = v_f_len + i_f_len */
  int32_t H;            /* ODEFunc hidden dim */
  int32_t n_hidden;     /* ODEFunc num_hidden_layers (reference default 3) -> n_hidden+1 Linears */
  int32_t L;            /* rnn_num_layers */
  int32_t activation;   /* ODEVIO_ACT_* */
  int32_t rnn_type;     /* ODEVIO_RNN_* */
  int32_t solver;       /* ODEVIO_SOLVER_* */
  int32_t substeps;     /* fixed-step solvers: steps per interval */
  float atol;           /* reference 1e-6  (PoseODERNN.py:57) */
  float rtol;           /* reference 1e-2  (PoseODERNN.py:57) */
  float dt0;            /* reference 1e-4  (PoseODERNN.py:72) */
  float safety;         /* 0.9  */
  float factor_min;     /* 0.2  */
  float factor_max;     /* 10.0 */
  int32_t accept_strict;      /* 1: accept iff ratio < 1 (torchode); 0: <= 1 */
  int32_t floor_factor;       /* 1: factor >= 1 after an accepted step (torchdiffeq rule) */
  int32_t endpoint_dense;     /* 1: end point = the step's fp32 quartic dense output at x=1 (literal
                                 torchode); 0: y1, its exact-arithmetic value (default) */
  int32_t max_steps;          /* per-interval guard; rows still running get STATUS_MAX_STEPS */
  int32_t precision;          /* ODEVIO_PRECISION_* */
  int32_t save_checkpoints;   /* 1: record what odevio_odernn_backward needs in `ckpt` (training;
                                 rows_per_tile 4 or 8 with rows_per_tile * L a multiple of 8,
                                 endpoint_dense = 0, D and H multiples of 128) */
  int32_t rows_per_tile;      /* 0 = auto; else 4, 8 or 16 sequences per CTA */
  int32_t exact_landing;      /* 1 (default): a step clamped to the remaining interval ends exactly at
                                 t_end; 0: literal fp32 t + (t_end - t), may need a 1-ulp extra step */
  int32_t trace_steps;        /* T >= 0: additionally record (dt, error ratio) of the first T steps of
                                 every solve; the stats row then has 2 + 2*T int32 (floats as bits) */
  int32_t ckpt_loops;         /* training: stored solver iterations per interval and tile (0 = 16, or
                                 `substeps` for the fixed-step solvers); overflow -> STATUS_CKPT_OVERFLOW */
  int32_t evolve_only;        /* 1: PoseODERNN.evolve_state (src/models/PoseODERNN.py:70-75): only the ODE solves of the
                                 S intervals, no jump, no pose head; hT returns the evolved states ([L,B,D]), pose is
                                 not written (may be NULL) and fv / fi are not read */
  int32_t weights_prepacked;  /* ODEVIO_PRECISION_FP16X3: 1 = the workspace still holds the packed weight images written by an
                                 earlier odevio_odernn_forward call with the same cfg (apart from this flag), the same
                                 workspace address and unchanged weights: the per-forward packing launches are skipped
                                 ("prepare the weights once").  0: pack on every call (stateless) */
  int32_t reserved[2];
} odevio_odernn_cfg;

/* PyTorch-layout parameters ([out, in] row-major), exactly the reference's state_dict tensors */
typedef struct odevio_odernn_weights {
  const float* ode_w[ODEVIO_MAX_ODE_LINEARS];   /* ode_func.net.{0,2,4,...}.weight */
  const float* ode_b[ODEVIO_MAX_ODE_LINEARS];   /* ode_func.net.{0,2,4,...}.bias   */
  const float* rnn_w_ih[ODEVIO_MAX_RNN_LAYERS]; /* rnn.weight_ih_l{k}: [G*D, D], G = 1 (rnn) | 3 (gru) */
  const float* rnn_w_hh[ODEVIO_MAX_RNN_LAYERS]; /* rnn.weight_hh_l{k} */
  const float* rnn_b_ih[ODEVIO_MAX_RNN_LAYERS]; /* rnn.bias_ih_l{k}:  [G*D] */
  const float* rnn_b_hh[ODEVIO_MAX_RNN_LAYERS]; /* rnn.bias_hh_l{k} */
  const float* reg_w0;  /* regressor.0.weight [128, D] */
  const float* reg_b0;  /* regressor.0.bias   [128]    */
  const float* reg_w1;  /* regressor.2.weight [6, 128] */
  const float* reg_b1;  /* regressor.2.bias   [6]      */
  /* FusionModule "soft" (src/models/FusionModule.py:20-23): fused = cat * (W cat + b), evaluated in the
   * forward kernel when both are non-NULL (fv / fi are then the RAW features); NULL = "cat".
   * odevio_odernn_backward ignores them: in training the host applies the fusion (autograd). */
  const float* fuse_w;  /* fuse.net.0.weight [D, D] or NULL */
  const float* fuse_b;  /* fuse.net.0.bias   [D]    or NULL */
} odevio_odernn_weights;

/* Gradient outputs of odevio_odernn_backward: same shapes as odevio_odernn_weights (overwritten). */
typedef struct odevio_odernn_grads {
  float* ode_w[ODEVIO_MAX_ODE_LINEARS];
  float* ode_b[ODEVIO_MAX_ODE_LINEARS];
  float* rnn_w_ih[ODEVIO_MAX_RNN_LAYERS];
  float* rnn_w_hh[ODEVIO_MAX_RNN_LAYERS];
  float* rnn_b_ih[ODEVIO_MAX_RNN_LAYERS];
  float* rnn_b_hh[ODEVIO_MAX_RNN_LAYERS];
  float* reg_w0;
  float* reg_b0;
  float* reg_w1;
  float* reg_b1;
} odevio_odernn_grads;

/* ABI version of the loaded library (== ODEVIO_ABI_VERSION of the header it was built from). */
ODEVIO_API int32_t odevio_version(void);

/* Human-readable text for a negative return code (static storage). */
ODEVIO_API const char* odevio_error_string(int32_t code);

/* Fill *cfg with the reference's defaults (scripts/config.py:50-69, PoseODERNN.py:57,72). */
ODEVIO_API void odevio_odernn_default_cfg(odevio_odernn_cfg* cfg);

/* Bytes of device workspace odevio_odernn_forward needs for this cfg (0 on invalid cfg). */
ODEVIO_API size_t odevio_odernn_workspace_bytes(const odevio_odernn_cfg* cfg);

/*
 * Fused ODE-RNN regressor forward.
 *   fv   [B,S,Dv], fi [B,S,D-Dv]   visual / inertial features (fusion "cat" happens in-kernel);
 *                                  pass fi = NULL and Dv = D when fv already holds fused features
 *   ts   [B,S+1]                   timestamps exactly as the solver must see them (the host layer
 *                                  has already applied `ts - ts[:, :1]` when prev is None)
 *   h0   [L,B,D] or NULL (zeros)
 *   pose [B,S,6]  hT [L,B,D]       outputs
 *   stats  [S,L,B,2+2T] int32 or NULL: (n_steps, n_accepted, then T x (dt, ratio) as float bits,
 *                                  T = cfg->trace_steps) per interval / layer / row; zero-fill it
 *   status [B] int32 or NULL:       worst ODEVIO_STATUS_* seen by the row
 *   ckpt: NULL, or (cfg.save_checkpoints = 1) >= odevio_odernn_ckpt_bytes(cfg), 256-byte aligned
 *   workspace: >= odevio_odernn_workspace_bytes(cfg), 256-byte aligned
 */
ODEVIO_API int32_t odevio_odernn_forward(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                              const float* fv, const float* fi, int32_t Dv,
                              const float* ts, const float* h0,
                              float* pose, float* hT, int32_t* stats, int32_t* status,
                              void* ckpt, size_t ckpt_bytes,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Launch geometry chosen for cfg: out[0] = sequences per tile (RT), out[1] = rows per tile
 * (RT * L), out[2] = number of tiles, out[3] = vector-field evaluations recorded per stored
 * solver iteration (stages entering y1), out[4] = stored iterations per interval (ckpt_loops
 * resolved), out[5] = CTAs launched; out[6..7] reserved.  out must hold 8 int32 (HOST).
 */
ODEVIO_API int32_t odevio_odernn_geometry(const odevio_odernn_cfg* cfg, int32_t* out);

/*
 * Training (replaces loss.backward() through PoseODERNN.forward, reference
 * scripts/train_model.py:78 -- torchode's AutoDiffAdjoint is plain autograd through the solver loop).
 *
 * 1. forward with cfg.save_checkpoints = 1 and a `ckpt` buffer of odevio_odernn_ckpt_bytes(cfg)
 *    bytes (256-byte aligned).  Its head is int32 nloops[ntiles * S]: the number of stored solver
 *    iterations of every (tile, interval).
 * 2. the caller turns nloops into rec_base[ntiles * S] (int64, DEVICE) = exclusive prefix sum of
 *    nloops * out[3] * out[1] and ode_rows = the total (this is the one host read of the step).
 * 3. odevio_odernn_backward with a workspace of odevio_odernn_backward_workspace_bytes(cfg, ode_rows).
 *
 * Discretise-then-optimise with the accepted step sizes treated as constants.
 *   grad_pose [B,S,6], grad_hT [L,B,D] or NULL (zero)      incoming gradients
 *   g                                                     parameter gradients (overwritten)
 *   grad_fused [B,S,D] or NULL                            gradient of the fused features cat(fv, fi)
 *   grad_h0 [L,B,D] or NULL                               gradient of the initial hidden state
 */
ODEVIO_API size_t odevio_odernn_ckpt_bytes(const odevio_odernn_cfg* cfg);
ODEVIO_API size_t odevio_odernn_backward_workspace_bytes(const odevio_odernn_cfg* cfg, int64_t ode_rows);
ODEVIO_API int32_t odevio_odernn_backward(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                               const float* fv, const float* fi, int32_t Dv,
                               const void* ckpt, size_t ckpt_bytes,
                               const int64_t* rec_base, int64_t ode_rows,
                               const float* grad_pose, const float* grad_hT,
                               const odevio_odernn_grads* g, float* grad_fused, float* grad_h0,
                               void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same backward, one RANGE of observation intervals [i_lo, i_hi] per call: bounds the record streams of the deferred
 * weight-gradient GEMMs (they are ~20 MB per sequence at configs[3]: 84 GB for B = 4096 in one piece).  Call with the same
 * workspace from the last range (i_hi = S - 1) down to the first (i_lo = 0); `rec_base` holds, for the (tile, interval)
 * pairs of the range, their first record row RELATIVE to the range; `ode_rows` = record rows of this range,
 * `ode_rows_plan` = the capacity the workspace was sized with (odevio_odernn_backward_workspace_bytes(cfg, ode_rows_plan),
 * >= every range's ode_rows).  The hidden-state gradient is carried between the calls inside the workspace; the ODEFunc
 * weight gradients accumulate over the calls (the first call overwrites), the rnn / regressor gradients, grad_fused and
 * grad_h0 are complete after the i_lo = 0 call.  odevio_odernn_backward == one range [0, S - 1].
 */
ODEVIO_API int32_t odevio_odernn_backward_range(const odevio_odernn_cfg* cfg, const odevio_odernn_weights* w,
                                                const float* fv, const float* fi, int32_t Dv,
                                                const void* ckpt, size_t ckpt_bytes,
                                                const int64_t* rec_base, int64_t ode_rows, int64_t ode_rows_plan,
                                                int32_t i_lo, int32_t i_hi,
                                                const float* grad_pose, const float* grad_hT,
                                                const odevio_odernn_grads* g, float* grad_fused, float* grad_h0,
                                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ Neural CDE (PoseCDE) ---- */

/* cde solver menu: torchdiffeq names reachable through src/models/PoseCDE.py:72,101 */
enum { ODEVIO_CDE_SOLVER_DOPRI5 = 0, ODEVIO_CDE_SOLVER_RK4 = 1 /* torchdiffeq fixed-grid 3/8 rule */ };
/* control path: the reference's rectilinear linear interpolation (PoseCDE.py:94-95) or the
 * north_star cubic path (Hermite cubics with backward differences on the integer knot grid) */
enum { ODEVIO_CDE_INTERP_LINEAR = 0, ODEVIO_CDE_INTERP_CUBIC = 1 };

typedef struct odevio_cde_cfg {
  int32_t B;            /* sequences */
  int32_t S;            /* output times (= seq_len - 1), <= 64 */
  int32_t So;           /* observations in the control path (>= 2; = S in training, grows with the
                           eval-mode history of PoseCDE.py:88-92) */
  int32_t Hc;           /* cde_hidden_dim; the fused feature width must equal it (PoseCDE.py:47,62) */
  int32_t n_layers;     /* cde_fn_num_layers: Hc->Hc Linears before the final Hc -> Hc*(Hc+1) */
  int32_t activation;   /* ODEVIO_ACT_* */
  int32_t solver;       /* ODEVIO_CDE_SOLVER_* */
  int32_t interp;       /* ODEVIO_CDE_INTERP_* */
  float atol;           /* reference 1e-6 (PoseCDE.py:101) */
  float rtol;           /* reference 1e-4 */
  double step_size;     /* rk4: > 0 = fixed grid spacing with linear output interpolation, 0 = output times */
  int32_t max_steps;    /* dopri5 guard (status 1 when hit) */
  int32_t rows_per_tile;/* 0 = auto; 8 or 16 (CUDA-core kernel) */
  int32_t precision;    /* ODEVIO_PRECISION_FP32 (default): CUDA-core FFMA kernel.  ODEVIO_PRECISION_FP16X3: the final
                           Hc -> Hc*(Hc+1) Linear of CDEFunc on tcgen05 as 3xFP16 (fp32-level accuracy), weights resident in
                           shared memory; Hc in {32, 64, 128}, B <= 32 * Hc, inference only (odevio_cde_forward_ckpt takes
                           the CUDA-core kernel); other shapes: odevio_cde_workspace_bytes returns 0 / ODEVIO_E_SHAPE */
  int32_t reserved[5];
} odevio_cde_cfg;

typedef struct odevio_cde_weights {
  const float* cde_w[ODEVIO_MAX_ODE_LINEARS];  /* cde_func.net.{0,2,..}.weight; the last is [Hc*(Hc+1), Hc] */
  const float* cde_b[ODEVIO_MAX_ODE_LINEARS];
  const float* init_w;  /* initial.0.weight [Hc, Hc+1] */
  const float* init_b;
  const float* reg_w0;  /* regressor.0.weight [128, Hc] */
  const float* reg_b0;
  const float* reg_w1;  /* regressor.2.weight [6, 128] */
  const float* reg_b1;
} odevio_cde_weights;

ODEVIO_API void odevio_cde_default_cfg(odevio_cde_cfg* cfg);
ODEVIO_API size_t odevio_cde_workspace_bytes(const odevio_cde_cfg* cfg);

/*
 * Fused PoseCDE forward (replaces src/models/PoseCDE.py:76-103 incl. torchcde / torchdiffeq).
 *   tobs [B,So]                     channel 0 of the observations (the timestamps the reference
 *                                   concatenates in front of the fused features, PoseCDE.py:83-85)
 *   fv [B,So,Dv], fi [B,So,Hc-Dv]   fused features (fi = NULL and Dv = Hc when already fused)
 *   tout [S] float64 (DEVICE)       output times in the integration variable: the reference passes
 *                                   batch row 0's times in seconds (PoseCDE.py:101); the cubic mode
 *                                   integrates over the knot grid
 *   z0_in [B,Hc] or NULL            NULL: z0 = initial(X(knot 0)) (PoseCDE.py:96)
 *   pose [B,S,6], z0_out [B,Hc]     outputs (the reference returns z0, PoseCDE.py:103)
 *   hidden [B,S,Hc] or NULL         optional: the integrated hidden states
 *   stats int32[4] or NULL          n_steps, n_accepted, n_f_evals, status (ODEVIO_STATUS_*)
 * Launches cooperatively (one grid-wide reduction per solver step: torchdiffeq's batch-joint
 * step control); the whole grid must be resident, which the workspace planner guarantees.
 */
ODEVIO_API int32_t odevio_cde_forward(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                                      const float* tobs, const float* fv, const float* fi, int32_t Dv,
                                      const double* tout, const float* z0_in,
                                      float* pose, float* z0_out, float* hidden, int32_t* stats,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* Gradients of the PoseCDE parameters, PyTorch shapes (same members as odevio_cde_weights). */
typedef struct odevio_cde_grads {
  float* cde_w[ODEVIO_MAX_ODE_LINEARS];
  float* cde_b[ODEVIO_MAX_ODE_LINEARS];
  float* init_w;
  float* init_b;
  float* reg_w0;
  float* reg_b0;
  float* reg_w1;
  float* reg_b1;
} odevio_cde_grads;

/*
 * Training (reference src/models/PoseCDE.py:98-101 `cdeint(..., adjoint=False)` + scripts/train_model.py:78
 * `loss.backward()`): plain autograd through torchdiffeq's solver loop = discretise-then-optimise with the accepted step
 * sizes as constants.  odevio_cde_forward_ckpt is odevio_cde_forward that also leaves, for each of up to `ckpt_steps`
 * accepted solver steps, the stage values and a log entry in `ckpt` (odevio_cde_ckpt_bytes(cfg, ckpt_steps) bytes;
 * stats[3] = ODEVIO_STATUS_CKPT_OVERFLOW when the solve accepted more steps than that).  The log starts with
 * int32 {n_accepted, n_pullbacks, status}; entry s (48 bytes, at byte 48 * (1 + s)) holds at int32 index 10 the number of
 * vector-field pullbacks of all earlier steps -- the caller copies those n_accepted + 1 prefix counts to the HOST
 * (`vjp_base`, last = n_pullbacks) and hands them to odevio_cde_backward, which walks the log backwards in chunks of at most
 * `chunk_vjps` pullbacks per launch (that bounds the record streams of the deferred weight-gradient GEMMs:
 * odevio_cde_backward_workspace_bytes(cfg, chunk_vjps)).
 *   hidden [B,S,Hc], z0 [B,Hc]      outputs of the forward
 *   has_prev                        the forward ran from z0_in (grad_prev [B,Hc] receives its gradient) instead of initial()
 *   grad_pose [B,S,6]               incoming gradient; grad_z0 [B,Hc] or NULL: gradient of the returned z0
 *   grad_x [B,So,Hc+1] or NULL      ZERO-INITIALISED by the caller; receives d loss / d observations (channel 0, the
 *                                   timestamps, is left untouched)
 */
ODEVIO_API size_t odevio_cde_ckpt_bytes(const odevio_cde_cfg* cfg, int32_t ckpt_steps);
ODEVIO_API int32_t odevio_cde_forward_ckpt(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                                           const float* tobs, const float* fv, const float* fi, int32_t Dv,
                                           const double* tout, const float* z0_in,
                                           float* pose, float* z0_out, float* hidden, int32_t* stats,
                                           void* ckpt, size_t ckpt_bytes, int32_t ckpt_steps,
                                           void* workspace, size_t workspace_bytes, void* stream);
ODEVIO_API size_t odevio_cde_backward_workspace_bytes(const odevio_cde_cfg* cfg, int32_t chunk_vjps);
ODEVIO_API int32_t odevio_cde_backward(const odevio_cde_cfg* cfg, const odevio_cde_weights* w,
                                       const float* tobs, const float* fv, const float* fi, int32_t Dv,
                                       const double* tout, int32_t has_prev, const float* hidden, const float* z0,
                                       const void* ckpt, size_t ckpt_bytes, int32_t ckpt_steps,
                                       const int32_t* vjp_base /* HOST [n_accepted + 1] */, int32_t n_accepted,
                                       int32_t chunk_vjps,
                                       const float* grad_pose, const float* grad_z0,
                                       const odevio_cde_grads* g, float* grad_x, float* grad_prev,
                                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------- ODEFunc.forward (tensor cores) ---- */

/*
 * Batched evaluation of the vector field, out = tanh(W_n a(... a(W_0 x + b_0) ...) + b_n), replacing
 * ODEFunc.forward (reference src/models/ODEFunc.py:38-39) for callers that evaluate the field
 * directly.  Runs on tcgen05 (kind::tf32 with the 3xTF32 hi/lo split: fp32-level accuracy, fp32
 * accumulation in TMEM); a 128-row tile is split by output columns over a cluster of 8 CTAs.
 *   weights[j] [out_j, in_j], biases[j] [out_j]  (HOST arrays of n_hidden + 1 DEVICE pointers,
 *                                                 ode_func.net.{0,2,..})
 *   x [M, D], out [M, D]
 * Supported: D / 8 and H / 8 multiples of 32 in [32, 128] (D, H in {256, 512, 768, 1024}).
 */
ODEVIO_API size_t odevio_odefunc_workspace_bytes(int32_t M, int32_t D, int32_t H, int32_t n_hidden);
ODEVIO_API int32_t odevio_odefunc_forward(int32_t M, int32_t D, int32_t H, int32_t n_hidden, int32_t activation,
                                          const float* const* weights, const float* const* biases,
                                          const float* x, float* out,
                                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * Generic MLP evaluation on CUDA cores (fp32 FMA) for the module interfaces that are called
 * directly and are not on a fused path: CDEFunc.forward (reference src/models/ODEFunc.py:81-84,
 * the caller views the [M, Hc*C] result as [M, Hc, C]) and ODEFunc.forward for shapes the
 * tensor-core kernel does not cover.  dims[n_linears + 1] and acts[n_linears] (ODEVIO_ACT_* or
 * 4 = identity) are HOST arrays; weights[j] [dims[j+1], dims[j]] / biases[j] are HOST arrays of
 * DEVICE pointers; x [M, dims[0]], out [M, dims[n_linears]].
 */
ODEVIO_API int32_t odevio_mlp_forward(int32_t M, int32_t n_linears, const int32_t* dims, const int32_t* acts,
                                      const float* const* weights, const float* const* biases,
                                      const float* x, float* out, void* stream);

/*
 * InertialEncoder.forward (reference src/models/Encoder.py:39-74), inference mode: raw IMU rows imu [B, 10*S + 1, 6]
 * -> windows of 11 samples (stride 10) -> 3 x {Conv1d(k=3, pad=1) + BatchNorm1d(running statistics) + LeakyReLU(0.1)}
 * (6 -> 64 -> 128 -> 256) -> Linear(256 * 11, i_f_len) -> out [B, S, i_f_len], the `fi` input of the regressors.
 * Weights are the reference's state_dict tensors (encoder_conv.{0,4,8}.{weight [C_out, C_in, 3], bias},
 * encoder_conv.{1,5,9}.{weight, bias, running_mean, running_var}, proj.{weight [i_f_len, 2816], bias}), DEVICE pointers.
 */
typedef struct odevio_imu_encoder_weights {
  const float* conv_w[3]; const float* conv_b[3];
  const float* bn_weight[3]; const float* bn_bias[3]; const float* bn_mean[3]; const float* bn_var[3];
  float bn_eps;                 /* nn.BatchNorm1d default 1e-5 */
  const float* proj_w; const float* proj_b;
} odevio_imu_encoder_weights;
ODEVIO_API size_t odevio_imu_encoder_workspace_bytes(int32_t i_f_len);
ODEVIO_API int32_t odevio_imu_encoder_forward(int32_t B, int32_t S, int32_t i_f_len, const odevio_imu_encoder_weights* w,
                                              const float* imu, float* out,
                                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Training-loop glue downstream of odevio_odernn_backward.
 * odevio_pose_loss: reference scripts/train_model.py:72-76 -- loss3[0] = w_angle * MSE(pose[:, :3], gts[:, :3]) +
 *   MSE(pose[:, 3:], gts[:, 3:]) (w_angle = 100), loss3[1] = the angle MSE, loss3[2] = the translation MSE (DEVICE
 *   float[3]); grad_pose (may be NULL) = grad_scale * d loss3[0] / d pose.  pose, gts, grad_pose: [n_rows, 6].
 * odevio_adam_step: torch.nn.utils.clip_grad_norm_(max_norm) (scripts/train_model.py:83-85; max_norm <= 0: none)
 *   followed by one torch.optim.Adam step (src/utils/utils.py:150-157: L2 weight decay, no amsgrad) on a flat fp32
 *   bucket of n values -- the bucket of the NCCL gradient all-reduce.  step = 1, 2, ... ; norm_coef (DEVICE float[2],
 *   required with max_norm > 0) receives the total gradient norm and the clip coefficient.  All four arrays 16-byte
 *   aligned.  No host synchronisation.
 * odevio_adam_step_groups: the same for the reference's TWO parameter groups (src/utils/utils.py:116-119: [other, regressor],
 *   re-scheduled separately by scripts/train_model.py:215-216): elements [0, split) step with lr_first, elements
 *   [split, n) with lr_rest; split a multiple of 4.  The clip norm is the whole bucket's.
 * workspace: odevio_train_glue_workspace_bytes() bytes, 16-byte aligned.
 */
ODEVIO_API size_t odevio_train_glue_workspace_bytes(void);
ODEVIO_API int32_t odevio_pose_loss(int64_t n_rows, const float* pose, const float* gts, float w_angle, float grad_scale,
                                    float* loss3, float* grad_pose, void* workspace, size_t workspace_bytes, void* stream);
ODEVIO_API int32_t odevio_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                    int32_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    float max_norm, float* norm_coef, void* workspace, size_t workspace_bytes, void* stream);
ODEVIO_API int32_t odevio_adam_step_groups(int64_t n, int64_t split, float* params, const float* grads, float* exp_avg,
                                           float* exp_avg_sq, int32_t step, float lr_first, float lr_rest, float beta1, float beta2,
                                           float eps, float weight_decay, float max_norm, float* norm_coef, void* workspace,
                                           size_t workspace_bytes, void* stream);

/*
 * Data-parallel training (the one collective of the path: the gradient all-reduce in front of clip + Adam,
 * scripts/train_model.py:78-86 under torch.distributed): gradient all-reduce + global-norm clip + Adam in ONE kernel over
 * NVLink peer memory.  Every rank's parameter bucket, gradient bucket and a 256-byte zero-initialised flag pad live in
 * symmetric memory mapped into every process; `params_peers` / `grads_peers` / `pad_peers` are HOST arrays of `world` DEVICE
 * pointers (index = rank).  Rank r reduces the r-th slice of the bucket by P2P loads (fixed order: bit-reproducible),
 * exchanges its sum of squares through the pads, applies clip + Adam to its slice (exp_avg / exp_avg_sq: full-size arrays of
 * which only this rank's slice is ever touched -- the optimiser state is sharded) and stores the new parameters into every
 * rank's bucket.  Every rank must call it with the same arguments and `epoch` = 1, 2, ... (the flags are monotonic);
 * grad_scale = 1 / world for the mean.  n a multiple of 4; norm_coef as in odevio_adam_step.  Launches cooperatively.
 */
ODEVIO_API size_t odevio_allreduce_adam_peer_workspace_bytes(int64_t n, int32_t world);
ODEVIO_API int32_t odevio_allreduce_adam_peer(int64_t n, int64_t split, int32_t rank, int32_t world,
                                              float* const* params_peers, const float* const* grads_peers,
                                              uint32_t* const* pad_peers, float* exp_avg, float* exp_avg_sq, int32_t step,
                                              uint32_t epoch, float grad_scale, float lr_first, float lr_rest, float beta1,
                                              float beta2, float eps, float weight_decay, float max_norm, float* norm_coef,
                                              void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostics / measurement hooks (odevio_debug_*, odevio_microbench_ffma) are NOT part of the product ABI: they are
 * declared in include/odevio_debug.h. */

#ifdef __cplusplus
}
#endif
#endif /* ODEVIO_H_ */
