// Internal interface of the second-generation tensor-core ODE solver (odernn_h3.cu) used by api.cu for
// ODEVIO_PRECISION_FP16X3: clusters of 4 CTAs around 64-row tiles (SPT = 64 / L sequences x L rnn layers), 3xFP16.  With a
// tanh rnn, "cat" fusion and L <= 2 ONE launch runs the whole forward of every tile -- the solver loops of all S
// intervals, the rnn jump at every observation and the pose head (PoseODERNN.py:108-122) -- without returning to the host;
// otherwise one launch per interval evolves the state and the FMA kernel runs the jump + head (skip_evolve = 1).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "odernn_params.h"

namespace odevio {

// workspace bytes (packed weight images, activation images, per-cluster stage vectors); 0 = unsupported shape
size_t odernn_h3_workspace_bytes(const odevio_odernn_cfg& c);
void odernn_h3_last_geometry(int* clusters, int* max_clusters, int* rows);
void odernn_h3_timing_enable(bool on);
int odernn_h3_timing_read(float* total_ms, int* launches);
// development (-DODEVIO_H3_TIMELINE builds): clock64 stamps of cluster 0 / CTA 0, last solver iteration
int odernn_h3_debug_timeline(long long* host_dst);

// the rnn jump + pose head of this configuration can run inside the cluster kernel (tanh rnn, L in {1, 2})
bool odernn_h3_can_fuse_jump(const odevio_odernn_cfg& c);

class H3Evolve {
 public:
  H3Evolve();
  ~H3Evolve();
  H3Evolve(const H3Evolve&) = delete;
  H3Evolve& operator=(const H3Evolve&) = delete;
  // Plans the launch and (pack_weights) packs the PyTorch-layout weights into the workspace (256-byte aligned) as fp16
  // hi / lo operand images -- ODEFunc always, rnn + regressor.0 with `with_jump`.  pack_weights = false: the workspace
  // still holds the images of an earlier call with the same cfg and unchanged weights.  0 or an ODEVIO_E_* / CUDA code.
  int prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const odevio_odernn_weights* w, bool with_jump,
              bool pack_weights, void* workspace, size_t workspace_bytes, cudaStream_t stream);
  int max_clusters();       // clusters of 4 CTAs that can be co-resident (cudaOccupancyMaxActiveClusters)
  // Training: the following run() also writes the checkpoints odevio_odernn_backward replays, in the FMA kernels' layout
  // (odernn_params.h) for tiles of RTf sequences (RTf = 4 or 8 divides 64 / L); nloops: [ntiles_f][S_total] stored iterations.
  void set_checkpoints(float* ckpt, int* nloops, size_t ckpt_floats_per_tile, int CK, int RTf, int ntiles_f, int S_total);
  // Integrates the intervals [interval0, interval0 + n_intervals) of all L * B rows: state from h0 ([L][B][D]; nullptr =
  // zeros; may alias hT) to hT; with_jump: after every interval the rnn jump on the features fv / fi ([B][S_io][Dv],
  // [B][S_io][D - Dv]) and the pose head -> pose [B][S_io][6].
  int run(const float* h0, float* hT, const float* ts, int ts_ld, int interval0, int n_intervals, const float* fv,
          const float* fi, int Dv, int S_io, float* pose, int* stats, int* status, cudaStream_t stream);

 private:
  struct Impl;
  Impl* impl;
};

}  // namespace odevio
