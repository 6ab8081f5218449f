// Internal interface of the tensor-core ODE solver (odernn_tc.cu) used by api.cu for
// ODEVIO_PRECISION_TF32X3: per observation interval one cluster kernel evolves all L*B rows of the hidden
// state in place; the jump + head of the interval run in the FMA kernel with skip_evolve = 1.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "odernn_params.h"

namespace odevio {

// workspace bytes of the solver (packed weight images, operand buffers, per-cluster stage vectors); 0 = unsupported shape
size_t odernn_tc_workspace_bytes(const odevio_odernn_cfg& c);

// development: clusters launched / co-resident maximum (cudaOccupancyMaxActiveClusters) / rows of the last evolve launch
void odernn_tc_last_geometry(int* clusters, int* max_clusters, int* rows);

// measurement hook: CUDA events around every solver launch (bench.py's live kernel duration); read() synchronises,
// returns the summed duration and the number of launches since enable / the last read
void odernn_tc_timing_enable(bool on);
int odernn_tc_timing_read(float* total_ms, int* launches);

// development (-DODEVIO_FT_TIMELINE builds): clock64 stamps of cluster 0 / CTA 0 / tile 0, last solver iteration
int odernn_tc_debug_timeline(long long* host_dst);

// seq[S][B] (device): per interval the sequences in increasing index order, except that the n_side sequences with the
// shortest interval are moved to the tail (the FFMA side launch takes the tail).  B <= 8192.
int odernn_tc_select(const float* ts, int B, int S, int n_side, int* seq, cudaStream_t stream);

class TcEvolve {
 public:
  TcEvolve();
  ~TcEvolve();
  TcEvolve(const TcEvolve&) = delete;
  TcEvolve& operator=(const TcEvolve&) = delete;
  // packs the ODEFunc weights (PyTorch [out][in] layout) into the workspace; returns 0 or an ODEVIO_E_* / CUDA code
  int prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const float* const* ode_w,
              const float* const* ode_b, void* workspace, size_t workspace_bytes, cudaStream_t stream);
  // clusters of 8 CTAs that can be co-resident on this GPU (cudaOccupancyMaxActiveClusters; 15-16 on B200)
  int max_clusters();
  int cluster_size() const;          // 8 (up to 16 tiles) or 4 (beyond)
  // evolves, in place over interval `interval`, the L * Bsub rows (l, b = seq[j]), j < Bsub, of Y[L][B][D]
  // (seq == nullptr: b = j); row b integrates ts[b * ts_ld + interval] -> ts[.. + 1].  Other rows are left to the caller.
  int evolve(float* Y, int Bsub, const int* seq, const float* ts, int ts_ld, int interval, int* stats, int* status,
             cudaStream_t stream);

 private:
  struct Impl;
  Impl* impl;
};

}  // namespace odevio
