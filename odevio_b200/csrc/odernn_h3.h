// Internal interface of the second-generation tensor-core ODE solver (odernn_h3.cu) used by api.cu for
// ODEVIO_PRECISION_FP16X3: per observation interval one cluster kernel (clusters of 4, 64-row tiles, 3xFP16) evolves all
// L*B rows of the hidden state in place; the jump + head of the interval run in the FMA kernel with skip_evolve = 1.
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

class H3Evolve {
 public:
  H3Evolve();
  ~H3Evolve();
  H3Evolve(const H3Evolve&) = delete;
  H3Evolve& operator=(const H3Evolve&) = delete;
  // packs the ODEFunc weights (PyTorch [out][in] layout) into the workspace (256-byte aligned); 0 or an ODEVIO_E_* / CUDA code
  int prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const float* const* ode_w,
              const float* const* ode_b, void* workspace, size_t workspace_bytes, cudaStream_t stream);
  int max_clusters();       // clusters of 4 CTAs that can be co-resident (cudaOccupancyMaxActiveClusters)
  // evolves, in place over interval `interval`, the L * Bsub rows (l, b = seq[j]), j < Bsub, of Y[L][B][D]
  int evolve(float* Y, int Bsub, const int* seq, const float* ts, int ts_ld, int interval, int* stats, int* status,
             cudaStream_t stream);

 private:
  struct Impl;
  Impl* impl;
};

}  // namespace odevio
