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

class TcEvolve {
 public:
  TcEvolve();
  ~TcEvolve();
  TcEvolve(const TcEvolve&) = delete;
  TcEvolve& operator=(const TcEvolve&) = delete;
  // packs the ODEFunc weights (PyTorch [out][in] layout) into the workspace; returns 0 or an ODEVIO_E_* / CUDA code
  int prepare(const odevio_odernn_cfg& c, const DevTableau& tab, bool adaptive, const float* const* ode_w,
              const float* const* ode_b, void* workspace, size_t workspace_bytes, cudaStream_t stream);
  // evolves Y[L*B][D] in place over interval `interval` (row b: ts[b * ts_ld + interval] -> [.. + 1])
  int evolve(float* Y, const float* ts, int ts_ld, int interval, int* stats, int* status, cudaStream_t stream);

 private:
  struct Impl;
  Impl* impl;
};

}  // namespace odevio
