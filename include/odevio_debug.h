/*
 * odevio_debug.h -- diagnostics and measurement hooks of libodevio_b200.so.
 *
 * NOT part of the product ABI (include/odevio.h): nothing here has a reference counterpart, the product path never
 * calls these, and they may change between builds.  Used by bench.py (live kernel durations, FFMA peak) and tools/.
 */
#ifndef ODEVIO_DEBUG_H_
#define ODEVIO_DEBUG_H_

#include "odevio.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Diagnostics (no reference counterpart): launch a dense FFMA loop on `blocks` CTAs of 512 threads
 * and report the FLOPs it performs in *flops_out (HOST); time it with events around the call to
 * obtain this GPU's fp32 FMA peak, the roofline denominator of ODEVIO_PRECISION_FP32.
 */
ODEVIO_API int32_t odevio_microbench_ffma(int32_t iters, int32_t blocks, float* sink, double* flops_out,
                                          void* stream);

/* Diagnostics: out[0] = clusters launched, out[1] = co-resident cluster maximum (cudaOccupancyMaxActiveClusters),
 * out[2] = rows taken by the cluster kernel (the rest ran in the FMA side launch) of the last tensor-core
 * solver launch (ODEVIO_PRECISION_TF32X3 / FP16X3) of this process.  out: int32[3] (HOST). */
ODEVIO_API int32_t odevio_debug_tc_geometry(int32_t* out);
/* Development (-DODEVIO_FT_TIMELINE builds): 64 clock64 stamps of cluster 0 / CTA 0 / tile 0 of the last solver iteration. */
ODEVIO_API int32_t odevio_debug_tc_timeline(long long* host_dst);
/* Measurement hook: enable = 1 / 0 switches CUDA-event timing of every tensor-core solver launch (ODEVIO_PRECISION_TF32X3 / FP16X3) on / off
 * (events on the launching stream); enable = -1 synchronises and returns the summed kernel duration (ms, HOST) and
 * the number of launches since the last read in *total_ms / *launches.  Not thread-safe; used by bench.py. */
ODEVIO_API int32_t odevio_debug_tc_timing(int32_t enable, float* total_ms, int32_t* launches);

/* Development (-DODEVIO_FT_TIMELINE builds): clock64 stamps of one 128-row evaluation of odefunc_tc_kernel (64 slots). */
ODEVIO_API int32_t odevio_debug_odefunc_timeline(long long* host_dst);
/* Development (-DODEVIO_H3_TIMELINE builds): 96 clock64 stamps / wait sums of cluster 0 / CTA 0 of the last solver iteration
 * of the ODEVIO_PRECISION_FP16X3 kernel (odernn_h3.cu). */
ODEVIO_API int32_t odevio_debug_h3_timeline(long long* host_dst);

/* Development: 32 clock64 sums of CTA 0 over all vector-field evaluations of the last tensor-core CDE launch (cde_tc.cu;
 * slot meanings in tools/cde_tc_timeline.py).  Synchronises. */
ODEVIO_API int32_t odevio_debug_cde_tc_timeline(long long* host_dst);

#ifdef __cplusplus
}
#endif
#endif /* ODEVIO_DEBUG_H_ */
