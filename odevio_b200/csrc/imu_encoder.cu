// InertialEncoder.forward (reference src/models/Encoder.py:39-74), inference mode, in one kernel:
//   imu [B, 10*S + 1, 6]  ->  S windows of 11 samples (stride 10)  ->  3 x {Conv1d(k=3, pad=1) + BatchNorm1d (running
//   statistics) + LeakyReLU(0.1)} (6 -> 64 -> 128 -> 256 channels; Dropout is the identity in eval mode)  ->  flatten
//   channel-major [256 * 11]  ->  Linear(2816, i_f_len)  ->  fi [B, S, i_f_len]
// This is the step immediately upstream of the regressor: with it the raw "IMU at 10 samples per
// frame" enters the path and `fi` never exists as a separate PyTorch op.  4.2 MFLOP per window against 26-140 MFLOP per
// sequence-step in the regressor: not performance-critical, so plain CUDA-core FFMA -- one CTA per 8 windows keeps
// every intermediate activation in shared memory, weights are read K-major (pre-transposed) so that thread = output
// channel loads are coalesced, activations are broadcast 128-bit shared-memory reads.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/odevio.h"
#include "common.cuh"

namespace odevio {

cudaError_t transpose_pack(const float* src, int N, int K, float* dst, int ldN, int k_off, int n_off, cudaStream_t stream);

namespace {

constexpr int IE_W = 8;          // windows per CTA
constexpr int IE_T = 11;         // samples per window (Encoder.py:57-60)
constexpr int IE_TP = 16;        // padded row: [0] = left zero pad, [1..11] = samples, [12] = right zero pad, float4 aligned
constexpr int IE_THREADS = 256;
constexpr int IE_C0 = 6, IE_C1 = 64, IE_C2 = 128, IE_C3 = 256;

struct IeParams {
  int B, S, T, F;                // T = 10 * S + 1 imu rows per sequence, F = i_f_len
  const float* imu;              // [B][T][6]
  float* out;                    // [B][S][F]
  const float* Wt[3];            // packed conv weights [C_in * 3][C_out]
  const float* cb[3];            // conv bias [C_out]
  const float* bn_w[3]; const float* bn_b[3]; const float* bn_m[3]; const float* bn_v[3];
  float eps;
  const float* Wp;               // packed projection [256 * 11][F]
  const float* bp;               // [F]
};

// One conv layer for NW windows of this thread's output channel `co`:
//   out[w][co][t] = leaky( bn( b[co] + sum_{ci, j} W[co][ci][j] * in[w][ci][t + j - 1] ) ),  t = 0..10
// `in` / `outb`: shared [IE_W][C][IE_TP] with the zero pads in place.
template <int NW>
__device__ __forceinline__ void ie_conv(const float* __restrict__ in, int Cin, float* __restrict__ outb, int Cout, int co,
                                        int w0, const float* __restrict__ Wt, float bias, float scale, float shift) {
  float acc[NW][IE_T];
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int t = 0; t < IE_T; ++t) acc[w][t] = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float k0 = Wt[static_cast<size_t>(ci * 3 + 0) * Cout + co];
    const float k1 = Wt[static_cast<size_t>(ci * 3 + 1) * Cout + co];
    const float k2 = Wt[static_cast<size_t>(ci * 3 + 2) * Cout + co];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float* row = in + (static_cast<size_t>(w0 + w) * Cin + ci) * IE_TP;
      float x[IE_TP];
#pragma unroll
      for (int q = 0; q < IE_TP / 4; ++q) {
        const float4 v = ld4(row + 4 * q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < IE_T; ++t) {                     // taps in PyTorch's order j = 0, 1, 2 (input t-1, t, t+1)
        acc[w][t] = fmaf(k0, x[t], acc[w][t]);
        acc[w][t] = fmaf(k1, x[t + 1], acc[w][t]);
        acc[w][t] = fmaf(k2, x[t + 2], acc[w][t]);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    float* row = outb + (static_cast<size_t>(w0 + w) * Cout + co) * IE_TP;
    row[0] = 0.f;
#pragma unroll
    for (int t = 0; t < IE_T; ++t) {
      const float v = fmaf(acc[w][t] + bias, scale, shift);          // BatchNorm1d, running statistics
      row[1 + t] = v > 0.f ? v : 0.1f * v;                           // LeakyReLU(0.1)
    }
#pragma unroll
    for (int t = IE_T + 1; t < IE_TP; ++t) row[t] = 0.f;
  }
}

__global__ void __launch_bounds__(IE_THREADS, 1) imu_encoder_kernel(const IeParams p) {
  extern __shared__ __align__(16) float ie_smem[];
  float* a2 = ie_smem;                                   // [8][128][16]
  float* a3 = a2 + IE_W * IE_C2 * IE_TP;                 // [8][256][16]; the input and the first activation share its
  float* a0 = a3;                                        // [8][6][16]     space (both dead before conv 3 writes a3)
  float* a1 = a0 + IE_W * IE_C0 * IE_TP;                 // [8][64][16]
  const int tid = threadIdx.x;
  const long long nwin = static_cast<long long>(p.B) * p.S;

  for (long long win0 = static_cast<long long>(blockIdx.x) * IE_W; win0 < nwin; win0 += static_cast<long long>(gridDim.x) * IE_W) {
    // ---- the CTA's windows, channel-major with zero pads (Encoder.py:57-66: window i = rows 10 i .. 10 i + 10, permuted)
    for (int e = tid; e < IE_W * IE_C0 * IE_TP; e += IE_THREADS) {
      const int w = e / (IE_C0 * IE_TP), rem = e - w * (IE_C0 * IE_TP), c = rem / IE_TP, tp = rem - c * IE_TP;
      const long long win = win0 + w;
      float v = 0.f;
      if (win < nwin && tp >= 1 && tp <= IE_T) {
        const long long b = win / p.S, s = win - b * p.S;
        v = p.imu[(b * p.T + s * 10 + (tp - 1)) * IE_C0 + c];
      }
      a0[e] = v;
    }
    __syncthreads();
    // ---- conv stack: thread = output channel x window group
    auto bn = [&](int l, int co, float& bias, float& scale, float& shift) {
      bias = p.cb[l][co];
      scale = p.bn_w[l][co] * rsqrtf(p.bn_v[l][co] + p.eps);
      shift = p.bn_b[l][co] - p.bn_m[l][co] * scale;
    };
    {
      const int co = tid % IE_C1, wg = tid / IE_C1;                     // 64 channels x 4 groups of 2 windows
      float b_, s_, h_;
      bn(0, co, b_, s_, h_);
      ie_conv<2>(a0, IE_C0, a1, IE_C1, co, 2 * wg, p.Wt[0], b_, s_, h_);
    }
    __syncthreads();
    {
      const int co = tid % IE_C2, wg = tid / IE_C2;                     // 128 channels x 2 groups of 4 windows
      float b_, s_, h_;
      bn(1, co, b_, s_, h_);
      ie_conv<4>(a1, IE_C1, a2, IE_C2, co, 4 * wg, p.Wt[1], b_, s_, h_);
    }
    __syncthreads();
    {
      float b_, s_, h_;                                                 // 256 channels, two passes of 4 windows
      bn(2, tid, b_, s_, h_);
      ie_conv<4>(a2, IE_C2, a3, IE_C3, tid, 0, p.Wt[2], b_, s_, h_);
      ie_conv<4>(a2, IE_C2, a3, IE_C3, tid, 4, p.Wt[2], b_, s_, h_);
    }
    __syncthreads();
    // ---- projection: out[w][o] = bp[o] + sum_{c, t} a3[w][c][t] * Wp[o][c * 11 + t]; thread = output feature
    for (int o = tid; o < p.F; o += IE_THREADS) {
      float acc[IE_W];
#pragma unroll
      for (int w = 0; w < IE_W; ++w) acc[w] = 0.f;
      for (int c = 0; c < IE_C3; ++c) {
        float wv[IE_T];
#pragma unroll
        for (int t = 0; t < IE_T; ++t) wv[t] = p.Wp[static_cast<size_t>(c * IE_T + t) * p.F + o];
#pragma unroll
        for (int w = 0; w < IE_W; ++w) {
          const float* row = a3 + (static_cast<size_t>(w) * IE_C3 + c) * IE_TP;
          const float4 v0 = ld4(row), v1 = ld4(row + 4), v2 = ld4(row + 8);
          const float x[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
          for (int t = 0; t < IE_T; ++t) acc[w] = fmaf(wv[t], x[1 + t], acc[w]);
        }
      }
      const float bo = p.bp[o];
#pragma unroll
      for (int w = 0; w < IE_W; ++w)
        if (win0 + w < nwin) p.out[(win0 + w) * p.F + o] = acc[w] + bo;
    }
    __syncthreads();
  }
}

constexpr size_t ie_smem_bytes() {
  return static_cast<size_t>(IE_W) * (IE_C2 + IE_C3) * IE_TP * sizeof(float);       // 196 KB
}
size_t ie_ws_floats(int F, size_t off[4]) {
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = (o + n + 63) / 64 * 64; return r; };
  off[0] = take(static_cast<size_t>(IE_C0) * 3 * IE_C1);
  off[1] = take(static_cast<size_t>(IE_C1) * 3 * IE_C2);
  off[2] = take(static_cast<size_t>(IE_C2) * 3 * IE_C3);
  off[3] = take(static_cast<size_t>(IE_C3) * IE_T * F);
  return o;
}

}  // namespace
}  // namespace odevio

using namespace odevio;

extern "C" {

size_t odevio_imu_encoder_workspace_bytes(int32_t i_f_len) {
  if (i_f_len <= 0) return 0;
  size_t off[4];
  return ie_ws_floats(i_f_len, off) * sizeof(float);
}

int32_t odevio_imu_encoder_forward(int32_t B, int32_t S, int32_t i_f_len, const odevio_imu_encoder_weights* w,
                                   const float* imu, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!w || !imu || !out || !workspace) return ODEVIO_E_NULL;
  if (B <= 0 || S <= 0 || i_f_len <= 0) return ODEVIO_E_SHAPE;
  for (int l = 0; l < 3; ++l)
    if (!w->conv_w[l] || !w->conv_b[l] || !w->bn_weight[l] || !w->bn_bias[l] || !w->bn_mean[l] || !w->bn_var[l]) return ODEVIO_E_NULL;
  if (!w->proj_w || !w->proj_b) return ODEVIO_E_NULL;
  size_t off[4];
  const size_t need = ie_ws_floats(i_f_len, off) * sizeof(float);
  if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ODEVIO_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* ws = static_cast<float*>(workspace);
  const int cin[3] = {IE_C0, IE_C1, IE_C2}, cout[3] = {IE_C1, IE_C2, IE_C3};
  IeParams p;
  p.B = B; p.S = S; p.T = 10 * S + 1; p.F = i_f_len; p.imu = imu; p.out = out; p.eps = w->bn_eps;
  for (int l = 0; l < 3; ++l) {
    // conv weight [C_out][C_in][3] is a row-major [C_out][C_in * 3] matrix -> K-major [C_in * 3][C_out]
    cudaError_t e = transpose_pack(w->conv_w[l], cout[l], cin[l] * 3, ws + off[l], cout[l], 0, 0, stream);
    if (e != cudaSuccess) return static_cast<int32_t>(e);
    p.Wt[l] = ws + off[l]; p.cb[l] = w->conv_b[l];
    p.bn_w[l] = w->bn_weight[l]; p.bn_b[l] = w->bn_bias[l]; p.bn_m[l] = w->bn_mean[l]; p.bn_v[l] = w->bn_var[l];
  }
  cudaError_t e = transpose_pack(w->proj_w, i_f_len, IE_C3 * IE_T, ws + off[3], i_f_len, 0, 0, stream);
  if (e != cudaSuccess) return static_cast<int32_t>(e);
  p.Wp = ws + off[3]; p.bp = w->proj_b;
  const size_t smem = ie_smem_bytes();
  e = cudaFuncSetAttribute(imu_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int32_t>(e);
  const long long nwin = static_cast<long long>(B) * S;
  long long grid = (nwin + IE_W - 1) / IE_W;
  if (grid > 148 * 8) grid = 148 * 8;
  imu_encoder_kernel<<<static_cast<unsigned>(grid), IE_THREADS, smem, stream>>>(p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int32_t>(e);
}

}  // extern "C"
