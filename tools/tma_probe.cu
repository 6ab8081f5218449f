// Development probe (not product code): where do the ~450 clk per ring stage of a bulk-TMA -> mbarrier ring go?
//   variant 0: producer warp + consumer warp (full / empty barriers), per-phase clock sums
//   variant 1: ONE warp is producer and consumer (software pipeline, no empty barriers)
//   variant 2: NP producer warps (stage s served by warp s % NP) + one consumer warp
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tma_probe tools/tma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../odevio_b200/csrc/common.cuh"
using namespace odevio;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Prm { int variant, chunk, nstages, nchunks, np, src_chunks; const unsigned char* src; long long* out; };

__global__ void __launch_bounds__(256, 1) tma_probe(const Prm p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[16];
  __shared__ __align__(8) uint64_t empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  const unsigned char* src0 = p.src + static_cast<size_t>(blockIdx.x) * p.src_chunks * p.chunk;
  long long* out = p.out + blockIdx.x * 8;
  if (p.variant == 1) {
    if (warp == 0) {
      int sc = 0;
      for (int s = 0; s < p.nstages; ++s) {
        if (elect_one()) { mbar_arrive_expect_tx(&full[s], p.chunk); tma_load_1d(smem + s * p.chunk, src0 + static_cast<size_t>(sc) * p.chunk, p.chunk, &full[s]); }
        __syncwarp();
        if (++sc == p.src_chunks) sc = 0;
      }
      uint32_t st = 0, ph = 0;
      long long t0 = clock64(), twait = 0, tissue = 0;
      for (int ch = 0; ch < p.nchunks; ++ch) {
        long long a = clock64();
        mbar_wait(&full[st], ph);
        long long b = clock64();
        if (elect_one()) { mbar_arrive_expect_tx(&full[st], p.chunk); tma_load_1d(smem + st * p.chunk, src0 + static_cast<size_t>(sc) * p.chunk, p.chunk, &full[st]); }
        __syncwarp();
        long long c = clock64();
        twait += b - a; tissue += c - b;
        if (++st == static_cast<uint32_t>(p.nstages)) { st = 0; ph ^= 1u; }
        if (++sc == p.src_chunks) sc = 0;
      }
      long long t1 = clock64();
      if (lane == 0) { out[0] = t1 - t0; out[1] = twait; out[2] = tissue; }
      // drain
      for (int s = 0; s < p.nstages; ++s) { mbar_wait(&full[st], ph); if (++st == static_cast<uint32_t>(p.nstages)) { st = 0; ph ^= 1u; } }
    }
    return;
  }
  const int np = p.variant == 2 ? p.np : 1;
  if (warp >= 1 && warp <= np) {
    const int w = warp - 1;
    long long twait = 0, t_exp = 0, t_tma = 0;
    // this warp serves chunks ch = w, w + np, ...; stage = ch % nstages (nstages % np == 0)
    int sc = w % p.src_chunks;
    for (int ch = w; ch < p.nchunks; ch += np) {
      const uint32_t st = ch % p.nstages, ph = (ch / p.nstages) & 1u;
      long long a = clock64();
      mbar_wait(&empty[st], ph ^ 1u);
      long long b = clock64();
      long long c = b, d = b;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], p.chunk);
        c = clock64();
        tma_load_1d(smem + st * p.chunk, src0 + static_cast<size_t>(sc) * p.chunk, p.chunk, &full[st]);
        d = clock64();
      }
      __syncwarp();
      twait += b - a; t_exp += c - b; t_tma += d - c;
      sc += np; if (sc >= p.src_chunks) sc -= p.src_chunks;
    }
    if (w == 0 && lane == 0) { out[3] = twait; }
    // the elected lane's sums (lane unknown): reduce by max over lanes
    for (int o = 16; o; o >>= 1) { t_exp = max(t_exp, __shfl_xor_sync(0xffffffffu, t_exp, o)); t_tma = max(t_tma, __shfl_xor_sync(0xffffffffu, t_tma, o)); }
    if (w == 0 && lane == 0) { out[4] = t_exp; out[5] = t_tma; }
  } else if (warp == 0) {
    uint32_t st = 0, ph = 0;
    long long t0 = clock64(), twait = 0;
    for (int ch = 0; ch < p.nchunks; ++ch) {
      long long a = clock64();
      mbar_wait(&full[st], ph);
      twait += clock64() - a;
      if (lane == 0) mbar_arrive(&empty[st]);
      __syncwarp();
      if (++st == static_cast<uint32_t>(p.nstages)) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = twait; }
  }
}

int main() {
  long long* d_out; unsigned char* d_src;
  CK(cudaMalloc(&d_out, sizeof(long long) * 8 * 148));
  CK(cudaMalloc(&d_src, 256u << 20)); CK(cudaMemset(d_src, 0, 256u << 20));
  CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int chunks[] = {4096, 16384};
  for (int variant = 0; variant < 3; ++variant)
    for (int ci = 0; ci < 2; ++ci)
      for (int np = 1; np <= 4; np *= 2) {
        if (variant != 2 && np > 1) continue;
        if (variant == 2 && np == 1) continue;
        for (int grid = 1; grid <= 148; grid += 147) {
          Prm p; p.variant = variant; p.chunk = chunks[ci]; p.nstages = 8; p.nchunks = 4096; p.np = np;
          p.src_chunks = (512 * 1024) / p.chunk; p.src = d_src; p.out = d_out;
          CK(cudaMemset(d_out, 0, sizeof(long long) * 8 * 148));
          for (int rep = 0; rep < 3; ++rep) { tma_probe<<<grid, 256, 8 * p.chunk + 1024>>>(p); CK(cudaDeviceSynchronize()); }
          long long h[8];
          CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
          const double n = p.nchunks;
          printf("variant %d np %d chunk %5d grid %3d: clk/chunk %7.1f  consumer wait %7.1f  issue(v1) %7.1f | producer (warp 0 of np): wait empty %7.1f  expect_tx %6.1f  bulk copy %6.1f (per own chunk)\n",
                 variant, np, p.chunk, grid, h[0] / n, h[1] / n, h[2] / n, h[3] / (n / np), h[4] / (n / np), h[5] / (n / np));
        }
      }
  printf("done\n");
  return 0;
}
