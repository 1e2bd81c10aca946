// TMEM read / write bandwidth per SM: how long does tcgen05.ld / tcgen05.st of 32 lanes x 64 columns take with
// 1, 4 (one per lane quadrant) and 8 (two per quadrant) warps issuing back to back?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../graph_neural_net_b200/csrc/fgnn_ptx.cuh"
using namespace fgnn::ptx;

__global__ void __launch_bounds__(512, 1) bw_kernel(int nwarps, int iters, int mode, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 % 512);
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = lane + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        tmem_ld64(base, r);
        tmem_wait_ld();
        acc += r[0] ^ r[21] ^ r[42] ^ r[63];
      } else if (mode == 1) {
        tmem_st32(base, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_st32(base + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_wait_st();
      } else if (mode == 3) {   // the mma C-fragment shape: 2 x (16 lanes x 64 columns)
        uint32_t q[32];
        tmem_ld_16x256b_x8(base, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_16x256b_x8(base + (16u << 16), q);
        tmem_wait_ld();
        acc += r[0] ^ r[31] ^ q[0] ^ q[31];
      } else {   // two loads in flight
        uint32_t q[32];
        tmem_ld32(base, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(base + 64, q);
        tmem_wait_ld();
        acc += r[0] ^ r[31] ^ q[0] ^ q[31];
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMallocManaged(&out, 16 * sizeof(long long));
  cudaMalloc(&sink, 512 * 4);
  const int iters = 2000;
  const char* names[4] = {"ld 32x32b.x64 (8 KB/warp)", "st 2 x 32x32b.x32 (8 KB/warp)", "ld 2 x 32x32b.x32 in flight (8 KB/warp)", "ld 2 x 16x256b.x8 (8 KB/warp)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int nw : {1, 2, 4, 8, 16}) {
      bw_kernel<<<1, 512>>>(nw, iters, mode, out, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = out[w] > mx ? out[w] : mx;
      const double cyc = (double)mx / iters;
      printf("%-44s warps=%d  %.1f cycles/iter  -> %.1f B/cycle/SM\n", names[mode], nw, cyc, nw * 8192.0 / cyc);
    }
  return 0;
}
