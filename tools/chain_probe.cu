// chain_probe.cu -- cost of the serial chain of factor_ldl_ahead (warp 0) in isolation and next to
// busy neighbours: old shuffle-based diag_block + block_row vs chain_panel8 (all-lanes-redundant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pycllp_b200/csrc -o tools/chain_probe tools/chain_probe.cu
#include <cstdio>
#include "ipm_device.cuh"
using namespace pb200;

__global__ void __launch_bounds__(NT, 1) probe(double* out, long long* cyc, int m, int reps, int mode, int which) {
  extern __shared__ __align__(16) double sm[];
  __shared__ volatile int stop;
  Work W;
  W.red = sm;
  W.L = sm + RED_SIZE;
  W.P = W.L + packed_doubles(m);
  W.D = W.P + 2 * 12 * m + 512;
  W.prof = nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  for (int j = 0; j < m; j++)
    for (int i = j + tid; i < m; i += NT) W.L[cidx(i, j, m)] = (i == j) ? 4.0 + 0.01 * j : 0.01 / (1 + i - j);
  for (int i = tid; i < 2 * 12 * m + 512; i += NT) W.P[i] = 1e-3 * (i & 63);
  if (tid == 0) stop = 0;
  __syncthreads();
  double* blk = W.P + 2 * 12 * (m - 8);
  double* Wp = blk + 160;
  int* th = reinterpret_cast<int*>(W.red + RED_TH);
  double acc = 0;
  if (warp == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      const int j0 = 8 * (r % 20);
      if (which == 0) {
        diag_block(m, j0, 8, W, 1e-6, th, blk, blk + 64, blk + 72);
        __syncwarp();
        if (lane < 8) W.D[j0 + lane] = blk[64 + lane];
      } else if (which == 1) {
        diag_block(m, j0, 8, W, 1e-6, th, blk, blk + 64, blk + 72);
        __syncwarp();
        if (lane < 8) W.D[j0 + lane] = blk[64 + lane];
        block_row(m, j0, 8, W, blk, blk + 72, th, Wp);
      } else if (which == 2) {
        chain_panel8<false>(m, j0, 0, false, W, 1e-6, th, blk, Wp);
      } else {
        chain_panel8<true>(m, j0, 0, false, W, 1e-6, th, blk, Wp);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; stop = 1; }
  } else if (mode >= 1 && (warp & 3) != 0) {
    // DMMA streams on sub-partitions 1-3 (like the trailing update)
    const int g = lane >> 2, tg = lane & 3;
    double c0 = 0, c1 = 0, u0 = 0, u1 = 0;
    while (!stop) {
      old_update16(W.L, W.P, m, 0, 64, tg, g, 64 + 8 * (warp % 8), c0, c1, u0, u1);
      acc += c0 + c1 + u0 + u1;
    }
  } else if (mode >= 2) {
    // shared-memory traffic on sub-partition 0 (like the table builds of warps 4, 8, 12)
    while (!stop) {
      for (int e = lane; e < 8 * 64; e += 32) W.P[(e >> 3) * 12 + (e & 7)] = W.L[coff(e >> 3, m) + 100 + (e & 7)] * W.D[e >> 3];
    }
  }
  out[blockIdx.x * NT + tid] = acc;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64);
  const int m = 200, reps = 2000;
  size_t smem = (RED_SIZE + packed_doubles(m) + 2 * 12 * m + 512 + m + 64) * sizeof(double);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"diag_block (shuffles)", "diag_block + block_row", "chain_panel8<false> (diag)", "chain_panel8<true> (diag + row)"};
  for (int mode = 0; mode < 3; mode++)
    for (int which = 0; which < 4; which++) {
      probe<<<148, NT, smem>>>(out, cyc, m, reps, mode, which);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (%s) %-34s %8.0f cycles/panel  %s\n", mode,
             mode == 0 ? "alone" : mode == 1 ? "DMMA on SMSP1-3" : "DMMA + smem on SMSP0", names[which], (double)h / reps,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
