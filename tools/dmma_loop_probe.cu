// dmma_loop_probe.cu -- isolates the panel-update loops of factor_ldl_fast: how long does one
// k-iteration (loads from shared + DMMAs) take for (a) a lone warp with 4 split-K chains,
// (b) 12 "bulk" warps with 2 tiles each, (c) both together.
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

__device__ __forceinline__ int coff(int j, int m) { return j * (m - 1) - ((j * (j - 1)) >> 1); }

__global__ void __launch_bounds__(512, 1) probe(double* out, long long* cyc, int m, int j0, int reps, int mode) {
  extern __shared__ double sm[];
  double* L = sm;                    // packed lower, column-major
  double* P = sm + m * (m + 1) / 2;  // j0 x 8
  for (int i = threadIdx.x; i < m * (m + 1) / 2 + m * 8; i += blockDim.x) sm[i] = 1e-3 * (i & 255);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  const int wsub = warp & 3;
  double acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    if (wsub != 0 && (mode & 1)) {   // bulk: two tiles, two k-steps per iteration
      const int widx = (warp >> 2) * 3 + wsub - 1;
      const int rowa = j0 + 8 * (1 + widx) + g, rowb = rowa + 96;
      const int ra = rowa < m ? rowa : j0, rb = rowb < m ? rowb : j0;
      double c0 = 0, c1 = 0, e0 = 0, e1 = 0, u0 = 0, u1 = 0, v0 = 0, v1 = 0;
      int k = tg, off = coff(k, m);
      for (int k0 = 0; k0 < j0; k0 += 8) {
        const int off2 = off + 4 * m - 10 - 4 * k;
        const double b1 = P[k * 8 + g], b2 = P[(k + 4) * 8 + g];
        const double a1 = L[off + ra], a2 = L[off2 + ra], a3 = L[off + rb], a4 = L[off2 + rb];
        DMMA(c0, c1, a1, b1); DMMA(u0, u1, a3, b1); DMMA(e0, e1, a2, b2); DMMA(v0, v1, a4, b2);
        off = off2 + 4 * m - 10 - 4 * (k + 4); k += 8;
      }
      acc += c0 + c1 + e0 + e1 + u0 + u1 + v0 + v1;
    } else if (warp == 0 && (mode & 2)) {   // tile 0: four split-K chains
      const int rs = j0 + g;
      double c[4][2] = {};
      int k = tg, off = coff(k, m);
      for (int k0 = 0; k0 < j0; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double a = L[off + rs];
          DMMA(c[q][0], c[q][1], a, P[k * 8 + g]);
          off += 4 * m - 10 - 4 * k; k += 4;
        }
      }
      acc += c[0][0] + c[1][0] + c[2][1] + c[3][1];
    }
    if (mode & 4) __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32)) cyc[threadIdx.x == 0 ? 0 : 1] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h[2];
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64);
  const int m = 200, reps = 200;
  size_t smem = (m * (m + 1) / 2 + m * 8) * sizeof(double);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int j0 : {32, 96, 160}) {
    for (int mode : {1, 2, 3}) {
      probe<<<1, 512, smem>>>(out, cyc, m, j0, reps, mode);
      cudaDeviceSynchronize();
      probe<<<148, 512, smem>>>(out, cyc, m, j0, reps, mode);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("j0=%3d mode=%d (%s): warp0 %8.1f cyc/panel, warp1 %8.1f cyc/panel   [DMMA-bound bulk %.0f]\n", j0, mode,
             mode == 1 ? "bulk only" : mode == 2 ? "tile0 only" : "bulk+tile0", (double)h[0] / reps, (double)h[1] / reps,
             24.0 * (j0 / 4) * 4.0);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
