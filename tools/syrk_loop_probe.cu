// syrk_loop_probe.cu -- which ingredient of the SYRK inner loop costs DMMA throughput?
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(double* out, long long* cyc, int iters, int ldm, int n0, int n1, int n2) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4 * ldm + 64; i += blockDim.x) sm[i] = 1e-3 * (i & 255);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  double acc[3][4][2];
  for (int s = 0; s < 3; s++) for (int t = 0; t < 4; t++) acc[s][t][0] = acc[s][t][1] = 0;
  int segI[3] = {warp % 25, (warp + 5) % 25, (warp + 11) % 25};
  int segJ[3] = {0, 4, 8};
  int segN[3] = {n0, n1, n2};
  const double* col = sm + tg * ldm + g;
  const double* dg = sm + 4 * ldm;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double dk = dg[(it & 7) * 4 + tg];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      if (MODE == 3 ? (segN[s] > 0) : true) {
        double as;
        if (MODE == 2) as = dk; else as = col[8 * segI[s]];
        if (MODE == 0 || MODE == 3) as *= dk;
#pragma unroll
        for (int t = 0; t < 4; t++) {
          if (MODE == 3 ? (t < segN[s]) : true) {
            double b;
            if (MODE == 2) b = dk; else b = col[8 * (segJ[s] + t)];
            DMMA(acc[s][t][0], acc[s][t][1], as, b);
          }
        }
      }
    }
  }
  long long t1 = clock64();
  double sum = 0;
  for (int s = 0; s < 3; s++) for (int t = 0; t < 4; t++) sum += acc[s][t][0] + acc[s][t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64);
  const int ldm = 212, iters = 2000;
  size_t smem = (4 * ldm + 64) * sizeof(double);
#define RUN(MODE, name) probe<MODE><<<148, 512, smem>>>(out, cyc, iters, ldm, 4, 4, 4); cudaDeviceSynchronize(); \
  probe<MODE><<<148, 512, smem>>>(out, cyc, iters, ldm, 4, 4, 4); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("%-46s %7.1f cyc per k-step (12 DMMA/warp; DMMA-bound = 768)\n", name, (double)h / iters);
  RUN(0, "V0 LDS a, DMUL, LDS b, DMMA");
  RUN(1, "V1 LDS a, LDS b, DMMA (no DMUL)");
  RUN(2, "V2 register operands only");
  RUN(3, "V3 = V0 + runtime segment predicates");
#define RUNN(a, b, c, name) probe<3><<<148, 512, smem>>>(out, cyc, iters, ldm, a, b, c); cudaDeviceSynchronize(); \
  probe<3><<<148, 512, smem>>>(out, cyc, iters, ldm, a, b, c); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("%-46s %7.1f cyc per k-step (%d DMMA/warp; DMMA-bound = %d)\n", name, (double)h / iters, a + b + c, 64 * (a + b + c));
  RUNN(4, 4, 2, "V3 segments 4,4,2 (2 predicated-off DMMAs)");
  RUNN(4, 4, 0, "V3 segments 4,4,0 (third segment branched over)");
  RUNN(4, 2, 2, "V3 segments 4,2,2 (4 predicated-off DMMAs)");
  RUNN(2, 2, 2, "V3 segments 2,2,2 (6 predicated-off DMMAs)");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
