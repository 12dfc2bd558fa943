// fp64_probe.cu -- measures FP64 issue rates on the box: DFMA, DMMA m8n8k4 and (if the
// assembler accepts them) the larger f64 mma shapes. Used once to pick the roofline
// denominator for the FP64 kernels (MEASURED_PEAKS.json only has HBM and bf16).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef BIG_SHAPES
__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; i++) b[i] = 1.0 + i * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, smem optin %zu\n", p.name, sms, p.sharedMemPerBlockOptin);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32, blocks = sms * (warps <= 16 ? 2 : 1);
    int iters = 20000;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    double flops = 2.0 * 8 * iters * (double)threads * blocks;
    printf("DFMA   warps/blk=%2d blocks=%d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms / 1e9);
    ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 256 * 8 * iters * (double)warps * blocks;
    printf("DMMA884 warps/blk=%2d blocks=%d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms / 1e9);
#ifdef BIG_SHAPES
    ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters / 4); });
    flops = 2.0 * 2048 * 4 * (iters / 4) * (double)warps * blocks;
    printf("DMMA16816 warps/blk=%2d blocks=%d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms / 1e9);
#endif
  }
  return 0;
}
