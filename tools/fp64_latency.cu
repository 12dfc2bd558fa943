// fp64_latency.cu -- dependent-issue latency of DMMA.8x8x4 / DFMA / generic-vs-shared loads on
// one SM, and FP64-pipe interference between a DMMA stream and a DFMA chain on one SMSP.
#include <cstdio>
#include <cuda_runtime.h>

#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

template <int CH>
__global__ void k_dmma_chain(double* out, long long* cyc, int iters) {
  double c[CH][2];
  for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) DMMA(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_dfma_chain(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) a = fma(a, b, c);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_ddiv_chain(double* out, long long* cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0000001;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) a = b / a;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// warp 0: dependent DFMA chain; warps 4, 8, 12 (same SMSP): DMMA streams (if with_dmma)
__global__ void k_interfere(double* out, long long* cyc, int iters, int with_dmma, int other_smsp) {
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) a = fma(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
  } else if (with_dmma && ((other_smsp ? (warp & 3) == 1 : (warp & 3) == 0))) {
    double c[4][2] = {};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) DMMA(c[i][0], c[i][1], a, b);
    }
    out[threadIdx.x] = c[0][0] + c[1][0] + c[2][1] + c[3][1];
  }
}

__global__ void k_lds_chain(double* out, long long* cyc, int iters, int generic) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  volatile int* gp = idx;
  long long t0 = clock64();
  if (generic) { for (int it = 0; it < iters; it++) p = gp[p]; }
  else { for (int it = 0; it < iters; it++) p = idx[p]; }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  const int it = 4096;
#define RUN(name, launch, per) launch; cudaDeviceSynchronize(); launch; cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.1f cyc per %s\n", name, (double)h / it, per);
  RUN("DMMA 1 chain, 1 warp", (k_dmma_chain<1><<<1, 32>>>(out, cyc, it)), "DMMA");
  RUN("DMMA 2 chains, 1 warp (per iteration of 2)", (k_dmma_chain<2><<<1, 32>>>(out, cyc, it)), "iter");
  RUN("DMMA 4 chains, 1 warp (per iteration of 4)", (k_dmma_chain<4><<<1, 32>>>(out, cyc, it)), "iter");
  RUN("DMMA 8 chains, 1 warp (per iteration of 8)", (k_dmma_chain<8><<<1, 32>>>(out, cyc, it)), "iter");
  RUN("DMMA 1 chain, 4 warps (1/SMSP)", (k_dmma_chain<1><<<1, 128>>>(out, cyc, it)), "DMMA");
  RUN("DMMA 1 chain, 16 warps (4/SMSP)", (k_dmma_chain<1><<<1, 512>>>(out, cyc, it)), "DMMA");
  RUN("DMMA 4 chains, 16 warps (per iteration of 4)", (k_dmma_chain<4><<<1, 512>>>(out, cyc, it)), "iter");
  RUN("DFMA dependent chain", (k_dfma_chain<<<1, 32>>>(out, cyc, it)), "DFMA");
  RUN("DDIV dependent chain", (k_ddiv_chain<<<1, 32>>>(out, cyc, it)), "DDIV");
  RUN("DFMA chain alone (512 thr block)", (k_interfere<<<1, 512>>>(out, cyc, it, 0, 0)), "DFMA");
  RUN("DFMA chain + 3 DMMA warps SAME SMSP", (k_interfere<<<1, 512>>>(out, cyc, it, 1, 0)), "DFMA");
  RUN("DFMA chain + 4 DMMA warps OTHER SMSP", (k_interfere<<<1, 512>>>(out, cyc, it, 1, 1)), "DFMA");
  RUN("LDS pointer chase (shared)", (k_lds_chain<<<1, 32>>>(out, cyc, it, 0)), "load");
  RUN("LD generic->shared pointer chase", (k_lds_chain<<<1, 32>>>(out, cyc, it, 1)), "load");
  return 0;
}
