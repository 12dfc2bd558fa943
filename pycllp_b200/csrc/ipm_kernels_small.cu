// ipm_kernels_small.cu -- the solve kernel for SMALL problems (factor and vectors in shared memory,
// working set <= half an SM's shared memory), compiled for 64 registers per thread so that TWO
// 512-thread blocks are resident per SM: at m = 50 a Newton step is a chain of short,
// latency-bound phases and a lone block leaves the SM idle most of the time; a second LP in
// flight fills it (north_star: several problems per SM for small m).  Same code as
// ipm_kernels.cu (preset "cl"), only the register budget differs (nvcc -maxrregcount=64 for
// this translation unit, pycllp_b200/build.py).
#include "ipm_solve.cuh"
#include "ipm_small.cuh"

namespace pb200 {

cudaError_t launch_solve_small(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                               int grid, size_t smem_bytes, cudaStream_t stream) {
  solve_kernel_t k = ipm_solve_kernel<true, true, true, 2>;
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err != cudaSuccess) return err;
  k<<<grid, NT, smem_bytes, stream>>>(A, B, sc, p);
  return cudaGetLastError();
}

int small_kernel_max_blocks_per_sm(size_t smem_bytes) {
  int nb = 0;
  solve_kernel_t k = ipm_solve_kernel<true, true, true, 2>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, NT, smem_bytes);
  return nb;
}

// the 128-thread kernel of ipm_small.cuh (m <= 64, everything in shared memory)
cudaError_t launch_solve_tiny(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                              int grid, size_t smem_bytes, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(ipm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err != cudaSuccess) return err;
  ipm_small_kernel<<<grid, SNT, smem_bytes, stream>>>(A, B, sc, p);
  return cudaGetLastError();
}

size_t tiny_kernel_smem_bytes(const Matrix& A) { return small_smem_doubles(A.m, A.n, A.nd) * sizeof(double); }

int tiny_kernel_blocks_per_sm(size_t smem_bytes) {
  int nb = 0;
  if (cudaFuncSetAttribute(ipm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ipm_small_kernel, SNT, smem_bytes);
  return nb;
}

}  // namespace pb200
