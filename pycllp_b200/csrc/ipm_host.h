// ipm_host.h -- host-callable launch helpers shared by ipm_kernels.cu and cabi.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "ipm_types.h"

namespace pb200 {

size_t work_area_doubles(const Matrix& A);
size_t vec_area_doubles(const Matrix& A);
size_t smem_doubles(const Matrix& A, int L_in_smem, int vec_in_smem);
cudaError_t launch_solve(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                         int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_solve_small(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                               int grid, size_t smem_bytes, cudaStream_t stream);
int small_kernel_max_blocks_per_sm(size_t smem_bytes);
cudaError_t launch_solve_tiny(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                              int grid, size_t smem_bytes, cudaStream_t stream);
size_t tiny_kernel_smem_bytes(const Matrix& A);
int tiny_kernel_blocks_per_sm(size_t smem_bytes);
bool params_are_cl(const Params& p);
cudaError_t launch_solve_py(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                            int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_tiles_hook(const Matrix& A, int N, int nnz, const int* pat_i, const int* pat_j,
                              const double* AA, double* Ldata, double* D, double beta, double delta,
                              double* scratch, size_t slot, int grid, cudaStream_t stream);
cudaError_t launch_ldl_hook(int N, int m, const double* AA, double* L, double* D, int modified,
                            double beta, double delta, double* scratch, size_t slot, int grid,
                            cudaStream_t stream);
cudaError_t launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t stream);
int solve_kernel_max_blocks_per_sm(size_t smem_bytes, int L_in_smem, int vec_in_smem);
}  // namespace pb200
