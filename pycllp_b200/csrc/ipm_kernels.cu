// ipm_kernels.cu -- __global__ entry points of the batched IPM engine (sm_100a).
#include "ipm_solve.cuh"

namespace pb200 {

// (modified) LDL' of given dense matrices -- the reference's `ldl` / `modified_ldl`
// kernels (ldl.cl:28-107), test hook.
__global__ void __launch_bounds__(NT, 1)
ldl_hook_kernel(int N, int m, const double* AA, double* Lout, double* Dout, int modified,
                double beta, double delta, double* scratch, size_t slot) {
  extern __shared__ __align__(16) double smem[];
  Work W;
  double* s = scratch + (size_t)blockIdx.x * slot;
  const size_t lsz = (size_t)m * (m + 1) / 2;   // (output layout: row-major packed, as the reference)
  W.red = smem;
  W.P = s;
  W.L = s + (size_t)2 * m * NB + 512;
  W.D = W.L + packed_doubles(m);
  W.prof = nullptr;
  for (int q = blockIdx.x; q < N; q += gridDim.x) {
    W.M = const_cast<double*>(AA) + (size_t)q * m * m;
    if (modified) {
      for (int e = threadIdx.x; e < m * m; e += NT) {
        const int j = e / m, i = e - j * m;
        if (i >= j) W.L[cidx(i, j, m)] = W.M[(size_t)j * m + i];
      }
      __syncthreads();
      factor_ldl_fast(m, W, beta, delta, nullptr, nullptr);
    } else {
      factor_ldl(m, W, beta, delta, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += NT) {
      Dout[(size_t)q * m + i] = W.D[i];
      for (int j = 0; j <= i; j++)
        Lout[q * lsz + (size_t)i * (i + 1) / 2 + j] = W.L[cidx(i, j, m)];
    }
    __syncthreads();
  }
}

// Modified LDL' of given symmetric matrices ON A GIVEN SPARSE PATTERN -- the reference's prototype
// sparse_ldl.modified_ldl (sparse_ldl.py:72-152; the device twin is the factor loop of
// sparse_factor_primal_normal, ldl.cl:422-500), test hook of the tile-sparse factor.
// pat_i / pat_j: the nnz pattern entries (CSR-lower order, diagonal last in each row);
// A.me_pos[k]: where entry k lives in the tile storage.  Ldata (N, nnz) / D (N, m) out.
__global__ void __launch_bounds__(NT, 1)
tiles_hook_kernel(Matrix A, int N, int nnz, const int* pat_i, const int* pat_j, const double* AA,
                  double* Ldata, double* Dout, double beta_in, double delta, double* scratch, size_t slot) {
  extern __shared__ __align__(16) double smem[];
  Work W;
  W.red = smem;
  W.fb = smem + RED_SIZE;
  W.L = scratch + (size_t)blockIdx.x * slot;
  W.D = W.L + (size_t)A.ntiles * 64;
  W.prof = nullptr;
  const int m = A.m, tid = threadIdx.x;
  for (int q = blockIdx.x; q < N; q += gridDim.x) {
    const double* Aq = AA + (size_t)q * m * m;
    for (size_t e = tid; e < (size_t)A.ntiles * 64; e += NT) W.L[e] = 0.0;
    __syncthreads();
    for (int k = tid; k < nnz; k += NT) W.L[A.me_pos[k]] = Aq[(size_t)pat_i[k] * m + pat_j[k]];
    for (int i = m + tid; i < 8 * A.nbk; i += NT) W.L[(size_t)A.tl_colptr[i >> 3] * 64 + (i & 7) * 9] = 1.0;
    __syncthreads();
    const double beta = beta_in > 0.0 ? beta_in : sqrt(tiles_diag_absmax(A, W));   // sparse_ldl.py:86
    tiles_factor(A, W, beta, delta);
    for (int k = tid; k < nnz; k += NT)
      Ldata[(size_t)q * nnz + k] = (pat_i[k] == pat_j[k]) ? 1.0 : W.L[A.me_pos[k]];
    for (int i = tid; i < m; i += NT) Dout[(size_t)q * m + i] = W.D[i];
    __syncthreads();
  }
}

cudaError_t launch_tiles_hook(const Matrix& A, int N, int nnz, const int* pat_i, const int* pat_j,
                              const double* AA, double* Ldata, double* D, double beta, double delta,
                              double* scratch, size_t slot, int grid, cudaStream_t stream) {
  const size_t smem = (size_t)(RED_SIZE + FB_DOUBLES) * sizeof(double);
  cudaError_t err = cudaFuncSetAttribute(tiles_hook_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  tiles_hook_kernel<<<grid, NT, smem, stream>>>(A, N, nnz, pat_i, pat_j, AA, Ldata, D, beta, delta, scratch, slot);
  return cudaGetLastError();
}

// FP64 tensor-core (DMMA m8n8k4) issue-rate probe: eight independent accumulator tiles per warp,
// nothing but DMMAs in the loop.  The roofline denominator of bench.py is measured with this
// on the box the bench runs on (MEASURED_PEAKS.json holds only HBM and bf16).
__global__ void __launch_bounds__(NT, 1) fp64_probe_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------
// host-side launch helpers (called from cabi.cu)
// ---------------------------------------------------------------------------------------
cudaError_t launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t stream) {
  fp64_probe_kernel<<<blocks, NT, 0, stream>>>(out, iters);
  return cudaGetLastError();
}

size_t smem_doubles(const Matrix& A, int L_in_smem, int vec_in_smem) {
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t o = RED_SIZE;
  if (vec_in_smem) {
    o += al(work_area_doubles(A)) + al(vec_area_doubles(A));
  } else {
    o += 2 * TB * LDT + FB_DOUBLES;
  }
  const size_t g = (size_t)2 * (A.ldd > 0 ? A.ldd : 1);
  if (L_in_smem) o += packed_doubles(A.m) > g ? packed_doubles(A.m) : g;
  else if (vec_in_smem) o += al(g);
  return o;
}

size_t work_area_doubles(const Matrix& A) { return work_area(A); }
size_t vec_area_doubles(const Matrix& A) { return (size_t)4 * A.n + w_doubles(A) + 6 * A.m; }

// preset "cl" conventions -> the kernels with those constants folded in (this file); anything else
// -> the run-time variants (ipm_kernels_py.cu)
static bool is_cl(const Params& p) {
  return !p.nan_guard && !p.mu_mode && !p.refine_mode && p.theta_floor && !p.dz_mode;
}

bool params_are_cl(const Params& p) { return is_cl(p); }

cudaError_t launch_solve(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                         int grid, size_t smem_bytes, cudaStream_t stream) {
  if (is_cl(p) && sc.small) return launch_solve_small(A, B, sc, p, grid, smem_bytes, stream);
  if (is_cl(p)) return launch_solve_t<true>(A, B, sc, p, grid, smem_bytes, stream);
  return launch_solve_py(A, B, sc, p, grid, smem_bytes, stream);
}

cudaError_t launch_ldl_hook(int N, int m, const double* AA, double* L, double* D, int modified,
                            double beta, double delta, double* scratch, size_t slot, int grid,
                            cudaStream_t stream) {
  ldl_hook_kernel<<<grid, NT, RED_SIZE * sizeof(double), stream>>>(N, m, AA, L, D, modified, beta,
                                                              delta, scratch, slot);
  return cudaGetLastError();
}

int solve_kernel_max_blocks_per_sm(size_t smem_bytes, int L_in_smem, int vec_in_smem) {
  int nb = 0;
  solve_kernel_t k = pick_kernel<true>(L_in_smem, vec_in_smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, NT, smem_bytes);
  return nb;
}

}  // namespace pb200
