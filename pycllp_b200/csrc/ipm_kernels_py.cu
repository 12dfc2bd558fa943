// ipm_kernels_py.cu -- the solve kernels with every convention read from Params at run time
// (preset "py" = solvers/normal_eqns.py + _ldl.pyx, and any mixture set through set_params).
#include "ipm_solve.cuh"

namespace pb200 {

cudaError_t launch_solve_py(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                            int grid, size_t smem_bytes, cudaStream_t stream) {
  return launch_solve_t<false>(A, B, sc, p, grid, smem_bytes, stream);
}

}  // namespace pb200
