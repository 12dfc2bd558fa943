// ipm_solve.cuh -- the persistent solve kernel (template).  Instantiated in two translation
// units so that they compile in parallel: ipm_kernels.cu (CL = true: the constants of the
// reference's OpenCL path folded in) and ipm_kernels_py.cu (CL = false: every convention read
// from Params at run time -- the "py" preset and any mixture).
#pragma once
#include "ipm_device.cuh"
#include "ipm_host.h"

namespace pb200 {

__device__ __forceinline__ size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// Carve the block's working set out of shared memory / its scratch slot.  LS / VS (factor /
// vectors in shared memory) are compile-time so that the compiler can prove which pointers
// are shared and emit LDS/STS with 32-bit addresses instead of generic loads.
template <bool LS, bool VS>
static __device__ __forceinline__ void carve(const Matrix& A, const Scratch& sc, double* smem, Work& W) {
  const int m = A.m, n = A.n;
  double* slot = sc.base + (size_t)blockIdx.x * sc.slot;
  size_t o = 0;
  W.red = smem + o; o += RED_SIZE;
  const size_t psz = work_area(A);
  const size_t nw = w_doubles(A);
  double* v;
  if constexpr (VS) {
    W.P = smem + o; o += align16(psz);
    W.tiles = nullptr;
    W.fb = W.P;                 // (work_area() >= FB_DOUBLES whenever the big factor can run)
    v = smem + o; o += align16((size_t)4 * n + nw + 6 * m);
    W.dg = v + 4 * (size_t)n;   // = W.w (max(n, ldd) doubles): dead between prepare_rhs and step
  } else {   // large problems: only the reduction scratch and the SYRK macro tiles stay on-chip
    W.tiles = smem + o; o += 2 * TB * LDT;
    W.fb = smem + o; o += FB_DOUBLES;
    W.P = slot + sc.off_P;
    W.dg = slot + sc.off_dg;
    v = slot + sc.off_vec;
  }
  const size_t ldd1 = (size_t)(A.ldd > 0 ? A.ldd : 1);
  W.g1 = W.P;      // gather buffers of A_times2 (!VS): the panel/stage area is idle then
  W.g2 = W.P + ldd1;
  W.x = v; W.z = v + n; W.t = v + 2 * n; W.d = v + 3 * n; W.w = v + 4 * n; W.c = nullptr;
  double* u = v + 4 * (size_t)n + nw;
  W.y = u; W.b = u + m; W.dy = u + 2 * m; W.S = u + 3 * m; W.RHS = u + 4 * m; W.D = u + 5 * m;
  if constexpr (LS) W.L = smem + o;
  else W.L = slot + sc.off_L;
  if constexpr (VS) {
    // gather buffers of A_times2: the factor's storage is dead while the right-hand side is
    // prepared; without it in shared memory, 2 ldd more doubles after the vectors
    W.g1 = smem + o;
    W.g2 = W.g1 + ldd1;
  }
  W.M = slot;
  W.prof = sc.prof ? sc.prof + (size_t)blockIdx.x * 16 : nullptr;
}

template <bool LS, bool VS, bool CL>
static __device__ __forceinline__ void ipm_solve_one(const Matrix& A, const Batch& B, Work& W, const Params& p, int q) {
  const int m = A.m, n = A.n, tid = threadIdx.x;
  W.c = B.c + (size_t)q * n;
  const bool given = B.hook || B.warm;            // start from the caller's x, z, y (primal_normal.cl:213-219)
  const size_t ld0n = B.ld_0 ? B.ld_0 : (size_t)n, ld0m = B.ld_0 ? B.ld_0 : (size_t)m;
  for (int j = tid; j < n; j += NT) {
    double x0 = 1.0, z0 = 1.0;                       // initialize_xzyw, primal_normal.cl:14-28
    if (given) {
      x0 = B.x0[(size_t)q * ld0n + j];
      z0 = B.z0[(size_t)q * ld0n + j];
      if (B.warm) { x0 = fmax(x0, p.warm_floor); z0 = fmax(z0, p.warm_floor); }
    }
    W.x[j] = x0;
    W.z[j] = z0;
  }
  for (int i = tid; i < m; i += NT) {
    const int io = A.rperm ? A.rperm[i] : i;        // the caller's row (constraints reordered at setup)
    W.b[i] = B.b[(size_t)q * m + io];
    W.y[i] = given ? B.y0[(size_t)q * ld0m + io] : 1.0;
  }
  __syncthreads();

  if (B.hook) {   // one solve_primal_normal (ldl.cl:602-653) on the given state
    double nr, ns;
    prepare_rhs<VS>(A, W, B.mu, nr, ns);
    solve_normal<LS, VS, CL>(A, W, p);
    for (int i = tid; i < m; i += NT) B.dy_out[(size_t)q * m + (A.rperm ? A.rperm[i] : i)] = W.dy[i];
    __syncthreads();
    return;
  }

  double cscale;                                  // max(1, max_j |c_j|): the scale of v = A'y near the optimum
  {
    double cm = 0.0;
    for (int j = tid; j < n; j += NT) cm = fmax(cm, fabs(W.c[j]));
    cscale = fmax(1.0, block_max(cm, W.red));
  }
  int stat = 5;                                   // primal_normal.cl:225
  double normr0 = INFINITY, norms0 = INFINITY;    // HUGE_VALF/10, :227-228
  int iter;
  // v = A'y is carried over from the previous step (one pass over A less) for the first
  // p.carry_v iterations (default 64: every LP that converges is done long before) and recomputed
  // from y like the reference does afterwards, which bounds the drift on the long diverging runs
  // of infeasible LPs; carry_v = 0 recomputes it every iteration (primal_normal.cl:76-94)
  bool carry_v = false;
  for (iter = 0; iter < p.max_iter; iter++) {
    double g = 0.0;
    for (int j = tid; j < n; j += NT) g += W.z[j] * W.x[j];
    const double gamma = block_sum(g, W.red);
    const double mu = p.delta * gamma / (double)((!CL && p.mu_mode) ? n : n + m);   // :272 (normal_eqns.py:65)
    double normr, norms;
    long long t0 = phase_begin(W);
    prepare_rhs<VS>(A, W, mu, normr, norms, carry_v);
    // (only while the primal residual still shrinks by >= 10 % per step: on an infeasible LP it
    // stalls, y diverges and the rounding noise of A'y -- which the reference recomputes every
    // iteration -- is what eventually trips the |sigma| > 10 |sigma_0| test, primal_normal.cl:266;
    // a carried v has a different noise and misses it: tests/test_gpu_features.py).  A residual that
    // is already below the stop tolerance is not stalling, it is done: the end game of a converging
    // LP, where |rho| sits at rounding level, keeps the carried v.
    carry_v = iter + 1 < p.carry_v && (normr <= 0.9 * normr0 || normr < p.eps);
    phase_end(W, 0, t0);
    if (B.trace && tid == 0 && iter < B.trace_iters) {           // :250-252 (the kernel's verbose > 1 printf)
      double* tr = B.trace + ((size_t)q * B.trace_iters + iter) * 3;
      tr[0] = normr; tr[1] = norms; tr[2] = gamma;
    }
    if (normr < p.eps && norms < p.eps && gamma < p.eps) { stat = 0; break; }   // :256-259
    if (normr > 10 * normr0 && normr > p.eps) { stat = 2; break; }              // :261-264
    if (norms > 10 * norms0 && norms > p.eps) { stat = 4; break; }              // :266-269
    // the iteration's scalars wait in shared memory while the tensor-core phases run: the SYRK
    // and factor loops need every register they can get (block-uniform values, benign race)
    W.red[RED_KEEP] = normr; W.red[RED_KEEP + 1] = norms; W.red[RED_KEEP + 2] = mu;
    // Residual without a stored M (residual_free) while both infeasibilities still shrink or are
    // already below the tolerance.  On a diverging (infeasible / unbounded) LP the residual is
    // rounding noise of size eps |M| |dy| >> refine_tol, every refinement pass feeds that noise back
    // into dy, and which status the run ends in depends on the noise: there the residual is
    // evaluated from the stored M exactly as before (tests/test_gpu_features.py, the
    // 224-instance status fixture).
    const bool healthy = (normr < normr0 || normr < p.eps) && (norms < norms0 || norms < p.eps);
    const bool have_w = solve_normal<LS, VS, CL>(A, W, p, healthy);
    if (!CL && p.nan_guard) {                                            // normal_eqns.py:85-87
      int bad = 0;
      for (int i = tid; i < m; i += NT) bad |= isnan(W.dy[i]);
      if (__syncthreads_or(bad)) { stat = 3; break; }
    }
    t0 = phase_begin(W);
    step<VS, CL>(A, W, W.red[RED_KEEP + 2], p, have_w);
    // ... and only if mu/x stayed moderate in this step (the error of the v it left, see step())
    carry_v = carry_v && W.red[RED_KEEP + 3] <= 1.0e3 * cscale;
    phase_end(W, 5, t0);
    normr0 = W.red[RED_KEEP];
    norms0 = W.red[RED_KEEP + 1];
  }
  // (warm start: x0/z0/y0 may alias the outputs -- with the same leading dimensions)
  const size_t ldx = B.ld_x ? B.ld_x : (size_t)n, ldy = B.ld_y ? B.ld_y : (size_t)m, ldz = B.ld_z ? B.ld_z : (size_t)n;
  const int lds = B.ld_s ? B.ld_s : 1;
  if (B.x) for (int j = tid; j < n; j += NT) B.x[(size_t)q * ldx + j] = W.x[j];
  if (B.z) for (int j = tid; j < n; j += NT) B.z[(size_t)q * ldz + j] = W.z[j];
  if (B.y) for (int i = tid; i < m; i += NT) B.y[(size_t)q * ldy + (A.rperm ? A.rperm[i] : i)] = W.y[i];
  if (tid == 0) {
    if (B.status) B.status[(size_t)q * lds] = stat;
    if (B.iters) B.iters[(size_t)q * lds] = iter;
  }
  __syncthreads();
}

// MINB: resident blocks per SM the register allocation is sized for (1: 128 registers per thread;
// 2: 64 -- small problems whose working set lets two blocks share an SM, ipm_kernels_small.cu)
template <bool LS, bool VS, bool CL, int MINB>
__global__ void __launch_bounds__(NT, MINB)
ipm_solve_kernel(Matrix A, Batch B, Scratch sc, Params p) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next;
  Work W;
  carve<LS, VS>(A, sc, smem, W);
  if (W.prof && threadIdx.x < 16) reinterpret_cast<unsigned long long*>(W.red + RED_PROF)[threadIdx.x] = 0;
  ring_init(W);
  if (A.sparse && !A.tiles && p.max_refine > 0) {   // entries outside the pattern of A A' are never written again
    const size_t mm = (size_t)A.m * A.m;
    for (size_t e = threadIdx.x; e < mm; e += NT) W.M[e] = 0.0;
    __syncthreads();
  }
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(sc.counter, 1);
    __syncthreads();
    const int q = __shfl_sync(0xffffffffu, s_next, 0);   // warp-uniform for the compiler
    __syncthreads();
    if (q >= B.N) break;
    ipm_solve_one<LS, VS, CL>(A, B, W, p, q);
  }
  if (W.prof && threadIdx.x < 16)
    W.prof[threadIdx.x] += reinterpret_cast<unsigned long long*>(W.red + RED_PROF)[threadIdx.x];
}


typedef void (*solve_kernel_t)(Matrix, Batch, Scratch, Params);
template <bool CL>
static solve_kernel_t pick_kernel(int L_in_smem, int vec_in_smem) {
  if (L_in_smem && vec_in_smem) return ipm_solve_kernel<true, true, CL, 1>;
  if (vec_in_smem) return ipm_solve_kernel<false, true, CL, 1>;
  return ipm_solve_kernel<false, false, CL, 1>;
}

template <bool CL>
static cudaError_t launch_solve_t(const Matrix& A, const Batch& B, const Scratch& sc, const Params& p,
                                  int grid, size_t smem_bytes, cudaStream_t stream) {
  solve_kernel_t k = pick_kernel<CL>(sc.L_in_smem, sc.vec_in_smem);
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
  if (err != cudaSuccess) return err;
  k<<<grid, NT, smem_bytes, stream>>>(A, B, sc, p);
  return cudaGetLastError();
}

}  // namespace pb200
