// ipm_device.cuh -- device side of the batched primal normal-equations IPM (sm_100a).
//
// One thread block solves one LP from start to finish (all <= 200 iterations),
// then fetches the next LP from a global work counter: a persistent grid of
// (#SMs x resident blocks) CTAs, no host round trips, no inter-block
// synchronisation, and all per-problem state (x, z, y, the normal matrix M, its
// LDL' factor) stays on-chip / in an L2-resident scratch slot owned by the block.
//
// What is computed follows the reference's OpenCL kernels (same constants, same stop
// rule, same modified LDL', same refinement rule; the operation ORDER differs -- see the
// list in DESIGN.md section 1 -- so results agree to rounding, not bit for bit):
//   primal_normal.cl:201-284  standard_primal_normal          -> ipm_solve_one
//   primal_normal.cl:30-120   primal/dual infeasibility       -> stage "residual norms"
//   ldl.cl:110-138,280-294    A (X/Z) A' entries, beta        -> form_M_dense_tma (ipm_syrk.cuh) / form_M_* (once/iteration)
//   ldl.cl:314-378            factor_primal_normal            -> factor_ldl_ahead / factor_ldl_fast (ipm_factor.cuh)
//   ldl.cl:198-219            primal_normal_rhs_i             -> RHS from the stored t
//   ldl.cl:505-537            forward_backward_primal_normal  -> forward part inside factor_ldl_*, back_solve_fast
//   ldl.cl:577-599            residual_primal_normal          -> residual_M
//   ldl.cl:602-653            solve_primal_normal             -> solve_normal
//   primal_normal.cl:122-156  primal_normal_step              -> step
// How it is computed is different: M is formed ONCE per iteration (FP64 tensor-core
// SYRK over the columns of A with >= 2 non-zeros, singleton/slack columns add to the
// diagonal only), t = c - A'y + mu/x is evaluated ONCE and reused by the RHS and the
// step, and all loops are block-cooperative.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "ipm_types.h"

namespace pb200 {

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
// Warp index as a value the compiler KNOWS to be warp-uniform (broadcast from lane 0).  With
// the plain threadIdx.x >> 5 every `if (warp == ...)` is treated as potentially divergent and
// each __shfl_sync / mma.sync inside it is wrapped in WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE
// (measured: the 8x8 pivot block ran 6x slower that way).
__device__ __forceinline__ int warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Deterministic block reductions; every thread gets the result. red: >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = warp_id();
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = (lane < NWARP) ? red[lane] : 0.0;
  r = warp_sum(r);
  __syncthreads();
  return __shfl_sync(0xffffffffu, r, 0);     // provably warp-uniform (loop exits depend on it)
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = warp_id();
  v = warp_max(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = (lane < NWARP) ? red[lane] : 0.0;   // all callers reduce non-negative values
  r = warp_max(r);
  __syncthreads();
  return __shfl_sync(0xffffffffu, r, 0);     // provably warp-uniform
}

// two maxima of non-negative values with one exchange through shared memory (red: >= 64 doubles)
__device__ __forceinline__ double block_max2(double v, double& w, double* red) {
  const int lane = threadIdx.x & 31, warp = warp_id();
  v = warp_max(v);
  w = warp_max(w);
  if (lane == 0) { red[warp] = v; red[NWARP + warp] = w; }
  __syncthreads();
  double r = (lane < NWARP) ? red[lane] : 0.0;
  double q = (lane < NWARP) ? red[NWARP + lane] : 0.0;
  r = warp_max(r);
  q = warp_max(q);
  __syncthreads();
  w = __shfl_sync(0xffffffffu, q, 0);
  return __shfl_sync(0xffffffffu, r, 0);
}

// minimum over the block of values of any sign (used as max = -min(-v))
__device__ __forceinline__ double block_max_signed(double negv, double* red) {
  const int lane = threadIdx.x & 31, warp = warp_id();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) negv = fmin(negv, __shfl_xor_sync(0xffffffffu, negv, o));
  if (lane == 0) red[warp] = negv;
  __syncthreads();
  double r = (lane < NWARP) ? red[lane] : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fmin(r, __shfl_xor_sync(0xffffffffu, r, o));
  __syncthreads();
  return __shfl_sync(0xffffffffu, r, 0);
}

// packed lower triangle, COLUMN-major: column j holds rows j..m-1 contiguously, so that a
// "thread per row" sweep over a column is a unit-stride access.
// (valid while the packed size fits 31 bits; layout: packed_off in ipm_types.h)
__device__ __forceinline__ int coff(int j, int m) { return packed_off(j, m); }
__device__ __forceinline__ int cidx(int i, int j, int m) { return packed_off(j, m) + i; }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor core (DMMA).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 8-byte shared-memory load that keeps its place in the instruction stream
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Per-block working set (pointers into shared memory or the block's scratch slot).
struct Work {
  double *x, *z, *t, *d, *w;            // n each
  const double* c;                      // n: this problem's objective, read from the batch in global memory
  double *y, *b, *dy, *S, *RHS, *D;     // m each
  double* P;                            // max(m*NB, 2*TB*LDT): panel multipliers / SYRK tiles
  double* dg;                           // ldd : d gathered on the packed SYRK columns
  double* g1;                           // ldd : gather buffers of A_times2 (alias the work area)
  double* g2;
  double* tiles;                        // 2*TB*LDT : staging of the macro-tile SYRK (large problems only)
  double* fb;                           // FB_DOUBLES of shared memory: multiplier-table chunks of factor_ldl_big
  double* L;                            // m(m+1)/2 packed column-major
  double* M;                            // m*m full symmetric (global scratch)
  double* red;                          // 256: reductions [0,32) + panel scratch (ipm_factor.cuh)
  unsigned long long* prof;             // per-phase cycle counters of this block (or null)
  int ring_g;                           // chunks that have gone through the TMA ring so far (ipm_syrk.cuh)
};

// phase ids: 0 rhs/norms, 1 form M, 2 factor, 3 triangular solves, 4 residual, 5 step
// Counters accumulate in shared memory (W.red[224..240)) and are flushed once per kernel,
// so that reading the clock does not put a global round trip on the critical path.
constexpr int RED_PROF = 224;
constexpr int RED_KEEP = 208;   // [4] per-iteration scalars parked across solve_normal (ipm_kernels.cu)
// (Every thread reads the clock -- a branch on threadIdx.x == 0 here would make the code that
// follows look divergent to the compiler and to the hardware.)
__device__ __forceinline__ long long phase_begin(const Work& W) {
  return W.prof ? clock64() : 0;
}
__device__ __forceinline__ void phase_end(const Work& W, int id, long long t0, int who = 0) {
  if (W.prof) {
    const unsigned long long dt = (unsigned long long)(clock64() - t0);
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(W.red + RED_PROF) + id;
    if (threadIdx.x == who) *slot += dt;
  }
}

// ---------------------------------------------------------------------------------------
// mat-vecs with the shared matrix.  A lives in L2 (shared by all blocks); a warp takes four
// rows x two vectors (or eight columns) at a time so that many independent loads per lane are
// in flight and the L2 latency is paid once per eight dot products, which are then reduced
// together (warp_sum8).  Dense mode works on the packed operand
// (columns of A with >= 2 non-zeros): Ad (m x ldd, row-major) for A u, sy_A (ldd x ldm,
// column k contiguous) for A' u; singleton columns (slacks) are a single multiply.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_sum4(double& a0, double& a1, double& a2, double& a3) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
}
// Eight warp-wide sums with 9 shuffles instead of 40: each exchange step halves the number
// of values a lane still carries.  Returns the total of v[q], q = (lane >> 2) & 7, in every
// lane of the group of four lanes that share q.
__device__ __forceinline__ double warp_sum8(const double (&v)[8], int lane) {
  const unsigned FULL = 0xffffffffu;
  double w[4], u[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double send = h16 ? v[i] : v[4 + i];
    const double keep = h16 ? v[4 + i] : v[i];
    w[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double send = h8 ? w[i] : w[2 + i];
    const double keep = h8 ? w[2 + i] : w[i];
    u[i] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  double t = (h4 ? u[1] : u[0]) + __shfl_xor_sync(FULL, h4 ? u[0] : u[1], 4);
  t += __shfl_xor_sync(FULL, t, 2);
  t += __shfl_xor_sync(FULL, t, 1);
  return t;
}

// out = A' u (n outputs).  Ends with __syncthreads().
__device__ __forceinline__ void At_times(const Matrix& A, const double* __restrict__ u,
                                         double* __restrict__ out) {
  const int m = A.m, n = A.n;
  const int lane = threadIdx.x & 31, warp = warp_id();
  if (!A.sparse) {
    const int ldm = A.sy_ldm;
    // singleton/empty columns: fetch the table entries now, use them after the packed loop
    const int jt = threadIdx.x;
    int r0 = -2;
    double cv0 = 0.0;
    if (jt < n) { r0 = A.colrow[jt]; cv0 = A.colval[jt]; }
    const int q = (lane >> 2) & 7;
    for (int k0 = warp * 8; k0 < A.nd; k0 += NWARP * 8) {      // ldd >= nd rounded up to 16: no guards
      const double* __restrict__ p = A.sy_A + (size_t)k0 * ldm;
      double a[8];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] = 0.0;
#pragma unroll 2
      for (int i = lane; i < m; i += 32) {
        const double ui = u[i];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] += __ldcg(p + c * ldm + i) * ui;
      }
      const double t = warp_sum8(a, lane);
      if ((lane & 3) == 0 && k0 + q < A.nd) out[A.dcols[k0 + q]] = t;
    }
    if (jt < n && r0 != -2) out[jt] = (r0 >= 0) ? cv0 * u[r0] : 0.0;
    for (int j = jt + NT; j < n; j += NT) {
      const int r = A.colrow[j];
      if (r != -2) out[j] = (r >= 0) ? A.colval[j] * u[r] : 0.0;
    }
  } else {
    for (int j = threadIdx.x; j < n; j += NT) {
      double acc = 0.0;
      for (int k = A.Tp[j]; k < A.Tp[j + 1]; k++) acc += A.Tx[k] * u[A.Ti[k]];
      out[j] = acc;
    }
  }
  __syncthreads();
}

// o1 = A u1, o2 = A u2 (m outputs each), one pass over A.  g1/g2: scratch of ldd doubles
// (dense mode: u1/u2 gathered onto the packed columns).  Ends with __syncthreads().
__device__ __forceinline__ void A_times2(const Matrix& A, const double* __restrict__ u1,
                                         const double* __restrict__ u2, double* __restrict__ o1,
                                         double* __restrict__ o2, double* __restrict__ g1,
                                         double* __restrict__ g2) {
  const int m = A.m;
  const int lane = threadIdx.x & 31, warp = warp_id();
  if (!A.sparse) {
    const int ldd = A.ldd;
    for (int k = threadIdx.x; k < ldd; k += NT) {
      const bool ok = k < A.nd;
      const int j = ok ? A.dcols[k] : 0;
      g1[k] = ok ? u1[j] : 0.0;
      g2[k] = ok ? u2[j] : 0.0;
    }
    // singleton (slack) columns, one thread per row; the packed part is added below
    for (int i = threadIdx.x; i < m; i += NT) {
      double s1 = 0.0, s2 = 0.0;
      for (int e = A.sing_ptr[i]; e < A.sing_ptr[i + 1]; e++) {
        const double a = A.sing_a[e];
        const int j = A.sing_col[e];
        s1 += a * u1[j];
        s2 += a * u2[j];
      }
      o1[i] = s1;
      o2[i] = s2;
    }
    __syncthreads();
    // four rows x two vectors per warp and pass: eight dot products, one 9-shuffle reduction
    const int q = (lane >> 2) & 7;
    for (int i0 = warp * 4; i0 < m; i0 += NWARP * 4) {
      const double* __restrict__ r0 = A.Ad + (size_t)i0 * ldd;
      const double* __restrict__ r1 = A.Ad + (size_t)min(i0 + 1, m - 1) * ldd;
      const double* __restrict__ r2 = A.Ad + (size_t)min(i0 + 2, m - 1) * ldd;
      const double* __restrict__ r3 = A.Ad + (size_t)min(i0 + 3, m - 1) * ldd;
      double a[8];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] = 0.0;
#pragma unroll 2
      for (int k = lane; k < ldd; k += 32) {
        const double x = g1[k], w = g2[k];
        const double v0 = __ldcg(r0 + k), v1 = __ldcg(r1 + k), v2 = __ldcg(r2 + k), v3 = __ldcg(r3 + k);
        a[0] += v0 * x; a[1] += v0 * w;
        a[2] += v1 * x; a[3] += v1 * w;
        a[4] += v2 * x; a[5] += v2 * w;
        a[6] += v3 * x; a[7] += v3 * w;
      }
      const double t = warp_sum8(a, lane);
      const int i = i0 + (q >> 1);
      if ((lane & 3) == 0 && i < m) {
        if (q & 1) o2[i] += t;
        else o1[i] += t;
      }
    }
  } else {
    for (int i = warp; i < m; i += NWARP) {
      double a1 = 0.0, a2 = 0.0;
      for (int k = A.Ap[i] + lane; k < A.Ap[i + 1]; k += 32) {
        double a = A.Ax[k];
        int j = A.Ai[k];
        a1 += a * u1[j];
        a2 += a * u2[j];
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if (lane == 0) { o1[i] = a1; o2[i] = a2; }
    }
  }
  __syncthreads();
}

// o = A u (dense operator only), one pass over A: eight rows per warp, one 9-shuffle reduction.
// g: scratch of ldd doubles (u gathered onto the packed columns).  Ends with __syncthreads().
__device__ __forceinline__ void A_times1(const Matrix& A, const double* __restrict__ u, double* __restrict__ o,
                                         double* __restrict__ g) {
  const int m = A.m, ldd = A.ldd;
  const int lane = threadIdx.x & 31, warp = warp_id();
  for (int k = threadIdx.x; k < ldd; k += NT) g[k] = (k < A.nd) ? u[A.dcols[k]] : 0.0;
  for (int i = threadIdx.x; i < m; i += NT) {
    double s = 0.0;
    for (int e = A.sing_ptr[i]; e < A.sing_ptr[i + 1]; e++) s += A.sing_a[e] * u[A.sing_col[e]];
    o[i] = s;
  }
  __syncthreads();
  const int q = (lane >> 2) & 7;
  for (int i0 = warp * 8; i0 < m; i0 += NWARP * 8) {
    const double* __restrict__ r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = A.Ad + (size_t)min(i0 + c, m - 1) * ldd;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[c] = 0.0;
#pragma unroll 2
    for (int k = lane; k < ldd; k += 32) {
      const double v = g[k];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] += __ldcg(r[c] + k) * v;
    }
    const double t = warp_sum8(a, lane);
    if ((lane & 3) == 0 && i0 + q < m) o[i0 + q] += t;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// M = A diag(d) A'  (full symmetric m x m into W.M), returns nothing; caller syncs.
// ---------------------------------------------------------------------------------------
static __device__ __forceinline__ void form_M_dense(const Matrix& A, Work& W) {
  const int m = A.m, ldd = A.ldd;
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int g = lane >> 2, tg = lane & 3;
  const int wr = warp >> 2, wc = warp & 3;
  double* As = W.tiles;
  double* Bs = W.tiles + TB * LDT;
  for (int k = tid; k < ldd; k += NT) W.dg[k] = (k < A.nd) ? W.d[A.dcols[k]] : 0.0;
  const int nmt = (m + TB - 1) / TB;
  for (int I = 0; I < nmt; I++) {
    for (int J = 0; J <= I; J++) {
      const bool active = !(I == J && wc > wr);
      double acc[2][2][2];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b2 = 0; b2 < 2; b2++) acc[a][b2][0] = acc[a][b2][1] = 0.0;
      for (int k0 = 0; k0 < ldd; k0 += KC) {
        __syncthreads();   // previous chunk consumed (and dg written on the first pass)
        for (int e = tid; e < TB * KC; e += NT) {
          int r = e / KC, kk = e % KC;
          int gi = I * TB + r, gj = J * TB + r;
          As[r * LDT + kk] = (gi < m) ? A.Ad[(size_t)gi * ldd + k0 + kk] : 0.0;
          Bs[r * LDT + kk] = (gj < m) ? A.Ad[(size_t)gj * ldd + k0 + kk] * W.dg[k0 + kk] : 0.0;
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int ks = 0; ks < KC / 4; ks++) {
            double a0 = As[(wr * 16 + g) * LDT + ks * 4 + tg];
            double a1 = As[(wr * 16 + 8 + g) * LDT + ks * 4 + tg];
            double b0 = Bs[(wc * 16 + g) * LDT + ks * 4 + tg];
            double b1 = Bs[(wc * 16 + 8 + g) * LDT + ks * 4 + tg];
            dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
          }
        }
      }
      if (active) {
#pragma unroll
        for (int ri = 0; ri < 2; ri++)
#pragma unroll
          for (int ci = 0; ci < 2; ci++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              int i = I * TB + wr * 16 + ri * 8 + g;
              int j = J * TB + wc * 16 + ci * 8 + tg * 2 + h;
              if (i < m && j <= i) {
                double v = acc[ri][ci][h];
                W.M[(size_t)i * m + j] = v;
                W.M[(size_t)j * m + i] = v;
              }
            }
      }
    }
  }
  __syncthreads();
  // singleton (slack) columns: diagonal only
  for (int i = tid; i < m; i += NT) {
    double s = 0.0;
    for (int e = A.sing_ptr[i]; e < A.sing_ptr[i + 1]; e++) s += A.sing_w[e] * W.d[A.sing_col[e]];
    if (A.sing_ptr[i + 1] > A.sing_ptr[i]) W.M[(size_t)i * m + i] += s;
  }
}

// Sparse A: the entries of the shared pattern of A A' (lower triangle) from their (k, A_ik A_jk)
// lists, straight into the packed factor storage, which is cleared first (the factorisation of
// the previous step left its fill there).  keepM: also into the full symmetric M for the
// refinement residual (the reference's sparse path has none, ldl.cl:698-711).
template <int FQ = 4>
static __device__ __forceinline__ void form_M_sparse(const Matrix& A, Work& W, bool keepM) {
  const int m = A.m;
  {
    double2* L2 = reinterpret_cast<double2*>(W.L);
    const size_t n2 = packed_doubles(m) / 2;
    for (size_t e = threadIdx.x; e < n2; e += NT) L2[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  // FQ entries per thread and round (four in the shared-memory kernel, eight with the factor in
  // global memory), their dependent loads (list bounds -> k -> d_k) issued
  // side by side: one entry at a time is three serialised L2 round trips per entry.  The entries
  // come in the order of the packed storage (cabi.cu), so a warp writes neighbouring addresses.
  // (eight per round in the shared-memory kernel as well made the untouched SYRK of that kernel 13 %
  // slower, DESIGN.md section 6 on the code generator: its instance stays at four)
  for (int e0 = threadIdx.x; e0 < A.nme; e0 += FQ * NT) {
    int t0[FQ], t1[FQ], i[FQ], j[FQ], k[FQ];
    double w[FQ], s[FQ];
#pragma unroll
    for (int q = 0; q < FQ; q++) {
      const int e = min(e0 + q * NT, A.nme - 1);
      t0[q] = A.me_ptr[e]; t1[q] = A.me_ptr[e + 1];
      i[q] = A.me_i[e]; j[q] = A.me_j[e];
    }
#pragma unroll
    for (int q = 0; q < FQ; q++) { k[q] = A.mt_k[t0[q]]; w[q] = A.mt_w[t0[q]]; }   // every entry has >= 1 term
#pragma unroll
    for (int q = 0; q < FQ; q++) s[q] = w[q] * W.d[k[q]];
#pragma unroll
    for (int q = 0; q < FQ; q++)
      for (int t = t0[q] + 1; t < t1[q]; t++) s[q] += A.mt_w[t] * W.d[A.mt_k[t]];
#pragma unroll
    for (int q = 0; q < FQ; q++) {
      if (e0 + q * NT < A.nme) {
        W.L[cidx(i[q], j[q], m)] = s[q];
        if (keepM) {
          W.M[(size_t)i[q] * m + j[q]] = s[q];
          W.M[(size_t)j[q] * m + i[q]] = s[q];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// (modified) LDL' of W.M (lower part) -> W.L (unit lower, packed column-major), W.D
//   ldl.cl:349-376 : Dj = M_jj - sum_k D_k L_jk^2 ; c_ij = M_ij - sum_k L_ik L_jk D_k ;
//                    theta = max_i |c_ij| ; Dj = max(|Dj|, (theta/beta)^2, delta) ; L_ij = c_ij/Dj
//   plain != 0     : ldl.cl:28-55 (no clamping).
// Left-looking by panels of NB columns, one thread per row.
// ---------------------------------------------------------------------------------------
static __device__ __forceinline__ void factor_ldl(int m, Work& W, double beta, double delta, int plain) {
  const int tid = threadIdx.x;
  double* __restrict__ L = W.L;
  double* __restrict__ D = W.D;
  double* __restrict__ P = W.P;
  const double* __restrict__ M = W.M;
  for (int j0 = 0; j0 < m; j0 += NB) {
    const int nb = min(NB, m - j0);
    // P[k][jj] = L(j0+jj, k) * D[k]
    for (int e = tid; e < j0 * NB; e += NT) {
      int k = e / NB, jj = e % NB;
      P[e] = (jj < nb) ? L[cidx(j0 + jj, k, m)] * D[k] : 0.0;
    }
    __syncthreads();
    for (int i = j0 + tid; i < m; i += NT) {
      double acc[NB];
#pragma unroll
      for (int jj = 0; jj < NB; jj++) acc[jj] = (jj < nb) ? M[(size_t)i * m + j0 + jj] : 0.0;
      for (int k = 0; k < j0; k++) {
        double lik = L[cidx(i, k, m)];
        const double* pk = P + k * NB;
#pragma unroll
        for (int jj = 0; jj < NB; jj++) acc[jj] -= lik * pk[jj];
      }
#pragma unroll
      for (int jj = 0; jj < NB; jj++)
        if (jj < nb && j0 + jj <= i) L[cidx(i, j0 + jj, m)] = acc[jj];
    }
    __syncthreads();
    // eliminate inside the panel, one column at a time
    for (int jj = 0; jj < nb; jj++) {
      const int j = j0 + jj;
      const double djraw = L[cidx(j, j, m)];
      double Dj;
      if (plain) {
        Dj = djraw;
        __syncthreads();   // everyone has read L(j,j)
      } else {
        double th = 0.0;
        for (int i = j + 1 + tid; i < m; i += NT) th = fmax(th, fabs(L[cidx(i, j, m)]));
        th = block_max(th, W.red);
        double q = th / beta;
        Dj = fmax(fabs(djraw), fmax(q * q, delta));
      }
      for (int i = j + 1 + tid; i < m; i += NT) L[cidx(i, j, m)] /= Dj;
      if (tid == 0) { D[j] = Dj; L[cidx(j, j, m)] = 1.0; }
      __syncthreads();
      if (jj + 1 < nb) {
        for (int i = j + 1 + tid; i < m; i += NT) {
          double lij = L[cidx(i, j, m)];
          for (int j2 = j + 1; j2 < j0 + nb && j2 <= i; j2++)
            L[cidx(i, j2, m)] -= lij * (Dj * L[cidx(j2, j, m)]);
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace pb200
#include "ipm_factor.cuh"
#include "ipm_syrk.cuh"
#include "ipm_tiles.cuh"
namespace pb200 {

// S = RHS - M dy ; returns max |S|   (ldl.cl:577-599); M is the block's full symmetric copy
// in its L2-resident scratch slot, eight rows per warp at a time.
static __device__ __forceinline__ double residual_M(int m, Work& W, bool signed_max = false) {
  const int lane = threadIdx.x & 31, warp = warp_id();
  const int q = (lane >> 2) & 7;
  double mx = 0.0;
  for (int i0 = warp * 8; i0 < m; i0 += NWARP * 8) {
    const double* __restrict__ r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = W.M + (size_t)min(i0 + c, m - 1) * m;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[c] = 0.0;
#pragma unroll 2
    for (int j = lane; j < m; j += 32) {
      const double v = W.dy[j];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] += __ldcg(r[c] + j) * v;
    }
    const double t = warp_sum8(a, lane);
    if ((lane & 3) == 0 && i0 + q < m) {
      const double res = W.RHS[i0 + q] - t;
      W.S[i0 + q] = res;
      mx = fmax(mx, signed_max ? res : fabs(res));
    }
  }
  // (signed: _ldl.pyx:144 tests np.max(r); a maximum below zero never passes a positive
  // tolerance, so flooring it at 0 changes nothing)
  return block_max(mx, W.red);
}

// The same residual WITHOUT M:  S = RHS - A ((x/z) o (A' dy)),  M = A diag(x/z) A'.  Two passes over
// the L2-resident operand instead of writing M (m^2 doubles per Newton step into the block's
// scratch slot, whose dirty lines L2 keeps spilling to DRAM) and reading it back; and it leaves
// w = A' dy in W.w, which is exactly what the step needs next -- so the step's own pass over A
// goes away whenever the residual was computed for the final dy.  Dense operator, VS only.
static __device__ __forceinline__ double residual_free(const Matrix& A, Work& W, bool signed_max = false) {
  const int m = A.m, n = A.n, tid = threadIdx.x;
  At_times(A, W.dy, W.w);
  for (int j = tid; j < n; j += NT) W.d[j] = W.x[j] * W.w[j] / W.z[j];      // (d = x/z has done its duty in the SYRK)
  __syncthreads();
  A_times1(A, W.d, W.S, W.P);               // (gather buffer: the panel / stage area is idle; g1 aliases L here)
  double mx = 0.0;
  for (int i = tid; i < m; i += NT) {
    const double res = W.RHS[i] - W.S[i];
    W.S[i] = res;
    mx = fmax(mx, signed_max ? res : fabs(res));
  }
  return block_max(mx, W.red);
}

// factor + solve + refinement (ldl.cl:602-653); requires W.d, W.RHS set. Leaves dy.
// CL: the constants of the reference's OpenCL path (preset "cl") are compile-time here -- the
// branches of the "py" conventions cost the common path registers and issue slots otherwise.
// Returns true if it leaves w = A' dy (of the final dy) in W.w (residual_free).
template <bool LS, bool VS, bool CL>
static __device__ __forceinline__ bool solve_normal(const Matrix& A, Work& W, const Params& p, bool free_ok = false) {
  const int m = A.m, tid = threadIdx.x;
  if (!LS && A.tiles) {                     // genuinely sparse factor: tiles of the symbolic pattern only
    solve_normal_tiles(A, W, p);
    return false;
  }
  long long t0 = phase_begin(W);
  const bool refine = p.max_refine > 0;
  // dense operator with the vectors on chip, iteration in good health (caller): M is not stored, the
  // residual goes through A (residual_free)
  const bool matfree = VS && !A.sparse && free_ok;
  double* const Mst = matfree ? nullptr : W.M;
  if (A.sparse) form_M_sparse<LS ? 4 : 8>(A, W, refine);
  else if (VS)                              // operand staged by TMA, needs the shared work area
    W.ring_g = form_M_dense_tma_call(A.sy_A, A.sy_seg, A.dcols, A.sing_ptr, A.sing_col, A.sing_w, A.m, A.nd,
                                     A.ldd, A.sy_ldm, A.sy_npass, W.d, W.dg, W.P, W.red, W.L, Mst, W.prof,
                                     W.ring_g);
  else form_M_dense(A, W);                  // large problems: macro-tile SYRK, 20 KB of staging
  __syncthreads();
  phase_end(W, 1, t0);
  t0 = phase_begin(W);
  // the factorisation works in place on the packed lower triangle (the TMA SYRK and the sparse
  // formation wrote it; the macro-tile SYRK of the largest dense problems wrote M only)
  if (!A.sparse && !VS) {
    // (four loads in flight per thread: written element by element, the copy is one memory round
    // trip per element, since the compiler may not move a load of M across a store to L)
    for (int e0 = tid; e0 < m * m; e0 += 4 * NT) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int e = min(e0 + q * NT, m * m - 1);
        v[q] = W.M[e];                                   // M[j][i], e = j m + i
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int e = e0 + q * NT;
        const int j = e / m, i = e - j * m;
        if (e < m * m && i >= j) W.L[cidx(i, j, m)] = v[q];
      }
    }
  }
  for (int i = tid; i < m; i += NT) W.dy[i] = 0.0;
  __syncthreads();
  double bmax = 0.0;
  for (int i = tid; i < m; i += NT) bmax = fmax(bmax, fabs(W.L[cidx(i, i, m)]));
  const double beta = sqrt(block_max(bmax, W.red));
  phase_end(W, 6, t0);
  bool redo = true;
  const bool ahead = LS && m <= 208;                  // (L in shared memory: m <= ~202 in practice)
  if (ahead) redo = factor_ldl_ahead_call(m, W.L, W.D, W.P, W.red, W.prof, beta, p.ldl_delta, W.RHS, W.S);
  const bool big = !LS && m > SB;                     // factor in global memory: super-panel sweep
  if (big) redo = factor_ldl_big(m, W, beta, p.ldl_delta, W.RHS, W.S);
  if (redo) {
    if (ahead || big) {          // speculation failed somewhere: restore M and take the exact-capable path
      if (W.prof && tid == 0)    // (counted in a phase slot its path never times: big -> 7, ahead -> 11)
        reinterpret_cast<unsigned long long*>(W.red + RED_PROF)[big ? 7 : 11] += 1;
      if (A.sparse && !refine) {
        form_M_sparse<LS ? 4 : 8>(A, W, false);
      } else if (matfree) {      // (no stored M: form it again)
        W.ring_g = form_M_dense_tma_call(A.sy_A, A.sy_seg, A.dcols, A.sing_ptr, A.sing_col, A.sing_w, A.m, A.nd,
                                         A.ldd, A.sy_ldm, A.sy_npass, W.d, W.dg, W.P, W.red, W.L, nullptr, W.prof,
                                         W.ring_g);
      } else {
        for (int e = tid; e < m * m; e += NT) {
          const int j = e / m, i = e - j * m;
          if (i >= j) W.L[cidx(i, j, m)] = W.M[(size_t)j * m + i];
        }
      }
      __syncthreads();
    }
    factor_ldl_fast(m, W, beta, p.ldl_delta, W.RHS, W.S);   // also S <- (L D)^-1 RHS
  }
  phase_end(W, 2, t0);
  t0 = phase_begin(W);
  back_solve_fast<!LS>(m, W);
  phase_end(W, 3, t0);
  if (!refine) return false;                          // (nobody would look at the residual)
  const bool pymode = !CL && p.refine_mode != 0;      // _ldl.pyx:144-148: signed test, dy -= correction
  t0 = phase_begin(W);
  double maxr = matfree ? residual_free(A, W, pymode) : residual_M(m, W, pymode);
  phase_end(W, 4, t0);
  int nref = 0;
  while (maxr > p.refine_tol && nref < p.max_refine) {
    t0 = phase_begin(W);
    fwd_solve_fast(m, W);
    back_solve_fast<!LS>(m, W, pymode ? -1.0 : 1.0);
    phase_end(W, 3, t0);
    t0 = phase_begin(W);
    maxr = matfree ? residual_free(A, W, pymode) : residual_M(m, W, pymode);
    phase_end(W, 4, t0);
    nref++;
  }
  return matfree;
}

// Given x, z, y (and mu): v = A'y -> W.w ; t, d ; q -> W.w ; RHS ; also rho/sigma norms.
// Returns through refs. After this W.t holds t = c - A'y + mu/x (the ONE evaluation).
// have_v: W.t already holds v = A'y, carried over from the previous step as v + theta A'dy
// (one pass over A less per iteration; an SM only gets 20-35 B/cycle from L2).
template <bool VS>
static __device__ __forceinline__ void prepare_rhs(const Matrix& A, Work& W, double mu, double& normr, double& norms,
                                                   bool have_v = false) {
  const int m = A.m, n = A.n, tid = threadIdx.x;
  const double c_first = (tid < n) ? W.c[tid] : 0.0;   // c comes straight from the batch (global): fetch early
  if (!have_v) At_times(A, W.y, W.w);
  double ss = 0.0;
  for (int j = tid; j < n; j += NT) {
    double v = have_v ? W.t[j] : W.w[j], xj = W.x[j], zj = W.z[j], cj = (j == tid) ? c_first : W.c[j];
    double sig = cj - v + zj;
    ss += sig * sig;
    double tj = cj - v + mu / xj;
    W.t[j] = tj;
    W.d[j] = xj / zj;
    W.w[j] = xj * tj / zj;                     // q_j
  }
  norms = sqrt(block_sum(ss, W.red));          // (syncs: t, d, q visible)
  // S <- A x ; RHS <- A q
  A_times2(A, W.x, W.w, W.S, W.RHS, W.g1, W.g2);
  double rr = 0.0;
  for (int i = tid; i < m; i += NT) {
    double rho = W.b[i] - W.S[i];
    rr += rho * rho;
    W.RHS[i] = W.RHS[i] - rho;                 // -(b - Ax - Aq)
  }
  normr = sqrt(block_sum(rr, W.red));
}

// dx, dz, ratio test, update (primal_normal.cl:122-156) using the stored t.  Leaves
// v = A'y of the UPDATED y in W.t: v_new = (c - t + mu/x) + theta A'dy.
// have_w: W.w already holds A' dy (left there by residual_free).
template <bool VS, bool CL>
static __device__ __forceinline__ void step(const Matrix& A, Work& W, double mu, const Params& p, bool have_w = false) {
  const double r = p.r;
  const bool dz1 = !CL && p.dz_mode != 0;
  const bool floor0 = CL || p.theta_floor != 0;
  const int m = A.m, n = A.n, tid = threadIdx.x;
  const double c_first = (tid < n) ? W.c[tid] : 0.0;
  if (!have_w) At_times(A, W.dy, W.w);
  double th = floor0 ? 0.0 : -INFINITY;          // primal_normal.cl:134 / normal_eqns.py:92
  double big = 0.0;                              // max_j mu / x_j: the cancellation in the carried v
  for (int j = tid; j < n; j += NT) {
    double xj = W.x[j], zj = W.z[j];
    double dx = (W.t[j] - W.w[j]) * xj / zj;
    double dz = dz1 ? (mu - xj * zj - zj * dx) / xj : (mu - zj * dx) / xj - zj;
    th = fmax(th, fmax(-dz / zj, -dx / xj));
    big = fmax(big, mu / xj);
    W.d[j] = dx;                                 // (d = x/z is dead by now; t and w are still needed)
  }
  // v_new = (c - t + mu/x) + theta A'dy below recovers v = A'y from t = c - v + mu/x with an error
  // of eps max_j(mu/x_j); on the central path mu/x_j ~ z_j and that is rounding level, far from
  // it (infeasible LPs) it is not: the caller stops carrying v then
  if (floor0) th = block_max2(th, big, W.red);
  else { th = -block_max_signed(-th, W.red); big = block_max(big, W.red); }
  if (tid == 0) W.red[RED_KEEP + 3] = big;
  const double theta = fmin(r / th, 1.0);
  for (int j = tid; j < n; j += NT) {
    const double xj = W.x[j], zj = W.z[j], dx = W.d[j];
    const double dz = dz1 ? (mu - xj * zj - zj * dx) / xj : (mu - zj * dx) / xj - zj;   // same expression, same operands as above
    const double cj = (j == tid) ? c_first : W.c[j];
    W.t[j] = (cj - W.t[j] + mu / xj) + theta * W.w[j];
    W.z[j] = zj + theta * dz;
    W.x[j] = xj + theta * dx;
  }
  for (int i = tid; i < m; i += NT) W.y[i] += theta * W.dy[i];
  __syncthreads();
}

}  // namespace pb200
