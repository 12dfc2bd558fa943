// ipm_types.h -- plain structs and tile constants shared by host (cabi.cu) and device code.
#pragma once
#include <stddef.h>
#include <vector_types.h>

namespace pb200 {

constexpr int NT = 512;          // threads per block
constexpr int NWARP = NT / 32;   // 16 warps
constexpr int NB = 8;            // LDL' panel width
constexpr int TB = 64;           // SYRK macro tile (TB x TB outputs per pass)
constexpr int KC = 16;           // SYRK k-chunk staged in shared memory
constexpr int LDT = KC + 4;      // padded leading dimension of a staged tile (conflict-free)
constexpr int SY_KC = 8;         // TMA-staged SYRK: packed columns per chunk
constexpr int SY_STAGES = 3;     //   chunks in flight (ring of shared-memory stages)
constexpr int SY_SEG = 3;        //   segments per warp per pass
constexpr int SY_CW = 4;         //   8x8 tiles per segment
constexpr int FB_DOUBLES = 2 * 32 * 68 + 16 * 8 * 64 + 64 * 68 + 64 + 32;   // shared memory of factor_ldl_big: two T chunks, 16 A rings, row-solve tables

struct Params {
  double eps, delta, r, ldl_delta, refine_tol;
  int max_iter, max_refine;
  // (see pycllp_b200_params in include/pycllp_b200.h)
  int nan_guard;     // NaN in dy => status 3 (normal_eqns.py:85-87)
  int carry_v;       // iterations during which v = A'y is carried over instead of recomputed (0: never)
  int mu_mode;       // 0: mu = delta gamma / (n+m) (primal_normal.cl:272) ; 1: / n (normal_eqns.py:65)
  int refine_mode;   // 0: while max|r| > tol: dy += solve(r) (ldl.cl:645-652) ; 1: while max r > tol: dy -= solve(r) (_ldl.pyx:144-148)
  int theta_floor;   // 1: theta = max(0, ...) (primal_normal.cl:134) ; 0: no floor (normal_eqns.py:92)
  int dz_mode;       // 0: dz = (mu - z dx)/x - z (primal_normal.cl:143) ; 1: (mu - x z - z dx)/x (normal_eqns.py:90)
  double warm_floor; // warm start: x0, z0 <- max(., warm_floor) (0: the raw previous point, primal_normal.cl:213-219)
};

// The shared constraint matrix and everything precomputed from it at setup.
struct Matrix {
  int m, n, sparse;
  int big;                 // the factor lives in global memory (factor_ldl_big needs FB_DOUBLES of work area)
  // dense operator (row-major m x n)
  const double* A;
  // SYRK operand: the nd columns of A with >= 2 non-zeros, packed m x ldd row-major
  int nd, ldd;
  const double* Ad;
  const int* dcols;        // [ldd] original column of packed column k (padding -> 0)
  // TMA-staged SYRK (ipm_syrk.cuh): operand pre-tiled in chunks of SY_KC columns, k-major,
  // rows padded to sy_ldm; per-(pass, warp, slot) tile segments {tile row, first tile col, count}
  const double* sy_A;
  int sy_ldm, sy_npass;
  const int4* sy_seg;
  // singleton columns (exactly one non-zero a in row i): M_ii += a^2 d_k ; CSR by row
  const int* sing_ptr;     // [m+1]
  const int* sing_col;     // original column index
  const double* sing_w;    // a^2
  const double* sing_a;    // a
  // per column j: colrow[j] = -2 general (>= 2 non-zeros), -1 empty, else the row of its only
  // non-zero colval[j]
  const int* colrow;
  const double* colval;
  // sparse operator: CSR of A and of A'
  const int *Ap, *Ai;
  const double* Ax;
  const int *Tp, *Ti;
  const double* Tx;
  // sparse M formation: lower-triangle entries e=(i,j) of the pattern of A A' and the
  // list of (k, w = A_ik A_jk) contributing to each
  int nme;
  const int *me_ptr, *me_i, *me_j, *mt_k;
  const double* mt_w;
  // tile-sparse numeric factor (ipm_tiles.cuh): L stored as the 8x8 tiles of the symbolic block
  // fill pattern, memory ~ nnz(L).  Block column J owns tiles tl_colptr[J] .. tl_colptr[J+1]-1
  // (diagonal tile first, then block rows ascending); element (r, c) of tile t lives at
  // 64 t + 8 c + r.  Target tile t receives  - L(tl_upda[p]) D L(tl_updb[p])'  for
  // p in [tl_updptr[t], tl_updptr[t+1]).  me_pos[e]: where entry e of the pattern of A A' goes.
  int tiles, nbk, ntiles;
  // optional symmetric reordering of the constraints (rows of A) chosen at setup to reduce the fill of
  // L: internal row i is the caller's row rperm[i] (b is read, y written through it); null = identity
  const int* rperm;
  const int *tl_colptr, *tl_row, *tl_col, *tl_updptr, *tl_upda, *tl_updb, *tl_updk, *me_pos;   // tl_updk[p]: block column K of pair p
};

// doubles of the shared work area W.P: two panel-multiplier tables + split-K partials
// (ipm_factor.cuh), the TMA stages of the SYRK, or the two gather buffers of A_times2
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t work_area(const Matrix& A) {
  size_t psz = (size_t)2 * A.m * NB + 512;
  const size_t st = (size_t)SY_STAGES * SY_KC * A.sy_ldm;
  const size_t g = (size_t)2 * (A.ldd > 0 ? A.ldd : 1);
  const size_t tb = (size_t)2 * 12 * (A.m > 16 ? A.m - 8 : 8) + 512;   // two tables of factor_ldl_ahead
  if (psz < st) psz = st;
  if (psz < g) psz = g;
  if (psz < tb) psz = tb;
  if (A.big && psz < (size_t)FB_DOUBLES) psz = FB_DOUBLES;   // factor_ldl_big (ipm_factor.cuh)
  return psz;
}
// doubles of the vector W.w, which also serves as the gather of d on the packed SYRK columns
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t w_doubles(const Matrix& A) { return (size_t)(A.ldd > A.n ? A.ldd : A.n); }

// Packed lower triangle, COLUMN-major with every column starting at an EVEN offset (so that
// 16-byte shared-memory loads of two consecutive rows are always aligned): L(i, j) lives at
// packed_off(j, m) + i.  Consecutive columns are m-1-j apart, rounded up to even.
#ifdef __CUDACC__
__host__ __device__
#endif
inline int packed_off(int j, int m) { return j * (m - 1) - ((j * (j - 1)) >> 1) + ((j + 1 - (m & 1)) >> 1); }
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t packed_doubles(int m) { return (size_t)m * (m + 1) / 2 + m / 2 + 16; }   // + slack: 16-row loads may run past the last column

struct Batch {
  int N;
  const double *b, *c;     // (N, m), (N, n)
  double *x, *y, *z;       // (N, n), (N, m), (N, n)   (may be null)
  int *status, *iters;     // (N)                      (may be null)
  // leading dimensions of the outputs (0 = contiguous: n, m, n doubles; 1, 1 ints): a packed
  // per-problem record [x | y | z | status, iters] is written in place for the multi-GPU gather
  size_t ld_x, ld_y, ld_z;
  int ld_s;
  size_t ld_0;             // leading dimension of x0, z0, y0 (0 = contiguous)
  // warm start: begin from x0, z0 (N,n), y0 (N,m) instead of x = z = y = 1 (they may alias x, z, y)
  int warm;
  // optional (N, trace_iters, 3): |rho|, |sigma|, gamma of every iteration (primal_normal.cl:250-252)
  double* trace;
  int trace_iters;
  // hook mode (one normal-equations solve on given state): x0,z0 (N,n), y0 (N,m) in, dy out
  int hook;
  const double *x0, *z0, *y0;
  double* dy_out;
  double mu;
};

struct Scratch {
  double* base;       // per-block slots
  size_t slot;        // doubles per slot
  size_t off_L;       // offset of L inside the slot (if not in smem)
  size_t off_vec;     // offset of the vectors inside the slot (if not in smem)
  size_t off_P;       // offset of the panel work area and of dg (if the vectors are not in smem)
  size_t off_dg;
  int* counter;       // work counter
  int L_in_smem, vec_in_smem;
  int small;          // two blocks per SM: launch the 64-register build of the kernel (ipm_kernels_small.cu)
  unsigned long long* prof;   // optional [grid][8] per-phase cycle counters (null = off)
};

}  // namespace pb200
