// ipm_small.cuh -- the whole interior-point iteration for SMALL dense problems (m <= 64, the
// reference's own example size: m = 50, n = 100, examples/random_problem.py), one WARP GROUP
// of 128 threads per LP and several LPs per SM.
//
// The 512-thread kernel (ipm_solve.cuh) spends 35 us per Newton step at m = 50: every phase is a
// chain of short latency-bound steps (named-barrier pipelines, TMA stages, L2 round trips for
// A) sized for m = 200 .. 2000.  At this size everything fits in ~50 KB of shared memory --
// the packed operand of A included -- so this kernel keeps A, M, L and all vectors on chip,
// runs four to seven blocks per SM, and uses the simplest exact form of every step:
//   v = A'y, A x, A q      from the shared-memory copy of A (no L2 traffic inside the iteration)
//   M = A diag(x/z) A'     DMMA m8n8k4, one 8x8 tile of the lower triangle per warp and pass
//   modified LDL'          right-looking by 8-column panels: the sequential rule of ldl.cl:349-376
//                          as is (theta_j by an exact warp reduction, no speculation) with every
//                          row's panel entries in registers, DMMA trailing update; the right-hand
//                          side rides along as row m, so the forward solve (ldl.cl:519-527) is a
//                          by-product
//   back substitution      one warp, two rows per lane, the pivot broadcast by a shuffle
//   refinement             ldl.cl:645-652 on a packed copy of M
// Same constants, stop rule and status codes as ipm_solve_one (primal_normal.cl:201-284);
// v = A'y is recomputed every iteration like the reference does (it is free here).
// Preset "cl" only; hook launches and every other shape use the 512-thread kernel.
#pragma once
#include "ipm_device.cuh"

namespace pb200 {

constexpr int SNT = 128;          // threads per block
constexpr int SNW = SNT / 32;     // 4 warps
constexpr int SMALL_MAX_M = 63;    // rows 0 .. m (the right-hand side rides along as row m) on two warps
constexpr int SRED = 32;          // doubles of reduction scratch

// leading dimension of the shared-memory copy of the packed operand: a multiple of 4 (k-steps),
// = 4 (mod 16) so that the 8 rows x 4 columns of a DMMA fragment fall into disjoint banks
#ifdef __CUDACC__
__host__ __device__
#endif
inline int small_lda(int nd) {
  int l = 4;
  while (l < nd) l += 16;
  return l;
}
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t small_smem_doubles(int m, int n, int nd) {
  auto al = [](size_t v) { return (v + 1) & ~(size_t)1; };
  const int T = (m + 7) / 8, lda = small_lda(nd);
  size_t o = SRED;
  o += (size_t)8 * T * lda;                 // As
  o += 2 * al(packed_doubles(m));           // Mp, L (each 16-byte aligned: double2 loads)
  o += al(m + 1);                           // colbuf
  o += 6 * al(n);                           // x z t d w c
  o += 6 * al(m + 1);                       // y b dy S RHS D
  o += 3 * (size_t)lda;                     // dg g1 g2
  o += al(n) + al((n + 1) / 2);             // colval, colrow (ints)
  o += 2 * al((m + 2) / 2);                 // column offsets, singleton row pointers (ints)
  return o;
}

struct SmallWork {
  double *red, *As, *Mp, *L, *colbuf;
  double *x, *z, *t, *d, *w, *c;
  double *y, *b, *dy, *S, *RHS, *D;
  double *dg, *g1, *g2, *colval;
  int *colrow, *offs, *sptr;     // sptr: sing_ptr (CSR by row of the singleton columns)
  int lda, T;
  unsigned long long* prof;      // optional per-block phase counters (thread 0)
  bool exact;                    // test switch: every panel by the sequential rule (Scratch::small == 4)
};

__device__ __forceinline__ long long s_t0(const SmallWork& W) { return W.prof ? clock64() : 0; }
__device__ __forceinline__ void s_t1(const SmallWork& W, int id, long long t0) {
  if (W.prof && threadIdx.x == 0) W.prof[id] += (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ void s_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

// block reductions over 4 warps; every thread gets the result
__device__ __forceinline__ double s_block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  s_bar();
  const double r = (red[0] + red[1]) + (red[2] + red[3]);
  s_bar();
  return r;
}
__device__ __forceinline__ double s_block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) red[warp] = v;
  s_bar();
  const double r = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
  s_bar();
  return r;
}
// one barrier only: the caller guarantees another barrier before `red + off` is written again
__device__ __forceinline__ double s_block_max1(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) red[8 + warp] = v;
  s_bar();
  return fmax(fmax(red[8], red[9]), fmax(red[10], red[11]));
}

// out[j] = (A' u)_j for all n columns (primal_normal.cl:76-94 inner sum)
static __device__ __forceinline__ void s_At_times(const Matrix& A, const SmallWork& W, const double* u, double* out) {
  const int m = A.m, n = A.n, tid = threadIdx.x;
  for (int k = tid; k < A.nd; k += SNT) {
    const double* a = W.As + k;
    double s0 = 0.0, s1 = 0.0;
    int i = 0;
    for (; i + 1 < m; i += 2) {
      s0 = fma(a[(size_t)i * W.lda], u[i], s0);
      s1 = fma(a[(size_t)(i + 1) * W.lda], u[i + 1], s1);
    }
    if (i < m) s0 = fma(a[(size_t)i * W.lda], u[i], s0);
    out[__ldg(A.dcols + k)] = s0 + s1;
  }
  for (int j = tid; j < n; j += SNT) {
    const int cr = W.colrow[j];
    if (cr >= 0) out[j] = W.colval[j] * u[cr];
    else if (cr == -1) out[j] = 0.0;
  }
  s_bar();
}

// o1 = A u1, o2 = A u2 in one pass over the shared-memory operand: a row is split over the four
// lanes of a quad (lane kq takes the columns k = kq mod 4 -- the conflict-free access pattern of
// the DMMA fragments), eight rows per warp and round, two shuffles per sum
static __device__ __forceinline__ void s_A_times2(const Matrix& A, const SmallWork& W, const double* u1, const double* u2,
                                                  double* o1, double* o2) {
  const unsigned FULL = 0xffffffffu;
  const int m = A.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, kq = lane & 3;
  for (int k = tid; k < W.lda; k += SNT) {
    const int j = (k < A.nd) ? __ldg(A.dcols + k) : 0;
    W.g1[k] = (k < A.nd) ? u1[j] : 0.0;
    W.g2[k] = (k < A.nd) ? u2[j] : 0.0;
  }
  s_bar();
  for (int i0 = 8 * warp; i0 < m; i0 += 8 * SNW) {
    const int i = i0 + g;                                   // (rows >= m of As are zero)
    const double* a = W.As + (size_t)i * W.lda + kq;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < W.lda; k += 4) {
      const double av = a[k];
      s1 = fma(av, W.g1[k + kq], s1);
      s2 = fma(av, W.g2[k + kq], s2);
    }
    // singleton columns of this row (slacks): the row's list split over the quad
    if (i < m) {
      for (int p = W.sptr[i] + kq; p < W.sptr[i + 1]; p += 4) {
        const double av = __ldg(A.sing_a + p);
        const int j = __ldg(A.sing_col + p);
        s1 = fma(av, u1[j], s1);
        s2 = fma(av, u2[j], s2);
      }
    }
    s1 += __shfl_xor_sync(FULL, s1, 1);
    s2 += __shfl_xor_sync(FULL, s2, 1);
    s1 += __shfl_xor_sync(FULL, s1, 2);
    s2 += __shfl_xor_sync(FULL, s2, 2);
    if (kq == 0 && i < m) { o1[i] = s1; o2[i] = s2; }
  }
  s_bar();
}

// M = A diag(d) A' -> L (packed lower, the factorisation works in place) and Mp (kept for the residual)
static __device__ __forceinline__ void s_form_M(const Matrix& A, const SmallWork& W) {
  const int m = A.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  for (int k = tid; k < W.lda; k += SNT) W.dg[k] = (k < A.nd) ? W.d[__ldg(A.dcols + k)] : 0.0;
  s_bar();
  const int ntile = W.T * (W.T + 1) / 2, nks = W.lda >> 2;
  for (int t = warp; t < ntile; t += SNW) {
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= t) I++;
    const int J = t - I * (I + 1) / 2;
    const double* pa = W.As + (size_t)(8 * I + g) * W.lda + tg;
    const double* pb = W.As + (size_t)(8 * J + g) * W.lda + tg;
    const double* pd = W.dg + tg;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    int ks = 0;
    for (; ks + 1 < nks; ks += 2) {
      const double a1 = pa[4 * ks], b1 = pb[4 * ks] * pd[4 * ks];
      const double a2 = pa[4 * ks + 4], b2 = pb[4 * ks + 4] * pd[4 * ks + 4];
      dmma884(c0, c1, a1, b1);
      dmma884(e0, e1, a2, b2);
    }
    if (ks < nks) dmma884(c0, c1, pa[4 * ks], pb[4 * ks] * pd[4 * ks]);
    const int r = 8 * I + g, q0 = 8 * J + 2 * tg;
    if (r < m) {
      if (q0 <= r) { const double v = c0 + e0; W.L[W.offs[q0] + r] = v; W.Mp[W.offs[q0] + r] = v; }
      if (q0 + 1 <= r) { const double v = c1 + e1; W.L[W.offs[q0 + 1] + r] = v; W.Mp[W.offs[q0 + 1] + r] = v; }
    }
  }
  s_bar();
  // singleton columns add a^2 d to the diagonal only
  for (int i = tid; i < m; i += SNT) {
    const int p0 = W.sptr[i], p1 = W.sptr[i + 1];
    double s = 0.0;
    for (int p = p0; p < p1; p++) s = fma(__ldg(A.sing_w + p), W.d[__ldg(A.sing_col + p)], s);
    if (p1 > p0) {
      const double v = W.L[W.offs[i] + i] + s;
      W.L[W.offs[i] + i] = v;
      W.Mp[W.offs[i] + i] = v;
    }
  }
  s_bar();
}

// exact maximum of NON-NEGATIVE doubles over the warp with two integer reductions (their bit
// patterns order like unsigned integers)
__device__ __forceinline__ double warp_max_pos(double v) {
  const unsigned FULL = 0xffffffffu;
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(FULL, hi);
  const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// Modified LDL' in place (ldl.cl:349-376), right-looking by panels of eight columns; row m = the
// right-hand side kept in S: on exit S = (L D)^-1 RHS.
//   panel: one thread per row (rows j0 .. m), the row's eight entries in registers.  Speculative
//          pass without any communication: every lane eliminates the 8x8 diagonal block itself
//          (D_j = max(|d_j|, delta)) and solves its own row against it; theta_j is accumulated by
//          integer warp reductions and checked once (one 64-thread barrier.red when the rows span
//          two warps).  Should (theta_j/beta)^2 exceed a D_j -- it cannot for a positive
//          semi-definite M, up to rounding -- the panel is redone from memory by the sequential
//          rule, column by column with a block-wide theta_j;
//   trailing matrix: 8x8 tiles of the lower triangle -= L(I, panel) D L(J, panel)', two DMMAs each.
static __device__ __forceinline__ void s_factor(int m, const SmallWork& W, double beta, double delta, bool exact) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nthr = (m + 1 + 31) & ~31;             // threads that own a row (32 or 64)
  const int np = (m + 7) >> 3;
  const double inv_beta = 1.0 / beta;
  double* __restrict__ L = W.L;
  double* wsm = W.red + 16;                         // [2][8]
  double* rsm = W.red + 8;                          // [2][2]
  for (int p = 0; p < np; p++) {
    const int j0 = 8 * p, nb = min(8, m - j0);
    long long tq = s_t0(W);
    if (tid < nthr) {
      const int r = tid;
      const bool live = r >= j0 && r <= m;
      const int jlim = (r < m) ? r - j0 : 7;        // last column of the panel this row has (rhs row: all)
      const bool spec = !exact && nb == 8;          // (a ragged last panel goes by the sequential rule)
      double c[8];
      if (!spec) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          c[jj] = 0.0;
          if (live && jj < nb && jj <= jlim) c[jj] = (r < m) ? L[W.offs[j0 + jj] + r] : W.S[j0 + jj];
        }
      }
      // ---- speculative pass: no communication at all.  Every lane eliminates the 8x8 diagonal
      // block itself (36 entries in registers, D_j = max(|d_j|, delta): the theta term of
      // ldl.cl:368 assumed inactive, as it is for a positive semi-definite M), then solves its own
      // row against the block; theta_j is only accumulated and checked at the end.
      bool bad = false;
      if (spec) {
        double e[8][8], rr[8];
        int ob[8];
#pragma unroll
        for (int j = 0; j < 8; j++) ob[j] = W.offs[j0 + j] + j0;      // (row j0 of column j0 + j; even)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (j & 1) e[j][j] = L[ob[j] + j];
#pragma unroll
          for (int i = j + (j & 1); i < 8; i += 2) {                    // 16-byte aligned pairs
            const double2 v = *reinterpret_cast<const double2*>(L + ob[j] + i);
            e[i][j] = v.x;
            e[i + 1][j] = v.y;
          }
        }
        // this thread's row (a row above the panel reads whatever lies there and never uses it)
        const int ri = r - j0;                      // index inside the panel (rhs row: > 7)
#pragma unroll
        for (int jj = 0; jj < 8; jj++) c[jj] = (r < m) ? L[ob[jj] + ri] : W.S[j0 + jj];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          rr[j] = rcp_pos64(fmax(fabs(e[j][j]), delta));
#pragma unroll
          for (int i = j + 1; i < 8; i++) {
            const double l = e[i][j] * rr[j];
#pragma unroll
            for (int k = j + 1; k <= i; k++) e[i][k] = fma(-l, e[k][j], e[i][k]);
          }
        }
        const bool below = live && r < m;           // rows that count for theta (not the rhs row)
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double Dk = fmax(fabs(e[k][k]), delta);
          // (theta_k/beta)^2 > D_k would activate the clamp of ldl.cl:368; theta_k is a maximum over
          // the rows below the pivot, so it exceeds the threshold iff some row does: no reduction
          const double q = fabs(c[k]) * inv_beta;
          bad |= below && ri > k && !(q * q <= Dk);
          if (live && ri > k) {
            const double lk = c[k] * rr[k];
            c[k] = lk;
#pragma unroll
            for (int j2 = k + 1; j2 < 8; j2++)
              if (j2 <= jlim) c[j2] = fma(-lk, e[j2][k], c[j2]);
          } else if (ri == k) {
            c[k] = 1.0;
            W.D[j0 + k] = Dk;
          }
        }
        // any warp's partial maximum over the threshold <=> the block maximum is
        if (nthr > 32) {
          unsigned any;
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, 1, 64, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(any) : "r"((unsigned)bad) : "memory");
          bad = any != 0;
        } else {
          bad = __any_sync(0xffffffffu, bad);
        }
        if (bad) {                                  // nothing has been stored yet: reload and redo exactly
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            c[jj] = 0.0;
            if (live && jj < nb && jj <= jlim) c[jj] = (r < m) ? L[W.offs[j0 + jj] + r] : W.S[j0 + jj];
          }
        }
      }
      if (!spec || bad) {
        // ---- the sequential rule, one column at a time with a block-wide theta ----
#pragma unroll 1
        for (int jj = 0; jj < nb; jj++) {
          double* wb = wsm + 8 * (jj & 1);
          double* rb = rsm + 2 * (jj & 1);
          double cj = c[0];
#pragma unroll
          for (int q2 = 1; q2 < 8; q2++) cj = (jj == q2) ? c[q2] : cj;
          if (r >= j0 && r < j0 + 8) wb[r - j0] = cj;      // unscaled c_{r j} of the diagonal block
          double th = (live && r > j0 + jj && r < m) ? fabs(cj) : 0.0;
          th = warp_max_pos(th);
          if (nthr > 32) {
            if (lane == 0) rb[warp] = th;
            asm volatile("barrier.sync 1, 64;" ::: "memory");
            th = fmax(rb[0], rb[1]);
          } else {
            __syncwarp();
          }
          const double q = th * inv_beta;
          const double Dj = fmax(fabs(wb[jj]), fmax(q * q, delta));   // ldl.cl:368
          if (live && r > j0 + jj) {
            const double l = cj / Dj;
#pragma unroll
            for (int j2 = 0; j2 < 8; j2++) {
              if (j2 == jj) c[j2] = l;
              else if (j2 > jj && j2 <= jlim && j2 < nb) c[j2] = fma(-l, wb[j2], c[j2]);
            }
          } else if (r == j0 + jj) {
#pragma unroll
            for (int j2 = 0; j2 < 8; j2++)
              if (j2 == jj) c[j2] = 1.0;
            W.D[j0 + jj] = Dj;
          }
        }
      }
      if (live) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++)
          if (jj < nb && jj <= jlim) {
            if (r < m) L[W.offs[j0 + jj] + r] = c[jj];
            else W.S[j0 + jj] = c[jj];
          }
      }
    }
    s_bar();
    s_t1(W, 9, tq);                                 // (profile slots: f_diag = panels, f_old = trailing updates)
    tq = s_t0(W);
    if (j0 + 8 < m) {
      const int Tt = W.T - p - 1, ntile = Tt * (Tt + 1) / 2;
      const int oa1 = W.offs[j0 + tg], oa2 = W.offs[j0 + 4 + tg];
      const double d1 = W.D[j0 + tg], d2 = W.D[j0 + 4 + tg];
      for (int t = warp; t < ntile; t += SNW) {
        int Ii = 0;
        while ((Ii + 1) * (Ii + 2) / 2 <= t) Ii++;
        const int I = p + 1 + Ii, J = p + 1 + (t - Ii * (Ii + 1) / 2);
        const double a1 = L[oa1 + 8 * I + g], a2 = L[oa2 + 8 * I + g];
        const double b1 = L[oa1 + 8 * J + g] * d1, b2 = L[oa2 + 8 * J + g] * d2;
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, a1, b1);
        dmma884(c0, c1, a2, b2);
        const int r = 8 * I + g, q0 = 8 * J + 2 * tg;
        if (r < m) {
          if (q0 <= r) L[W.offs[q0] + r] -= c0;
          if (q0 + 1 <= r) L[W.offs[q0 + 1] + r] -= c1;
        }
      }
      // the right-hand-side row: S_k -= sum_j (S_j D_j) L(k, j) over the panel (S_j is already scaled)
      for (int k = j0 + 8 + tid; k < m; k += SNT) {
        double acc = 0.0;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) acc = fma(W.S[j0 + jj] * W.D[j0 + jj], L[W.offs[j0 + jj] + k], acc);
        W.S[k] -= acc;
      }
      s_bar();
    }
    s_t1(W, 13, tq);
  }
}

// warp 0: S <- L^-T S (second half of ldl.cl:529-536); two rows per lane (m <= 64)
static __device__ __forceinline__ void s_back_warp(int m, const SmallWork& W) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int r0 = lane, r1 = lane + 32;
  double s0 = (r0 < m) ? W.S[r0] : 0.0, s1 = (r1 < m) ? W.S[r1] : 0.0;
  const int o0 = (r0 < m) ? W.offs[r0] : 0, o1 = (r1 < m) ? W.offs[r1] : 0;
  for (int j = m - 1; j >= 1; j--) {
    const double sj = __shfl_sync(FULL, (j >= 32) ? s1 : s0, j & 31);
    if (r0 < j) s0 = fma(-W.L[o0 + j], sj, s0);               // L(j, r0)
    if (r1 < j) s1 = fma(-W.L[o1 + j], sj, s1);
  }
  if (r0 < m) W.S[r0] = s0;
  if (r1 < m) W.S[r1] = s1;
}
// warp 0: S <- (L D)^-1 S (first half, refinement passes only)
static __device__ __forceinline__ void s_fwd_warp(int m, const SmallWork& W) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int r0 = lane, r1 = lane + 32;
  double s0 = (r0 < m) ? W.S[r0] : 0.0, s1 = (r1 < m) ? W.S[r1] : 0.0;
  for (int j = 0; j + 1 < m; j++) {
    const double sj = __shfl_sync(FULL, (j >= 32) ? s1 : s0, j & 31);
    const int oj = W.offs[j];
    if (r0 > j && r0 < m) s0 = fma(-W.L[oj + r0], sj, s0);
    if (r1 > j && r1 < m) s1 = fma(-W.L[oj + r1], sj, s1);
  }
  if (r0 < m) W.S[r0] = s0 / W.D[r0];
  if (r1 < m) W.S[r1] = s1 / W.D[r1];
}

// S <- RHS - M dy ; returns max |S| (ldl.cl:577-599)
static __device__ __forceinline__ double s_residual(int m, const SmallWork& W) {
  double mx = 0.0;
  for (int i = threadIdx.x; i < m; i += SNT) {
    double acc = 0.0;
    const int oi = W.offs[i];
    for (int j = 0; j <= i; j++) acc = fma(W.Mp[W.offs[j] + i], W.dy[j], acc);
    for (int j = i + 1; j < m; j++) acc = fma(W.Mp[oi + j], W.dy[j], acc);
    const double r = W.RHS[i] - acc;
    W.S[i] = r;
    mx = fmax(mx, fabs(r));
  }
  return s_block_max(mx, W.red);
}

// factor + solve + refinement (ldl.cl:602-653); needs W.d and W.RHS; leaves dy
static __device__ __forceinline__ void s_solve_normal(const Matrix& A, const SmallWork& W, const Params& p) {
  const int m = A.m, tid = threadIdx.x;
  long long t0 = s_t0(W);
  s_form_M(A, W);
  s_t1(W, 1, t0);
  t0 = s_t0(W);
  double bmax = 0.0;
  for (int i = tid; i < m; i += SNT) {
    bmax = fmax(bmax, fabs(W.L[W.offs[i] + i]));
    W.dy[i] = 0.0;
    W.S[i] = W.RHS[i];
  }
  const double beta = sqrt(s_block_max(bmax, W.red));
  s_factor(m, W, beta, p.ldl_delta, W.exact);
  s_t1(W, 2, t0);
  t0 = s_t0(W);
  if (tid < 32) s_back_warp(m, W);
  s_bar();
  for (int i = tid; i < m; i += SNT) W.dy[i] += W.S[i];
  s_bar();
  s_t1(W, 3, t0);
  if (p.max_refine <= 0) return;
  t0 = s_t0(W);
  double maxr = s_residual(m, W);
  s_t1(W, 4, t0);
  int nref = 0;
  while (maxr > p.refine_tol && nref < p.max_refine) {
    if (tid < 32) { s_fwd_warp(m, W); __syncwarp(); s_back_warp(m, W); }
    s_bar();
    for (int i = tid; i < m; i += SNT) W.dy[i] += W.S[i];
    s_bar();
    maxr = s_residual(m, W);
    nref++;
  }
}

static __device__ __forceinline__ void s_solve_one(const Matrix& A, const Batch& B, const SmallWork& W, const Params& p, int q) {
  const int m = A.m, n = A.n, tid = threadIdx.x;
  const bool warm = B.warm != 0;
  const size_t ld0n = B.ld_0 ? B.ld_0 : (size_t)n, ld0m = B.ld_0 ? B.ld_0 : (size_t)m;
  for (int j = tid; j < n; j += SNT) {
    double x0 = 1.0, z0 = 1.0;                       // initialize_xzyw, primal_normal.cl:14-28
    if (warm) {
      x0 = fmax(B.x0[(size_t)q * ld0n + j], p.warm_floor);
      z0 = fmax(B.z0[(size_t)q * ld0n + j], p.warm_floor);
    }
    W.x[j] = x0;
    W.z[j] = z0;
    W.c[j] = B.c[(size_t)q * n + j];
  }
  for (int i = tid; i < m; i += SNT) {
    W.b[i] = B.b[(size_t)q * m + i];
    W.y[i] = warm ? B.y0[(size_t)q * ld0m + i] : 1.0;
  }
  s_bar();
  int stat = 5;                                   // primal_normal.cl:225
  double normr0 = INFINITY, norms0 = INFINITY;    // :227-228
  int iter;
  for (iter = 0; iter < p.max_iter; iter++) {
    double gsum = 0.0;
    for (int j = tid; j < n; j += SNT) gsum = fma(W.z[j], W.x[j], gsum);
    long long t0 = s_t0(W);
    const double gamma = s_block_sum(gsum, W.red);
    const double mu = p.delta * gamma / (double)(n + m);                       // :272
    // v = A'y ; sigma ; t ; d ; q                                              (:76-120, ldl.cl:198-219)
    s_At_times(A, W, W.y, W.w);
    double ss = 0.0;
    for (int j = tid; j < n; j += SNT) {
      const double v = W.w[j], xj = W.x[j], zj = W.z[j], cj = W.c[j];
      const double sig = cj - v + zj;
      ss = fma(sig, sig, ss);
      const double tj = cj - v + mu / xj;
      W.t[j] = tj;
      W.d[j] = xj / zj;
      W.w[j] = xj * tj / zj;
    }
    const double norms = sqrt(s_block_sum(ss, W.red));
    s_A_times2(A, W, W.x, W.w, W.S, W.RHS);
    double rr = 0.0;
    for (int i = tid; i < m; i += SNT) {
      const double rho = W.b[i] - W.S[i];
      rr = fma(rho, rho, rr);
      W.RHS[i] = W.RHS[i] - rho;
    }
    const double normr = sqrt(s_block_sum(rr, W.red));
    s_t1(W, 0, t0);
    if (B.trace && tid == 0 && iter < B.trace_iters) {           // :250-252
      double* tr = B.trace + ((size_t)q * B.trace_iters + iter) * 3;
      tr[0] = normr; tr[1] = norms; tr[2] = gamma;
    }
    if (normr < p.eps && norms < p.eps && gamma < p.eps) { stat = 0; break; }   // :256-259
    if (normr > 10 * normr0 && normr > p.eps) { stat = 2; break; }              // :261-264
    if (norms > 10 * norms0 && norms > p.eps) { stat = 4; break; }              // :266-269
    s_solve_normal(A, W, p);
    // dx, dz, ratio test, update                                               (:122-156)
    t0 = s_t0(W);
    s_At_times(A, W, W.dy, W.w);
    double th = 0.0;
    for (int j = tid; j < n; j += SNT) {
      const double xj = W.x[j], zj = W.z[j];
      const double dx = (W.t[j] - W.w[j]) * xj / zj;
      const double dz = (mu - zj * dx) / xj - zj;
      th = fmax(th, fmax(-dz / zj, -dx / xj));
      W.d[j] = dx;
      W.t[j] = dz;
    }
    th = s_block_max(th, W.red);
    const double theta = fmin(p.r / th, 1.0);
    for (int j = tid; j < n; j += SNT) {
      W.z[j] = W.z[j] + theta * W.t[j];
      W.x[j] = W.x[j] + theta * W.d[j];
    }
    for (int i = tid; i < m; i += SNT) W.y[i] += theta * W.dy[i];
    s_bar();
    s_t1(W, 5, t0);
    normr0 = normr;
    norms0 = norms;
  }
  const size_t ldx = B.ld_x ? B.ld_x : (size_t)n, ldy = B.ld_y ? B.ld_y : (size_t)m, ldz = B.ld_z ? B.ld_z : (size_t)n;
  const int lds = B.ld_s ? B.ld_s : 1;
  if (B.x) for (int j = tid; j < n; j += SNT) B.x[(size_t)q * ldx + j] = W.x[j];
  if (B.z) for (int j = tid; j < n; j += SNT) B.z[(size_t)q * ldz + j] = W.z[j];
  if (B.y) for (int i = tid; i < m; i += SNT) B.y[(size_t)q * ldy + i] = W.y[i];
  if (tid == 0) {
    if (B.status) B.status[(size_t)q * lds] = stat;
    if (B.iters) B.iters[(size_t)q * lds] = iter;
  }
  s_bar();
}

__global__ void __launch_bounds__(SNT, 4)
ipm_small_kernel(Matrix A, Batch B, Scratch sc, Params p) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next;
  const int m = A.m, n = A.n, tid = threadIdx.x;
  auto al = [](size_t v) { return (v + 1) & ~(size_t)1; };
  SmallWork W;
  W.T = (m + 7) / 8;
  W.lda = small_lda(A.nd);
  size_t o = 0;
  W.red = smem + o; o += SRED;
  W.As = smem + o; o += (size_t)8 * W.T * W.lda;
  W.Mp = smem + o; o += al(packed_doubles(m));
  W.L = smem + o; o += al(packed_doubles(m));
  W.colbuf = smem + o; o += al(m + 1);
  W.x = smem + o; o += al(n); W.z = smem + o; o += al(n); W.t = smem + o; o += al(n);
  W.d = smem + o; o += al(n); W.w = smem + o; o += al(n); W.c = smem + o; o += al(n);
  W.y = smem + o; o += al(m + 1); W.b = smem + o; o += al(m + 1); W.dy = smem + o; o += al(m + 1);
  W.S = smem + o; o += al(m + 1); W.RHS = smem + o; o += al(m + 1); W.D = smem + o; o += al(m + 1);
  W.dg = smem + o; o += W.lda; W.g1 = smem + o; o += W.lda; W.g2 = smem + o; o += W.lda;
  W.colval = smem + o; o += al(n);
  W.colrow = reinterpret_cast<int*>(smem + o); o += al((n + 1) / 2);
  W.offs = reinterpret_cast<int*>(smem + o); o += al((m + 2) / 2);
  W.sptr = reinterpret_cast<int*>(smem + o); o += al((m + 2) / 2);
  W.prof = sc.prof ? sc.prof + (size_t)blockIdx.x * 16 : nullptr;
  W.exact = sc.small == 4;
  // the shared matrix, once per block
  for (int e = tid; e < 8 * W.T * W.lda; e += SNT) {
    const int i = e / W.lda, k = e - i * W.lda;
    W.As[e] = (i < m && k < A.nd) ? A.Ad[(size_t)i * A.ldd + k] : 0.0;
  }
  for (int j = tid; j < n; j += SNT) { W.colrow[j] = A.colrow[j]; W.colval[j] = A.colval[j]; }
  for (int j = tid; j <= m; j += SNT) { W.offs[j] = (j < m) ? packed_off(j, m) : 0; W.sptr[j] = A.sing_ptr[j]; }
  for (size_t e = tid; e < packed_doubles(m); e += SNT) { W.L[e] = 0.0; W.Mp[e] = 0.0; }
  s_bar();
  for (;;) {
    if (tid == 0) s_next = atomicAdd(sc.counter, 1);
    s_bar();
    const int q = s_next;
    s_bar();
    if (q >= B.N) break;
    s_solve_one(A, B, W, p, q);
  }
}

}  // namespace pb200
