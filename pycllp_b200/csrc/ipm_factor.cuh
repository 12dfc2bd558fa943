// ipm_factor.cuh -- the fast modified LDL' and the blocked triangular solves.
//
// factor_ldl_fast: same arithmetic as factor_ldl (ldl.cl:314-378) -- left-looking, panels of
// 8 columns, D_j = max(|D_j|, (theta_j/beta)^2, delta) -- organised for one CTA:
//   A. panel update  C = M[:, panel] - L[:, :j0] * (D L[panel, :j0])'   on the FP64 tensor
//      cores (DMMA m8n8k4, one 8-row tile per warp, split-K accumulators);
//   B. the 8x8 diagonal block is eliminated by warp 0 (lane-redundant, in registers) while
//      the other warps still run step A;
//   C. every row below the block is solved against the block by its own thread;
//   D. the theta_j clamp is checked AFTER the fact: steps B/C assume it is inactive (it is,
//      for a positive semi-definite M), each thread tracks max |c_ij| per column and if
//      (theta_ub/beta)^2 could exceed a D_j the panel is restored and redone by the exact
//      column-by-column code (panel_exact).  Same results as the sequential rule.
// The right-hand side rides along as an extra row of the matrix ("row m"), so the forward
// substitution S <- (L D)^-1 RHS of ldl.cl:519-527 is a by-product of the factorisation.
#pragma once

namespace pb200 {

// scratch layout inside W.red (256 doubles)
constexpr int RED_W = 32;       // [64]  W[jj][k] = D_k * L11[jj][k]
constexpr int RED_D1 = 96;      // [8]   D of the panel
constexpr int RED_RINV = 104;   // [8]   1/D
constexpr int RED_CRHS = 112;   // [8]   panel entries of the rhs row (raw)
constexpr int RED_TH = 120;     // [2][8] ints: hi-words of max |c_ij| per column (double buffered)
constexpr int RED_RAWD = 128;   // [64]  raw diagonal block (for the exact redo)
constexpr int RED_SIZE = 256;

__device__ __forceinline__ int dbl_hi(double v) { return __double2hiint(v); }

// Exact, column-by-column elimination of one panel whose updated (unscaled) entries are in
// L storage; rhs row entries in crhs[]. Used when the speculative path cannot be proven
// equivalent, and for whole factorizations in "plain" mode by the caller's v1 code.
static __device__ void panel_exact(int m, int j0, int nb, Work& W, double beta, double delta,
                                   double* crhs, double* Sf, bool with_rhs) {
  const int tid = threadIdx.x;
  double* __restrict__ L = W.L;
  for (int jj = 0; jj < nb; jj++) {
    const int j = j0 + jj;
    const double djraw = L[cidx(j, j, m)];
    double th = 0.0;
    for (int i = j + 1 + tid; i < m; i += NT) th = fmax(th, fabs(L[cidx(i, j, m)]));
    th = block_max(th, W.red);
    const double q = th / beta;
    const double Dj = fmax(fabs(djraw), fmax(q * q, delta));
    for (int i = j + 1 + tid; i < m; i += NT) L[cidx(i, j, m)] /= Dj;
    if (tid == 0) {
      W.D[j] = Dj;
      L[cidx(j, j, m)] = 1.0;
      if (with_rhs) Sf[j] = crhs[jj] / Dj;
    }
    __syncthreads();
    if (jj + 1 < nb) {
      for (int i = j + 1 + tid; i < m; i += NT) {
        const double lij = L[cidx(i, j, m)];
        for (int j2 = j + 1; j2 < j0 + nb && j2 <= i; j2++)
          L[cidx(i, j2, m)] -= lij * (Dj * L[cidx(j2, j, m)]);
      }
      if (with_rhs && tid == 0) {
        const double s = Sf[j];
        for (int j2 = j + 1; j2 < j0 + nb; j2++) crhs[j2 - j0] -= s * (Dj * L[cidx(j2, j, m)]);
      }
      __syncthreads();
    }
  }
}

// 32-bit packed index: L(i, j) lives at coff(j, m) + i   (valid while m(m+1)/2 < 2^31)
__device__ __forceinline__ int coff(int j, int m) { return j * (m - 1) - ((j * (j - 1)) >> 1); }

// Pre-condition: the lower triangle of M is stored in W.L (packed column-major).
// rhs != nullptr: also computes Sf = (L D)^-1 rhs.
static __device__ void factor_ldl_fast(int m, Work& W, double beta, double delta,
                                       const double* __restrict__ rhs, double* __restrict__ Sf) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  double* __restrict__ L = W.L;
  double* __restrict__ D = W.D;
  double* __restrict__ P = W.P;
  double* Wm = W.red + RED_W;
  double* D1 = W.red + RED_D1;
  double* rinv = W.red + RED_RINV;
  double* crhs = W.red + RED_CRHS;
  int* thbuf = reinterpret_cast<int*>(W.red + RED_TH);
  double* rawd = W.red + RED_RAWD;
  const bool with_rhs = rhs != nullptr;
  const double inv_beta2 = 1.0 / (beta * beta);
  int parity = 0;

  for (int j0 = 0; j0 < m; j0 += NB, parity ^= 1) {
    const int nb = min(NB, m - j0);
    const int R = m - j0;
    const int ntile = (R + 7) >> 3;
    int* th = thbuf + parity * 8;
    long long tq = phase_begin(W);
    // P[k][jj] = D_k L(j0+jj, k)
    for (int e = tid; e < j0 * NB; e += NT) {
      const int k = e >> 3, jj = e & 7;
      P[e] = (jj < nb) ? L[coff(k, m) + j0 + jj] * D[k] : 0.0;
    }
    if (tid < 8) th[tid] = 0;
    __syncthreads();
    phase_end(W, 7, tq);
    tq = phase_begin(W);

    // ---- step A (+ B on warp 0) ----
    // FP64 FMAs queue behind DMMAs of the same SM sub-partition (measured: a dependent DFMA
    // chain runs 20x slower next to three DMMA warps, profiles/fp64_latency_r01.txt), so the
    // latency-critical elimination of the diagonal block (step B) gets sub-partition 0 to
    // itself: warp 0 updates tile 0 and eliminates it, warp 4 updates the rhs row, and the
    // twelve warps of sub-partitions 1-3 update all other row tiles.
    const int wsub = warp & 3;
    if (wsub != 0) {
      const int widx = (warp >> 2) * 3 + wsub - 1;            // 0..11
      for (int rt0 = 1 + widx; rt0 < ntile; rt0 += 24) {
        const int rt1 = rt0 + 12;
        const bool two = rt1 < ntile;
        const int rowa = j0 + 8 * rt0 + g, rowb = j0 + 8 * rt1 + g;
        const bool oka = rowa < m, okb = two && rowb < m;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;     // tile rt0: two split-K accumulators
        double u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;     // tile rt1
        const int ra = oka ? rowa : j0, rb = okb ? rowb : j0;   // safe rows for the loads
        int k = tg;
        int off = coff(k, m);
        for (int k0 = 0; k0 < j0; k0 += 8) {
          const int off2 = off + 4 * m - 10 - 4 * k;          // coff(k + 4)
          const double b1 = P[k * NB + g];
          const double b2 = P[(k + 4) * NB + g];
          double a1 = L[off + ra], a2 = L[off2 + ra];
          double a3 = L[off + rb], a4 = L[off2 + rb];
          if (!oka) { a1 = 0.0; a2 = 0.0; }
          if (!okb) { a3 = 0.0; a4 = 0.0; }
          dmma884(c0, c1, a1, b1);
          dmma884(u0, u1, a3, b1);
          dmma884(e0, e1, a2, b2);
          dmma884(v0, v1, a4, b2);
          off = off2 + 4 * m - 10 - 4 * (k + 4);              // coff(k + 8)
          k += 8;
        }
        c0 += e0; c1 += e1; u0 += v0; u1 += v1;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int col = 2 * tg + h;
          if (col < nb) {
            const int cb = coff(j0 + col, m);
            if (oka) L[cb + rowa] -= (h ? c1 : c0);
            if (okb) L[cb + rowb] -= (h ? u1 : u0);
          }
        }
      }
    } else if (warp == 0 || (warp == 4 && with_rhs)) {
      // tile 0 (warp 0) / the rhs row (warp 4): four split-K chains to keep the chain short
      const bool is_rhs = warp == 4;
      const int row = j0 + g;
      const bool ok = !is_rhs && row < m;
      const int rs = ok ? row : j0;
      double c[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      int k = tg;
      int off = coff(k, m);
      for (int k0 = 0; k0 < j0; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (k0 + 4 * q < j0) {
            double a = is_rhs ? Sf[k] : L[off + rs];
            if (is_rhs ? (g != 0) : !ok) a = 0.0;
            dmma884(c[q][0], c[q][1], a, P[k * NB + g]);
          }
          off += 4 * m - 10 - 4 * k;
          k += 4;
        }
      }
      const double s0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
      const double s1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int col = 2 * tg + h;
        if (col < nb) {
          if (is_rhs) { if (g == 0) crhs[col] = rhs[j0 + col] - (h ? s1 : s0); }
          else if (ok && row >= j0 + col) L[coff(j0 + col, m) + row] -= (h ? s1 : s0);
        }
      }
      if (warp == 0) {
        // ---- step B: 8x8 diagonal block, every lane of warp 0 redundantly ----
        __syncwarp();
        long long tb = phase_begin(W);
        int cb[8];
#pragma unroll
        for (int j = 0; j < 8; j++) cb[j] = coff(min(j0 + j, m - 1), m) + j0;
        double a[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j <= i; j++)
            a[i][j] = (i < nb) ? L[cb[j] + i] : (i == j ? 1.0 : 0.0);
        if (lane < 8) {   // lane i keeps the raw row i for the exact redo
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (lane == i) {
#pragma unroll
              for (int j = 0; j <= i; j++) rawd[i * 8 + j] = a[i][j];
            }
        }
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          double t = 0.0;
#pragma unroll
          for (int i = jj + 1; i < 8; i++) t = fmax(t, fabs(a[i][jj]));
          const double Dj = fmax(fabs(a[jj][jj]), delta);   // speculative: theta clamp inactive
          const double r = 1.0 / Dj;
          if (lane == jj) {                                   // spread the stores over lanes
            if (jj < nb) {
              atomicMax(&th[jj], dbl_hi(t));
              D[j0 + jj] = Dj;
              D1[jj] = Dj;
              rinv[jj] = r;
            } else {
              D1[jj] = 1.0;
              rinv[jj] = 1.0;
            }
          }
#pragma unroll
          for (int i = jj + 1; i < 8; i++) a[i][jj] *= r;
#pragma unroll
          for (int j2 = jj + 1; j2 < 8; j2++) {
            const double w = Dj * a[j2][jj];
            if (lane == 8 + jj) Wm[j2 * 8 + jj] = w;
#pragma unroll
            for (int i = j2; i < 8; i++) a[i][j2] -= a[i][jj] * w;
          }
        }
        // lane i (< 8) writes row i of the unit-lower block
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (lane == i && i < nb) {
#pragma unroll
            for (int j = 0; j < i; j++) L[cb[j] + i] = a[i][j];
            L[cb[i] + i] = 1.0;
          }
        }
        phase_end(W, 9, tb);
      }
    }
    __syncthreads();
    phase_end(W, 8, tq);
    tq = phase_begin(W);

    // ---- step C: rows below the diagonal block (+ the rhs row), one thread each ----
    const int nbelow = (R > 8) ? R - 8 : 0;
    const int nthr_rows = nbelow + (with_rhs ? 1 : 0);
    int hmax[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) hmax[jj] = 0;
    if (tid < nthr_rows || nthr_rows > NT) {
      double wreg[28];                                        // W[jj][k], k < jj
      double rv[8];
      {
        int q = 0;
#pragma unroll
        for (int jj = 1; jj < 8; jj++)
#pragma unroll
          for (int k = 0; k < jj; k++) wreg[q++] = Wm[jj * 8 + k];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) rv[jj] = rinv[jj];
      }
      int cb[8];
#pragma unroll
      for (int j = 0; j < 8; j++) cb[j] = coff(min(j0 + j, m - 1), m);
      for (int t = tid; t < nthr_rows; t += NT) {
        const bool is_rhs = (t == nbelow);
        const int row = is_rhs ? m - 1 : j0 + 8 + t;
        double l[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          double cv = (jj < nb) ? L[cb[jj] + row] : 0.0;
          if (is_rhs) cv = (jj < nb) ? crhs[jj] : 0.0;
          P[t * 8 + jj] = cv;                                  // kept for the exact redo
          l[jj] = cv;
        }
        {
          int q = 0;
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            double v = l[jj];
#pragma unroll
            for (int k = 0; k < jj; k++) v -= l[k] * wreg[q++];
            if (!is_rhs) hmax[jj] = max(hmax[jj], dbl_hi(fabs(v)));
            l[jj] = v * rv[jj];
          }
        }
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          if (jj < nb) {
            if (is_rhs) Sf[j0 + jj] = l[jj];
            else L[cb[jj] + row] = l[jj];
          }
        }
      }
    }
    {
      int hm = 0;   // lane jj of each warp publishes column jj
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int r = __reduce_max_sync(0xffffffffu, hmax[jj]);
        if (lane == jj) hm = r;
      }
      if (lane < 8 && hm > 0) atomicMax(&th[lane], hm);
    }
    __syncthreads();
    phase_end(W, 10, tq);
    tq = phase_begin(W);

    // ---- step D: was the speculation exact? ----
    bool bad = false;
    for (int jj = 0; jj < nb; jj++) {
      // theta_ub > theta : bump the hi-word by one (covers the dropped low word)
      const double tub = __hiloint2double(th[jj] + 1, 0);
      if (!(tub * tub * inv_beta2 * 1.0000001 <= D1[jj])) bad = true;
    }
    if (bad) {
      // restore the updated-but-uneliminated panel and redo it by the sequential rule
      if (tid == 0) {
        for (int i = 0; i < nb; i++)
          for (int j = 0; j <= i; j++) L[cidx(j0 + i, j0 + j, m)] = rawd[i * 8 + j];
      }
      for (int t = tid; t < nthr_rows; t += NT) {
        const bool is_rhs = (t == nbelow);
        const int row = j0 + 8 + t;
        for (int jj = 0; jj < nb; jj++) {
          if (is_rhs) crhs[jj] = P[(size_t)t * 8 + jj];
          else L[cidx(row, j0 + jj, m)] = P[(size_t)t * 8 + jj];
        }
      }
      __syncthreads();
      panel_exact(m, j0, nb, W, beta, delta, crhs, Sf, with_rhs);
      __syncthreads();
    }
    phase_end(W, 11, tq);
  }
}

// S <- L^-T S ; dy += S     (second half of ldl.cl:529-536), blocks of 32 columns:
// the part of each dot product below the block is a warp-per-column reduction on all
// warps, the 32x32 triangle is back-substituted by warp 0 in registers with shuffles.
static __device__ void back_solve_fast(int m, Work& W) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ L = W.L;
  double* __restrict__ S = W.S;
  const int nblk = (m + 31) >> 5;
  for (int kb = nblk - 1; kb >= 0; kb--) {
    const int c0 = kb << 5;
    const int c1 = min(m, c0 + 32);
    if (c1 < m) {
      for (int j = c0 + warp; j < c1; j += NWARP) {
        const double* col = L + cidx(j, j, m) - j;     // col[i] = L(i, j)
        double acc = 0.0;
        for (int i = c1 + lane; i < m; i += 32) acc += col[i] * S[i];
        acc = warp_sum(acc);
        if (lane == 0) S[j] -= acc;
      }
      __syncthreads();
    }
    if (warp == 0) {
      const int j = c0 + lane;
      const bool valid = j < c1;
      double w = valid ? S[j] : 0.0;
      double c[32];
      const double* col = L + cidx(valid ? j : c0, valid ? j : c0, m) - (valid ? j : c0);
#pragma unroll
      for (int ii = 1; ii < 32; ii++)
        c[ii] = (valid && ii > lane && c0 + ii < c1) ? col[c0 + ii] : 0.0;
#pragma unroll
      for (int ii = 31; ii >= 1; ii--) {
        const double vi = __shfl_sync(0xffffffffu, w, ii);
        w -= c[ii] * vi;
      }
      if (valid) {
        S[j] = w;
        W.dy[j] += w;
      }
    }
    __syncthreads();
  }
}

// S <- (L D)^-1 S   (first half, ldl.cl:519-527); only used by refinement passes -- the
// first solve of every iteration gets this from factor_ldl_fast.
static __device__ void fwd_solve_fast(int m, Work& W) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ L = W.L;
  double* __restrict__ S = W.S;
  const int nblk = (m + 31) >> 5;
  for (int kb = 0; kb < nblk; kb++) {
    const int c0 = kb << 5;
    const int c1 = min(m, c0 + 32);
    if (warp == 0) {
      const int i = c0 + lane;
      const bool valid = i < c1;
      double u = valid ? S[i] : 0.0;
      double r[32];
#pragma unroll
      for (int jj = 0; jj < 31; jj++)
        r[jj] = (valid && jj < lane) ? L[cidx(i, c0 + jj, m)] : 0.0;
#pragma unroll
      for (int jj = 0; jj < 31; jj++) {
        const double uj = __shfl_sync(0xffffffffu, u, jj);
        u -= r[jj] * uj;
      }
      if (valid) S[i] = u;
    }
    __syncthreads();
    if (c1 < m) {
      for (int i = c1 + tid; i < m; i += NT) {
        double acc = 0.0;
        for (int j = c0; j < c1; j++) acc += L[cidx(i, j, m)] * S[j];
        S[i] -= acc;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < m; i += NT) S[i] /= W.D[i];
  __syncthreads();
}

}  // namespace pb200
