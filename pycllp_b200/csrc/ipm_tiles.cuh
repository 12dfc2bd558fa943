// ipm_tiles.cuh -- batched SPARSE numeric LDL' on the symbolic pattern (block-sparse, 8x8 tiles).
//
// Replaces sparse_factor_primal_normal / sparse_forward_backward_primal_normal (ldl.cl:381-502,
// 540-574; prototype sparse_ldl.py:72-152) for constraint matrices whose factor is genuinely
// sparse.  The reference stores L on the pattern of the Cholesky factor of A A' in CSR-lower
// (diagonal last) and finds every L_ij by a merge of rows i and j; here the pattern is analysed
// once per engine at the granularity of 8x8 tiles (cabi.cu: block elimination tree, block fill,
// the list of tile pairs that update every tile), and each LP keeps ONLY the tiles of that
// pattern in its scratch slot -- memory ~ nnz(L), no m x m array anywhere.
//
// Per block column J (left-looking, natural order like the reference):
//   A. every tile (I, J) -= sum_K L(I,K) D_K L(J,K)'  over its precomputed pair list: one warp per
//      tile, two DMMA m8n8k4 per pair, fragments straight from the tiles (a tile column is 8
//      consecutive doubles, a fragment load is one coalesced 256-byte request);
//   B. the eight columns of the block are eliminated one at a time across ALL rows of the block
//      column by the exact sequential rule of ldl.cl:349-376:
//        theta_j = max_i |c_ij| ; D_j = max(|D_j|, (theta_j/beta)^2, delta) ; L_ij = c_ij / D_j ;
//        c_ij' -= L_ij c_j'j   (j' = j+1 .. 7 of this block)
//      one thread per row, the theta reduction over a named barrier that only the warps owning
//      rows take part in.  No speculation, hence no fallback path.
// The triangular solves walk the same tiles (one warp, column-oriented forward, dot-product
// backward) -- they are latency chains of length m either way.
// Entries of M outside the pattern of A A' are structural zeros; tiles inside the block fill hold
// explicit zeros where the scalar fill has none, which changes nothing in the arithmetic.
#pragma once

namespace pb200 {

constexpr int BAR_TILES = 9;      // named barrier of the panel elimination (ids 1..8: ipm_factor.cuh)
constexpr int RED_TW = 160;       // [2][8]  unscaled diagonal-tile column, double-buffered (W.red scratch)
constexpr int RED_TR = 176;       // [2][16] per-warp maxima                     (RED_KEEP starts at 208)

// M = A diag(d) A' scattered into the tiles (zero fill first); rows m .. 8 nbk - 1 of the last
// block get a unit diagonal so that the padded system factors trivially.
static __device__ __forceinline__ void tiles_form_M(const Matrix& A, Work& W) {
  const int tid = threadIdx.x;
  double* __restrict__ Lt = W.L;
  {
    double2* L2 = reinterpret_cast<double2*>(Lt);
    const size_t n2 = (size_t)A.ntiles * 32;
    for (size_t e = tid; e < n2; e += NT) L2[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  for (int e0 = tid; e0 < A.nme; e0 += 4 * NT) {
    int t0[4], t1[4], pos[4], k[4];
    double w[4], s[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int e = min(e0 + q * NT, A.nme - 1);
      t0[q] = A.me_ptr[e]; t1[q] = A.me_ptr[e + 1];
      pos[q] = A.me_pos[e];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { k[q] = A.mt_k[t0[q]]; w[q] = A.mt_w[t0[q]]; }   // every entry has >= 1 term
#pragma unroll
    for (int q = 0; q < 4; q++) s[q] = w[q] * W.d[k[q]];
#pragma unroll
    for (int q = 0; q < 4; q++)
      for (int t = t0[q] + 1; t < t1[q]; t++) s[q] += A.mt_w[t] * W.d[A.mt_k[t]];
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (e0 + q * NT < A.nme) Lt[pos[q]] = s[q];
  }
  const int mp = 8 * A.nbk;
  for (int i = A.m + tid; i < mp; i += NT) Lt[(size_t)A.tl_colptr[i >> 3] * 64 + (i & 7) * 9] = 1.0;
}

// max_i |M_ii| (ldl.cl:280-294) from the diagonal tiles
static __device__ __forceinline__ double tiles_diag_absmax(const Matrix& A, Work& W) {
  double bmax = 0.0;
  for (int i = threadIdx.x; i < A.m; i += NT)
    bmax = fmax(bmax, fabs(W.L[(size_t)A.tl_colptr[i >> 3] * 64 + (i & 7) * 9]));
  return block_max(bmax, W.red);
}

// Per-block cache of D (and of S during the solves) in the shared-memory area that the dense
// super-panel factor would use (W.fb, FB_DOUBLES): the tile updates read D_K of every pair and
// the solves are latency chains over S, neither should pay an L2 round trip per access.
__device__ __forceinline__ bool tiles_cache_ok(const Matrix& A) { return 2 * 8 * A.nbk <= FB_DOUBLES; }

static __device__ __forceinline__ void tiles_factor(const Matrix& A, Work& W, double beta, double delta) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int g = lane >> 2, tg = lane & 3;
  const unsigned FULL = 0xffffffffu;
  double* __restrict__ Lt = W.L;
  const bool cached = tiles_cache_ok(A);
  double* __restrict__ D = cached ? W.fb : W.D;      // (copied to W.D at the end when cached)
  double* wsm = W.red + RED_TW;                      // [2][8]
  double* rsm = W.red + RED_TR;                      // [2][16]
  const double inv_beta = 1.0 / beta;
  for (int J = 0; J < A.nbk; J++) {
    const int c0 = __ldg(A.tl_colptr + J), c1 = __ldg(A.tl_colptr + J + 1);
    long long tq = phase_begin(W);
    // ---- A. left-looking tile updates: one warp per tile ----
    for (int t = c0 + warp; t < c1; t += NWARP) {
      const int p0 = __ldg(A.tl_updptr + t), p1 = __ldg(A.tl_updptr + t + 1);
      if (p0 == p1) continue;
      double* T = Lt + (size_t)t * 64;
      const double v0 = T[(2 * tg) * 8 + g], v1 = T[(2 * tg + 1) * 8 + g];   // (in flight during the loop)
      double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
      for (int base = p0; base < p1; base += 32) {
        // the pair list of this tile, 32 pairs per round trip (one per lane), broadcast by shuffles
        const int mine = min(base + lane, p1 - 1);
        const int ia = __ldg(A.tl_upda + mine), ib = __ldg(A.tl_updb + mine), ik = __ldg(A.tl_updk + mine);
        const int cnt = min(32, p1 - base);
        for (int q = 0; q < cnt; q += 4) {
          double fa1[4], fa2[4], fb1[4], fb2[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {              // four pairs' fragments in flight
            const int qq = min(q + u, cnt - 1);
            const int a = __shfl_sync(FULL, ia, qq), b = __shfl_sync(FULL, ib, qq), K = __shfl_sync(FULL, ik, qq);
            const double* pa = Lt + (size_t)a * 64 + tg * 8 + g;
            const double* pb = Lt + (size_t)b * 64 + tg * 8 + g;
            const bool on = q + u < cnt;
            fa1[u] = on ? pa[0] : 0.0;
            fa2[u] = on ? pa[32] : 0.0;
            fb1[u] = pb[0] * D[8 * K + tg];
            fb2[u] = pb[32] * D[8 * K + tg + 4];
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            dmma884(x0, x1, fa1[u], fb1[u]);
            dmma884(y0, y1, fa2[u], fb2[u]);
          }
        }
      }
      T[(2 * tg) * 8 + g] = v0 - (x0 + y0);
      T[(2 * tg + 1) * 8 + g] = v1 - (x1 + y1);
    }
    __syncthreads();
    phase_end(W, 13, tq);                            // (profile slots of the dense factor: f_old = tile updates,
    tq = phase_begin(W);                             //  f_solve = panel elimination)
    // ---- B. the eight columns, one at a time, over all rows of the block column ----
    const int rows = 8 * (c1 - c0);
    double* Tc = Lt + (size_t)c0 * 64;               // the tiles of this block column are contiguous
    if (rows <= NT) {
      // one thread per row, the row's eight entries in registers from start to finish
      const int nthr = (rows + 31) & ~31;
      if (tid < nthr) {
        const int nw = nthr >> 5;
        const bool live = tid < rows;
        double* pr = Tc + (size_t)(tid >> 3) * 64 + (tid & 7);
        double c[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) c[jj] = live ? pr[jj * 8] : 0.0;
        const int jmax = (tid < 8) ? tid : 7;        // diagonal tile: lower triangle only
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          double* wb = wsm + 8 * (jj & 1);
          double* rb = rsm + 16 * (jj & 1);
          if (tid < 8) wb[tid] = c[jj];              // unscaled c_{j' j} of the diagonal tile
          double th = (live && tid > jj) ? fabs(c[jj]) : 0.0;
          th = warp_max(th);
          if (lane == 0) rb[warp] = th;
          nbar_sync(BAR_TILES, nthr);
          double theta = 0.0;
          for (int w = 0; w < nw; w++) theta = fmax(theta, rb[w]);
          const double q = theta * inv_beta;
          const double Dj = fmax(fabs(wb[jj]), fmax(q * q, delta));        // ldl.cl:368 / :479
          if (tid > jj) {
            const double l = c[jj] * rcp_pos64(Dj);  // (reciprocal-multiply, <= 1 ulp from the division, as the dense factor)
            c[jj] = l;
#pragma unroll
            for (int j2 = jj + 1; j2 < 8; j2++)
              if (j2 <= jmax) c[j2] = fma(-l, wb[j2], c[j2]);
          } else if (tid == jj) {
            c[jj] = 1.0;
            if (cached || 8 * J + jj < A.m) D[8 * J + jj] = Dj;
          }
        }
        if (live) {
#pragma unroll
          for (int jj = 0; jj < 8; jj++)
            if (jj <= jmax) pr[jj * 8] = c[jj];
        }
      }
    } else {
      // (block columns with more than NT rows: same rule, rows in memory)
#pragma unroll 1
      for (int jj = 0; jj < 8; jj++) {
        if (tid < 8) wsm[tid] = Tc[jj * 8 + tid];
        double th = 0.0;
        for (int R = tid; R < rows; R += NT)
          if (R > jj) th = fmax(th, fabs(Tc[(size_t)(R >> 3) * 64 + jj * 8 + (R & 7)]));
        const double theta = block_max(th, W.red);
        const double q = theta * inv_beta;
        const double Dj = fmax(fabs(wsm[jj]), fmax(q * q, delta));
        double wj[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++) wj[j2] = wsm[j2];
        for (int R = tid; R < rows; R += NT) {
          double* pr = Tc + (size_t)(R >> 3) * 64 + (R & 7);
          if (R > jj) {
            const double l = pr[jj * 8] / Dj;
            const int jmax = (R < 8) ? R : 7;
            double v[8];
#pragma unroll
            for (int j2 = 1; j2 < 8; j2++) v[j2] = pr[j2 * 8];     // (loads before the stores)
            pr[jj * 8] = l;
#pragma unroll
            for (int j2 = 1; j2 < 8; j2++)
              if (j2 > jj && j2 <= jmax) pr[j2 * 8] = v[j2] - l * wj[j2];
          } else if (R == jj) {
            pr[jj * 8] = 1.0;
            if (cached || 8 * J + jj < A.m) D[8 * J + jj] = Dj;
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    phase_end(W, 10, tq);
  }
  if (cached) {
    for (int i = tid; i < A.m; i += NT) W.D[i] = D[i];
    __syncthreads();
  }
}

// S <- (L D)^-1 RHS  (first half of ldl.cl:540-574), column-oriented, one warp; S is kept in
// shared memory while the chain runs (tiles_cache_ok)
static __device__ __forceinline__ void tiles_forward(const Matrix& A, Work& W, const double* rhs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int m = A.m, mp = 8 * A.nbk;
  const double* __restrict__ Lt = W.L;
  const bool cached = tiles_cache_ok(A);
  double* S = cached ? W.fb + mp : W.S;
  if (cached) {
    for (int i = tid; i < mp; i += NT) S[i] = (i < m) ? rhs[i] : 0.0;
  } else if (rhs != S) {
    for (int i = tid; i < m; i += NT) S[i] = rhs[i];
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned FULL = 0xffffffffu;
    const int r = lane & 7, slot = lane >> 3;
    int c0 = __ldg(A.tl_colptr);
    for (int J = 0; J < A.nbk; J++) {
      const int c1 = __ldg(A.tl_colptr + J + 1);
      const double* Td = Lt + (size_t)c0 * 64;
      double lcol[7];
#pragma unroll
      for (int jj = 0; jj < 7; jj++) lcol[jj] = Td[jj * 8 + r];      // L(r, jj), used for r > jj
      // the first sub-diagonal tile of each slot is fetched together with the diagonal tile
      const int t1 = c0 + 1 + slot;
      double tl[8];
#pragma unroll
      for (int jj = 0; jj < 8; jj++) tl[jj] = (t1 < c1) ? Lt[(size_t)t1 * 64 + jj * 8 + r] : 0.0;
      const int row = 8 * J + r;
      double s = (cached || row < m) ? S[row] : 0.0;
#pragma unroll
      for (int jj = 0; jj < 7; jj++) {
        const double sj = __shfl_sync(FULL, s, jj);
        if (r > jj) s = fma(-lcol[jj], sj, s);
      }
      if (lane < 8 && (cached || row < m)) S[row] = s;
      double sb[8];
#pragma unroll
      for (int jj = 0; jj < 8; jj++) sb[jj] = __shfl_sync(FULL, s, jj);
      for (int t = t1; t < c1; t += 4) {
        if (t != t1) {
#pragma unroll
          for (int jj = 0; jj < 8; jj++) tl[jj] = Lt[(size_t)t * 64 + jj * 8 + r];
        }
        double acc = 0.0;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) acc = fma(tl[jj], sb[jj], acc);
        const int ri = 8 * __ldg(A.tl_row + t) + r;
        if (cached || ri < m) S[ri] -= acc;
      }
      __syncwarp();
      c0 = c1;
    }
  }
  __syncthreads();
  for (int i = tid; i < m; i += NT) S[i] /= W.D[i];
  __syncthreads();
}

// S <- L^-T S ; dy += sign S  (second half of ldl.cl:540-574), one warp
static __device__ __forceinline__ void tiles_backward(const Matrix& A, Work& W, double sign = 1.0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int m = A.m, mp = 8 * A.nbk;
  const double* __restrict__ Lt = W.L;
  const bool cached = tiles_cache_ok(A);
  double* S = cached ? W.fb + mp : W.S;
  if (warp == 0) {
    const unsigned FULL = 0xffffffffu;
    const int c = lane & 7, slot = lane >> 3;
    int c1 = __ldg(A.tl_colptr + A.nbk);
    for (int J = A.nbk - 1; J >= 0; J--) {
      const int c0 = __ldg(A.tl_colptr + J);
      const double* Td = Lt + (size_t)c0 * 64;
      double lrow[8];
#pragma unroll
      for (int jj = 1; jj < 8; jj++) lrow[jj] = Td[c * 8 + jj];        // L(jj, c), used for jj > c
      double part = 0.0;
      for (int t = c0 + 1 + slot; t < c1; t += 4) {
        const double* T = Lt + (size_t)t * 64 + c * 8;
        const int r0 = 8 * __ldg(A.tl_row + t);
        double tv[8];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) tv[rr] = T[rr];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
          const double sv = (cached || r0 + rr < m) ? S[r0 + rr] : 0.0;
          part = fma(tv[rr], sv, part);
        }
      }
      part += __shfl_xor_sync(FULL, part, 8);
      part += __shfl_xor_sync(FULL, part, 16);
      const int row = 8 * J + c;
      double s = ((cached || row < m) ? S[row] : 0.0) - part;
#pragma unroll
      for (int jj = 7; jj >= 1; jj--) {
        const double sj = __shfl_sync(FULL, s, jj);
        if (c < jj) s = fma(-lrow[jj], sj, s);
      }
      if (lane < 8 && (cached || row < m)) S[row] = s;
      __syncwarp();
      c1 = c0;
    }
  }
  __syncthreads();
  for (int i = tid; i < m; i += NT) W.dy[i] += sign * S[i];
  __syncthreads();
}

// factor + solve on the tile pattern (sparse_solve_primal_normal, ldl.cl:655-712: no refinement).
// Out of line, and its arguments BY VALUE: a reference to the caller's Matrix / Work would force
// those structs into local memory for the whole kernel (measured: the stack frame of the
// LS = false kernels grew from 672 to 1160 bytes and configs 4 and 5 lost 6-8 %).
struct TilesArgs {
  int m, nme, nbk, ntiles;
  const int *me_ptr, *me_pos, *mt_k, *tl_colptr, *tl_row, *tl_updptr, *tl_upda, *tl_updb, *tl_updk;
  const double* mt_w;
  double *L, *D, *S, *RHS, *dy, *d, *fb, *red;
  unsigned long long* prof;
  double ldl_delta;
};

static __device__ __noinline__ void solve_normal_tiles_call(TilesArgs a) {
  Matrix A;
  A.m = a.m; A.nme = a.nme; A.nbk = a.nbk; A.ntiles = a.ntiles;
  A.me_ptr = a.me_ptr; A.me_pos = a.me_pos; A.mt_k = a.mt_k; A.mt_w = a.mt_w;
  A.tl_colptr = a.tl_colptr; A.tl_row = a.tl_row; A.tl_updptr = a.tl_updptr;
  A.tl_upda = a.tl_upda; A.tl_updb = a.tl_updb; A.tl_updk = a.tl_updk;
  Work W;
  W.L = a.L; W.D = a.D; W.S = a.S; W.RHS = a.RHS; W.dy = a.dy; W.d = a.d; W.fb = a.fb; W.red = a.red;
  W.prof = a.prof;
  const int m = A.m, tid = threadIdx.x;
  long long t0 = phase_begin(W);
  tiles_form_M(A, W);
  for (int i = tid; i < m; i += NT) W.dy[i] = 0.0;
  __syncthreads();
  phase_end(W, 1, t0);
  t0 = phase_begin(W);
  const double beta = sqrt(tiles_diag_absmax(A, W));
  tiles_factor(A, W, beta, a.ldl_delta);
  phase_end(W, 2, t0);
  t0 = phase_begin(W);
  tiles_forward(A, W, W.RHS);
  tiles_backward(A, W);
  phase_end(W, 3, t0);
}

static __device__ __forceinline__ void solve_normal_tiles(const Matrix& A, Work& W, const Params& p) {
  TilesArgs a;
  a.m = A.m; a.nme = A.nme; a.nbk = A.nbk; a.ntiles = A.ntiles;
  a.me_ptr = A.me_ptr; a.me_pos = A.me_pos; a.mt_k = A.mt_k; a.mt_w = A.mt_w;
  a.tl_colptr = A.tl_colptr; a.tl_row = A.tl_row; a.tl_updptr = A.tl_updptr;
  a.tl_upda = A.tl_upda; a.tl_updb = A.tl_updb; a.tl_updk = A.tl_updk;
  a.L = W.L; a.D = W.D; a.S = W.S; a.RHS = W.RHS; a.dy = W.dy; a.d = W.d; a.fb = W.fb; a.red = W.red;
  a.prof = W.prof;
  a.ldl_delta = p.ldl_delta;
  solve_normal_tiles_call(a);
}

}  // namespace pb200
