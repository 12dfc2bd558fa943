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
static __device__ __forceinline__ void panel_exact(int m, int j0, int nb, Work& W, double beta, double delta,
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


// approximate single-precision reciprocal (one MUFU), seed of rcp_pos
__device__ __forceinline__ float rcp_approx_f32(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 1/d to within an ulp for delta <= d < 1e30 (callers guarantee the range): single-precision
// seed + two Newton steps.
__device__ __forceinline__ double rcp_pos(double d) {
  double r = (double)rcp_approx_f32((float)d);
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Step B of factor_ldl_fast: eliminate the 8x8 diagonal block of panel j0 (its updated,
// unscaled entries are in L storage).  Run by ONE warp, and it is the serial bottleneck of the
// whole factorisation (a lone warp issues a dependent instruction every ~5 cycles), so the
// instruction stream is kept minimal: the 36 lower-triangle entries are spread over the lanes
// (entry q = lane, column-major inside the block: column starts 0,8,15,21,26,30,33,35; the
// four entries 32..35 ride in a second register of lanes 0..3), columns 0..4 are eliminated
// with two shuffles per entry, the trailing 3x3 block redundantly in every lane.
// Speculative D_j = max(|D_j|, delta) (theta clamp assumed inactive, checked in step D; a
// pivot above 1e30 also routes to the exact path).  Outputs, all to shared scratch: D1, rinv,
// W[j][k] = D_k L11[j][k] (= the unscaled entry), th[] = in-block part of theta_j (hi words),
// the raw block (for the exact redo); the scaled L11 goes to L storage.
static __device__ __forceinline__ void diag_block(int m, int j0, int nb, Work& W, double delta,
                                                  int* th, double* Wm, double* D1, double* rinv) {
  const int lane = threadIdx.x & 31;
  double* __restrict__ L = W.L;
  double* rawd = W.red + RED_RAWD;
  const unsigned FULL = 0xffffffffu;
  // lane -> (row ei, column ej) of entry q = lane
  const int ej = (lane >= 8) + (lane >= 15) + (lane >= 21) + (lane >= 26) + (lane >= 30);
  const int cs_mine = (0x1e1a150f0800ull >> (8 * ej)) & 0xff;      // {0,8,15,21,26,30}
  const int ei = ej + lane - cs_mine;
  // second entry of lanes 0..3: (7,5) (6,6) (7,6) (7,7)
  const bool has2 = lane < 4;
  const int fi = (lane == 1) ? 6 : 7;
  const int fj = (lane == 0) ? 5 : (lane == 3) ? 7 : 6;
  const int idx0 = coff(min(j0 + ej, m - 1), m) + j0 + ei;
  const int idx1 = coff(min(j0 + fj, m - 1), m) + j0 + fi;
  double e0 = (ei == ej) ? 1.0 : 0.0;
  double e1 = (has2 && fi == fj) ? 1.0 : 0.0;
  if (ei < nb) e0 = L[idx0];
  if (has2 && fi < nb) e1 = L[idx1];
  rawd[ei * 8 + ej] = e0;
  if (has2) rawd[fi * 8 + fj] = e1;
  bool big = false;
  const int hdelta = dbl_hi(delta);

#pragma unroll
  for (int jj = 0; jj < 5; jj++) {
    const int c0 = (jj == 0) ? 0 : (jj == 1) ? 8 : (jj == 2) ? 15 : (jj == 3) ? 21 : 26;
    const bool upd = ej > jj;
    const bool incol = (ej == jj) && (ei > jj);
    const int srcl = c0 + (upd ? ei - jj : 0), srcw = c0 + (upd ? ej - jj : 0);
    const double pj = __shfl_sync(FULL, e0, c0);
    const double li = __shfl_sync(FULL, e0, srcl);
    const double wj = __shfl_sync(FULL, e0, srcw);
    const double li2 = __shfl_sync(FULL, e0, c0 + fi - jj);
    const double wj2 = __shfl_sync(FULL, e0, c0 + fj - jj);
    const int ht = __reduce_max_sync(FULL, incol ? (dbl_hi(e0) & 0x7fffffff) : 0);
    // D_j = max(|pivot|, delta) decided on the high words (ties, NaN and huge pivots -> exact path)
    const int hp = dbl_hi(pj) & 0x7fffffff;
    big |= (hp == hdelta) | (hp >= 0x46293e59);                  // 0x46293e59 ~ hi word of 1e30
    const double Dj = (hp < hdelta) ? delta : fabs(pj);
    const double r = rcp_pos(Dj);
    if (upd) e0 = fma(-(li * r), wj, e0);
    e1 = fma(-(li2 * r), wj2, e1);
    if (incol) {
      Wm[ei * 8 + jj] = e0;        // D_j * l_ij == the unscaled entry
      e0 *= r;
    }
    if (lane == c0) {
      th[jj] = ht;
      D1[jj] = Dj;
      rinv[jj] = r;
    }
  }
  // trailing 3x3 block (rows/cols 5..7), redundantly in every lane
  double b55 = __shfl_sync(FULL, e0, 30), b65 = __shfl_sync(FULL, e0, 31);
  double b75 = __shfl_sync(FULL, e1, 0), b66 = __shfl_sync(FULL, e1, 1);
  double b76 = __shfl_sync(FULL, e1, 2), b77 = __shfl_sync(FULL, e1, 3);
  const double D5 = fmax(fabs(b55), delta);
  const double r5 = rcp_pos(D5);
  const int ht5 = max(dbl_hi(b65) & 0x7fffffff, dbl_hi(b75) & 0x7fffffff);
  const double w65 = b65, w75 = b75;
  b65 *= r5;
  b75 *= r5;
  b66 = fma(-b65, w65, b66);
  b76 = fma(-b75, w65, b76);
  b77 = fma(-b75, w75, b77);
  const double D6 = fmax(fabs(b66), delta);
  const double r6 = rcp_pos(D6);
  const int ht6 = dbl_hi(b76) & 0x7fffffff;
  const double w76 = b76;
  b76 *= r6;
  b77 = fma(-b76, w76, b77);
  const double D7 = fmax(fabs(b77), delta);
  const double r7 = rcp_pos(D7);
  big |= !(D5 < 1e30) | !(D6 < 1e30) | !(D7 < 1e30);
  if (lane == 5) { th[5] = ht5; D1[5] = D5; rinv[5] = r5; Wm[6 * 8 + 5] = w65; Wm[7 * 8 + 5] = w75; }
  if (lane == 6) { th[6] = ht6; D1[6] = D6; rinv[6] = r6; Wm[7 * 8 + 6] = w76; }
  if (lane == 7) { th[7] = 0; D1[7] = D7; rinv[7] = r7; if (big) th[0] = 0x7ff00000; }
  if (lane == 31) e0 = b65;
  if (lane == 0) e1 = b75;
  if (lane == 2) e1 = b76;
  // scaled unit-lower block back to L storage
  if (ei == ej) e0 = 1.0;
  if (fi == fj) e1 = 1.0;
  if (ei < nb) L[idx0] = e0;
  if (has2 && fi < nb) L[idx1] = e1;
}

// The inner loop of step A for one warp: acc (8x8, rows ra+0..7 [and ra+drow+0..7]) =
// sum_k L(row, k) * P[k][:] over k < j0, as DMMAs with two split-K chains per tile.  A lone
// warp issues roughly one dependent instruction per 5 cycles, so the loop carries nothing but
// the loads, the DMMAs and three pointer updates; operands of the next iteration are loaded
// before the DMMAs of the current one.
template <bool TWO>
__device__ __forceinline__ void panel_update_tiles(const double* __restrict__ L,
                                                   const double* __restrict__ P, int m, int kbase, int j0,
                                                   int tg, int g, int ra, int drow, double& c0,
                                                   double& c1, double& u0, double& u1) {
  double e0 = 0.0, e1 = 0.0, v0 = 0.0, v1 = 0.0;
  c0 = c1 = u0 = u1 = 0.0;
  if (j0 == kbase) return;
  const double* pa = L + coff(kbase + tg, m) + ra;      // column k = kbase + tg, row ra
  int d = 4 * m - 8 - 4 * (kbase + tg);         // coff(k + 4) - coff(k); decreases by 16 per step
  const double* pb = P + tg * NB + g;
  double a1, a2, a3 = 0.0, a4 = 0.0, b1, b2;
  a1 = pa[0]; if (TWO) a3 = pa[drow]; pa += d; d -= 16;
  a2 = pa[0]; if (TWO) a4 = pa[drow]; pa += d; d -= 16;
  b1 = pb[0]; b2 = pb[4 * NB]; pb += 8 * NB;
  for (int k0 = kbase + 8; k0 < j0; k0 += 8) {
    double n1, n2, n3 = 0.0, n4 = 0.0;
    n1 = pa[0]; if (TWO) n3 = pa[drow]; pa += d; d -= 16;
    n2 = pa[0]; if (TWO) n4 = pa[drow]; pa += d; d -= 16;
    const double m1 = pb[0], m2 = pb[4 * NB]; pb += 8 * NB;
    dmma884(c0, c1, a1, b1);
    if (TWO) dmma884(u0, u1, a3, b1);
    dmma884(e0, e1, a2, b2);
    if (TWO) dmma884(v0, v1, a4, b2);
    a1 = n1; a2 = n2; a3 = n3; a4 = n4; b1 = m1; b2 = m2;
  }
  dmma884(c0, c1, a1, b1);
  if (TWO) dmma884(u0, u1, a3, b1);
  dmma884(e0, e1, a2, b2);
  if (TWO) dmma884(v0, v1, a4, b2);
  c0 += e0; c1 += e1; u0 += v0; u1 += v1;
}

// Pre-condition: the lower triangle of M is stored in W.L (packed column-major).
// rhs != nullptr: also computes Sf = (L D)^-1 rhs  (rhs == Sf, in place, is allowed).
// Panels [jbeg, jend) (jbeg a multiple of 8); the columns k < kbase have ALREADY been applied
// to them (by super_update below), the loop applies the columns kbase <= k < j0.
static __device__ __forceinline__ bool factor_panels(int m, Work& W, double beta, double delta,
                                                     const double* rhs, double* Sf, int jbeg, int jend,
                                                     int kbase) {
  bool any_exact = false;
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
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

  for (int j0 = jbeg; j0 < jend; j0 += NB, parity ^= 1) {
    const int nb = min(NB, m - j0);
    const int R = m - j0;
    const int kw = j0 - kbase;                                // columns still to apply
    const int ntile = (R + 7) >> 3;
    int* th = thbuf + parity * 8;
    long long tq = phase_begin(W);
    // P[k][jj] = D_k L(j0+jj, k)
    // (four entries per thread and round: the loads are issued together -- the compiler cannot move a
    // load of L across a store to P on its own, and with L in global memory every round would be
    // one exposed L2 round trip)
    for (int e0 = tid; e0 < kw * NB; e0 += 4 * NT) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int e = e0 + q * NT;
        const int k = kbase + (min(e, kw * NB - 1) >> 3), jj = e & 7;
        v[q] = (jj < nb) ? L[coff(k, m) + j0 + jj] * D[k] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int e = e0 + q * NT;
        if (e < kw * NB) P[e] = v[q];
      }
    }
    if (tid < 8) th[tid] = 0;
    __syncthreads();
    phase_end(W, 7, tq);
    tq = phase_begin(W);

    // ---- step A (+ B on warp 0) ----
    // Measured on B200 (profiles/fp64_latency_r01.txt, dmma_loop_probe_r01.txt): FP64 FMAs
    // queue behind the DMMAs of their own SM sub-partition (a dependent DFMA chain runs 20x
    // slower next to three DMMA warps) and a DMMA queues behind every DMMA already issued on
    // its sub-partition.  So the latency-critical chain  tile 0 -> step B  is kept apart:
    //   A1. the 8x8 diagonal tile and the rhs row are updated first, split-K over warps 0-3
    //       and 4-7 (one warp per sub-partition each), while nobody else issues DMMAs;
    //   A2. warp 0 (alone on sub-partition 0) reduces the partials and eliminates the block
    //       (step B) while the twelve warps of sub-partitions 1-3 update all other row tiles.
    double* part = P + (size_t)2 * m * NB;                   // [8][64] split-K partials
    const int nks = kw >> 2;                                  // k-steps of 4 columns
    if (warp < 8 && (warp < 4 || with_rhs)) {
      const long long t0w = phase_begin(W);
      const int qw = warp & 3;
      const bool is_rhs = warp >= 4;
      const int row = j0 + g;
      const bool ok = is_rhs ? (g == 0) : (row < m);
      const int rs = (row < m) ? row : j0;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      for (int ks = qw; ks < nks; ks += 8) {
        const int ka = kbase + 4 * ks + tg, kb = ka + 16;
        const bool hasb = ks + 4 < nks;
        double a1 = is_rhs ? Sf[ka] : L[coff(ka, m) + rs];
        double a2 = hasb ? (is_rhs ? Sf[kb] : L[coff(kb, m) + rs]) : 0.0;
        const double b1 = P[(ka - kbase) * NB + g];
        const double b2 = hasb ? P[(kb - kbase) * NB + g] : 0.0;
        if (!ok) { a1 = 0.0; a2 = 0.0; }
        dmma884(c0, c1, a1, b1);
        if (hasb) dmma884(e0, e1, a2, b2);
      }
      part[warp * 64 + g * 8 + 2 * tg] = c0 + e0;
      part[warp * 64 + g * 8 + 2 * tg + 1] = c1 + e1;
      if (warp == 0) phase_end(W, 12, t0w);
    }
    __syncthreads();
    const int wsub = warp & 3;
    if (wsub != 0) {
      const long long tw = W.prof ? clock64() : 0;
      const int widx = (warp >> 2) * 3 + wsub - 1;            // 0..11
      for (int rt0 = 1 + widx; rt0 < ntile; rt0 += 24) {
        const int rt1 = rt0 + 12;
        const bool two = rt1 < ntile;
        const int rowa = j0 + 8 * rt0 + g, rowb = j0 + 8 * rt1 + g;
        const bool oka = rowa < m, okb = two && rowb < m;
        // rows past the end read row j0 instead (always valid); a DMMA row only feeds the same
        // row of the result, which is then simply not stored -- no masking in the loop
        const int ra = oka ? rowa : j0, rb = okb ? rowb : j0;
        double c0, c1, u0, u1;
        if (two) panel_update_tiles<true>(L, P, m, kbase, j0, tg, g, ra, rb - ra, c0, c1, u0, u1);
        else panel_update_tiles<false>(L, P, m, kbase, j0, tg, g, ra, 0, c0, c1, u0, u1);
        // L -= acc: all four loads first, then the stores (written as four read-modify-writes the
        // compiler keeps them in order, and with L in global memory that is four L2 round trips
        // per tile pair -- 7 % of all warp samples at config 5)
        {
          const int cA = coff(min(j0 + 2 * tg, m - 1), m), cB = coff(min(j0 + 2 * tg + 1, m - 1), m);
          const bool h0 = 2 * tg < nb, h1 = 2 * tg + 1 < nb;
          double* pa0 = L + cA + (oka ? rowa : j0);
          double* pa1 = L + cB + (oka ? rowa : j0);
          double* pb0 = L + cA + (okb ? rowb : j0);
          double* pb1 = L + cB + (okb ? rowb : j0);
          const double va0 = *pa0, va1 = *pa1, vb0 = *pb0, vb1 = *pb1;
          if (oka && h0) *pa0 = va0 - c0;
          if (oka && h1) *pa1 = va1 - c1;
          if (okb && h0) *pb0 = vb0 - u0;
          if (okb && h1) *pb1 = vb1 - u1;
        }
      }
      if (W.prof) {
        const unsigned long long dt = (unsigned long long)(clock64() - tw);
        if (tid == 32) reinterpret_cast<unsigned long long*>(W.red + RED_PROF)[13] += dt;
      }
    } else if (warp == 0) {
      // reduce the split-K partials of tile 0, apply, eliminate
      // (both loads of L before the stores: with L in global memory two read-modify-writes in
      // program order are two memory round trips on the critical chain)
      const int row = j0 + g;
      const int e0 = g * 8 + 2 * tg, e1 = e0 + 1;
      const double s0 = (part[e0] + part[64 + e0]) + (part[128 + e0] + part[192 + e0]);
      const double s1 = (part[e1] + part[64 + e1]) + (part[128 + e1] + part[192 + e1]);
      const bool w0 = 2 * tg < nb && row < m && row >= j0 + 2 * tg;
      const bool w1 = 2 * tg + 1 < nb && row < m && row >= j0 + 2 * tg + 1;
      double* q0 = L + coff(min(j0 + 2 * tg, m - 1), m) + min(row, m - 1);
      double* q1 = L + coff(min(j0 + 2 * tg + 1, m - 1), m) + min(row, m - 1);
      const double v0 = *q0, v1 = *q1;
      if (w0) *q0 = v0 - s0;
      if (w1) *q1 = v1 - s1;
      // ---- step B: 8x8 diagonal block on warp 0, one matrix entry per lane ----
      __syncwarp();
      long long tb = phase_begin(W);
      diag_block(m, j0, nb, W, delta, th, Wm, D1, rinv);
      phase_end(W, 9, tb);
    } else if (warp == 4 && with_rhs) {
      if (lane < 8 && lane < nb) {
        const int e = lane;   // row g = 0 of the rhs tile
        const double sum = (part[256 + e] + part[320 + e]) + (part[384 + e] + part[448 + e]);
        crhs[lane] = rhs[j0 + lane] - sum;
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
    if (tid >= NT - 8 && tid - (NT - 8) < nb) D[j0 + tid - (NT - 8)] = D1[tid - (NT - 8)];
    {
      int cb[8];
#pragma unroll
      for (int j = 0; j < 8; j++) cb[j] = coff(min(j0 + j, m - 1), m);
      for (int t = tid; t < nthr_rows; t += NT) {
        const bool is_rhs = (t == nbelow);
        const int row = is_rhs ? m - 1 : j0 + 8 + t;
        double c[8];
        // (all eight loads before the first store: eight L2 round trips in parallel, not in series)
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          double cv = (jj < nb) ? L[cb[jj] + row] : 0.0;
          if (is_rhs) cv = (jj < nb) ? crhs[jj] : 0.0;
          c[jj] = cv;
        }
#pragma unroll
        for (int jj = 0; jj < 8; jj++) P[jj * nthr_rows + t] = c[jj];   // kept for the exact redo (conflict-free)
        // right-looking inside the row: l_k = c_k / D_k, then c_jj -= l_k * (D_k L11[jj][k])
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (!is_rhs) hmax[k] = max(hmax[k], dbl_hi(c[k]) & 0x7fffffff);
          const double lk = c[k] * rinv[k];
#pragma unroll
          for (int jj = k + 1; jj < 8; jj++) c[jj] -= lk * Wm[jj * 8 + k];
          c[k] = lk;
        }
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          if (jj < nb) {
            if (is_rhs) Sf[j0 + jj] = c[jj];
            else L[cb[jj] + row] = c[jj];
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
    bool bad;
    {
      const int jj = lane & 7;
      // theta_ub > theta : bump the hi-word by one (covers the dropped low word)
      const double tub = __hiloint2double(th[jj] + 1, 0);
      const bool mine = (jj < nb) && !(tub * tub * inv_beta2 * 1.0000001 <= D1[jj]);
      bad = __any_sync(0xffffffffu, mine);
    }
    if (bad) {
      any_exact = true;
      // restore the updated-but-uneliminated panel and redo it by the sequential rule
      if (tid == 0) {
        for (int i = 0; i < nb; i++)
          for (int j = 0; j <= i; j++) L[cidx(j0 + i, j0 + j, m)] = rawd[i * 8 + j];
      }
      for (int t = tid; t < nthr_rows; t += NT) {
        const bool is_rhs = (t == nbelow);
        const int row = j0 + 8 + t;
        for (int jj = 0; jj < nb; jj++) {
          if (is_rhs) crhs[jj] = P[jj * nthr_rows + t];
          else L[cidx(row, j0 + jj, m)] = P[jj * nthr_rows + t];
        }
      }
      __syncthreads();
      panel_exact(m, j0, nb, W, beta, delta, crhs, Sf, with_rhs);
      __syncthreads();
    }
    phase_end(W, 11, tq);
  }
  return any_exact;
}

static __device__ __forceinline__ void factor_ldl_fast(int m, Work& W, double beta, double delta,
                                                       const double* rhs, double* Sf) {
  factor_panels(m, W, beta, delta, rhs, Sf, 0, m, 0);
}

// ---------------------------------------------------------------------------------------
// factor_ldl_big: the same factorisation for factors that live in GLOBAL memory (m > ~208:
// config 5, m = 500, L = 1 MB; config 4, m = 2000, L = 16 MB).  Left-looking by SUPER-PANELS
// of SB = 64 columns: one GEMM  L(J0.., J0..J0+63) -= L(J0.., 0..J0-1) T,  T[k][c] =
// D_k L(J0+c, k),  brings in every earlier column at once, then the eight 8-column panels of
// the super-panel are factorised by factor_panels with kbase = J0 (columns inside the
// super-panel only).  L(:, k < J0) is thus streamed m/64 times per factorisation instead of
// m/8 times (config 4: 166 MB instead of 1.3 GB per LP and Newton step).
//
// super_update, the GEMM (FP64 tensor cores, DMMA m8n8k4):
//   * a warp owns a unit of 16 rows x 64 columns: 16 accumulator tiles = 64 registers; the A
//     fragment of its two interleaved 8-row tiles (even / odd rows) is ONE 16-byte word per lane
//     and k-step (4 columns x 16 consecutive rows = four 128-byte runs per warp), copied by
//     cp.async into a per-lane ring of 8 slots seven k-steps ahead of its use (28 KB in flight per
//     SM: enough for HBM latency at config 4);
//   * the multiplier table T is built cooperatively in chunks of 32 k-rows, double buffered in
//     shared memory (row stride 68 = 4 mod 16: conflict-free B fragments), one __syncthreads per
//     chunk; the global loads of the next chunk are issued before the DMMAs of the current one;
//   * the rows of a super-panel are processed in passes of 16 units (256 rows); what is left over
//     (and the diagonal block) goes as 16x32 items (SU_ITEM), two warps per row unit, so that
//     partial passes still use all warps -- 16x16 items, four warps per unit, were bound by the
//     operand stream from L2, not by the DMMAs; the right-hand-side row (Sf, kept in place) is
//     accumulated by the threads that build T, four FMAs per chunk each.
// Per k-step a warp issues 16 DMMAs for 512 bytes of L: 8 B/cycle per SM at the DMMA rate,
// well under the 20-35 B/cycle an SM gets from L2.
// ---------------------------------------------------------------------------------------
constexpr int SB = 64;                               // super-panel width
constexpr int SB_KCH = 32;                           // k-rows of T per shared-memory chunk
constexpr int SB_LDT = SB + 4;                       // row stride of T
constexpr int SB_NST = SB_KCH / 4;                   // A-operand ring: one stage per k-step of a chunk
constexpr int SB_RING = SB_NST * 32 * 2;             // doubles per warp (16 bytes per lane and stage)
constexpr int FB_T = 2 * SB_KCH * SB_LDT;            // the two T buffers (34 KB)
static_assert(FB_T + NWARP * SB_RING + SB * SB_LDT + SB + SB / 2 == FB_DOUBLES, "FB_DOUBLES (ipm_types.h) out of date");

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

struct SuT { double2 t0, t1; double dk, sk; };   // one thread's share of a T chunk, in flight

// four 16-byte shared-memory loads 64 bytes apart as ONE statement: left to itself ptxas funnels
// them through one register quad (load, use, load, use ...) and every use waits a full
// shared-memory round trip
__device__ __forceinline__ void lds_f64x2_x4(uint32_t addr, double2 (&w)[4]) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%8];\n\tld.shared.v2.f64 {%2,%3}, [%8+64];\n\t"
               "ld.shared.v2.f64 {%4,%5}, [%8+128];\n\tld.shared.v2.f64 {%6,%7}, [%8+192];"
               : "=d"(w[0].x), "=d"(w[0].y), "=d"(w[1].x), "=d"(w[1].y), "=d"(w[2].x), "=d"(w[2].y),
                 "=d"(w[3].x), "=d"(w[3].y)
               : "r"(addr));
}

// Everything of super_update that is the same for all passes.
struct SuCtx {
  int m, J0, nbw, nch;
  double* L;
  const double* D;
  const double* Sf;
  double* T;
  uint32_t ring;       // shared-memory address of this lane's first ring slot
  double* blk;         // packed nbw x nbw diagonal block (shared memory)
  double* Wm;          // [64][SB_LDT] k-major: Wm[k][j] = -D_k L11[j][k] for j in a later tile than k; diagonal tiles: U_t
  double* rinv;        // [64] 1/D_k
  int* th;             // [64] hi-words of max |c_ik| over the rows below the block
  const Work* W;       // (phase counters)
  int kk, c4;          // T build: this thread's k-row inside a chunk and its four columns
  double r[4];         // right-hand-side row: partial sums of Sf[k] T[k][c4..c4+3]
  int buf;             // T buffer in use
  // T pipeline of the current group of K-loop passes: chunk g (of gtot, contents g mod nch) is
  // in buffer `buf`; with look-ahead 2 (the short chunks of the 16x16 passes) chunk g+1 waits in tn
  int g, gtot;
  bool rhs;            // this group accumulates the right-hand-side row (during its first sweep)
  SuT tn;
};

// chunk `ch` of T: the raw global loads (issued a whole chunk of DMMAs before their use) ...
__device__ __forceinline__ void su_t_load(const SuCtx& c, int ch, bool rhs, SuT& t) {
  const int k = ch * SB_KCH + c.kk;
  t.dk = c.D[k];
  t.sk = rhs ? c.Sf[k] : 0.0;
  const double* src = c.L + coff(k, c.m) + c.J0;
  if (c.nbw == SB) {
    t.t0 = *reinterpret_cast<const double2*>(src + c.c4);
    t.t1 = *reinterpret_cast<const double2*>(src + c.c4 + 2);
  } else {                                           // last super-panel: rows J0+c >= m do not exist
    const int nb1 = c.nbw - 1, c4 = c.c4;
    t.t0.x = src[min(c4, nb1)];     t.t0.y = src[min(c4 + 1, nb1)];
    t.t1.x = src[min(c4 + 2, nb1)]; t.t1.y = src[min(c4 + 3, nb1)];
  }
}
// ... and their use: T[k][c] = D_k L(J0+c, k) into buffer `buf`; the right-hand-side row
// accumulates Sf[k] T[k][c] (sk = 0 when it is not this pass' job)
__device__ __forceinline__ void su_t_store(SuCtx& c, int buf, SuT& t) {
  if (c.nbw != SB) {
    const int c4 = c.c4;
    if (c4 >= c.nbw) t.t0.x = 0.0;
    if (c4 + 1 >= c.nbw) t.t0.y = 0.0;
    if (c4 + 2 >= c.nbw) t.t1.x = 0.0;
    if (c4 + 3 >= c.nbw) t.t1.y = 0.0;
  }
  t.t0.x *= t.dk; t.t0.y *= t.dk; t.t1.x *= t.dk; t.t1.y *= t.dk;
  c.r[0] = fma(t.sk, t.t0.x, c.r[0]); c.r[1] = fma(t.sk, t.t0.y, c.r[1]);
  c.r[2] = fma(t.sk, t.t1.x, c.r[2]); c.r[3] = fma(t.sk, t.t1.y, c.r[3]);
  double* dst = c.T + buf * (SB_KCH * SB_LDT) + c.kk * SB_LDT + c.c4;
  *reinterpret_cast<double2*>(dst) = t.t0;
  *reinterpret_cast<double2*>(dst + 2) = t.t1;
}

static __device__ __forceinline__ bool factor_ldl_ahead_call(int m, double* L, double* D, double* P, double* red,
                                                             unsigned long long* prof, double beta, double delta,
                                                             const double* rhs, double* Sf);

// Row solve of a 16-row x 64-column unit against the factorised 64x64 diagonal block, in
// registers, in the accumulator layout of the DMMAs (cx: rows ra+2g, cy: rows ra+2g+1; tile t
// holds columns 8t+2tg, 8t+2tg+1), and store of the finished columns.  Blocked by 8-column tiles,
// everything on the FP64 tensor pipe:
//   c_t = b_t U_t           U_t = L11(tile t, tile t)^-T (unit upper triangular 8x8, built once per
//                           super-panel into the diagonal tiles of the table) -- the unscaled
//                           entries c_k of the tile's eight columns, all at once;
//   l_t = c_t / D           the finished columns;
//   b_j += l_t WmT(t, j)    for the tiles j > t behind it,  WmT[k][j] = -D_k L11[j][k].
// 4 + 4 (7 - t) DMMAs per tile instead of 64 column steps of a shuffle, a warp reduction, eight
// 16-byte table loads and 18 FMAs on average: DMMA and DFMA run at the same rate (DESIGN.md
// section 4), the row solve was bound by its instruction count and the per-column dependency.
// The accumulator tile becomes the A operand by four shuffles per row tile (entry (g, k) of a
// C tile lives in lane (g, k >> 1), register k & 1).
// rinv[k] = 1/D_k.  hmA/hmB: hi-words of max |c_k| over the rows of this warp for column lane /
// 32+lane (the theta check of the speculative factorisation); vx, vy: row exists.
__device__ __forceinline__ void c_to_a(double c0, double c1, int sl, bool odd, double& a0, double& a1) {
  const unsigned FULL = 0xffffffffu;
  const double p0 = __shfl_sync(FULL, c0, sl), p1 = __shfl_sync(FULL, c1, sl);
  const double q0 = __shfl_sync(FULL, c0, sl + 2), q1 = __shfl_sync(FULL, c1, sl + 2);
  a0 = odd ? p1 : p0;
  a1 = odd ? q1 : q0;
}
__device__ __forceinline__ void trsm_unit(double (&cx)[8][2], double (&cy)[8][2],
                                          const double* __restrict__ WmT,
                                          const double* __restrict__ rinv, bool vx, bool vy,
                                          int& hmA, int& hmB, double* __restrict__ L, int m, int J0,
                                          int r0, bool interior) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int sl = (lane & ~3) | (tg >> 1);            // source lane of A-fragment entry k = tg (k = tg + 4: sl + 2)
  const bool odd = tg & 1;
  const bool up = lane & 16;
  const int hsrc = ((lane & 1) << 4) | ((lane & 7) >> 1);   // a lane that ends up with column lane & 7 of a tile
  constexpr int KH = 4 * SB_LDT * 8;                 // bytes from table row k to row k + 4
  constexpr int TD = (8 * SB_LDT + 8) * 8;           // bytes from one diagonal tile of the table to the next
  const uint32_t wb0 = smem_u32(WmT + tg * SB_LDT + g);
  const double* rq = rinv + 2 * tg;
  // (fully unrolled: the tile counts are compile-time -- a predicated-off DMMA costs its full 16
  // cycles of the pipe, DESIGN.md section 4 -- and no accumulator moves between registers)
#pragma unroll
  for (int tk = 0; tk < 8; tk++) {
    const uint32_t wb = wb0 + tk * TD;
    const double u0 = lds_f64(wb), u1 = lds_f64(wb + KH);
    const double2 rv = *reinterpret_cast<const double2*>(rq + 8 * tk);
    double a0x, a1x, a0y, a1y;
    c_to_a(cx[tk][0], cx[tk][1], sl, odd, a0x, a1x);
    c_to_a(cy[tk][0], cy[tk][1], sl, odd, a0y, a1y);
    double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
    dmma884(x0, x1, a0x, u0);
    dmma884(y0, y1, a0y, u0);
    dmma884(x0, x1, a1x, u1);
    dmma884(y0, y1, a1y, u1);
    // theta check: column maxima of |c| over the sixteen rows (three exchange steps: the first
    // one leaves column 2tg to the lower half-warp and column 2tg+1 to the upper one)
    {
      const int h0 = max(vx ? (dbl_hi(x0) & 0x7fffffff) : 0, vy ? (dbl_hi(y0) & 0x7fffffff) : 0);
      const int h1 = max(vx ? (dbl_hi(x1) & 0x7fffffff) : 0, vy ? (dbl_hi(y1) & 0x7fffffff) : 0);
      int v = max(up ? h1 : h0, __shfl_xor_sync(FULL, up ? h0 : h1, 16));
      v = max(v, __shfl_xor_sync(FULL, v, 8));
      v = max(v, __shfl_xor_sync(FULL, v, 4));
      const int hm = __shfl_sync(FULL, v, hsrc);
      // column 8 tk + (lane & 7): lanes 8 (tk & 3) .. +7 of hmA (tk < 4) / hmB
      if ((lane >> 3) == (tk & 3)) {
        if (tk < 4) hmA = max(hmA, hm);
        else hmB = max(hmB, hm);
      }
    }
    x0 *= rv.x; y0 *= rv.x; x1 *= rv.y; y1 *= rv.y;
    if (tk < 7) {
      // the tiles behind: b_j += l_t (-WmT)(t, j)
      c_to_a(x0, x1, sl, odd, a0x, a1x);
      c_to_a(y0, y1, sl, odd, a0y, a1y);
#pragma unroll
      for (int j0 = tk + 1; j0 < 8; j0 += 4) {
        double b0[4], b1[4];
#pragma unroll
        for (int j = j0; j < j0 + 4 && j < 8; j++) {
          b0[j - j0] = lds_f64(wb + 64 * (j - tk));
          b1[j - j0] = lds_f64(wb + KH + 64 * (j - tk));
        }
#pragma unroll
        for (int j = j0; j < j0 + 4 && j < 8; j++) {
          dmma884(cx[j][0], cx[j][1], a0x, b0[j - j0]);
          dmma884(cy[j][0], cy[j][1], a0y, b0[j - j0]);
          dmma884(cx[j][0], cx[j][1], a1x, b1[j - j0]);
          dmma884(cy[j][0], cy[j][1], a1y, b1[j - j0]);
        }
      }
    }
    // finished tile -> L
    {
      double* q0 = L + coff(J0 + 8 * tk + 2 * tg, m) + r0;
      double* q1 = L + coff(J0 + 8 * tk + 2 * tg + 1, m) + r0;
      if (interior) {
        *reinterpret_cast<double2*>(q0) = make_double2(x0, y0);
        *reinterpret_cast<double2*>(q1) = make_double2(x1, y1);
      } else {
        if (vx) { q0[0] = x0; q1[0] = x1; }
        if (vy) { q0[1] = y0; q1[1] = y1; }
      }
    }
  }
}

// What a pass does with its accumulators.
constexpr int SU_ITEM = 4;  // column tiles of an item of step 1 (2: 16x16, look-ahead 2 for T; 4: 16x32, half the operand bytes per DMMA)
constexpr int SU_RMW = 0;    // 16x16 item below the diagonal block: panel -= acc (global)
constexpr int SU_DIAG = 1;   // 16x16 item of the diagonal block: (panel - acc) -> shared-memory block
constexpr int SU_TRSM = 2;   // 16x64 unit below the diagonal block: (panel - acc), row solve, -> global

// One pass: this warp accumulates a 16-row x (8 NCT)-column item over all k < J0 (kloop) and
// finishes it according to `mode`.  rows ra .. ra+15, columns J0 + col0 ...  first: the
// right-hand-side row is accumulated by the T-build threads during this pass; more: another
// K-loop pass follows (its first T chunk is prefetched during this pass' last one).
template <int NCT>
__device__ __forceinline__ void su_pass(SuCtx& c, int mode, bool active, int ra, int col0, bool kloop) {
  constexpr int LA = (NCT == 2) ? 2 : 1;             // T look-ahead in chunks
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int m = c.m, J0 = c.J0;
  double cx[NCT][2], cy[NCT][2];                     // even-row / odd-row tiles x NCT column tiles
#pragma unroll
  for (int t = 0; t < NCT; t++) cx[t][0] = cx[t][1] = cy[t][0] = cy[t][1] = 0.0;
  if (kloop) {
    // A operand: rows ra+2g, ra+2g+1 of column k0+tg = one 16-byte word per lane and k-step, copied
    // global -> this lane's ring slot by cp.async (no registers held while in flight) SB_NST-1
    // k-steps ahead; every lane reads back exactly the word it copied, so no barrier is needed.
    // (the address is base + offset, rebuilt for every copy: incrementing the pointer the copy
    // was issued with waits until the copy has read it -- measured as a long-scoreboard stall
    // on every k-step)
    const double* abase = c.L + ra + 2 * g;
    int aoff = coff(tg, m);
    int d = 4 * m - 8 - 4 * tg;                      // coff(k + 4) - coff(k); decreases by 16 per k-step
    int ksleft = active ? J0 / 4 : 0;                // k-steps of this pass not yet issued
    auto a_issue = [&](int stage) {
      if (ksleft > 0) {
        const double* src = abase + aoff;
        if (NCT == 8) cp_async16_cg(c.ring + stage * 512, src);
        else cp_async16_ca(c.ring + stage * 512, src);  // four warps share these rows: keep them in L1
        aoff += d; d -= 16; ksleft--;
      }
      cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < SB_NST - 1; s++) a_issue(s);
    // one chunk: fetch T chunk g+LA into `tld`, the DMMAs of chunk g, publish chunk g+1 from `tst`
    // (LA == 1: tst is tld itself; LA == 2: the registers filled one chunk earlier)
    double2 an = make_double2(0.0, 0.0);             // A fragment of the next k-step (in registers)
    if (active) {
      cp_async_wait<SB_NST - 2>();
      an = lds_f64x2(c.ring);
    }
    auto chunk = [&](int ch, SuT& tld, SuT& tst) {
      const int gl = c.g + LA;                       // the chunk to fetch now
      const bool ld = gl < c.gtot;
      if (ld) su_t_load(c, ch + LA >= c.nch ? ch + LA - c.nch : ch + LA, c.rhs && gl < c.nch, tld);
      if (NCT == 8 && active && ch == c.nch - 1) {
        // the panel entries this unit reads right after its K-loop (64 columns x 128 bytes): into
        // L2 during the last chunk, two lines per lane (config 4 streams them from HBM)
        const double* q = c.L + ra;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + coff(J0 + 2 * lane, m)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + coff(J0 + 2 * lane + 1, m)));
      }
      // publish chunk g+1.  Look-ahead 1: half-way through the DMMAs (the loads were issued above),
      // so that the multiplications do not queue behind all of them on the FP64 pipe with the
      // barrier waiting.  Look-ahead 2 (short chunks): after the DMMAs -- storing before them was
      // measured slower at config 4, the words then have only one chunk to arrive from HBM.
      if (LA == 1 && !active && ld) su_t_store(c, c.buf ^ 1, tld);
      if (active) {
        const uint32_t tb = smem_u32(c.T + c.buf * (SB_KCH * SB_LDT) + tg * SB_LDT + col0 + g);
#pragma unroll
        for (int ks = 0; ks < SB_NST; ks++) {
          if (LA == 1 && ks == SB_NST / 2 && ld) su_t_store(c, c.buf ^ 1, tld);
          a_issue((ks + SB_NST - 1) % SB_NST);
          const double2 a = an;
          cp_async_wait<SB_NST - 2>();               // the copy of the NEXT k-step has landed
          an = lds_f64x2(c.ring + ((ks + 1) % SB_NST) * 512);
          if constexpr (NCT == 2) {
            double b0, b1;                           // (one asm: both loads in flight before the DMMAs)
            asm volatile("ld.shared.f64 %0, [%2];\n\tld.shared.f64 %1, [%2+64];"
                         : "=d"(b0), "=d"(b1) : "r"(tb + 8 * (ks * 4 * SB_LDT)));
            dmma884(cx[0][0], cx[0][1], a.x, b0);
            dmma884(cy[0][0], cy[0][1], a.y, b0);
            dmma884(cx[1][0], cx[1][1], a.x, b1);
            dmma884(cy[1][0], cy[1][1], a.y, b1);
          } else {
#pragma unroll
            for (int h = 0; h < NCT; h += 4) {
              double b[4];
              asm volatile("ld.shared.f64 %0, [%4];\n\tld.shared.f64 %1, [%4+64];\n\t"
                           "ld.shared.f64 %2, [%4+128];\n\tld.shared.f64 %3, [%4+192];"
                           : "=d"(b[0]), "=d"(b[1]), "=d"(b[2]), "=d"(b[3])
                           : "r"(tb + 8 * (ks * 4 * SB_LDT + 8 * h)));
#pragma unroll
              for (int t = 0; t < 4; t++) {
                dmma884(cx[h + t][0], cx[h + t][1], a.x, b[t]);
                dmma884(cy[h + t][0], cy[h + t][1], a.y, b[t]);
              }
            }
          }
        }
      }
      if (LA == 2 && c.g + 1 < c.gtot) su_t_store(c, c.buf ^ 1, tst);
      __syncthreads();
      c.buf ^= 1;
      c.g++;
    };
#pragma unroll 1
    for (int ch = 0; ch < c.nch; ch += 2) {          // (nch is even: J0 is a multiple of 64)
      if (LA == 1) {
        SuT t;
        t.t0 = t.t1 = make_double2(0.0, 0.0); t.dk = t.sk = 0.0;
        chunk(ch, t, t);
        chunk(ch + 1, t, t);
      } else {
        // look-ahead 2 without register copies: tn holds chunk g+1 on entry; the two register
        // sets swap roles every chunk
        SuT t;
        t.t0 = t.t1 = make_double2(0.0, 0.0); t.dk = t.sk = 0.0;
        chunk(ch, t, c.tn);                          // fetch g+2 -> t ; publish g+1 from tn
        chunk(ch + 1, c.tn, t);                      // fetch g+3 -> tn ; publish g+2 from t
      }
    }
    cp_async_wait<0>();
  }
  if (!active) return;
  if constexpr (NCT == 8) {
    // ---- SU_TRSM: (panel - acc), row solve against the diagonal block, store ----
    // (rows >= J0 + 64: every column of the super-panel exists for them)
    long long tp = phase_begin(*c.W);
    const bool interior = ra + 16 <= m;
    const int r0 = ra + 2 * g;
    const bool vx = r0 < m, vy = r0 + 1 < m;
    if (interior) {
#pragma unroll
      for (int t = 0; t < 8; t++) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int col = J0 + 8 * t + 2 * tg + hh;
          const double2 v = *reinterpret_cast<const double2*>(c.L + coff(col, m) + r0);
          cx[t][hh] = v.x - cx[t][hh];
          cy[t][hh] = v.y - cy[t][hh];
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; t++) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const double* q = c.L + coff(J0 + 8 * t + 2 * tg + hh, m);
          const double v0 = vx ? q[r0] : 0.0, v1 = vy ? q[r0 + 1] : 0.0;
          cx[t][hh] = v0 - cx[t][hh];
          cy[t][hh] = v1 - cy[t][hh];
        }
      }
    }
    int hmA = 0, hmB = 0;
    phase_end(*c.W, 11, tp);
    tp = phase_begin(*c.W);
    trsm_unit(cx, cy, c.Wm, c.rinv, vx, vy, hmA, hmB, c.L, m, J0, r0, interior);
    if (hmA > 0) atomicMax(&c.th[lane], hmA);
    if (hmB > 0) atomicMax(&c.th[32 + lane], hmB);
    phase_end(*c.W, 12, tp);
  } else if (mode == SU_DIAG) {
    // ---- diagonal block: (M - acc) into the packed nbw x nbw block in shared memory ----
    const int nbw = c.nbw;
#pragma unroll
    for (int t = 0; t < NCT; t++) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int cb = col0 + 8 * t + 2 * tg + (e & 1);          // block-local column
        const int rb = ra - J0 + 2 * g + (e >> 1);               // block-local row
        if (cb < nbw && rb < nbw && rb >= cb) {
          const double v = c.L[coff(J0 + cb, m) + J0 + rb];
          const double acc = (e & 2) ? cy[t][e & 1] : cx[t][e & 1];
          c.blk[packed_off(cb, nbw) + rb] = v - acc;
        }
      }
    }
  } else if (c.nbw == SB && ra + 16 <= m) {
    // ---- SU_RMW, interior item: rows ra+2g, ra+2g+1 of a column are one aligned 16-byte word
    // (four loads, then four stores: read-modify-writes in program order would be
    // serialised L2 round trips)
    double* base = c.L + ra + 2 * g;
#pragma unroll
    for (int t = 0; t < NCT; t += 2) {
      double2* q[4];
      double2 v[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int col = J0 + col0 + 8 * (t + (e >> 1)) + 2 * tg + (e & 1);
        q[e] = reinterpret_cast<double2*>(base + coff(col, m));
        v[e] = *q[e];
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        v[e].x -= cx[t + (e >> 1)][e & 1];
        v[e].y -= cy[t + (e >> 1)][e & 1];
        *q[e] = v[e];
      }
    }
  } else {
    // ---- SU_RMW at the ragged end: element-wise, only i < m ----
#pragma unroll
    for (int t = 0; t < NCT; t++) {
      double* q[4];
      double v[4];
      bool ok[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int cc = col0 + 8 * t + 2 * tg + (e & 1);
        const int row = ra + 2 * g + (e >> 1);
        ok[e] = cc < c.nbw && row < m && row >= J0 + cc;
        q[e] = c.L + coff(min(J0 + cc, m - 1), m) + (ok[e] ? row : m - 1);
        v[e] = *q[e];
      }
      if (ok[0]) *q[0] = v[0] - cx[t][0];
      if (ok[1]) *q[1] = v[1] - cx[t][1];
      if (ok[2]) *q[2] = v[2] - cy[t][0];
      if (ok[3]) *q[3] = v[3] - cy[t][1];
    }
  }
}

// start a group of K-loop passes (`npass` sweeps over the nch chunks, look-ahead LA)
__device__ __forceinline__ void su_t_prime(SuCtx& c, int npass, int LA, bool rhs) {
  c.g = 0; c.gtot = npass * c.nch; c.rhs = rhs;
  SuT t;
  su_t_load(c, 0, rhs, t);
  su_t_store(c, c.buf, t);
  __syncthreads();
  if (LA == 2) su_t_load(c, 1, rhs, c.tn);           // (nch >= 2)
}

// One super-panel (columns J0 .. J0+nbw-1) of factor_ldl_big.  Returns (block-uniform) true if
// the speculation "theta clamp inactive" could not be proven for one of its columns.
//   1. K-loop passes of 16x32 items: the left-over row units below the diagonal block (those that
//      do not fill a pass of sixteen 16x64 units; panel -= acc in global memory) and, in the last
//      of these passes, the items of the diagonal block, which land in shared memory;
//   2. the right-hand-side row (accumulated by the T-build threads) is reduced, the nbw x nbw
//      diagonal block is factorised in shared memory by factor_panels (right-hand side riding
//      along), written to L, and the row-solve tables are built: Wm = -D_k L11[j][k] for j in a
//      later 8-column tile than k, the inverted diagonal tiles U_t in the diagonal tiles, 1/D_k;
//   3. full passes of sixteen 16x64 units below the diagonal block: K-loop, then the row solve
//      against the diagonal block in registers (trsm_unit, on the tensor pipe) -- each entry of the panel is read
//      once and written once; the left-over units of step 1 get their row solve last.
static __device__ __noinline__ bool super_panel(int m, int J0, int nbw, Work& W, double beta, double delta,
                                                double* Sf) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  double* fb = W.fb;
  SuCtx c;
  c.m = m; c.J0 = J0; c.nbw = nbw; c.nch = J0 / SB_KCH;   // (J0 is a multiple of SB)
  c.L = W.L; c.D = W.D; c.Sf = Sf; c.T = fb; c.W = &W;
  c.ring = smem_u32(fb + FB_T + warp * SB_RING) + lane * 16;
  c.blk = fb;                                        // (aliases T: used between the K-loop groups)
  c.Wm = fb + FB_T + NWARP * SB_RING;
  c.rinv = c.Wm + SB * SB_LDT;
  c.th = reinterpret_cast<int*>(c.rinv + SB);
  c.kk = tid >> 4; c.c4 = (tid & 15) * 4;
  c.r[0] = c.r[1] = c.r[2] = c.r[3] = 0.0;
  c.buf = 0; c.g = c.gtot = 0; c.rhs = false;
  c.tn.t0 = c.tn.t1 = make_double2(0.0, 0.0); c.tn.dk = c.tn.sk = 0.0;
  const bool kloop = J0 > 0;
  const int nbelow = (m - J0 - nbw + 15) >> 4;       // 16-row units below the diagonal block
  const int nfull = nbelow >> 4, rem = nbelow & 15;
  const int nru = (nbw + 15) >> 4;                   // row units of the diagonal block
  // items of the diagonal block on or below the diagonal: 16 rows x (8 SU_ITEM) columns each
  const int ndiag = SU_ITEM == 2 ? nru * (nru + 1) / 2 : nru + (nru > 2 ? nru - 2 : 0);
  // (the diagonal items must all sit in the LAST pass -- their epilogue overwrites the T buffers --
  // so the left-over items are padded to a full pass when the two kinds would straddle one)
  constexpr int IPU = 8 / SU_ITEM;                   // items per 16-row unit
  int nrmw = kloop ? IPU * rem : 0;
  if ((nrmw & (NWARP - 1)) + ndiag > NWARP) nrmw = (nrmw + NWARP - 1) & ~(NWARP - 1);
  const int nit = nrmw + ndiag;
  const int nq = (nit + NWARP - 1) / NWARP;
  if (tid < SB) c.th[tid] = 0;
  long long t0 = phase_begin(W);
  // ---- 1. 16 x (8 SU_ITEM) items ----
  if (kloop) su_t_prime(c, nq, SU_ITEM == 2 ? 2 : 1, true);
  for (int q = 0; q < nq; q++) {
    const int it = q * NWARP + warp;
    // item order: left-over units first, diagonal block LAST (its epilogue overwrites T)
    const int id = it - (nit - ndiag);               // index among the diagonal items
    int mode = SU_RMW, ru, cg;
    bool active = it < nit && (id >= 0 || it < IPU * rem);
    if (id >= 0) {
      mode = SU_DIAG;
      if (SU_ITEM == 2) {
        ru = (id >= 6) ? 3 : (id >= 3) ? 2 : (id >= 1) ? 1 : 0;
        cg = id - ru * (ru + 1) / 2;
      } else {                                       // row units 0, 1: one item; 2, 3: two
        ru = id < 2 ? id : 2 + ((id - 2) >> 1);
        cg = id < 2 ? 0 : (id - 2) & 1;
      }
    } else {
      ru = 4 + it / IPU;                             // (below the diagonal block: nbw == 64)
      cg = it % IPU;
    }
    su_pass<SU_ITEM>(c, mode, active, J0 + 16 * ru, 8 * SU_ITEM * cg, kloop);
  }
  __syncthreads();
  phase_end(W, 14, t0);
  t0 = phase_begin(W);
  // ---- 2. right-hand-side row, diagonal block ----
  if (kloop) {
    double* part = fb + FB_T;                        // (the rings are idle)
#pragma unroll
    for (int e = 0; e < 4; e++) part[tid * 4 + e] = c.r[e];
    __syncthreads();
    if (tid < nbw) {
      double sum = 0.0;
#pragma unroll 8
      for (int k = 0; k < SB_KCH; k++) sum += part[(k * 16 + (tid >> 2)) * 4 + (tid & 3)];
      Sf[J0 + tid] -= sum;
    }
    __syncthreads();
  }
  bool bad;
  {
    Work Wb = W;
    Wb.L = c.blk;
    Wb.D = W.D + J0;
    Wb.P = fb + 2176;                                // (packed_doubles(64) = 2128) scratch of 1856 doubles
    Wb.prof = nullptr;
    // the look-ahead factorisation (serial chain on its own warp); speculative like this sweep:
    // a failure is reported to the caller, who redoes the whole factorisation
    bad = factor_ldl_ahead_call(nbw, Wb.L, Wb.D, Wb.P, Wb.red, nullptr, beta, delta, Sf + J0, Sf + J0);
  }
  // block -> L ; tables of the row solve
  for (int e = tid; e < nbw * nbw; e += NT) {
    const int jb = e / nbw, ib = e - jb * nbw;       // column jb, row ib
    if (ib >= jb) {
      const double l = c.blk[packed_off(jb, nbw) + ib];
      c.L[coff(J0 + jb, m) + J0 + ib] = l;
      if ((ib >> 3) > (jb >> 3)) c.Wm[jb * SB_LDT + ib] = -(l * W.D[J0 + jb]);
    }
  }
  // the diagonal tiles of the table: U_t = L11(tile t, tile t)^-T for the row solve (trsm_unit);
  // thread (t, j) solves column j of the inverse = row j of U_t
  if (nbw == SB && tid < SB) {
    const int t8 = tid & ~7, j = tid & 7;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; k++) s = fma(-c.blk[packed_off(t8 + k, nbw) + t8 + i], v[k], s);
      v[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) c.Wm[(t8 + j) * SB_LDT + t8 + i] = v[i];
  }
  if (tid < nbw) c.rinv[tid] = 1.0 / W.D[J0 + tid];
  __syncthreads();
  phase_end(W, 9, t0);
  if (nbelow == 0) return bad;
  t0 = phase_begin(W);
  // ---- 3. 16x64 units below the diagonal block ----
  if (kloop && nfull > 0) su_t_prime(c, nfull, 1, false);
  for (int pss = 0; pss < nfull; pss++) {
    const int u = rem + pss * NWARP + warp;
    su_pass<8>(c, SU_TRSM, true, J0 + SB + 16 * u, 0, kloop);
  }
  if (rem > 0) {
    if (kloop) __syncthreads();                      // (their panel entries were updated by other warps)
    su_pass<8>(c, SU_TRSM, warp < rem, J0 + SB + 16 * warp, 0, false);
  }
  __syncthreads();
  // ---- theta check: the columns' maxima over the rows below the block ----
  {
    const double inv_beta2 = 1.0 / (beta * beta);
    bool mine = false;
    if (tid < nbw) {
      const double tub = __hiloint2double(c.th[tid] + 1, 0);
      mine = !(tub * tub * inv_beta2 * 1.0000001 <= W.D[J0 + tid]);
    }
    bad |= __syncthreads_or(mine) != 0;
  }
  phase_end(W, 10, t0);
  return bad;
}

// Sf <- (L D)^-1 rhs rides along (in place).  Requires W.fb: FB_DOUBLES of shared memory.
// Returns true if the factorisation has to be redone by the sequential rule (factor_ldl_fast).
static __device__ __forceinline__ bool factor_ldl_big(int m, Work& W, double beta, double delta,
                                                      const double* __restrict__ rhs, double* Sf) {
  for (int i = threadIdx.x; i < m; i += NT) Sf[i] = rhs[i];
  __syncthreads();
  bool bad = false;
  for (int J0 = 0; J0 < m; J0 += SB) bad |= super_panel(m, J0, min(SB, m - J0), W, beta, delta, Sf);
  return bad;
}

// ---------------------------------------------------------------------------------------
// factor_ldl_ahead: the same factorisation with the SERIAL CHAIN ON ITS OWN WARP.
//
// The chain  diag_block(p) -> rows of block p+1 solved against it -> diagonal tile p+1 updated
// -> diag_block(p+1)  is the critical path of any LDL' (one warp, ~3k cycles per panel).  Here
// warp 0 runs nothing but that chain, one panel AHEAD of the other fifteen warps, and the two
// sides meet only at two named barriers per panel:
//   warp 0, panel p :  diag_block(p) ; [E3: others finished panel p-1] ; block_row(p) (rows
//                      j1..j1+7 of panel p, gives W' = D L(block p+1, panel p)) ; tile (p+1,p+1)
//                      -= L(block p+1, panel p) W'^T ; [E1: arrive]
//   others, panel p :  [E1] ; step 1, one thread per row i >= j1+8 (and the rhs row): solve the
//                      row against block p, then subtract its product with W' from panel p+1 ;
//                      [O1] theta check ; step 2: multiplier table D_k L(rows of block p+2, k),
//                      k < j1 ; [O2] ; step 3: panel p+2 -= L(:, k<j1) table^T on the FP64 tensor
//                      cores (twelve warps of sub-partitions 1-3, warp 0 keeps sub-partition 0
//                      free of DMMAs, see profiles/fp64_latency_r01.txt) ; [E3: arrive]
// so every panel q receives the columns k < 8(q-1) two panels early (step 3 of panel q-2) and
// the eight columns of panel q-1 one panel early (step 1 of panel q-1 / warp 0).
// The right-hand side rides along as an extra row kept in Sf (in place).  Speculative:
// returns true if some panel could not be proven equivalent to the sequential
// rule (theta clamp active, huge or tied pivot); the caller then redoes the factorisation.
// Requires L in shared memory and m <= 8*35.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void nbar_sync(int id, int nthreads) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int nthreads) {
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// named barriers; the three warp 0 <-> others events are double-buffered by panel parity so
// that an early arrival for panel p+1 can never be counted into the phase of panel p
constexpr int BAR_ED = 1, BAR_EB = 3, BAR_E3 = 5, BAR_O1 = 7, BAR_O2 = 8;

constexpr int PBS = 12;     // row stride of the multiplier tables: the four k-rows of a DMMA B fragment
                            // (8 doubles each) then fall into disjoint shared-memory banks

// warp 0: the eight rows j1..j1+7 of panel j0 against the eliminated block (Wm, rinv), two
// entries per lane (row = lane & 7, columns 2cg and 2cg+1 with cg = lane >> 3), the column-k
// multiplier broadcast by one shuffle per step.  Outputs: the scaled rows to L storage, the
// unscaled ones (D_k l_k) as the 8-column multiplier table Wp[k][row] (stride PBS); then the
// diagonal tile of panel j1 -= L(block, panel j0) Wp (two DMMAs).
static __device__ __forceinline__ void block_row(int m, int j0, int nb, Work& W, const double* Wm,
                                                 const double* rinv, int* th, double* Wp) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  double* __restrict__ L = W.L;
  const int j1 = j0 + 8;
  const int r = lane & 7, cg = lane >> 3;
  const int row = j1 + r;
  const bool valid = row < m;
  const int rs = valid ? row : j0;           // (always a valid address)
  const int ca = coff(min(j0 + 2 * cg, m - 1), m), cb = coff(min(j0 + 2 * cg + 1, m - 1), m);
  double c0 = (valid && 2 * cg < nb) ? L[ca + rs] : 0.0;
  double c1 = (valid && 2 * cg + 1 < nb) ? L[cb + rs] : 0.0;
  int hmine = 0;
  double own0 = 0.0, own1 = 0.0;                      // unscaled entries of this lane's two columns
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double ck = __shfl_sync(FULL, (k & 1) ? c1 : c0, r + 8 * (k >> 1));
    const int hk = __reduce_max_sync(FULL, dbl_hi(ck) & 0x7fffffff);
    if (lane == k) hmine = hk;
    const double lk = ck * rinv[k];
    if (2 * cg > k) c0 = fma(-lk, Wm[(2 * cg) * 8 + k], c0);
    if (2 * cg + 1 > k) c1 = fma(-lk, Wm[(2 * cg + 1) * 8 + k], c1);
    if (cg == (k >> 1)) {                           // the lanes that own column k
      if (k & 1) { own1 = ck; c1 = lk; }
      else { own0 = ck; c0 = lk; }
    }
  }
  // (the unscaled entries are stored after the chain: a store inside it pins the multiplier loads
  // of the later steps behind it)
  Wp[(2 * cg) * PBS + r] = own0;
  Wp[(2 * cg + 1) * PBS + r] = own1;
  if (valid && 2 * cg < nb) L[ca + row] = c0;
  if (valid && 2 * cg + 1 < nb) L[cb + row] = c1;
  if (lane < 8) atomicMax(&th[lane], hmine);          // (the other warps add their rows concurrently)
  __syncwarp();
  {
    const int ra = (j1 + g < m) ? j1 + g : j0;
    double d0 = 0.0, d1 = 0.0;
    dmma884(d0, d1, L[coff(j0 + tg, m) + ra], Wp[tg * PBS + g]);
    dmma884(d0, d1, L[coff(min(j0 + 4 + tg, m - 1), m) + ra], Wp[(4 + tg) * PBS + g]);
    const int r2 = j1 + g;
    const bool w0 = r2 < m && 2 * tg <= g, w1 = r2 < m && 2 * tg + 1 <= g;
    double* q0 = L + coff(min(j1 + 2 * tg, m - 1), m) + (r2 < m ? r2 : m - 1);
    double* q1 = L + coff(min(j1 + 2 * tg + 1, m - 1), m) + (r2 < m ? r2 : m - 1);
    const double v0 = *q0, v1 = *q1;               // (both loads before the stores)
    if (w0) *q0 = v0 - d0;
    if (w1) *q1 = v1 - d1;
  }
  __syncwarp();
}

// Sixteen rows ra..ra+15 against a multiplier table (row stride PBS), columns k0 .. k0+K-1 of
// L.  One 16-byte load per lane brings rows ra+2g and ra+2g+1 of a column -- two interleaved
// 8-row tiles (even rows / odd rows) -- and every column is a run of sixteen consecutive
// doubles, i.e. exactly one conflict-free wavefront (the packed layout keeps column starts
// even).  Two split-K chains per tile; operands of the next step are loaded before the DMMAs
// of the current one.  K = 8, 16, ...  Rows past the end of a column read whatever follows (a
// DMMA row only feeds the same row of the result, which is then not stored).
__device__ __forceinline__ void old_update16(const double* __restrict__ L, const double* __restrict__ PB,
                                             int m, int k0, int K, int tg, int g, int ra, double& c0,
                                             double& c1, double& u0, double& u1) {
  double e0 = 0.0, e1 = 0.0, v0 = 0.0, v1 = 0.0;
  c0 = c1 = u0 = u1 = 0.0;
  const double* pa = L + coff(k0 + tg, m) + ra + 2 * g;
  int d = 4 * m - 8 - 4 * (k0 + tg);            // coff(k + 4) - coff(k); decreases by 16 per step
  const double* pb = PB + tg * PBS + g;
  double2 a1 = *reinterpret_cast<const double2*>(pa); pa += d; d -= 16;
  double2 a2 = *reinterpret_cast<const double2*>(pa); pa += d; d -= 16;
  double b1 = pb[0], b2 = pb[4 * PBS]; pb += 8 * PBS;
  for (int kk = 8; kk < K; kk += 8) {
    const double2 n1 = *reinterpret_cast<const double2*>(pa); pa += d; d -= 16;
    const double2 n2 = *reinterpret_cast<const double2*>(pa); pa += d; d -= 16;
    const double m1 = pb[0], m2 = pb[4 * PBS]; pb += 8 * PBS;
    dmma884(c0, c1, a1.x, b1);
    dmma884(u0, u1, a1.y, b1);
    dmma884(e0, e1, a2.x, b2);
    dmma884(v0, v1, a2.y, b2);
    a1 = n1; a2 = n2; b1 = m1; b2 = m2;
  }
  dmma884(c0, c1, a1.x, b1);
  dmma884(u0, u1, a1.y, b1);
  dmma884(e0, e1, a2.x, b2);
  dmma884(v0, v1, a2.y, b2);
  c0 += e0; c1 += e1; u0 += v0; u1 += v1;
}
// panel jc (nbc columns), rows ra+2g / ra+2g+1  -=  the accumulators of old_update16
__device__ __forceinline__ void sub_unit16(double* __restrict__ L, int m, int jc, int nbc, int ra, int g,
                                           int tg, double c0, double c1, double u0, double u1) {
  // (the four loads first, then the four stores: as four read-modify-writes the compiler keeps them
  // in order -- four shared-memory round trips in series on a phase that is a few hundred cycles long)
  const int rowA = ra + 2 * g, rowB = rowA + 1;
  const int col0 = 2 * tg, col1 = col0 + 1;
  double* d0 = L + coff(min(jc + col0, m - 1), m);
  double* d1 = L + coff(min(jc + col1, m - 1), m);
  const bool a0 = col0 < nbc && rowA < m && rowA >= jc + col0, b0 = col0 < nbc && rowB < m && rowB >= jc + col0;
  const bool a1 = col1 < nbc && rowA < m && rowA >= jc + col1, b1 = col1 < nbc && rowB < m && rowB >= jc + col1;
  const int sA = rowA < m ? rowA : m - 1, sB = rowB < m ? rowB : m - 1;   // (always valid addresses)
  const double vA0 = d0[sA], vB0 = d0[sB], vA1 = d1[sA], vB1 = d1[sB];
  if (a0) d0[rowA] = vA0 - c0;
  if (b0) d0[rowB] = vB0 - u0;
  if (a1) d1[rowA] = vA1 - c1;
  if (b1) d1[rowB] = vB1 - u1;
}

// Step 1a of factor_ldl_ahead for one row (two threads per row do the same work, thread
// `half == 0` stores): solve the row against the eliminated block.  tab != nullptr: the row
// belongs to block p+2, its unscaled entries D_k l_k are the new columns of that block's
// multiplier table.
template <bool FULL>
__device__ __forceinline__ void step1_solve(double* __restrict__ L, int m, int j0, int nb, int row, int half,
                                            const double* __restrict__ Wm, const double* __restrict__ rinv,
                                            double* __restrict__ tab, int (&hmax)[8]) {
  // offsets of the panel's columns (block-uniform, incremental: coff(j+1) - coff(j) = m-1-j
  // rounded up to even; j0 is a multiple of 8)
  int cb[8];
  cb[0] = coff(j0, m);
  const int e = 1 - (m & 1);
#pragma unroll
  for (int jj = 1; jj < 8; jj++) cb[jj] = cb[jj - 1] + (m - j0 - jj) + ((jj - 1 + e) & 1);
  double* __restrict__ Lr = L + row;
  double c[8];
#pragma unroll
  for (int jj = 0; jj < 8; jj++) {
    if (!FULL) cb[jj] = coff(min(j0 + jj, m - 1), m);
    c[jj] = (FULL || jj < nb) ? Lr[cb[jj]] : 0.0;
  }
  double un[8];                                 // unscaled entries D_k l_k (table rows of block p+2)
#pragma unroll
  for (int k = 0; k < 8; k++) {
    hmax[k] = max(hmax[k], dbl_hi(c[k]) & 0x7fffffff);
    un[k] = c[k];
    const double lk = c[k] * rinv[k];
#pragma unroll
    for (int jj = k + 1; jj < 8; jj++) c[jj] -= lk * Wm[jj * 8 + k];
    c[k] = lk;
  }
  // (stores after the chain: a store inside it would pin the multiplier loads of the later steps
  // behind it, and this warp is the one everybody waits for)
  if (tab != nullptr && half == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++) tab[k * PBS] = un[k];
  }
  if (half == 0) {
#pragma unroll
    for (int jj = 0; jj < 8; jj++)
      if (FULL || jj < nb) Lr[cb[jj]] = c[jj];
  }
}

// 1/d for delta <= d < 1e30: MUFU.RCP64H seed (>= 20 bits, no conversions) + two Newton steps.
__device__ __forceinline__ double rcp_pos64(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Warp 0 of factor_ldl_ahead on a FULL panel (eight columns, j0 % 8 == 0): the serial chain with
// every value of the 8x8 block in the registers of EVERY lane.  A lone warp is bound by its
// instruction count and by the latency of whatever it waits for; the shuffle-based diag_block
// spends ten SHFLs per pivot on moving entries between lanes.  Here all lanes run the same
// scalar elimination (84 FMAs + 28 multiplications + 8 reciprocals, ~95 dependent cycles per
// pivot), so nothing moves; lane 0 stores.  The block row (rows j1..j1+7 against the block) then
// uses the multipliers straight from those registers, one row per group of four lanes, and feeds
// the two DMMAs of the diagonal-tile update from registers as well.
//   blk: Wm[64] (unscaled entries D_k l_ik), D1 at +64, rinv at +72 -- read by the other warps
// ROW: rows j1..j1+7 all exist (m - j1 >= 8) and are solved here; otherwise the caller runs block_row.
template <bool ROW>
static __device__ __forceinline__ void chain_panel8(int m, int j0, int par, bool wait_e3, Work& W, double delta,
                                                    int* th, double* __restrict__ blk, double* __restrict__ Wp) {
  const int lane = threadIdx.x & 31;
  double* __restrict__ L = W.L;
  const long long tb = phase_begin(W);
  // offsets of (row j0, column j0 + jj); coff(j+1) - coff(j) = m-1-j rounded up to even
  int cb[8];
  cb[0] = coff(j0, m) + j0;
  const int ev = 1 - (m & 1);
#pragma unroll
  for (int jj = 1; jj < 8; jj++) cb[jj] = cb[jj - 1] + (m - j0 - jj) + ((jj - 1 + ev) & 1);
  double e[8][8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (j & 1) e[j][j] = L[cb[j] + j];
#pragma unroll
    for (int i = j + (j & 1); i < 8; i += 2) {      // (cb even, i even: 16-byte aligned)
      const double2 v = *reinterpret_cast<const double2*>(L + cb[j] + i);
      e[i][j] = v.x;
      e[i + 1][j] = v.y;
    }
  }
  const int hdelta = dbl_hi(delta);
  const bool st = lane == 0;
  bool big = false;
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    // D_j = max(|pivot|, delta) decided on the high words (ties, NaN and huge pivots -> exact path)
    const int hp = dbl_hi(e[j][j]) & 0x7fffffff;
    big |= (hp == hdelta) | (hp >= 0x46293e59);                  // 0x46293e59 ~ hi word of 1e30
    const double Dv = (hp < hdelta) ? delta : fabs(e[j][j]);
    r[j] = rcp_pos64(Dv);
    int h = 0;
    double l[8];
#pragma unroll
    for (int i = j + 1; i < 8; i++) {
      h = max(h, dbl_hi(e[i][j]) & 0x7fffffff);
      l[i] = e[i][j] * r[j];
    }
#pragma unroll
    for (int i = j + 1; i < 8; i++)
#pragma unroll
      for (int k = j + 1; k <= i; k++) e[i][k] = fma(-l[i], e[k][j], e[i][k]);
    // scaled column j of the unit-lower block back to L storage; D, 1/D, in-block theta
    if (st) {
      blk[64 + j] = Dv;
      blk[72 + j] = r[j];
      th[j] = h;
      l[j] = 1.0;
      if (j & 1) L[cb[j] + j] = 1.0;
#pragma unroll
      for (int i = j + (j & 1); i < 8; i += 2)
        *reinterpret_cast<double2*>(L + cb[j] + i) = make_double2(l[i], l[i + 1]);
    }
  }
  if (st) {
    // Wm[i][k] = D_k l_ik = the unscaled entry (k < i); entries k >= i of a row are never read
#pragma unroll
    for (int i = 1; i < 8; i++)
#pragma unroll
      for (int k = 0; k < i; k += 2)
        *reinterpret_cast<double2*>(blk + i * 8 + k) = make_double2(e[i][k], (k + 1 < i) ? e[i][k + 1] : 0.0);
    if (big) th[0] = 0x7ff00000;
  }
  __syncwarp();
  if (lane < 8) W.D[j0 + lane] = blk[64 + lane];
  nbar_arrive(BAR_ED + par, NT);                  // block p eliminated
  phase_end(W, 9, tb);
  const long long tw = phase_begin(W);
  if (wait_e3) nbar_sync(BAR_E3 + (par ^ 1), NT); // others are done with panel p-1
  phase_end(W, 12, tw);
  const long long tr = phase_begin(W);
  if (ROW) {
    const unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, tg = lane & 3;
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = L[cb[k] + 8 + g];
    int hmine = 0;
    double b1 = 0.0, b2 = 0.0;                        // B fragments of the tile update: unscaled entries k = tg, 4 + tg
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int hk = __reduce_max_sync(FULL, dbl_hi(c[k]) & 0x7fffffff);
      if (lane == k) hmine = hk;
      if (tg == 0) Wp[k * PBS + g] = c[k];            // unscaled: the 8-column multiplier table W'
      if (tg == (k & 3)) {
        if (k < 4) b1 = c[k];
        else b2 = c[k];
      }
      const double lk = c[k] * r[k];
#pragma unroll
      for (int jj = k + 1; jj < 8; jj++) c[jj] = fma(-lk, e[jj][k], c[jj]);
      c[k] = lk;
    }
    if (tg == 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) L[cb[k] + 8 + g] = c[k];
    }
    if (lane < 8) atomicMax(&th[lane], hmine);          // (the other warps add their rows concurrently)
    // diagonal tile of panel j1 -= L(block, panel j0) Wp^T: fragments from this lane's row
    const double a1 = (tg == 0) ? c[0] : (tg == 1) ? c[1] : (tg == 2) ? c[2] : c[3];
    const double a2 = (tg == 0) ? c[4] : (tg == 1) ? c[5] : (tg == 2) ? c[6] : c[7];
    double d0 = 0.0, d1 = 0.0;
    dmma884(d0, d1, a1, b1);
    dmma884(d0, d1, a2, b2);
    // (row j1+g, columns j1+2tg, j1+2tg+1): cb of the next panel's columns
    const int j1 = j0 + 8;
    const int q0i = coff(j1 + 2 * tg, m) + j1 + g, q1i = coff(j1 + 2 * tg + 1, m) + j1 + g;
    const double v0 = L[q0i], v1 = L[q1i];             // (both loads before the stores)
    if (2 * tg <= g) L[q0i] = v0 - d0;
    if (2 * tg + 1 <= g) L[q1i] = v1 - d1;
    __syncwarp();
  }
  if (ROW) {
    nbar_arrive(BAR_EB + par, NT);                  // W' of panel p published
    phase_end(W, 8, tr);
  }
}

static __device__ __forceinline__ bool factor_ldl_ahead(int m, Work& W, double beta, double delta,
                                                        const double* rhs, double* __restrict__ Sf) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int g = lane >> 2, tg = lane & 3;
  double* __restrict__ L = W.L;
  double* __restrict__ D = W.D;
  const int TBL = PBS * max(8, m - 8);                // one multiplier table [k][PBS], k < j1 <= m-9
  double* xtra = W.P + 2 * TBL;                       // 512 more doubles of the work area
  // per-parity block data at xtra + 80 par: Wm[64], D1[8], rinv[8]
  double* Wp = xtra + 160;                            // D_k L(block p+1, panel p)   [k][PBS]
  int* thbuf = reinterpret_cast<int*>(W.red + RED_TH);
  const double inv_beta2 = 1.0 / (beta * beta);
  const int np = (m + 7) >> 3;
  const int wsub = warp & 3;
  const bool isK = wsub != 0;
  const int widx = (warp >> 2) * 3 + wsub - 1;        // 0..11 for the DMMA warps
  const int NOTH = NT - 32;                           // threads other than warp 0
  bool bad = false;

  if (rhs != Sf)
    for (int i = tid; i < m; i += NT) Sf[i] = rhs[i];
  if (tid < 16) thbuf[tid] = 0;
  __syncthreads();

  if (warp == 0) {
    // ------------------------------ the serial chain ------------------------------
    for (int p = 0; p < np; p++) {
      const int j0 = 8 * p, nb = min(NB, m - j0), par = p & 1;
      int* th = thbuf + par * 8;
      double* blk = xtra + 80 * par;
#ifdef PB200_NEW_CHAIN   // (measured in situ at config 3: factor 153 k cycles per step with it, 148 k without -- the other warps are the bottleneck of most panels)
      if (nb == NB && m - j0 >= 16) {                 // full panel, full block row below it
        chain_panel8<true>(m, j0, par, p > 0, W, delta, th, blk, Wp);
        continue;
      }
      if (nb == NB) {
        chain_panel8<false>(m, j0, par, p > 0, W, delta, th, blk, Wp);
      } else
#endif
      {
        const long long tb = phase_begin(W);
        diag_block(m, j0, nb, W, delta, th, blk, blk + 64, blk + 72);
        __syncwarp();
        if (lane < nb) D[j0 + lane] = blk[64 + lane];
        nbar_arrive(BAR_ED + par, NT);                  // block p eliminated
        phase_end(W, 9, tb);
        const long long tw = phase_begin(W);
        if (p > 0) nbar_sync(BAR_E3 + (par ^ 1), NT);   // others are done with panel p-1
        phase_end(W, 12, tw);
      }
      const long long tr = phase_begin(W);
      if (j0 + 8 < m) block_row(m, j0, nb, W, blk, blk + 72, th, Wp);
      nbar_arrive(BAR_EB + par, NT);                  // W' of panel p published
      phase_end(W, 8, tr);
    }
  } else {
    // ------------------------------ everything else -------------------------------
    // step 1: DMMA warp widx owns the sixteen rows j2 + 16 widx .. +15, two threads per row
    const int r = widx * 16 + (lane >> 1), half = lane & 1;
    for (int p = 0; p < np; p++) {
      const int j0 = 8 * p, j1 = j0 + 8, j2 = j0 + 16, j3 = j0 + 24;
      const int nb = min(NB, m - j0), par = p & 1;
      const double* blk = xtra + 80 * par;
      const double* Wm = blk;
      const double* rinv = blk + 72;
      int* th = thbuf + par * 8;
      double* PBcur = W.P + par * TBL;                // table of panel p+2 (used in step 3)
      double* PBnxt = W.P + (par ^ 1) * TBL;          // table of panel p+3 (old part built here)
      const int nrows = max(0, m - j2);
      const bool full = (nb == NB) && (m - j1 >= NB || j1 >= m);
      long long tq = phase_begin(W);
      nbar_sync(BAR_ED + par, NT);
      phase_end(W, 7, tq, 32);
      tq = phase_begin(W);
      // ---- step 1a: solve the rows below block p+1 against block p; rhs row on warp 4 ----
      double crhs[8];
      if (isK) {
        int hmax[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) hmax[jj] = 0;
        if (r < nrows) {
          double* tab = (r < 8) ? PBcur + j0 * PBS + r : nullptr;   // rows of block p+2
          if (full) step1_solve<true>(L, m, j0, nb, j2 + r, half, Wm, rinv, tab, hmax);
          else step1_solve<false>(L, m, j0, nb, j2 + r, half, Wm, rinv, tab, hmax);
        }
        int hm = 0;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          const int rr = __reduce_max_sync(0xffffffffu, hmax[jj]);
          if (lane == jj) hm = rr;
        }
        if (lane < 8 && hm > 0) atomicMax(&th[lane], hm);
      } else if (warp == 4) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) crhs[jj] = (jj < nb) ? Sf[j0 + jj] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double lk = crhs[k] * rinv[k];
#pragma unroll
          for (int jj = k + 1; jj < 8; jj++) crhs[jj] -= lk * Wm[jj * 8 + k];
          crhs[k] = lk;
        }
        __syncwarp();
        if (lane < 8) {
          double mine = crhs[0];
#pragma unroll
          for (int jj = 1; jj < 8; jj++) mine = (lane == jj) ? crhs[jj] : mine;
          if (lane < nb) Sf[j0 + lane] = mine;
        }
      }
      nbar_sync(BAR_O1, NOTH);                        // all rows solved, table of panel p+2 complete
      phase_end(W, 10, tq, 32);
      tq = phase_begin(W);
      // ---- theta check of panel p (all contributions are in: warp 0's came before BAR_ED..EB
      // of this panel at the latest -- checked again after BAR_EB below) ----
      if (j2 < m) {
        if (isK) {
          // ---- step 2: panel p+2 -= L(:, k<j1) table^T ; unit u = sixteen rows j2+16u.. ----
          if (16 * widx < nrows) {
            const int ra = j2 + 16 * widx;
            double c0, c1, u0, u1;
            old_update16(L, PBcur, m, 0, j1, tg, g, ra, c0, c1, u0, u1);
            sub_unit16(L, m, j2, min(NB, m - j2), ra, g, tg, c0, c1, u0, u1);
          }
          // ---- the old part of the NEXT table (panel p+3): D_k L(j3+jj, k), k < j1 -- by the twelve
          // DMMA warps (those without a row unit get to it first), NOT by the three spare warps of
          // sub-partition 0: anything that runs there takes issue slots and shared-memory bandwidth
          // from the chain warp (tools/chain_probe.cu: +35 % per panel) ----
          if (j3 < m) {
            const int total = j1 * NB;
            for (int e = widx * 32 + lane; e < total; e += 384) {
              const int k = e >> 3, jj = e & 7;
              PBnxt[k * PBS + jj] = (j3 + jj < m) ? L[coff(k, m) + j3 + jj] * D[k] : 0.0;
            }
          }
        } else {
          if (warp == 8) {
            // rhs row: Sf[j2+jj] -= sum_{k<j1} Sf[k] table[k][jj]; lane = (k slice, jj)
            const int jj = lane & 7, sl = lane >> 3;
            double acc = 0.0;
            for (int k = sl; k < j1; k += 4) acc += Sf[k] * PBcur[k * PBS + jj];
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane < 8 && j2 + lane < m) Sf[j2 + lane] -= acc;
          }
        }
      }
      phase_end(W, 13, tq, 32);
      tq = phase_begin(W);
      nbar_sync(BAR_EB + par, NT);                    // W' of panel p (warp 0 published it long ago)
      phase_end(W, 14, tq, 32);
      tq = phase_begin(W);
      {
        const int jj = lane & 7;
        const double tub = __hiloint2double(th[jj] + 1, 0);
        const bool mine = (jj < nb) && !(tub * tub * inv_beta2 * 1.0000001 <= blk[64 + jj]);
        bad |= __any_sync(0xffffffffu, mine);
      }
      // ---- step 3: panel p+1 -= (solved rows) W' : each DMMA warp for its own sixteen rows ----
      if (j1 < m) {
        if (isK) {
          if (16 * widx < nrows) {
            double c0, c1, u0, u1;
            old_update16(L, Wp, m, j0, 8, tg, g, j2 + 16 * widx, c0, c1, u0, u1);
            sub_unit16(L, m, j1, min(NB, m - j1), j2 + 16 * widx, g, tg, c0, c1, u0, u1);
          }
        } else if (warp == 4) {
          if (lane < 8 && j1 + lane < m) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) sacc += crhs[k] * Wp[k * PBS + lane];
            Sf[j1 + lane] -= sacc;
          }
        }
      }
      phase_end(W, 15, tq, 32);
      if (p + 1 < np) nbar_arrive(BAR_E3 + par, NT);
    }
  }
  return __syncthreads_or(bad) != 0;
}

// Out-of-line entry (like the SYRK): inlined into the persistent kernel the chain warp's code is
// at the mercy of the register allocation of everything around it (measured: +25 % on the 8x8
// pivot blocks after unrelated code was added to the kernel); as a real function it gets its own.
// SH: every pointer is into shared memory.  Generic pointers arriving through the call would make
// each access a generic LD/ST (ncu: 67 % of the row solve's stall samples on the long scoreboard);
// re-deriving them from the dynamic shared-memory base lets the compiler emit LDS/STS.
template <bool SH>
static __device__ __noinline__ bool factor_ldl_ahead_call_t(int m, double* L, double* D, double* P, double* red,
                                                            unsigned long long* prof, double beta, double delta,
                                                            const double* rhs, double* Sf) {
  Work W;
  if (SH) {
    extern __shared__ __align__(16) double dyn_smem[];
    const uint32_t b0 = smem_u32(dyn_smem);
    W.L = dyn_smem + ((smem_u32(L) - b0) >> 3);
    W.D = dyn_smem + ((smem_u32(D) - b0) >> 3);
    W.P = dyn_smem + ((smem_u32(P) - b0) >> 3);
    W.red = dyn_smem + ((smem_u32(red) - b0) >> 3);
    rhs = dyn_smem + ((smem_u32(rhs) - b0) >> 3);
    Sf = dyn_smem + ((smem_u32(Sf) - b0) >> 3);
  } else {
    W.L = L; W.D = D; W.P = P; W.red = red;
  }
  W.prof = prof;
  return factor_ldl_ahead(m, W, beta, delta, rhs, Sf);
}
static __device__ __forceinline__ bool factor_ldl_ahead_call(int m, double* L, double* D, double* P, double* red,
                                                             unsigned long long* prof, double beta, double delta,
                                                             const double* rhs, double* Sf) {
  // (block-uniform: D and Sf are in the block's global scratch slot when the vectors do not fit on chip)
  if (__isShared(D) && __isShared(Sf) && __isShared(rhs))
    return factor_ldl_ahead_call_t<true>(m, L, D, P, red, prof, beta, delta, rhs, Sf);
  return factor_ldl_ahead_call_t<false>(m, L, D, P, red, prof, beta, delta, rhs, Sf);
}

// S <- L^-T S ; dy += S     (second half of ldl.cl:529-536), blocks of 32 columns:
// the part of each dot product below the block is a warp-per-column reduction on all
// warps, the 32x32 triangle is back-substituted by warp 0 in registers with shuffles.
// GLOB: L lives in global memory (a load is an L2 / HBM round trip): the two columns of a warp
// together, four loads each in flight, instead of one load per loop iteration.
template <bool GLOB = false>
static __device__ __forceinline__ void back_solve_fast(int m, Work& W, double sign = 1.0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const double* __restrict__ L = W.L;
  double* __restrict__ S = W.S;
  const int nblk = (m + 31) >> 5;
  for (int kb = nblk - 1; kb >= 0; kb--) {
    const int c0 = kb << 5;
    const int c1 = min(m, c0 + 32);
    if (c1 < m) {
      if constexpr (GLOB) {
        const int ja = c0 + warp, jb = ja + NWARP;     // (ja < c1: a block below the last one is full)
        const bool hb = jb < c1;
        const double* ca = L + cidx(ja, ja, m) - ja;   // ca[i] = L(i, ja)
        const double* cb = hb ? L + cidx(jb, jb, m) - jb : ca;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int i = c1 + lane;
#pragma unroll 2
        for (; i + 96 < m; i += 128) {
          const double la0 = ca[i], la1 = ca[i + 32], la2 = ca[i + 64], la3 = ca[i + 96];
          const double lb0 = cb[i], lb1 = cb[i + 32], lb2 = cb[i + 64], lb3 = cb[i + 96];
          const double s0 = S[i], s1 = S[i + 32], s2 = S[i + 64], s3 = S[i + 96];
          a0 = fma(la0, s0, a0); a1 = fma(la1, s1, a1); a2 = fma(la2, s2, a2); a3 = fma(la3, s3, a3);
          b0 = fma(lb0, s0, b0); b1 = fma(lb1, s1, b1); b2 = fma(lb2, s2, b2); b3 = fma(lb3, s3, b3);
        }
        for (; i < m; i += 32) {
          const double la = ca[i], lb = cb[i], si = S[i];
          a0 = fma(la, si, a0); b0 = fma(lb, si, b0);
        }
        const double sa = warp_sum((a0 + a1) + (a2 + a3)), sb = warp_sum((b0 + b1) + (b2 + b3));
        if (lane == 0) {
          if (ja < c1) S[ja] -= sa;
          if (hb) S[jb] -= sb;
        }
      } else {
        for (int j = c0 + warp; j < c1; j += NWARP) {
          const double* col = L + cidx(j, j, m) - j;     // col[i] = L(i, j)
          double acc = 0.0;
          for (int i = c1 + lane; i < m; i += 32) acc += col[i] * S[i];
          acc = warp_sum(acc);
          if (lane == 0) S[j] -= acc;
        }
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
        W.dy[j] += sign * w;
      }
    }
    __syncthreads();
  }
}

// S <- (L D)^-1 S   (first half, ldl.cl:519-527); only used by refinement passes -- the
// first solve of every iteration gets this from factor_ldl_fast.
static __device__ __forceinline__ void fwd_solve_fast(int m, Work& W) {
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
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
