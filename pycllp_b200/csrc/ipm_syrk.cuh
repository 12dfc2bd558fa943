// ipm_syrk.cuh -- M = A diag(d) A' on the FP64 tensor cores, operand staged by TMA.
//
// Only the nd columns of A with >= 2 non-zeros take part (singleton/slack columns add to the
// diagonal afterwards).  The packed operand lives in global memory pre-tiled exactly as it
// is used: chunk ch = SY_KC consecutive packed columns, stored k-major with the m rows padded
// to SY ldm (ldm = 4 mod 16 makes every DMMA fragment load bank-conflict free), so one
// cp.async.bulk (TMA, 1-D) brings a whole chunk into shared memory; SY_STAGES stages, one
// mbarrier each; the last warp to finish a stage refills it.  The lower triangle of M is cut into 8x8 tiles; a warp owns up to SY_SEG
// "segments" per pass (one tile row x up to SY_CW consecutive tile columns) and keeps their
// accumulators in registers for the whole K loop, so every A fragment is loaded once per
// SY_CW DMMAs and the row fragment is the one scaled by d (one DMUL per segment).  The
// result goes straight into the factorisation's packed storage W.L and (both triangles)
// into W.M for the residual.
#pragma once
#include <type_traits>
#include <cstdio>

namespace pb200 {

constexpr int RED_MBAR = 216;   // SY_STAGES 8-byte mbarriers + SY_STAGES int counters inside W.red (216..223)

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Spin on the barrier's phase.  (C-level loop + __syncwarp so that the lanes are converged
// again before the warp-collective code that follows.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
  __syncwarp();
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// The TMA ring (SY_STAGES stages in W.P, one mbarrier + one counter each) is initialised ONCE
// per kernel: re-initialising a live mbarrier object is undefined behaviour (seen: a hang at
// one call site).  Every user continues the global chunk count W.ring_g, which fixes stage
// (g mod SY_STAGES) and phase parity ((g / SY_STAGES) & 1) of every chunk.
__device__ __forceinline__ void ring_init(Work& W) {
  uint64_t* full = reinterpret_cast<uint64_t*>(W.red + RED_MBAR);
  int* done = reinterpret_cast<int*>(full + SY_STAGES);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < SY_STAGES; s++) { mbar_init(&full[s], 1); done[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  W.ring_g = 0;
  __syncthreads();
}
// shared-memory counter: release this thread's (warp's, after __syncwarp) earlier accesses,
// acquire those of the threads that incremented before
__device__ __forceinline__ int atom_add_acq_rel(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}

// The DMMAs of one staged chunk for one warp.  Measured on B200 (tools/syrk_loop_probe.cu): a
// PREDICATED-OFF DMMA still occupies the FP64 tensor pipe for its full 16 cycles, so the tile
// counts must be compile-time or branched on, never predicated.  N2 >= 0: segments 0 and 1
// hold SY_CW tiles each and segment 2 holds N2 (the shape the host-side balancer produces
// for all but tiny problems) -> straight-line code.  N2 < 0: any shape; full segments take
// an unpredicated path, partial ones a predicated one.
template <int N2>
__device__ __forceinline__ void syrk_chunk(double (&acc)[SY_SEG][SY_CW][2], const double* __restrict__ S,
                                           const double* __restrict__ dgc, int ldm, int g, int tg,
                                           const int (&segI)[SY_SEG], const int (&segJ)[SY_SEG],
                                           const int (&segN)[SY_SEG]) {
#pragma unroll
  for (int ks = 0; ks < SY_KC / 4; ks++) {
    const double dk = dgc[ks * 4 + tg];
    const double* __restrict__ col = S + (ks * 4 + tg) * ldm + g;
    if constexpr (N2 >= 0) {
      // explicit software pipeline (volatile loads keep their program order): ptxas otherwise
      // tends to funnel all B fragments through ONE register pair, load -> DMMA -> load ..., which
      // exposes the shared-memory latency on every DMMA
      const uint32_t cb = smem_u32(col);
      double as[SY_SEG], b0[SY_CW], b1[SY_CW], b2[SY_CW];
#pragma unroll
      for (int s = 0; s < SY_SEG; s++) as[s] = lds_f64(cb + 64 * segI[s]);   // (empty segment: tile 0)
#pragma unroll
      for (int t = 0; t < SY_CW; t++) b0[t] = lds_f64(cb + 64 * (segJ[0] + t));
#pragma unroll
      for (int t = 0; t < SY_CW; t++) b1[t] = lds_f64(cb + 64 * (segJ[1] + t));
      as[0] *= dk;
#pragma unroll
      for (int t = 0; t < SY_CW; t++) dmma884(acc[0][t][0], acc[0][t][1], as[0], b0[t]);
#pragma unroll
      for (int t = 0; t < SY_CW; t++)
        if (t < N2) b2[t] = lds_f64(cb + 64 * (segJ[2] + t));
      as[1] *= dk;
#pragma unroll
      for (int t = 0; t < SY_CW; t++) dmma884(acc[1][t][0], acc[1][t][1], as[1], b1[t]);
      if (N2 > 0) as[2] *= dk;
#pragma unroll
      for (int t = 0; t < SY_CW; t++)
        if (t < N2) dmma884(acc[2][t][0], acc[2][t][1], as[2], b2[t]);
    } else {
#pragma unroll
      for (int s = 0; s < SY_SEG; s++) {
        if (segN[s] == SY_CW) {
          const double as = col[8 * segI[s]] * dk;
#pragma unroll
          for (int t = 0; t < SY_CW; t++) dmma884(acc[s][t][0], acc[s][t][1], as, col[8 * (segJ[s] + t)]);
        } else if (segN[s] > 0) {
          const double as = col[8 * segI[s]] * dk;
#pragma unroll
          for (int t = 0; t < SY_CW; t++) {
            if (t < segN[s]) dmma884(acc[s][t][0], acc[s][t][1], as, col[8 * (segJ[s] + t)]);
          }
        }
      }
    }
  }
}

static __device__ __forceinline__ void form_M_dense_tma(const Matrix& A, Work& W) {
  const int m = A.m, ldm = A.sy_ldm;
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_id();
  const int g = lane >> 2, tg = lane & 3;
  const int nch = A.ldd / SY_KC;                     // chunks per pass
  const int total = nch * A.sy_npass;                // chunks streamed in all (A is re-read per pass)
  const uint32_t chunk_bytes = (uint32_t)(SY_KC * ldm * sizeof(double));
  const int stage_doubles = SY_KC * ldm;
  uint64_t* full = reinterpret_cast<uint64_t*>(W.red + RED_MBAR);
  // done[s]: warps that have finished with the chunk in stage s.  The LAST warp to finish
  // refills the stage, so no warp ever waits for the others to release a stage.
  int* done = reinterpret_cast<int*>(full + SY_STAGES);

  for (int k = tid; k < A.ldd; k += NT) W.dg[k] = (k < A.nd) ? W.d[A.dcols[k]] : 0.0;
  const int g0 = W.ring_g;                            // the ring's chunk count so far
  __syncthreads();
  long long tk = phase_begin(W);
  if (tid == 0) {
    // the staging area was last touched through the generic proxy (the factorisation's P)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int c = 0; c < SY_STAGES && c < total; c++) {
      const int st = (g0 + c) % SY_STAGES;
      mbar_expect_tx(&full[st], chunk_bytes);
      tma_load_1d(W.P + st * stage_doubles, A.sy_A + (size_t)(c % nch) * stage_doubles, chunk_bytes, &full[st]);
    }
  }

  int gch = 0;                                        // chunk counter over both passes
  for (int pass = 0; pass < A.sy_npass; pass++) {
    // this warp's segments: (tile row, first tile col, count)
    int segI[SY_SEG], segJ[SY_SEG], segN[SY_SEG];
#pragma unroll
    for (int s = 0; s < SY_SEG; s++) {
      const int4 e = A.sy_seg[(pass * NWARP + warp) * SY_SEG + s];
      segI[s] = __shfl_sync(0xffffffffu, e.x, 0);      // broadcast: known warp-uniform
      segJ[s] = __shfl_sync(0xffffffffu, e.y, 0);
      segN[s] = __shfl_sync(0xffffffffu, e.z, 0);
    }
    double acc[SY_SEG][SY_CW][2];
#pragma unroll
    for (int s = 0; s < SY_SEG; s++)
#pragma unroll
      for (int t = 0; t < SY_CW; t++) acc[s][t][0] = acc[s][t][1] = 0.0;

    // the chunk loop, instantiated for the shape of this warp's segments
    auto chunk_loop = [&](auto n2c) {
      constexpr int N2 = decltype(n2c)::value;
      for (int ch = 0; ch < nch; ch++, gch++) {
        const int st = (g0 + gch) % SY_STAGES;
        mbar_wait(&full[st], ((g0 + gch) / SY_STAGES) & 1);
        syrk_chunk<N2>(acc, W.P + st * stage_doubles, W.dg + ch * SY_KC, ldm, g, tg, segI, segJ, segN);
        __syncwarp();
        if (lane == 0) {                               // this warp is done with stage st
          if (atom_add_acq_rel(&done[st], 1) == NWARP - 1) {
            done[st] = 0;
            const int c = gch + SY_STAGES;
            if (c < total) {
              mbar_expect_tx(&full[st], chunk_bytes);
              tma_load_1d(W.P + st * stage_doubles, A.sy_A + (size_t)(c % nch) * stage_doubles,
                          chunk_bytes, &full[st]);
            }
          }
        }
      }
    };
    const bool regular = (segN[0] == SY_CW) && (segN[1] == SY_CW);
    if (!regular) chunk_loop(std::integral_constant<int, -1>{});
    else if (segN[2] == 0) chunk_loop(std::integral_constant<int, 0>{});
    else if (segN[2] == 1) chunk_loop(std::integral_constant<int, 1>{});
    else if (segN[2] == 2) chunk_loop(std::integral_constant<int, 2>{});
    else if (segN[2] == 3) chunk_loop(std::integral_constant<int, 3>{});
    else chunk_loop(std::integral_constant<int, 4>{});
    tk = phase_begin(W);
    // epilogue: tiles -> packed L storage and full M
#pragma unroll
    for (int s = 0; s < SY_SEG; s++) {
#pragma unroll
      for (int t = 0; t < SY_CW; t++) {
        if (t < segN[s]) {
          const int i = 8 * segI[s] + g;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int j = 8 * (segJ[s] + t) + 2 * tg + h;
            if (i < m && j <= i) {
              const double v = acc[s][t][h];
              W.L[coff(j, m) + i] = v;
              if (W.M != nullptr) {                  // (only callers that keep a copy of M for the residual)
                W.M[(size_t)i * m + j] = v;
                W.M[(size_t)j * m + i] = v;
              }
            }
          }
        }
      }
    }
    tk = phase_begin(W);
  }
  W.ring_g = g0 + total;
  __syncthreads();
  // singleton (slack) columns: diagonal only
  for (int i = tid; i < m; i += NT) {
    if (A.sing_ptr[i + 1] > A.sing_ptr[i]) {
      double s = 0.0;
      for (int e = A.sing_ptr[i]; e < A.sing_ptr[i + 1]; e++) s += A.sing_w[e] * W.d[A.sing_col[e]];
      const double v = W.L[coff(i, m) + i] + s;
      if (W.M != nullptr) W.M[(size_t)i * m + i] = v;
      W.L[coff(i, m) + i] = v;
    }
  }
}


// Out-of-line entry: the SYRK loop needs ~100 registers of its own (12 accumulator tiles,
// prefetched fragments).  Inlined into the persistent kernel its register allocation is at the
// mercy of everything that is live around it (seen: the fragment loads collapse onto one
// register and serialise, +16 % time); as a real function it gets its own allocation and the
// caller parks its live values around the single call per Newton step.
template <bool LSH>
static __device__ __noinline__ int form_M_dense_tma_call_t(
    const double* sy_A, const int4* sy_seg, const int* dcols, const int* sing_ptr, const int* sing_col,
    const double* sing_w, int m, int nd, int ldd, int ldm, int npass, double* d, double* dg, double* P,
    double* red, double* L, double* M, unsigned long long* prof, int ring_g) {
  Matrix A;
  A.m = m; A.nd = nd; A.ldd = ldd; A.sy_ldm = ldm; A.sy_npass = npass;
  A.sy_A = sy_A; A.sy_seg = sy_seg; A.dcols = dcols;
  A.sing_ptr = sing_ptr; A.sing_col = sing_col; A.sing_w = sing_w;
  // the pointers arrive generic; re-derived from the dynamic shared-memory base they give LDS/STS
  // (d, dg, the stage area and the reduction scratch are always on chip here; L only if LSH)
  extern __shared__ __align__(16) double dyn_smem[];
  const uint32_t b0 = smem_u32(dyn_smem);
  Work W;
  W.d = dyn_smem + ((smem_u32(d) - b0) >> 3);
  W.dg = dyn_smem + ((smem_u32(dg) - b0) >> 3);
  W.P = dyn_smem + ((smem_u32(P) - b0) >> 3);
  W.red = dyn_smem + ((smem_u32(red) - b0) >> 3);
  W.L = LSH ? dyn_smem + ((smem_u32(L) - b0) >> 3) : L;
  W.M = M; W.prof = prof; W.ring_g = ring_g;
  form_M_dense_tma(A, W);
  return W.ring_g;
}
static __device__ __forceinline__ int form_M_dense_tma_call(
    const double* sy_A, const int4* sy_seg, const int* dcols, const int* sing_ptr, const int* sing_col,
    const double* sing_w, int m, int nd, int ldd, int ldm, int npass, double* d, double* dg, double* P,
    double* red, double* L, double* M, unsigned long long* prof, int ring_g) {
  if (__isShared(L))
    return form_M_dense_tma_call_t<true>(sy_A, sy_seg, dcols, sing_ptr, sing_col, sing_w, m, nd, ldd, ldm, npass, d,
                                         dg, P, red, L, M, prof, ring_g);
  return form_M_dense_tma_call_t<false>(sy_A, sy_seg, dcols, sing_ptr, sing_col, sing_w, m, nd, ldd, ldm, npass, d,
                                        dg, P, red, L, M, prof, ring_g);
}

}  // namespace pb200
