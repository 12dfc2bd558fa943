/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference's own OpenCL kernels (pycllp/cl/primal_normal.cl and
 * pycllp/cl/ldl.cl) as plain C, *from where they lie* under /root/reference
 * (include path given on the gcc command line, see oracle/Makefile), and wraps
 * them in a host driver that does what pycllp/solvers/cl.py does:
 *   - lay b, c out interleaved "problem-minor" (cl.py:99,102 upload lp.b.T / lp.c.T),
 *   - launch initialize_xzyw then (sparse_)standard_primal_normal once per
 *     work-item (cl.py:108-111, 263-264),
 *   - read x (and, beyond the reference, y and z) back (cl.py:117-121).
 *
 * No reference source is copied: the two .cl files are #included.  The kernels
 * use no barriers / local memory / vector types, so the OpenCL qualifiers can be
 * defined away.  Work-item ids become thread-local variables so that chunks of
 * problems can be run on several host threads (problems are independent).
 *
 * MUST be built without FMA contraction: the Makefile uses
 * plain x86-64 -O2 -ffp-contract=off.
 *
 * The output of this build lives in oracle/_ref/ (git-ignored, travels to the GPU
 * box with the snapshot).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <stdarg.h>

#define __kernel
#define __global
#define __constant
#define inline static inline

static __thread int ref_gid__, ref_gsize__;
#define get_global_id(d)   (ref_gid__)
#define get_global_size(d) (ref_gsize__)

/* OpenCL pown(x, n): exact repeated multiplication (x*x for n == 2). */
static double pown(double x, int n) {
  double r = 1.0;
  for (int i = 0; i < n; i++) r *= x;
  return r;
}

/* The kernels printf() one line per IPM iteration when verbose > 1
 * (primal_normal.cl:250-252).  Intercept it to record the iteration count and the
 * last (|rho|, |sigma|, gamma) per work-item without touching the sources. */
static __thread int *ref_iters__;      /* [gsize] or NULL */
static __thread double *ref_trace__;   /* [gsize*3] or NULL */
static int ref_printf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int gid = va_arg(ap, int);
  (void)va_arg(ap, int);
  int iter = va_arg(ap, int);
  double normr = va_arg(ap, double);
  double norms = va_arg(ap, double);
  double gamma = va_arg(ap, double);
  va_end(ap);
  (void)fmt;
  if (ref_iters__) ref_iters__[gid] = iter;
  if (ref_trace__) {
    ref_trace__[3 * gid + 0] = normr;
    ref_trace__[3 * gid + 1] = norms;
    ref_trace__[3 * gid + 2] = gamma;
  }
  return 0;
}
#define printf ref_printf

/* same concatenation order as cl.py:72 */
#include "primal_normal.cl"
#include "ldl.cl"

#undef printf
#undef inline

/* ------------------------------------------------------------------------- */
/* Raw kernel hooks: caller supplies interleaved arrays exactly as the          */
/* reference's tests do (tests/test_ldl.py:172,184,265,352); one "launch" over   */
/* global size N.                                                               */
/* ------------------------------------------------------------------------- */

void ref_ldl(int N, int m, int n, double *A, double *L, double *D) {
  ref_gsize__ = N;
  for (ref_gid__ = 0; ref_gid__ < N; ref_gid__++) ldl(m, n, A, L, D);
}

void ref_modified_ldl(int N, int m, int n, double *A, double *L, double *D, double beta,
                      double delta) {
  ref_gsize__ = N;
  for (ref_gid__ = 0; ref_gid__ < N; ref_gid__++) modified_ldl(m, n, A, L, D, beta, delta);
}

void ref_solve_primal_normal(int N, int m, int n, double *A, double *x, double *z, double *y,
                             double *b, double *c, double mu, double *L, double *D, double *S,
                             double *dy, double delta) {
  ref_gsize__ = N;
  for (ref_gid__ = 0; ref_gid__ < N; ref_gid__++)
    solve_primal_normal(m, n, A, x, z, y, b, c, mu, L, D, S, dy, delta);
}

void ref_sparse_solve_primal_normal(int N, int m, int n, double *Adata, int *Aindptr,
                                    int *Aindices, double *ATdata, int *ATindptr, int *ATindices,
                                    double *x, double *z, double *y, double *b, double *c,
                                    double mu, double *Ldata, int *Lindptr, int *Lindices,
                                    int *LTindptr, int *LTindices, int *LTmap, double *D,
                                    double *S, double *dy, double delta) {
  ref_gsize__ = N;
  for (ref_gid__ = 0; ref_gid__ < N; ref_gid__++)
    sparse_solve_primal_normal(m, n, Adata, Aindptr, Aindices, ATdata, ATindptr, ATindices, x, z,
                               y, b, c, mu, Ldata, Lindptr, Lindices, LTindptr, LTindices, LTmap,
                               D, S, dy, delta);
}

/* ------------------------------------------------------------------------- */
/* Whole-solve drivers (what cl.py's solve() does), chunked over host threads.  */
/* b is (N, m) row-major, c is (N, n) row-major (the lp.b / lp.c arrays); the     */
/* driver interleaves each chunk the way cl.py uploads lp.b.T / lp.c.T, and       */
/* de-interleaves x, y, z into (N, n), (N, m), (N, n).                           */
/* iters[p] = index of the last iteration entered (the one whose stop test       */
/* fired, or 199); trace[p*3..] = its (|rho|, |sigma|, gamma).                    */
/* ------------------------------------------------------------------------- */

#define REF_CHUNK_MAX 8
/* problems per chunk: small enough that every host thread gets work */
static int ref_chunk(int N, int nthreads) {
  int c = (N + nthreads - 1) / nthreads;
  if (c > REF_CHUNK_MAX) c = REF_CHUNK_MAX;
  if (c < 1) c = 1;
  return c;
}

typedef struct {
  int nnzL;
  const int *Lindptr, *Lindices, *LTindptr, *LTindices, *LTmap;
  const double *Adata, *ATdata;
  const int *Aindptr, *Aindices, *ATindptr, *ATindices;
} ref_sparse_t;

static void ref_run_chunk(int p0, int cn, int m, int n, const double *A, const ref_sparse_t *sp,
                          const double *b, const double *c, double *x, double *y, double *z,
                          int *status, int *iters, double *trace) {
  size_t lsz = sp ? (size_t)sp->nnzL : (size_t)m * (m + 1) / 2;
  double *bi = malloc(sizeof(double) * m * cn), *ci = malloc(sizeof(double) * n * cn);
  double *xi = malloc(sizeof(double) * n * cn), *zi = malloc(sizeof(double) * n * cn);
  double *yi = malloc(sizeof(double) * m * cn);
  double *dxi = malloc(sizeof(double) * n * cn), *dzi = malloc(sizeof(double) * n * cn);
  double *dyi = malloc(sizeof(double) * m * cn);
  double *L = malloc(sizeof(double) * lsz * cn), *D = malloc(sizeof(double) * m * cn);
  double *S = malloc(sizeof(double) * m * cn);
  int *st = malloc(sizeof(int) * cn), *it = malloc(sizeof(int) * cn);
  double *tr = malloc(sizeof(double) * 3 * cn);
  for (int g = 0; g < cn; g++) {
    for (int i = 0; i < m; i++) bi[i * cn + g] = b[(size_t)(p0 + g) * m + i];
    for (int j = 0; j < n; j++) ci[j * cn + g] = c[(size_t)(p0 + g) * n + j];
    it[g] = -1;
  }
  ref_gsize__ = cn;
  ref_iters__ = it;
  ref_trace__ = tr;
  for (ref_gid__ = 0; ref_gid__ < cn; ref_gid__++) initialize_xzyw(m, n, xi, zi, yi);
  for (ref_gid__ = 0; ref_gid__ < cn; ref_gid__++) {
    if (sp)
      sparse_standard_primal_normal(
          m, n, (double *)sp->Adata, (int *)sp->Aindptr, (int *)sp->Aindices,
          (double *)sp->ATdata, (int *)sp->ATindptr, (int *)sp->ATindices, xi, zi, yi, dxi, dzi,
          dyi, bi, ci, L, (int *)sp->Lindptr, (int *)sp->Lindices, (int *)sp->LTindptr,
          (int *)sp->LTindices, (int *)sp->LTmap, D, S, st, 2);
    else
      standard_primal_normal(m, n, (double *)A, xi, zi, yi, dxi, dzi, dyi, bi, ci, L, D, S, st, 2);
  }
  ref_iters__ = NULL;
  ref_trace__ = NULL;
  for (int g = 0; g < cn; g++) {
    size_t p = (size_t)(p0 + g);
    for (int j = 0; j < n; j++) {
      x[p * n + j] = xi[j * cn + g];
      z[p * n + j] = zi[j * cn + g];
    }
    for (int i = 0; i < m; i++) y[p * m + i] = yi[i * cn + g];
    status[p] = st[g];
    if (iters) iters[p] = it[g];
    if (trace) memcpy(trace + 3 * p, tr + 3 * g, sizeof(double) * 3);
  }
  free(bi); free(ci); free(xi); free(zi); free(yi); free(dxi); free(dzi); free(dyi);
  free(L); free(D); free(S); free(st); free(it); free(tr);
}

int ref_run_dense(int N, int m, int n, const double *A, const double *b, const double *c,
                  double *x, double *y, double *z, int *status, int *iters, double *trace,
                  int nthreads) {
  if (nthreads < 1) nthreads = 1;
  const int CH = ref_chunk(N, nthreads);
  int nchunks = (N + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int ch = 0; ch < nchunks; ch++) {
    int p0 = ch * CH;
    int cn = N - p0 < CH ? N - p0 : CH;
    ref_run_chunk(p0, cn, m, n, A, NULL, b, c, x, y, z, status, iters, trace);
  }
  return 0;
}

int ref_run_sparse(int N, int m, int n, const double *Adata, const int *Aindptr,
                   const int *Aindices, const double *ATdata, const int *ATindptr,
                   const int *ATindices, int nnzL, const int *Lindptr, const int *Lindices,
                   const int *LTindptr, const int *LTindices, const int *LTmap, const double *b,
                   const double *c, double *x, double *y, double *z, int *status, int *iters,
                   double *trace, int nthreads) {
  ref_sparse_t sp = {nnzL,  Lindptr, Lindices, LTindptr, LTindices, LTmap,
                     Adata, ATdata,  Aindptr,  Aindices, ATindptr,  ATindices};
  if (nthreads < 1) nthreads = 1;
  const int CH = ref_chunk(N, nthreads);
  int nchunks = (N + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int ch = 0; ch < nchunks; ch++) {
    int p0 = ch * CH;
    int cn = N - p0 < CH ? N - p0 : CH;
    ref_run_chunk(p0, cn, m, n, NULL, &sp, b, c, x, y, z, status, iters, trace);
  }
  return 0;
}
