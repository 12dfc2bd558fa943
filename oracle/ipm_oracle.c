/*
 * oracle/ipm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement, in plain C, of the algorithm of the reference's batched
 * primal normal-equations interior-point path (pycllp/cl/primal_normal.cl and
 * pycllp/cl/ldl.cl; host driver pycllp/solvers/cl.py).  It is the checker for the
 * CUDA engine: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (pycllp_b200/) never does.
 *
 * Pinning: this file reproduces the reference's floating-point operation ORDER
 * exactly (every sum runs over the same index in the same direction, every
 * product/division is associated the same way), so when both are compiled without
 * FMA contraction its results are BIT-IDENTICAL to the reference kernels built by
 * oracle/ref_shim.c (oracle/_ref/libpycllp_ref.so).  tests/test_oracle.py checks
 * that on config 1, config 2 and random cases, and checks both against the
 * reference's known-answer LPs (tests/vanderbei_problems.py, tests/test_simple.py).
 *
 * What differs from the reference is only *when* things are evaluated: the
 * reference recomputes every entry of M = A (X/Z) A', every (A'y)_j and every
 * RHS_i each time it is used (O(m^2 n) per use); here each is evaluated once per
 * IPM iteration with the same expression and cached -- same bits, far fewer flops.
 * Terms with an exactly-zero A entry are skipped where adding the resulting +-0.0
 * cannot change any bit of the accumulator.
 *
 * Layout: unlike the kernels (interleaved, problem-minor) every per-problem vector
 * here is contiguous; b is (N, m), c is (N, n) row-major, exactly lp.b / lp.c.
 *
 * Build: oracle/Makefile (gcc -O3 -ffp-contract=off, no -ffast-math, no -mfma).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double eps;         /* primal_normal.cl:8   EPS 1.0e-7f (a float literal)      */
  double delta;       /* primal_normal.cl:10  DELTA 0.02                         */
  double r;           /* primal_normal.cl:11  R 0.9                              */
  double ldl_delta;   /* primal_normal.cl:275 / :366  solve_...(…, 1e-6)          */
  double refine_tol;  /* ldl.cl:645  maxr > 1e-8                                  */
  int max_iter;       /* primal_normal.cl:9   MAX_ITER 200                        */
  int max_refine;     /* ldl.cl:645  nref < 5 (dense); sparse: refinement is      */
                      /* commented out, ldl.cl:698-711 -> 0                       */
} oracle_params;

void oracle_default_params(oracle_params *p, int sparse) {
  p->eps = (double)1.0e-7f;
  p->delta = 0.02;
  p->r = 0.9;
  p->ldl_delta = 1e-6;
  p->refine_tol = 1e-8;
  p->max_iter = 200;
  p->max_refine = sparse ? 0 : 5;
}

static inline int tri(int i, int j) { return i * (i + 1) / 2 + j; } /* ldl.cl:12-18 */

/* ========================================================================= */
/* Dense path                                                                  */
/* ========================================================================= */

typedef struct {
  int m, n;
  const double *A;  /* m x n row-major (cl.py:39) */
  double *AT;       /* n x m copy, only to make the M loop contiguous */
  double *M;        /* m x m : AXZAt_ij(i, j) for every ordered pair  */
  double *Msq;      /* m     : AXZAt_ii(i)                             */
  double *L, *D, *S, *RHS, *dy, *Aty, *t, *dx, *dz;
} dense_work;

static void dense_work_alloc(dense_work *w, int m, int n, const double *A) {
  w->m = m; w->n = n; w->A = A;
  w->AT = malloc(sizeof(double) * (size_t)m * n);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) w->AT[(size_t)j * m + i] = A[(size_t)i * n + j];
  w->M = malloc(sizeof(double) * (size_t)m * m);
  w->Msq = malloc(sizeof(double) * m);
  w->L = malloc(sizeof(double) * ((size_t)m * (m + 1) / 2));
  w->D = malloc(sizeof(double) * m);
  w->S = malloc(sizeof(double) * m);
  w->RHS = malloc(sizeof(double) * m);
  w->dy = malloc(sizeof(double) * m);
  w->Aty = malloc(sizeof(double) * n);
  w->t = malloc(sizeof(double) * n);
  w->dx = malloc(sizeof(double) * n);
  w->dz = malloc(sizeof(double) * n);
}

static void dense_work_free(dense_work *w) {
  free(w->AT); free(w->M); free(w->Msq); free(w->L); free(w->D); free(w->S);
  free(w->RHS); free(w->dy); free(w->Aty); free(w->t); free(w->dx); free(w->dz);
}

/* primal_normal.cl:30-48 */
static double d_primal_infeasibility(const dense_work *w, const double *x, const double *b) {
  int m = w->m, n = w->n;
  double normr = 0.0;
  for (int i = 0; i < m; i++) {
    double rho = b[i];
    const double *Ai = w->A + (size_t)i * n;
    for (int j = 0; j < n; j++) rho -= Ai[j] * x[j];
    normr += rho * rho;
  }
  return sqrt(normr);
}

/* A'y with the accumulation order of primal_normal.cl:139-140 / ldl.cl:208-211
 * (i ascending from 0.0): shared by the RHS, the step and nothing else.       */
static void d_At_times(const dense_work *w, const double *v, double *out) {
  int m = w->m, n = w->n;
  for (int j = 0; j < n; j++) out[j] = 0.0;
  for (int i = 0; i < m; i++) {
    const double *Ai = w->A + (size_t)i * n;
    double vi = v[i];
    for (int j = 0; j < n; j++) out[j] += Ai[j] * vi;
  }
}

/* primal_normal.cl:76-94 : sigma = c + z; sigma += -A_ij*y_i (i ascending) */
static double d_dual_infeasibility(const dense_work *w, const double *z, const double *y,
                                   const double *c, double *tmp) {
  int m = w->m, n = w->n;
  for (int j = 0; j < n; j++) tmp[j] = c[j] + z[j];
  for (int i = 0; i < m; i++) {
    const double *Ai = w->A + (size_t)i * n;
    double yi = y[i];
    for (int j = 0; j < n; j++) tmp[j] += -Ai[j] * yi;
  }
  double norms = 0.0;
  for (int j = 0; j < n; j++) norms += tmp[j] * tmp[j];
  return sqrt(norms);
}

/* ldl.cl:110-138 : every AXZAt_ij(i, j) (ordered pair!) and AXZAt_ii(i).
 * a += A[i,k]*x[k]*A[j,k]/z[k]  ==  ((A_ik*x_k)*A_jk)/z_k ;  k ascending.     */
static void d_form_M(dense_work *w, const double *x, const double *z) {
  int m = w->m, n = w->n;
  for (int i = 0; i < m; i++) {
    double *Mi = w->M + (size_t)i * m;
    const double *Ai = w->A + (size_t)i * n;
    for (int j = 0; j < m; j++) Mi[j] = 0.0;
    double sq = 0.0;
    for (int k = 0; k < n; k++) {
      double aik = Ai[k];
      if (aik == 0.0) continue; /* adds +-0.0 to a non-negative-zero accumulator */
      double p = aik * x[k], zk = z[k];
      const double *ATk = w->AT + (size_t)k * m;
      for (int j = 0; j < m; j++) Mi[j] += p * ATk[j] / zk;
      sq += aik * aik * x[k] / zk; /* pown(A,2)*x/z */
    }
    w->Msq[i] = sq;
  }
}

/* ldl.cl:314-378 factor_primal_normal (beta: ldl.cl:280-294) */
static void d_factor(dense_work *w, double delta) {
  int m = w->m;
  double *L = w->L, *D = w->D;
  double beta = 0.0;
  for (int j = 0; j < m; j++) beta = fmax(beta, fabs(w->Msq[j]));
  beta = sqrt(beta);
  for (int j = 0; j < m; j++) {
    double Dj = w->Msq[j];
    for (int k = 0; k < j; k++) Dj -= D[k] * (L[tri(j, k)] * L[tri(j, k)]);
    double theta = 0.0;
    for (int i = j + 1; i < m; i++) {
      double Lij = w->M[(size_t)i * m + j];
      const double *Li = L + tri(i, 0), *Lj = L + tri(j, 0);
      for (int k = 0; k < j; k++) Lij -= Li[k] * Lj[k] * D[k];
      theta = fmax(theta, fabs(Lij));
      L[tri(i, j)] = Lij;
    }
    double q = theta / beta;
    Dj = fmax(fabs(Dj), fmax(q * q, delta));
    for (int i = j + 1; i < m; i++) L[tri(i, j)] /= Dj;
    D[j] = Dj;
    L[tri(j, j)] = 1.0;
  }
}

/* ldl.cl:198-219 primal_normal_rhs_i for every i; caches A'y and
 * t_j = c_j - (A'y)_j + mu/x_j  (the vector the step must see bit-identically). */
static void d_rhs(dense_work *w, const double *x, const double *z, const double *y,
                  const double *b, const double *c, double mu) {
  int m = w->m, n = w->n;
  d_At_times(w, y, w->Aty);
  for (int j = 0; j < n; j++) w->t[j] = c[j] - w->Aty[j] + mu / x[j];
  for (int i = 0; i < m; i++) {
    double rhs = b[i];
    const double *Ai = w->A + (size_t)i * n;
    for (int j = 0; j < n; j++) {
      double a = Ai[j];
      if (a == 0.0) continue; /* += -0.0 twice: no bit changes */
      rhs += -a * x[j];
      rhs += -a * x[j] * w->t[j] / z[j];
    }
    w->RHS[i] = -rhs;
  }
}

/* ldl.cl:505-537 forward_backward_primal_normal */
static void d_forward_backward(dense_work *w) {
  int m = w->m;
  double *L = w->L, *D = w->D, *S = w->S, *dy = w->dy;
  for (int i = 0; i < m; i++) {
    double Si = S[i];
    for (int j = 0; j < i; j++) Si -= S[j] * L[tri(i, j)] * D[j];
    S[i] = Si / D[i];
  }
  for (int j = m - 1; j >= 0; j--) {
    double Sj = S[j];
    for (int i = j + 1; i < m; i++) Sj -= S[i] * L[tri(i, j)];
    S[j] = Sj;
    dy[j] += Sj;
  }
}

/* ldl.cl:577-599 residual_primal_normal */
static double d_residual(dense_work *w) {
  int m = w->m;
  double maxr = 0.0;
  for (int i = 0; i < m; i++) {
    double r = w->RHS[i];
    const double *Mi = w->M + (size_t)i * m;
    for (int j = 0; j < m; j++) r -= Mi[j] * w->dy[j];
    w->S[i] = r;
    maxr = fmax(maxr, fabs(r));
  }
  return maxr;
}

/* ldl.cl:602-653 solve_primal_normal; returns the number of refinement passes */
static int d_solve_normal(dense_work *w, const double *x, const double *z, const double *y,
                          const double *b, const double *c, double mu, const oracle_params *p) {
  int m = w->m;
  d_form_M(w, x, z);
  d_factor(w, p->ldl_delta);
  d_rhs(w, x, z, y, b, c, mu);
  for (int i = 0; i < m; i++) { w->dy[i] = 0.0; w->S[i] = w->RHS[i]; }
  d_forward_backward(w);
  double maxr = d_residual(w);
  int nref = 0;
  while (maxr > p->refine_tol && nref < p->max_refine) {
    d_forward_backward(w);
    maxr = d_residual(w);
    nref += 1;
  }
  return nref;
}

/* primal_normal.cl:122-156 primal_normal_step */
static void d_step(dense_work *w, double *x, double *z, double *y, const double *c, double r,
                   double mu) {
  int m = w->m, n = w->n;
  double *dx = w->dx, *dz = w->dz, *Atdy = w->t; /* t is dead once the RHS is formed */
  /* The reference recomputes A'y here with the same summation order as inside
   * primal_normal_rhs_i, i.e. the same bits as the cached w->Aty. */
  d_At_times(w, w->dy, Atdy);
  double theta = 0.0;
  for (int j = 0; j < n; j++) {
    dx[j] = (c[j] - w->Aty[j] + mu / x[j] - Atdy[j]) * x[j] / z[j];
    dz[j] = (mu - z[j] * dx[j]) / x[j] - z[j];
    theta = fmax(theta, fmax(-dz[j] / z[j], -dx[j] / x[j]));
  }
  theta = fmin(r / theta, 1.0);
  for (int i = 0; i < m; i++) y[i] += theta * w->dy[i];
  for (int j = 0; j < n; j++) {
    z[j] += theta * dz[j];
    x[j] += theta * dx[j];
  }
}

/* primal_normal.cl:201-284 standard_primal_normal (after initialize_xzyw, :14-28) */
/* warm != 0: start from the x, z, y passed in (the kernel's doc comment, primal_normal.cl:213-219;
 * the reference's host code always re-initialises, cl.py:108).  itrace (may be NULL):
 * (|rho|, |sigma|, gamma) of every iteration, what the kernel prints at verbose > 1
 * (primal_normal.cl:250-252). */
static int d_solve_one_ex(dense_work *w, const double *b, const double *c, double *x, double *y,
                          double *z, const oracle_params *p, int *iters, int *nrefs,
                          double *trace, int warm, double *itrace) {
  int m = w->m, n = w->n;
  if (!warm) {
    for (int i = 0; i < m; i++) y[i] = 1.0;
    for (int j = 0; j < n; j++) { x[j] = 1.0; z[j] = 1.0; }
  }
  int stat = 5;
  double normr0 = HUGE_VALF / 10, norms0 = HUGE_VALF / 10;
  int iter, refs = 0;
  double normr = 0, norms = 0, gamma = 0;
  for (iter = 0; iter < p->max_iter; iter++) {
    normr = d_primal_infeasibility(w, x, b);
    norms = d_dual_infeasibility(w, z, y, c, w->t);
    gamma = 0.0;
    for (int j = 0; j < n; j++) gamma += z[j] * x[j];
    if (itrace) { itrace[3 * iter] = normr; itrace[3 * iter + 1] = norms; itrace[3 * iter + 2] = gamma; }
    if (normr < p->eps && norms < p->eps && gamma < p->eps) { stat = 0; break; }
    if (normr > 10 * normr0 && normr > p->eps) { stat = 2; break; }
    if (norms > 10 * norms0 && norms > p->eps) { stat = 4; break; }
    double mu = p->delta * gamma / (n + m);
    refs += d_solve_normal(w, x, z, y, b, c, mu, p);
    d_step(w, x, z, y, c, p->r, mu);
    normr0 = normr;
    norms0 = norms;
  }
  if (iters) *iters = iter; /* = number of Newton steps taken */
  if (nrefs) *nrefs = refs;
  if (trace) { trace[0] = normr; trace[1] = norms; trace[2] = gamma; }
  return stat;
}

static int d_solve_one(dense_work *w, const double *b, const double *c, double *x, double *y,
                       double *z, const oracle_params *p, int *iters, int *nrefs,
                       double *trace) {
  return d_solve_one_ex(w, b, c, x, y, z, p, iters, nrefs, trace, 0, NULL);
}

/* Test hook for the warm-start / iteration-trace features of the engine: x, y, z are in/out when
 * warm != 0; itrace is (N, max_iter, 3) or NULL. */
int oracle_solve_dense_ex(int N, int m, int n, const double *A, const double *b, const double *c,
                          double *x, double *y, double *z, int *status, int *iters,
                          const oracle_params *params, int nthreads, int warm, double *itrace) {
  oracle_params p;
  if (params) p = *params; else oracle_default_params(&p, 0);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    dense_work w;
    dense_work_alloc(&w, m, n, A);
#pragma omp for schedule(dynamic, 1)
    for (int q = 0; q < N; q++) {
      status[q] = d_solve_one_ex(&w, b + (size_t)q * m, c + (size_t)q * n, x + (size_t)q * n,
                                 y + (size_t)q * m, z + (size_t)q * n, &p, iters ? iters + q : NULL,
                                 NULL, NULL, warm,
                                 itrace ? itrace + 3 * (size_t)q * p.max_iter : NULL);
    }
    dense_work_free(&w);
  }
  return 0;
}

int oracle_solve_dense(int N, int m, int n, const double *A, const double *b, const double *c,
                       double *x, double *y, double *z, int *status, int *iters, int *nrefs,
                       double *trace, const oracle_params *params, int nthreads) {
  oracle_params p;
  if (params) p = *params; else oracle_default_params(&p, 0);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    dense_work w;
    dense_work_alloc(&w, m, n, A);
#pragma omp for schedule(dynamic, 1)
    for (int q = 0; q < N; q++) {
      status[q] = d_solve_one(&w, b + (size_t)q * m, c + (size_t)q * n, x + (size_t)q * n,
                              y + (size_t)q * m, z + (size_t)q * n, &p,
                              iters ? iters + q : NULL, nrefs ? nrefs + q : NULL,
                              trace ? trace + 3 * (size_t)q : NULL);
    }
    dense_work_free(&w);
  }
  return 0;
}

/* Kernel-level hook: one solve_primal_normal (ldl.cl:602-653) per problem on
 * caller-supplied x, z, y, b, c (each (N, len) row-major) -> dy (N, m). Mirrors the
 * reference's tests/test_ldl.py:219-273. Also returns L (packed) and D if non-NULL. */
int oracle_solve_primal_normal(int N, int m, int n, const double *A, const double *x,
                               const double *z, const double *y, const double *b,
                               const double *c, double mu, double delta, int max_refine,
                               double *dy, double *Lout, double *Dout) {
  oracle_params p;
  oracle_default_params(&p, 0);
  p.ldl_delta = delta;
  p.max_refine = max_refine;
  dense_work w;
  dense_work_alloc(&w, m, n, A);
  size_t lsz = (size_t)m * (m + 1) / 2;
  for (int q = 0; q < N; q++) {
    d_solve_normal(&w, x + (size_t)q * n, z + (size_t)q * n, y + (size_t)q * m,
                   b + (size_t)q * m, c + (size_t)q * n, mu, &p);
    memcpy(dy + (size_t)q * m, w.dy, sizeof(double) * m);
    if (Lout) memcpy(Lout + q * lsz, w.L, sizeof(double) * lsz);
    if (Dout) memcpy(Dout + (size_t)q * m, w.D, sizeof(double) * m);
  }
  dense_work_free(&w);
  return 0;
}

/* ldl.cl:28-55 (plain LDL') and ldl.cl:57-107 (modified LDL') of given dense
 * matrices, AA is (N, m, m) row-major; L packed (N, m(m+1)/2), D (N, m).
 * modified != 0 selects the modified variant with (beta, delta).             */
int oracle_ldl(int N, int m, const double *AA, double *L, double *D, int modified, double beta,
               double delta) {
  size_t lsz = (size_t)m * (m + 1) / 2;
  for (int q = 0; q < N; q++) {
    const double *A = AA + (size_t)q * m * m;
    double *Lq = L + q * lsz, *Dq = D + (size_t)q * m;
    if (!modified) {
      for (int i = 0; i < m; i++) {
        int j;
        for (j = 0; j < i; j++) {
          double l = A[(size_t)i * m + j];
          for (int k = 0; k < j; k++) l -= Lq[tri(i, k)] * Lq[tri(j, k)] * Dq[k];
          Lq[tri(i, j)] = l / Dq[j];
        }
        double d = A[(size_t)i * m + i];
        for (int k = 0; k < j; k++) d -= Dq[k] * (Lq[tri(i, k)] * Lq[tri(i, k)]);
        Dq[i] = d;
        Lq[tri(i, i)] = 1.0;
      }
    } else {
      for (int j = 0; j < m; j++) {
        double Dj = A[(size_t)j * m + j];
        for (int k = 0; k < j; k++) Dj -= Dq[k] * (Lq[tri(j, k)] * Lq[tri(j, k)]);
        double theta = 0.0;
        for (int i = j + 1; i < m; i++) {
          double Lij = A[(size_t)i * m + j];
          for (int k = 0; k < j; k++) Lij -= Lq[tri(i, k)] * Lq[tri(j, k)] * Dq[k];
          theta = fmax(theta, fabs(Lij));
          Lq[tri(i, j)] = Lij;
        }
        double qq = theta / beta;
        Dj = fmax(fabs(Dj), fmax(qq * qq, delta));
        for (int i = j + 1; i < m; i++) Lq[tri(i, j)] /= Dj;
        Dq[j] = Dj;
        Lq[tri(j, j)] = 1.0;
      }
    }
  }
  return 0;
}

/* ========================================================================= */
/* Sparse path (CSR A, CSR A', CSR-lower pattern of L with the diagonal LAST in  */
/* each row, CSR of L' + LTmap: exactly the structures cl.py:175-196 builds)     */
/* ========================================================================= */

typedef struct {
  int m, n, nnzL;
  const double *Adata, *ATdata;
  const int *Aindptr, *Aindices, *ATindptr, *ATindices;
  const int *Lindptr, *Lindices, *LTindptr, *LTindices, *LTmap;
  double *Ldata, *D, *S, *dy, *Aty, *t, *dx, *dz, *rowj;
  int *mark;
} sparse_work;

static void sparse_work_alloc(sparse_work *w) {
  int m = w->m, n = w->n;
  w->Ldata = malloc(sizeof(double) * (size_t)w->nnzL);
  w->D = malloc(sizeof(double) * m);
  w->S = malloc(sizeof(double) * m);
  w->dy = malloc(sizeof(double) * m);
  w->Aty = malloc(sizeof(double) * n);
  w->t = malloc(sizeof(double) * n);
  w->dx = malloc(sizeof(double) * n);
  w->dz = malloc(sizeof(double) * n);
  w->rowj = malloc(sizeof(double) * m);
  w->mark = malloc(sizeof(int) * m);
  for (int i = 0; i < m; i++) w->mark[i] = -1;
}

static void sparse_work_free(sparse_work *w) {
  free(w->Ldata); free(w->D); free(w->S); free(w->dy); free(w->Aty); free(w->t);
  free(w->dx); free(w->dz); free(w->rowj); free(w->mark);
}

/* primal_normal.cl:50-74 */
static double s_primal_infeasibility(const sparse_work *w, const double *x, const double *b) {
  double normr = 0.0;
  for (int i = 0; i < w->m; i++) {
    double rho = b[i];
    for (int k = w->Aindptr[i]; k < w->Aindptr[i + 1]; k++) rho -= w->Adata[k] * x[w->Aindices[k]];
    normr += rho * rho;
  }
  return sqrt(normr);
}

/* primal_normal.cl:96-120 */
static double s_dual_infeasibility(const sparse_work *w, const double *z, const double *y,
                                   const double *c) {
  double norms = 0.0;
  for (int j = 0; j < w->n; j++) {
    double sigma = c[j] + z[j];
    for (int k = w->ATindptr[j]; k < w->ATindptr[j + 1]; k++)
      sigma += -w->ATdata[k] * y[w->ATindices[k]];
    norms += sigma * sigma;
  }
  return sqrt(norms);
}

/* ldl.cl:140-172 sparse_AXZAt_ij : a += A_ik*A_jk*x_k/z_k over common columns */
static double s_Mij(const sparse_work *w, int i, int j, const double *x, const double *z) {
  int ik = w->Aindptr[i], ikk = w->Aindptr[i + 1], jk = w->Aindptr[j], jkk = w->Aindptr[j + 1];
  double a = 0.0;
  while (ik < ikk && jk < jkk) {
    int icol = w->Aindices[ik], jcol = w->Aindices[jk];
    if (icol == jcol) {
      a += w->Adata[ik] * w->Adata[jk] * x[icol] / z[icol];
      ik++; jk++;
    } else if (icol < jcol) ik++;
    else jk++;
  }
  return a;
}

/* ldl.cl:174-196 sparse_AXZAt_ii */
static double s_Mii(const sparse_work *w, int i, const double *x, const double *z) {
  double a = 0.0;
  for (int k = w->Aindptr[i]; k < w->Aindptr[i + 1]; k++) {
    int col = w->Aindices[k];
    a += w->Adata[k] * w->Adata[k] * x[col] / z[col];
  }
  return a;
}

/* ldl.cl:381-502 sparse_factor_primal_normal.  The reference finds the slot of
 * (i, j) by scanning row i (ldl.cl:435-445) and forms L_ij by a sorted merge of rows
 * i and j (ldl.cl:456-469); here the slot comes from the L' structure (LTmap) and
 * row j is scattered into a dense work row -- the products L_ik*L_jk*D_k are the
 * same and are subtracted in the same (ascending column) order.               */
static void s_factor(sparse_work *w, const double *x, const double *z, double delta) {
  int m = w->m;
  double *Ld = w->Ldata, *D = w->D;
  double beta = 0.0;
  for (int j = 0; j < m; j++) beta = fmax(beta, fabs(s_Mii(w, j, x, z)));
  beta = sqrt(beta);
  for (int j = 0; j < m; j++) {
    int j0 = w->Lindptr[j], jd = w->Lindptr[j + 1] - 1; /* jd: diagonal slot */
    double Dj = s_Mii(w, j, x, z);
    for (int k = j0; k < jd; k++) Dj -= D[w->Lindices[k]] * (Ld[k] * Ld[k]);
    for (int k = j0; k < jd; k++) { w->mark[w->Lindices[k]] = j; w->rowj[w->Lindices[k]] = Ld[k]; }
    double theta = 0.0;
    for (int kt = w->LTindptr[j] + 1; kt < w->LTindptr[j + 1]; kt++) {
      int i = w->LTindices[kt], slot = w->LTmap[kt];
      double Lij = s_Mij(w, i, j, x, z);
      for (int ik = w->Lindptr[i]; ik < w->Lindptr[i + 1]; ik++) {
        int col = w->Lindices[ik];
        if (col >= j) break;
        if (w->mark[col] == j) Lij -= Ld[ik] * w->rowj[col] * D[col];
      }
      Ld[slot] = Lij;
      theta = fmax(theta, fabs(Lij));
    }
    double q = theta / beta;
    Dj = fmax(fabs(Dj), fmax(q * q, delta));
    for (int kt = w->LTindptr[j] + 1; kt < w->LTindptr[j + 1]; kt++) Ld[w->LTmap[kt]] /= Dj;
    D[j] = Dj;
    Ld[jd] = 1.0;
  }
}

/* A'v with the order of primal_normal.cl:177-185 / ldl.cl:244-249 */
static void s_At_times(const sparse_work *w, const double *v, double *out) {
  for (int j = 0; j < w->n; j++) {
    double a = 0.0;
    for (int k = w->ATindptr[j]; k < w->ATindptr[j + 1]; k++) a += w->ATdata[k] * v[w->ATindices[k]];
    out[j] = a;
  }
}

/* ldl.cl:221-257 sparse_primal_normal_rhs_i -> S */
static void s_rhs(sparse_work *w, const double *x, const double *z, const double *y,
                  const double *b, const double *c, double mu) {
  s_At_times(w, y, w->Aty);
  for (int j = 0; j < w->n; j++) w->t[j] = c[j] - w->Aty[j] + mu / x[j];
  for (int i = 0; i < w->m; i++) {
    double rhs = b[i];
    for (int k = w->Aindptr[i]; k < w->Aindptr[i + 1]; k++) {
      int j = w->Aindices[k];
      double a = w->Adata[k];
      rhs += -a * x[j];
      rhs += -a * x[j] * w->t[j] / z[j];
    }
    w->S[i] = -rhs;
  }
}

/* ldl.cl:540-574 sparse_forward_backward_primal_normal */
static void s_forward_backward(sparse_work *w) {
  int m = w->m;
  double *Ld = w->Ldata, *D = w->D, *S = w->S;
  for (int i = 0; i < m; i++) {
    double Si = S[i];
    for (int k = w->Lindptr[i]; k < w->Lindptr[i + 1] - 1; k++) {
      int j = w->Lindices[k];
      Si -= S[j] * Ld[k] * D[j];
    }
    S[i] = Si / D[i];
  }
  for (int j = m - 1; j >= 0; j--) {
    double Sj = S[j];
    for (int k = w->LTindptr[j] + 1; k < w->LTindptr[j + 1]; k++)
      Sj -= S[w->LTindices[k]] * Ld[w->LTmap[k]];
    S[j] = Sj;
    w->dy[j] += Sj;
  }
}

/* ldl.cl:656-712 sparse_solve_primal_normal (no refinement: ldl.cl:698-711) */
static void s_solve_normal(sparse_work *w, const double *x, const double *z, const double *y,
                           const double *b, const double *c, double mu, double delta) {
  s_factor(w, x, z, delta);
  for (int i = 0; i < w->m; i++) w->dy[i] = 0.0;
  s_rhs(w, x, z, y, b, c, mu);
  s_forward_backward(w);
}

/* primal_normal.cl:158-198 sparse_primal_normal_step */
static void s_step(sparse_work *w, double *x, double *z, double *y, const double *c, double r,
                   double mu) {
  int m = w->m, n = w->n;
  double *Atdy = w->t; /* t is dead once the RHS is formed */
  s_At_times(w, w->dy, Atdy);
  double theta = 0.0;
  for (int j = 0; j < n; j++) {
    w->dx[j] = (c[j] - w->Aty[j] + mu / x[j] - Atdy[j]) * x[j] / z[j];
    w->dz[j] = (mu - z[j] * w->dx[j]) / x[j] - z[j];
    theta = fmax(theta, fmax(-w->dz[j] / z[j], -w->dx[j] / x[j]));
  }
  theta = fmin(r / theta, 1.0);
  for (int i = 0; i < m; i++) y[i] += theta * w->dy[i];
  for (int j = 0; j < n; j++) {
    z[j] += theta * w->dz[j];
    x[j] += theta * w->dx[j];
  }
}

/* primal_normal.cl:287-375 sparse_standard_primal_normal */
static int s_solve_one(sparse_work *w, const double *b, const double *c, double *x, double *y,
                       double *z, const oracle_params *p, int *iters, double *trace) {
  int m = w->m, n = w->n;
  for (int i = 0; i < m; i++) y[i] = 1.0;
  for (int j = 0; j < n; j++) { x[j] = 1.0; z[j] = 1.0; }
  int stat = 5, iter;
  double normr0 = HUGE_VALF / 10, norms0 = HUGE_VALF / 10;
  double normr = 0, norms = 0, gamma = 0;
  for (iter = 0; iter < p->max_iter; iter++) {
    normr = s_primal_infeasibility(w, x, b);
    norms = s_dual_infeasibility(w, z, y, c);
    gamma = 0.0;
    for (int j = 0; j < n; j++) gamma += z[j] * x[j];
    if (normr < p->eps && norms < p->eps && gamma < p->eps) { stat = 0; break; }
    if (normr > 10 * normr0 && normr > p->eps) { stat = 2; break; }
    if (norms > 10 * norms0 && norms > p->eps) { stat = 4; break; }
    double mu = p->delta * gamma / (n + m);
    s_solve_normal(w, x, z, y, b, c, mu, p->ldl_delta);
    s_step(w, x, z, y, c, p->r, mu);
    normr0 = normr;
    norms0 = norms;
  }
  if (iters) *iters = iter;
  if (trace) { trace[0] = normr; trace[1] = norms; trace[2] = gamma; }
  return stat;
}

int oracle_solve_sparse(int N, int m, int n, const double *Adata, const int *Aindptr,
                        const int *Aindices, const double *ATdata, const int *ATindptr,
                        const int *ATindices, int nnzL, const int *Lindptr, const int *Lindices,
                        const int *LTindptr, const int *LTindices, const int *LTmap,
                        const double *b, const double *c, double *x, double *y, double *z,
                        int *status, int *iters, double *trace, const oracle_params *params,
                        int nthreads) {
  oracle_params p;
  if (params) p = *params; else oracle_default_params(&p, 1);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    sparse_work w;
    w.m = m; w.n = n; w.nnzL = nnzL;
    w.Adata = Adata; w.Aindptr = Aindptr; w.Aindices = Aindices;
    w.ATdata = ATdata; w.ATindptr = ATindptr; w.ATindices = ATindices;
    w.Lindptr = Lindptr; w.Lindices = Lindices;
    w.LTindptr = LTindptr; w.LTindices = LTindices; w.LTmap = LTmap;
    sparse_work_alloc(&w);
#pragma omp for schedule(dynamic, 1)
    for (int q = 0; q < N; q++) {
      status[q] = s_solve_one(&w, b + (size_t)q * m, c + (size_t)q * n, x + (size_t)q * n,
                              y + (size_t)q * m, z + (size_t)q * n, &p,
                              iters ? iters + q : NULL, trace ? trace + 3 * (size_t)q : NULL);
    }
    sparse_work_free(&w);
  }
  return 0;
}

/* Kernel-level hook mirroring tests/test_ldl.py:276-361 */
int oracle_sparse_solve_primal_normal(int N, int m, int n, const double *Adata,
                                      const int *Aindptr, const int *Aindices,
                                      const double *ATdata, const int *ATindptr,
                                      const int *ATindices, int nnzL, const int *Lindptr,
                                      const int *Lindices, const int *LTindptr,
                                      const int *LTindices, const int *LTmap, const double *x,
                                      const double *z, const double *y, const double *b,
                                      const double *c, double mu, double delta, double *dy) {
  sparse_work w;
  w.m = m; w.n = n; w.nnzL = nnzL;
  w.Adata = Adata; w.Aindptr = Aindptr; w.Aindices = Aindices;
  w.ATdata = ATdata; w.ATindptr = ATindptr; w.ATindices = ATindices;
  w.Lindptr = Lindptr; w.Lindices = Lindices;
  w.LTindptr = LTindptr; w.LTindices = LTindices; w.LTmap = LTmap;
  sparse_work_alloc(&w);
  for (int q = 0; q < N; q++) {
    s_solve_normal(&w, x + (size_t)q * n, z + (size_t)q * n, y + (size_t)q * m,
                   b + (size_t)q * m, c + (size_t)q * n, mu, delta);
    memcpy(dy + (size_t)q * m, w.dy, sizeof(double) * m);
  }
  sparse_work_free(&w);
  return 0;
}
