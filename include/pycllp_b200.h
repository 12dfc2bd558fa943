/*
 * pycllp_b200.h -- C ABI of the B200-native batched interior-point LP engine.
 *
 * This is the drop-in boundary for the ONE hot path of jetuk/pycllp: many LPs that
 * share a constraint matrix A (max c'x, A x = b, x >= 0) solved by the primal
 * normal-equations path-following method.  It replaces everything the reference
 * does through pyopencl in pycllp/solvers/cl.py:
 *
 *   reference (file:line)                                   this ABI
 *   ------------------------------------------------------  ---------------------------
 *   cl.create_some_context / CommandQueue  (cl.py:18-25)     pycllp_b200_create
 *   cl.Buffer(A ...), state buffers, Program.build
 *     ClDensePrimalNormalSolver.init       (cl.py:28-83)     pycllp_b200_setup_dense
 *     ClSparsePrimalNormalSolver.init      (cl.py:143-238)   pycllp_b200_setup_sparse
 *   #define EPS/MAX_ITER/DELTA/R  (primal_normal.cl:8-11)    pycllp_b200_set_params
 *   enqueue_copy(b.T), enqueue_copy(c.T), initialize_xzyw,
 *   standard_primal_normal / sparse_standard_primal_normal,
 *   enqueue_copy(x), enqueue_copy(status)
 *     ClDense...solve (cl.py:85-124), ClSparse...solve
 *     (cl.py:240-278)                                        pycllp_b200_solve_host
 *                                                            pycllp_b200_solve_device
 *   kernels the reference's tests launch directly:
 *     ldl, modified_ldl        (ldl.cl:28-107; test_ldl.py:172,184)  pycllp_b200_ldl
 *     solve_primal_normal      (ldl.cl:602-653; test_ldl.py:265)     pycllp_b200_solve_primal_normal
 *     sparse_solve_primal_normal (ldl.cl:656-712; test_ldl.py:352)   (same entry, sparse engine)
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / numpy types.
 *   - every function returns 0 on success, <0 on error; pycllp_b200_last_error()
 *     gives the message (per engine; pass NULL for the error of a failed create).
 *   - all problem data is FP64, indices and status int32.
 *   - LAYOUT: every per-problem vector is CONTIGUOUS ("problem-major"): b is (N, m),
 *     c is (N, n), x/z are (N, n), y is (N, m), row-major -- i.e. exactly lp.b, lp.c
 *     and solver.x of the reference's Python API.  (The reference's *device* layout is
 *     interleaved problem-minor because it runs one work-item per LP, cl.py:99-118;
 *     this engine runs one thread block per LP, for which contiguous vectors are the
 *     coalesced layout, so no host transpose is needed.)
 *   - status codes are the reference's (primal_normal.cl:225,256-269): 0 optimal,
 *     2 primal infeasible, 4 dual infeasible, 5 iteration limit.
 *   - an engine is bound to one CUDA device and is not thread-safe.
 *   - there is NO CPU fallback: with no usable CUDA device create() fails.
 */
#ifndef PYCLLP_B200_H
#define PYCLLP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pycllp_b200_engine pycllp_b200_engine;

/* Algorithm constants; defaults = the reference's OpenCL path (preset "cl"). */
typedef struct {
  double eps;        /* stop tolerance       primal_normal.cl:8   1.0e-7f            */
  double delta;      /* centring parameter   primal_normal.cl:10  0.02               */
  double r;          /* step damping         primal_normal.cl:11  0.9                */
  double ldl_delta;  /* LDL' diagonal floor  primal_normal.cl:275 1e-6               */
  double refine_tol; /* refinement threshold ldl.cl:645           1e-8               */
  int max_iter;      /*                      primal_normal.cl:9   200                */
  int max_refine;    /* ldl.cl:645: 5 (dense); 0 (sparse, ldl.cl:698-711)            */
} pycllp_b200_params;

#define PYCLLP_B200_OK 0
#define PYCLLP_B200_ERR_ARG (-1)
#define PYCLLP_B200_ERR_CUDA (-2)
#define PYCLLP_B200_ERR_STATE (-3)

/* Create an engine on CUDA device `device`. */
int pycllp_b200_create(int device, pycllp_b200_engine **out);
int pycllp_b200_destroy(pycllp_b200_engine *e);
const char *pycllp_b200_last_error(const pycllp_b200_engine *e);
const char *pycllp_b200_version(void);

/* Upload the shared constraint matrix (equality form, m rows, n columns incl. slacks)
 * and size the device state for up to max_problems LPs per solve.
 * Dense: A is m*n row-major (what cl.py:39 builds with lp.A.todense()).
 * Sparse: CSR (indptr[m+1], indices[nnz], data[nnz]); the symbolic analysis
 * (pattern of A A', cl.py:185-196) is done once here. Calling setup again replaces
 * the matrix. Parameters are reset to the defaults of the chosen path. */
int pycllp_b200_setup_dense(pycllp_b200_engine *e, int m, int n, const double *A, int max_problems);
int pycllp_b200_setup_sparse(pycllp_b200_engine *e, int m, int n, const int *indptr,
                             const int *indices, const double *data, int max_problems);

int pycllp_b200_set_params(pycllp_b200_engine *e, const pycllp_b200_params *p);
int pycllp_b200_get_params(const pycllp_b200_engine *e, pycllp_b200_params *p);

/* Solve N <= max_problems LPs from the cold start x = z = y = 1 (cl.py:108).
 * HOST buffers: b (N, m), c (N, n) in; x (N, n), y (N, m), z (N, n), status (N),
 * iters (N) out (any output may be NULL). Copies host->device, runs, copies back and
 * returns when the results are in the caller's buffers. */
int pycllp_b200_solve_host(pycllp_b200_engine *e, int N, const double *b, const double *c,
                           double *x, double *y, double *z, int *status, int *iters);

/* Same with DEVICE buffers on the engine's device; work is enqueued on `stream`
 * (a cudaStream_t, NULL = the legacy default stream) and the call returns without
 * waiting. Outputs may be NULL. */
int pycllp_b200_solve_device(pycllp_b200_engine *e, int N, const double *d_b, const double *d_c,
                             double *d_x, double *d_y, double *d_z, int *d_status, int *d_iters,
                             void *stream);

/* Kernel-level hook: one normal-equations solve per problem (ldl.cl:602-653 /
 * 656-712) on caller-supplied HOST state: x, z, c (N, n); y, b (N, m); out dy (N, m).
 * Uses the engine's matrix, ldl_delta, refine_tol and max_refine. */
int pycllp_b200_solve_primal_normal(pycllp_b200_engine *e, int N, const double *x,
                                    const double *z, const double *y, const double *b,
                                    const double *c, double mu, double *dy);

/* Kernel-level hook: (modified) LDL' of N dense symmetric m x m matrices
 * (ldl.cl:28-55 plain when modified == 0; ldl.cl:57-107 with beta, delta otherwise).
 * HOST buffers: AA (N, m, m) row-major in; L (N, m(m+1)/2) packed lower ROW-major
 * (entry (i, j) at i(i+1)/2 + j, unit diagonal) and D (N, m) out. Needs no setup. */
int pycllp_b200_ldl(pycllp_b200_engine *e, int N, int m, const double *AA, double *L, double *D,
                    int modified, double beta, double delta);

/* Profiling aid: per-phase SM cycle counters summed over blocks (thread 0 of each block,
 * clock64). enable != 0 switches counting on (and zeroes the counters), 0 switches it off;
 * out16 (may be NULL, 16 entries) receives the counters accumulated since the previous call:
 * [0] rhs+norms [1] form M [2] factor [3] triangular solves [4] residual [5] step,
 * [6..15] sub-phases of the factorisation / SYRK (see ipm_factor.cuh). */
int pycllp_b200_phase_profile(pycllp_b200_engine *e, int enable, unsigned long long *out16);

/* Introspection for the bench harness. */
long long pycllp_b200_launch_count(const pycllp_b200_engine *e); /* kernels launched so far */
int pycllp_b200_info(const pycllp_b200_engine *e, int *num_sms, int *grid, int *block,
                     size_t *smem_bytes, size_t *scratch_bytes, int *factor_in_smem);

#ifdef __cplusplus
}
#endif
#endif /* PYCLLP_B200_H */
