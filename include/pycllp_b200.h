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
  /* -- the knobs in which the reference's two implementations of this algorithm differ
   *    (preset "cl" = pycllp/cl/*.cl, preset "py" = solvers/normal_eqns.py + _ldl.pyx) -- */
  int nan_guard;     /* 1: NaN in dy => status 3 (normal_eqns.py:85-87); cl: 0       */
  int carry_v;       /* iterations during which v = A'y is carried from step to step  */
                     /* instead of recomputed (64); 0 = recompute always, the         */
                     /* reference's operation order (primal_normal.cl:76-94,139)      */
  int mu_mode;       /* 0: mu = delta gamma/(n+m) (primal_normal.cl:272); 1: /n (normal_eqns.py:65) */
  int refine_mode;   /* 0: max|r| > tol, dy += solve(r) (ldl.cl:645-652);             */
                     /* 1: max r > tol, dy -= solve(r) (_ldl.pyx:144-148)             */
  int theta_floor;   /* 1: theta = max(0, ...) (primal_normal.cl:134); 0 (normal_eqns.py:92) */
  int dz_mode;       /* 0: (mu - z dx)/x - z (primal_normal.cl:143); 1: (mu - x z - z dx)/x (normal_eqns.py:90) */
  double warm_floor; /* warm starts only: x0, z0 <- max(x0, warm_floor), max(z0, warm_floor).     */
                     /* 0 (default) = the raw end point of the previous solve, as the kernel's     */
                     /* usage note describes (primal_normal.cl:213-219); a converged point lies on  */
                     /* the boundary, where the method restarts badly -- ~1e-2 pulls it back inside */
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

/* Kernels for SMALL problems (takes effect at the next setup; default 1, env PB200_SMALL overrides):
 *   1  auto.  Dense problems with m <= 63 whose whole working set (A included) fits in a third of an
 *      SM's shared memory run on the 128-thread kernel of csrc/ipm_small.cuh, four or more LPs per
 *      SM -- "one warp group per problem" at the reference's own example size
 *      (examples/random_problem.py: m = 50, n = 100) -- whenever a solve has more than 1.5 LPs per
 *      SM (with fewer, the 512-thread kernel has the shorter latency per LP and is used).  Other
 *      shapes whose working set is below half an SM's shared memory run the 512-thread kernel two
 *      blocks per SM (its 64-register build);
 *   3  as 1, but the 128-thread kernel for every batch size;   4  as 3, with every panel of its
 *      factorisation done by the sequential rule (test switch for the fallback path);
 *   2  only the two-blocks-per-SM build;   0  neither (one 512-thread block per SM always).
 * Same algorithm, constants and results (to rounding) in every mode. */
int pycllp_b200_set_small_kernels(pycllp_b200_engine *e, int mode);

/* Numeric factor of the SPARSE path (takes effect at the next setup_sparse):
 *   1  tiles: L is stored and computed on the symbolic fill pattern only, at 8x8-tile granularity
 *      (block elimination tree + block fill analysed once per engine; memory ~ nnz(L); DMMA tile
 *      updates) -- the counterpart of sparse_factor_primal_normal on the Lindptr/Lindices pattern
 *      (ldl.cl:381-502, cl.py:185-196, prototype sparse_ldl.py:72-152);
 *   2  dense: the packed dense kernels (exact too: entries outside the fill stay zero);
 *   0  auto (default): tiles when m > 512 and the block fill is below 40 % of the triangle.
 * The tiles mode keeps no copy of M and therefore has no iterative refinement (max_refine must
 * be 0, which is the sparse default and what the reference's sparse path does, ldl.cl:698-711). */
int pycllp_b200_set_sparse_factor(pycllp_b200_engine *e, int mode);
/* Ordering of the constraints for the tile-sparse factor (takes effect at the next setup_sparse):
 *   0  auto (default): reverse Cuthill-McKee on the graph of A A' when that gives fewer tiles than
 *      the order the constraints come in;   1  natural order always (what the reference does,
 *      cl.py:185-196);   2  RCM whenever the tile factor is used.
 * A symmetric reordering of the constraints leaves the LP and its iterates unchanged up to rounding;
 * b is read and y written through the permutation, callers never see it.
 * pycllp_b200_sparse_reordered: 1 if the current engine solves with reordered constraints. */
int pycllp_b200_set_sparse_ordering(pycllp_b200_engine *e, int mode);
/* Host-only: that RCM ordering for the CSR pattern of A; perm[i] = the caller's row at position i. */
int pycllp_b200_rcm_ordering(int m, int n, const int *indptr, const int *indices, int *perm);
int pycllp_b200_sparse_reordered(const pycllp_b200_engine *e);
/* After setup_sparse: which factor is in use, its doubles per LP vs the dense m(m+1)/2, the tile
 * pair updates per factorisation and the block fill fraction. Any pointer may be NULL. */
int pycllp_b200_sparse_info(const pycllp_b200_engine *e, int *tiles_mode, long long *factor_doubles,
                            long long *dense_factor_doubles, long long *update_pairs,
                            double *tile_fill);

/* Host-only (no device needed): the symbolic analysis behind mode 1 for the CSR pattern of A.
 * Block column J of L owns tiles colptr[J] .. colptr[J+1]-1 (diagonal tile first; row[t], col[t] =
 * block row / column of tile t); tile t receives - L(upda[p]) D L(updb[p])' for
 * p in [updptr[t], updptr[t+1]).  Call with NULL arrays to get nbk, ntiles and pairs first. */
int pycllp_b200_tile_analysis(int m, int n, const int *indptr, const int *indices, int *nbk,
                              int *ntiles, long long *pairs, int *colptr, int *row, int *col,
                              int *updptr, int *upda, int *updb);

int pycllp_b200_set_params(pycllp_b200_engine *e, const pycllp_b200_params *p);
int pycllp_b200_get_params(const pycllp_b200_engine *e, pycllp_b200_params *p);
/* All constants of one of the reference's implementations at once: "cl" (the OpenCL
 * kernels; the default after setup) or "py" (DensePrimalNormalSolver, normal_eqns.py:12-97:
 * EPS 1e-8, delta 0.1, mu = delta gamma / n, refinement tolerance 1e-6 with its sign
 * convention, no floor on theta, NaN => status 3). */
int pycllp_b200_set_preset(pycllp_b200_engine *e, const char *name);

/* Solve N <= max_problems LPs from the cold start x = z = y = 1 (cl.py:108).
 * HOST buffers: b (N, m), c (N, n) in; x (N, n), y (N, m), z (N, n), status (N),
 * iters (N) out (any output may be NULL). Copies host->device, runs, copies back and
 * returns when the results are in the caller's buffers. */
int pycllp_b200_solve_host(pycllp_b200_engine *e, int N, const double *b, const double *c,
                           double *x, double *y, double *z, int *status, int *iters);

/* Extended form.  warm_start != 0: start from the x, z, y the PREVIOUS solve_host[_ex] call
 * left on the device (same N) instead of x = z = y = 1 -- the repeat-solve use the library was
 * written for (README.md:6; primal_normal.cl:213-219; cl.py:108 never enabled it).  trace
 * (may be NULL): host buffer (N, trace_iters, 3) receiving |rho|, |sigma|, gamma of every
 * iteration (what the kernels print at verbose > 1, primal_normal.cl:250-252); entries of
 * iterations a problem did not reach are NaN. */
int pycllp_b200_solve_host_ex(pycllp_b200_engine *e, int N, const double *b, const double *c,
                              int warm_start, double *x, double *y, double *z, int *status,
                              int *iters, double *trace, int trace_iters);

/* Same with DEVICE buffers on the engine's device; work is enqueued on `stream`
 * (a cudaStream_t, NULL = the legacy default stream) and the call returns without
 * waiting. Outputs may be NULL.  An engine has ONE set of scratch slots and one work
 * counter: launches of one engine are serialised on the device even when they are issued
 * on different streams (each launch waits for the previous one's completion event). */
int pycllp_b200_solve_device(pycllp_b200_engine *e, int N, const double *d_b, const double *d_c,
                             double *d_x, double *d_y, double *d_z, int *d_status, int *d_iters,
                             void *stream);

/* Device-buffer form that writes ONE packed record per problem, d_rec (N, 2n+m+1) doubles:
 * [x (n) | y (m) | z (n) | status (int32), iterations (int32)] -- the unit of the multi-GPU
 * exchange (one all-gather of the records collects everything).
 * warm_start != 0: the records hold the x, y, z of a previous solve, which is the start. */
int pycllp_b200_solve_device_packed(pycllp_b200_engine *e, int N, const double *d_b,
                                    const double *d_c, double *d_rec, int warm_start, void *stream);

/* Device-buffer form with a warm start (d_x0, d_z0, d_y0: all three or none; they may alias
 * d_x, d_z, d_y) and an optional device trace buffer (N, trace_iters, 3). */
int pycllp_b200_solve_device_ex(pycllp_b200_engine *e, int N, const double *d_b, const double *d_c,
                                const double *d_x0, const double *d_z0, const double *d_y0,
                                double *d_x, double *d_y, double *d_z, int *d_status, int *d_iters,
                                double *d_trace, int trace_iters, void *stream);

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

/* Kernel-level hook: modified LDL' of N symmetric m x m matrices ON A GIVEN SPARSE PATTERN -- the
 * reference's prototype sparse_ldl.modified_ldl (sparse_ldl.py:72-152; tests/test_ldl.py:92-108)
 * and the factor loop of sparse_factor_primal_normal (ldl.cl:422-500), run by the tile-sparse
 * factor.  Lindptr[m+1] / Lindices[nnz]: CSR of the LOWER pattern of L, diagonal LAST in every
 * row (what cl.py:185-196 builds from tril(csr(L))); the pattern must be closed under
 * elimination, as that one is.  HOST buffers: AA (N, m, m) row-major in (entries outside the
 * pattern are not read); Ldata (N, nnz) in the order of Lindices (unit diagonal) and D (N, m) out.
 * beta <= 0: beta = sqrt(max |diag|) as the prototype does.  Needs no setup. */
int pycllp_b200_sparse_ldl(pycllp_b200_engine *e, int N, int m, const int *Lindptr,
                           const int *Lindices, const double *AA, double *Ldata, double *D,
                           double beta, double delta);

/* Profiling aid: per-phase SM cycle counters summed over blocks (thread 0 of each block,
 * clock64). enable != 0 switches counting on (and zeroes the counters), 0 switches it off;
 * out16 (may be NULL, 16 entries) receives the counters accumulated since the previous call:
 * [0] rhs+norms [1] form M [2] factor [3] triangular solves [4] residual [5] step,
 * [6..15] sub-phases of the factorisation / SYRK (see ipm_factor.cuh). */
int pycllp_b200_phase_profile(pycllp_b200_engine *e, int enable, unsigned long long *out16);

/* Page-locked host memory for the layer above (the solver classes keep their result arrays in
 * it and register lp.b / lp.c in place), so that the copies of solve_host are plain DMA:
 * replaces the host side of cl.Buffer(COPY_HOST_PTR) / enqueue_copy (cl.py:46,99-121). */
int pycllp_b200_host_alloc(pycllp_b200_engine *e, size_t bytes, void **out);
int pycllp_b200_host_free(pycllp_b200_engine *e, void *p);
int pycllp_b200_host_register(pycllp_b200_engine *e, void *p, size_t bytes);
int pycllp_b200_host_unregister(pycllp_b200_engine *e, void *p);

/* Bench aid: the FP64 tensor-core (DMMA m8n8k4) rate of this device in TFLOP/s, measured now
 * (a loop of nothing but DMMAs on every SM; best of three after a warm-up). */
int pycllp_b200_fp64_probe(pycllp_b200_engine *e, double *dmma_tflops);

/* Introspection for the bench harness. */
long long pycllp_b200_launch_count(const pycllp_b200_engine *e); /* kernels launched so far */
int pycllp_b200_info(const pycllp_b200_engine *e, int *num_sms, int *grid, int *block,
                     size_t *smem_bytes, size_t *scratch_bytes, int *factor_in_smem);

#ifdef __cplusplus
}
#endif
#endif /* PYCLLP_B200_H */
