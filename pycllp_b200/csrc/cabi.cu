// cabi.cu -- the C ABI declared in include/pycllp_b200.h (engine object, setup-time
// analysis of the shared matrix, host/device solve entry points, test hooks).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <exception>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/pycllp_b200.h"
#include "ipm_types.h"
#include "ipm_host.h"

using namespace pb200;

static thread_local std::string g_create_error;

struct pycllp_b200_engine {
  int device = 0, num_sms = 0;
  size_t smem_optin = 0;
  std::string err;
  bool ready = false;
  Matrix A{};
  Params p{};
  Scratch sc{};
  std::vector<void*> matrix_allocs;   // freed on re-setup / destroy
  int grid = 0, max_problems = 0;
  size_t smem_bytes = 0;
  // staging buffers of the host-buffer entry points
  double *d_b = nullptr, *d_c = nullptr, *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
  int *d_status = nullptr, *d_iters = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t last_done = nullptr;    // completion of the most recent launch (any stream)
  bool launched = false;
  int resident = 0;                   // problems whose x, y, z the staging buffers hold (warm start)
  double* d_trace = nullptr;
  size_t trace_cap = 0;               // doubles
  double* d_dy = nullptr;             // output of the solve_primal_normal hook
  size_t dy_cap = 0;
  long long launches = 0;
  unsigned long long* d_prof = nullptr;   // phase counters (debug/profiling aid)
  // sparse numeric factor: 0 auto, 1 tiles of the symbolic pattern (ipm_tiles.cuh), 2 dense packed
  int small_mode = 1;                     // 0: never use the small-problem kernels (env PB200_SMALL=0);
                                          // 1: both; 2: only the 64-register build of the big kernel
  int tiny_grid = 0;                      // > 0: the 128-thread kernel of ipm_small.cuh is usable (m <= 64)
  size_t tiny_smem = 0;
  int sparse_order_mode = 0;              // 0 auto (RCM when it gives fewer tiles), 1 natural order, 2 RCM
  int sparse_factor_mode = 0;
  long long tile_pairs = 0;               // tile-pair updates per factorisation (tiles mode)
  double tile_fill = 0.0;                 // tiles of the block fill / tiles of the full lower triangle
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int fail(pycllp_b200_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  return code;
}

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t _err = (call);                                                           \
    if (_err != cudaSuccess)                                                             \
      return fail(e, PYCLLP_B200_ERR_CUDA,                                               \
                  std::string(#call) + ": " + cudaGetErrorString(_err));                 \
  } while (0)

void default_params(Params& p, bool sparse) {
  p.eps = (double)1.0e-7f;      // primal_normal.cl:8 -- a float literal
  p.delta = 0.02;               // :10
  p.r = 0.9;                    // :11
  p.ldl_delta = 1e-6;           // :275
  p.refine_tol = 1e-8;          // ldl.cl:645
  p.max_iter = 200;             // primal_normal.cl:9
  p.max_refine = sparse ? 0 : 5;  // ldl.cl:645 / ldl.cl:698-711 (commented out)
  p.nan_guard = 0;              // the kernels have no NaN test
  p.carry_v = 64;
  p.mu_mode = 0;                // :272
  p.refine_mode = 0;            // ldl.cl:645-652
  p.theta_floor = 1;            // primal_normal.cl:134
  p.dz_mode = 0;                // :143
  p.warm_floor = 0.0;           // :213-219: "begin from the position the previous call ended in"
}

// solvers/normal_eqns.py + _ldl.pyx
void python_params(Params& p, bool sparse) {
  default_params(p, sparse);
  p.eps = 1.0e-8;               // normal_eqns.py:12
  p.delta = 0.1;                // :52
  p.refine_tol = 1e-6;          // _ldl.pyx:132
  p.max_refine = 5;             // _ldl.pyx:144
  p.nan_guard = 1;              // normal_eqns.py:85-87
  p.mu_mode = 1;                // :65
  p.refine_mode = 1;            // _ldl.pyx:144-148
  p.theta_floor = 0;            // normal_eqns.py:92
  p.dz_mode = 1;                // :90
}

// Run an API body; C++ exceptions (bad_alloc, length_error from the setup-time vectors) must
// not cross the C boundary.
template <class F>
int guarded(pycllp_b200_engine* e, F&& body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return fail(e, PYCLLP_B200_ERR_ARG, "out of host memory");
  } catch (const std::exception& ex) {
    return fail(e, PYCLLP_B200_ERR_ARG, std::string("invalid argument: ") + ex.what());
  } catch (...) {
    return fail(e, PYCLLP_B200_ERR_ARG, "unknown C++ exception");
  }
}

template <class T>
int upload(pycllp_b200_engine* e, const std::vector<T>& h, const T** out) {
  T* d = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  CU(cudaMalloc(&d, bytes));
  e->matrix_allocs.push_back(d);
  if (!h.empty()) CU(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

void free_matrix(pycllp_b200_engine* e) {
  for (void* p : e->matrix_allocs) cudaFree(p);
  e->matrix_allocs.clear();
  void* bufs[] = {e->d_b, e->d_c, e->d_x, e->d_y, e->d_z, e->d_status, e->d_iters};
  for (void* p : bufs)
    if (p) cudaFree(p);
  e->d_b = e->d_c = e->d_x = e->d_y = e->d_z = nullptr;
  e->d_status = e->d_iters = nullptr;
  if (e->d_trace) cudaFree(e->d_trace);
  if (e->d_dy) cudaFree(e->d_dy);
  e->d_trace = e->d_dy = nullptr;
  e->trace_cap = e->dy_cap = 0;
  e->resident = 0;
  e->d_prof = nullptr;
  e->sc.prof = nullptr;
  e->ready = false;
}


// Reverse Cuthill-McKee ordering of the graph of A A' (lower pattern me_i > me_j): perm[i] = the
// original row placed at position i.  A band-reducing ordering keeps the 8x8 tiles of the block
// fill dense, which is what the tile-sparse factor needs (a minimum-degree ordering reduces the
// scalar fill but scatters it over many nearly empty tiles).  The reference factorises in the
// order the constraints come in (cl.py:185-196); a symmetric reordering of the constraints leaves
// the LP and its interior-point iterates unchanged up to rounding.
static std::vector<int> rcm_ordering(int m, const std::vector<int>& me_i, const std::vector<int>& me_j) {
  std::vector<int> deg(m, 0), ptr(m + 1, 0);
  for (size_t e = 0; e < me_i.size(); e++)
    if (me_i[e] != me_j[e]) { deg[me_i[e]]++; deg[me_j[e]]++; }
  for (int i = 0; i < m; i++) ptr[i + 1] = ptr[i] + deg[i];
  std::vector<int> adj(ptr[m]), fill(ptr.begin(), ptr.end() - 1);
  for (size_t e = 0; e < me_i.size(); e++)
    if (me_i[e] != me_j[e]) { adj[fill[me_i[e]]++] = me_j[e]; adj[fill[me_j[e]]++] = me_i[e]; }
  for (int i = 0; i < m; i++)
    std::sort(adj.begin() + ptr[i], adj.begin() + ptr[i + 1], [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
  std::vector<int> order, level(m, -1);
  order.reserve(m);
  std::vector<char> seen(m, 0);
  std::vector<int> byDeg(m);
  for (int i = 0; i < m; i++) byDeg[i] = i;
  std::stable_sort(byDeg.begin(), byDeg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
  auto bfs = [&](int root, std::vector<int>& out) {      // level structure from root; returns its last node
    out.clear();
    out.push_back(root);
    level[root] = 0;
    for (size_t h = 0; h < out.size(); h++) {
      const int v = out[h];
      for (int q = ptr[v]; q < ptr[v + 1]; q++)
        if (level[adj[q]] < 0 && !seen[adj[q]]) { level[adj[q]] = level[v] + 1; out.push_back(adj[q]); }
    }
    const int last = out.back();
    return last;
  };
  std::vector<int> comp;
  for (int s0 : byDeg) {
    if (seen[s0]) continue;
    // pseudo-peripheral start: repeat the BFS from the (lowest-degree) node of the last level while it deepens
    int root = s0, depth = -1;
    for (int it = 0; it < 8; it++) {
      const int last = bfs(root, comp);
      const int d = level[last];
      int cand = last;
      for (int v : comp)
        if (level[v] == d && deg[v] < deg[cand]) cand = v;
      for (int v : comp) level[v] = -1;
      if (d <= depth) break;
      depth = d;
      root = cand;
    }
    bfs(root, comp);
    for (int v : comp) { seen[v] = 1; level[v] = -1; order.push_back(v); }
  }
  std::reverse(order.begin(), order.end());
  return order;
}

// Symbolic analysis of the tile-sparse factor (ipm_tiles.cuh) from the lower pattern (me_i >= me_j)
// of A A': block elimination tree and block fill at 8x8-tile granularity, natural order (the
// reference does not reorder either, cl.py:185-196), tile ids, and for every tile (I, J) the
// pairs (tile (I,K), tile (J,K)), K < J ascending, whose product it receives.
struct TileSym {
  int nbk = 0, ntiles = 0;
  std::vector<int> colptr, row, col, updptr, upda, updb, me_pos;
  size_t pairs = 0;
};

// returns false if the structure exceeds max_tiles (auto mode: not worth it) or 32-bit positions
static bool tile_symbolic(int m, const std::vector<int>& me_i, const std::vector<int>& me_j,
                          size_t max_tiles, TileSym& ts) {
  const int nbk = (m + 7) / 8;
  ts.nbk = nbk;
  std::vector<std::vector<int>> st(nbk), ch(nbk);
  for (size_t e = 0; e < me_i.size(); e++) {
    const int I = me_i[e] >> 3, J = me_j[e] >> 3;
    if (I > J) st[J].push_back(I);
  }
  std::vector<int> mark(nbk, -1);
  size_t total = (size_t)nbk;
  for (int J = 0; J < nbk; J++) {
    std::vector<int> list;
    mark[J] = J;
    for (int I : st[J])
      if (mark[I] != J) { mark[I] = J; list.push_back(I); }
    for (int K : ch[J])
      for (int I : st[K])
        if (I > J && mark[I] != J) { mark[I] = J; list.push_back(I); }
    std::sort(list.begin(), list.end());
    st[J].swap(list);
    if (!st[J].empty()) ch[st[J][0]].push_back(J);
    total += st[J].size();
    if (total > max_tiles || total * 64 > (size_t)INT_MAX) return false;
  }
  ts.ntiles = (int)total;
  ts.colptr.assign(nbk + 1, 0);
  ts.row.resize(total);
  ts.col.resize(total);
  for (int J = 0; J < nbk; J++) {
    int t = ts.colptr[J];
    ts.row[t] = J; ts.col[t] = J; t++;
    for (int I : st[J]) { ts.row[t] = I; ts.col[t] = J; t++; }
    ts.colptr[J + 1] = t;
  }
  // rl[J]: block columns K < J that hold a tile (J, K), ascending, with its position in st[K]
  std::vector<std::vector<std::pair<int, int>>> rl(nbk);
  for (int K = 0; K < nbk; K++)
    for (size_t q = 0; q < st[K].size(); q++) rl[st[K][q]].push_back({K, (int)q});
  ts.updptr.assign(total + 1, 0);
  std::vector<int> pos(nbk, -1), fillpos;
  // two passes over the same enumeration: count, then fill (the lists can be large)
  for (int pass = 0; pass < 2; pass++) {
    for (int J = 0; J < nbk; J++) {
      for (int t = ts.colptr[J]; t < ts.colptr[J + 1]; t++) pos[ts.row[t]] = t;
      for (auto& kq : rl[J]) {
        const int K = kq.first, q = kq.second;
        const int b = ts.colptr[K] + 1 + q;
        for (size_t q2 = q; q2 < st[K].size(); q2++) {
          const int a = ts.colptr[K] + 1 + (int)q2, t = pos[st[K][q2]];
          if (pass == 0) ts.updptr[t + 1]++;
          else { ts.upda[fillpos[t]] = a; ts.updb[fillpos[t]] = b; fillpos[t]++; }
        }
      }
    }
    if (pass == 0) {
      size_t acc = 0;
      for (size_t t = 0; t < total; t++) {
        acc += (size_t)ts.updptr[t + 1];
        if (acc > (size_t)INT_MAX) return false;
        ts.updptr[t + 1] = (int)acc;
      }
      ts.pairs = acc;
      ts.upda.resize(acc);
      ts.updb.resize(acc);
      fillpos.assign(ts.updptr.begin(), ts.updptr.end() - 1);
    }
  }
  // where every entry of the pattern of A A' goes
  ts.me_pos.resize(me_i.size());
  for (size_t e = 0; e < me_i.size(); e++) {
    const int i = me_i[e], j = me_j[e], I = i >> 3, J = j >> 3;
    int t = ts.colptr[J];
    if (I != J) t += 1 + (int)(std::lower_bound(st[J].begin(), st[J].end(), I) - st[J].begin());
    ts.me_pos[e] = t * 64 + (j & 7) * 8 + (i & 7);
  }
  return true;
}

size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }

// choose the shared-memory configuration, the grid and allocate scratch + staging
int finish_setup(pycllp_b200_engine* e, int max_problems) {
  const int m = e->A.m, n = e->A.n;
  const size_t limit = e->smem_optin - 64;     // static smem (16 B) + margin
  int Ls = 1, Vs = 1;
  e->A.big = 0;
  if (e->A.tiles || smem_doubles(e->A, 1, 1) * 8 > limit) { Ls = 0; e->A.big = 1; }
  if (smem_doubles(e->A, Ls, 1) * 8 > limit) { Vs = 0; }
  if (smem_doubles(e->A, Ls, Vs) * 8 > limit)
    return fail(e, PYCLLP_B200_ERR_ARG, "problem too large for the shared-memory work area");
  e->smem_bytes = smem_doubles(e->A, Ls, Vs) * 8;
  // the factor: packed dense lower triangle, or (tiles mode) the tiles of the symbolic pattern
  // only -- no m x m matrix in that case
  const size_t lsz = e->A.tiles ? (size_t)e->A.ntiles * 64 : packed_doubles(m);
  size_t slot = e->A.tiles ? 0 : al16((size_t)m * m);
  e->sc.off_L = slot;
  if (!Ls) slot += al16(lsz);
  e->sc.off_vec = slot;
  if (!Vs) slot += al16(vec_area_doubles(e->A));
  e->sc.off_P = slot;
  if (!Vs) slot += al16(work_area_doubles(e->A));
  e->sc.off_dg = slot;
  if (!Vs) slot += al16(e->A.ldd > 0 ? e->A.ldd : 1);
  e->sc.slot = slot;
  e->sc.L_in_smem = Ls;
  e->sc.vec_in_smem = Vs;
  int per_sm = solve_kernel_max_blocks_per_sm(e->smem_bytes, Ls, Vs);
  if (per_sm < 1) return fail(e, PYCLLP_B200_ERR_CUDA, "kernel does not fit on an SM");
  // small dense problems: the 128-thread kernel, everything in shared memory, >= 3 blocks per SM
  e->tiny_grid = 0;
  if (!e->A.sparse && m <= 63 && (e->small_mode == 1 || e->small_mode >= 3)) {
    const size_t ts = tiny_kernel_smem_bytes(e->A);
    if (ts <= e->smem_optin - 64) {
      const int nb = tiny_kernel_blocks_per_sm(ts);
      if (nb >= 3) {
        e->tiny_smem = ts;
        e->tiny_grid = std::max(1, std::min(max_problems, e->num_sms * nb));
      }
    }
  }
  // small problems: two blocks per SM with the 64-register build of the kernel
  e->sc.small = 0;
  if (Ls && Vs && e->small_mode != 0 && max_problems > e->num_sms) {
    const int two = small_kernel_max_blocks_per_sm(e->smem_bytes);
    if (two >= 2) { e->sc.small = 1; per_sm = 2; }
  }
  e->grid = std::max(1, std::min(max_problems, e->num_sms * per_sm));
  e->max_problems = max_problems;
  double* base = nullptr;
  CU(cudaMalloc(&base, slot * sizeof(double) * e->grid));
  e->matrix_allocs.push_back(base);
  e->sc.base = base;
  int* counter = nullptr;
  CU(cudaMalloc(&counter, sizeof(int)));
  e->matrix_allocs.push_back(counter);
  e->sc.counter = counter;
  const size_t N = (size_t)max_problems;
  CU(cudaMalloc(&e->d_b, N * m * sizeof(double)));
  CU(cudaMalloc(&e->d_c, N * n * sizeof(double)));
  CU(cudaMalloc(&e->d_x, N * n * sizeof(double)));
  CU(cudaMalloc(&e->d_y, N * m * sizeof(double)));
  CU(cudaMalloc(&e->d_z, N * n * sizeof(double)));
  CU(cudaMalloc(&e->d_status, N * sizeof(int)));
  CU(cudaMalloc(&e->d_iters, N * sizeof(int)));
  e->ready = true;
  return 0;
}

int run(pycllp_b200_engine* e, Batch& B, cudaStream_t stream) {
  // one work counter and one set of scratch slots per engine: a launch may not start (not even
  // its counter reset) before the previous one, on whatever stream, has finished
  if (e->launched) CU(cudaStreamWaitEvent(stream, e->last_done, 0));
  CU(cudaMemsetAsync(e->sc.counter, 0, sizeof(int), stream));
  // (one LP per SM or less: the 512-thread kernel has the shorter latency per LP; the 128-thread
  // kernel wins as soon as there are more LPs than SMs to keep busy)
  if (e->tiny_grid > 0 && !B.hook && params_are_cl(e->p) &&
      (e->small_mode >= 3 || B.N > e->num_sms + e->num_sms / 2)) {
    Scratch sct = e->sc;
    sct.small = e->small_mode == 4 ? 4 : 0;          // 4: test switch, every panel by the sequential rule
    CU(launch_solve_tiny(e->A, B, sct, e->p, std::min(e->tiny_grid, std::max(1, B.N)), e->tiny_smem, stream));
  } else {
    int grid = std::min(e->grid, std::max(1, B.N));
    CU(launch_solve(e->A, B, e->sc, e->p, grid, e->smem_bytes, stream));
  }
  CU(cudaEventRecord(e->last_done, stream));
  e->launched = true;
  e->launches += 1;
  return 0;
}

}  // namespace

extern "C" {

const char* pycllp_b200_version(void) { return "pycllp_b200 0.1 (sm_100a)"; }

const char* pycllp_b200_last_error(const pycllp_b200_engine* e) {
  return e ? e->err.c_str() : g_create_error.c_str();
}

int pycllp_b200_create(int device, pycllp_b200_engine** out) {
  if (!out) return PYCLLP_B200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    g_create_error = std::string("no usable CUDA device: ") +
                     (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                     " (this engine has no CPU fallback)";
    return PYCLLP_B200_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return PYCLLP_B200_ERR_ARG;
  }
  pycllp_b200_engine* e = new pycllp_b200_engine();
  e->device = device;
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  err = cudaGetDeviceProperties(&prop, device);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->last_done, cudaEventDisableTiming);
  if (err != cudaSuccess) {
    g_create_error = cudaGetErrorString(err);
    delete e;
    return PYCLLP_B200_ERR_CUDA;
  }
  e->num_sms = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  default_params(e->p, false);
  if (const char* sm = getenv("PB200_SMALL")) e->small_mode = atoi(sm);
  *out = e;
  return 0;
}

int pycllp_b200_destroy(pycllp_b200_engine* e) {
  if (!e) return 0;
  {
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    free_matrix(e);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->last_done) cudaEventDestroy(e->last_done);
  }
  delete e;
  return 0;
}

static int setup_dense_impl(pycllp_b200_engine* e, int m, int n, const double* A, int max_problems) {
  DeviceGuard guard(e->device);
  free_matrix(e);
  Matrix& M = e->A;
  M = Matrix{};
  M.m = m; M.n = n; M.sparse = 0;
  default_params(e->p, false);
  // classify columns by their number of non-zeros
  std::vector<int> cnt(n, 0), one_row(n, -1);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++)
      if (A[(size_t)i * n + j] != 0.0) { cnt[j]++; one_row[j] = i; }
  std::vector<int> dcols;
  std::vector<std::vector<std::pair<int, double>>> sing(m);
  for (int j = 0; j < n; j++) {
    if (cnt[j] >= 2) dcols.push_back(j);
    else if (cnt[j] == 1) {
      double a = A[(size_t)one_row[j] * n + j];
      sing[one_row[j]].push_back({j, a * a});
    }
  }
  M.nd = (int)dcols.size();
  M.ldd = std::max(KC, (M.nd + KC - 1) / KC * KC);
  std::vector<double> Ad((size_t)m * M.ldd, 0.0);
  for (int i = 0; i < m; i++)
    for (int k = 0; k < M.nd; k++) Ad[(size_t)i * M.ldd + k] = A[(size_t)i * n + dcols[k]];
  dcols.resize(M.ldd, 0);
  std::vector<int> sptr(m + 1, 0), scol, colrow(n, -1);
  std::vector<double> sw, sa, colval(n, 0.0);
  for (int i = 0; i < m; i++) {
    for (auto& pr : sing[i]) {
      scol.push_back(pr.first);
      sw.push_back(pr.second);
      sa.push_back(A[(size_t)i * n + pr.first]);
      colrow[pr.first] = i;
      colval[pr.first] = A[(size_t)i * n + pr.first];
    }
    sptr[i + 1] = (int)scol.size();
  }
  for (int k = 0; k < M.nd; k++) colrow[dcols[k]] = -2;
  // TMA-staged SYRK operand: chunks of SY_KC packed columns, k-major, rows padded to ldm
  // with ldm = 4 (mod 16) so that the four k-rows of a DMMA fragment hit disjoint banks
  int ldm = (m + 7) / 8 * 8;
  while (ldm % 16 != 4) ldm++;
  M.sy_ldm = ldm;
  std::vector<double> Apk((size_t)M.ldd * ldm, 0.0);
  for (int k = 0; k < M.nd; k++)
    for (int i = 0; i < m; i++) Apk[(size_t)k * ldm + i] = A[(size_t)i * n + dcols[k]];
  // tile segments: tile row I holds tiles J = 0..I, cut into runs of SY_CW; longest first,
  // each to the least-loaded (pass, warp) that still has a free slot
  {
    const int T = (m + 7) / 8;
    std::vector<int4> segs;
    for (int I = 0; I < T; I++)
      for (int J = 0; J <= I; J += SY_CW) segs.push_back(make_int4(I, J, std::min(SY_CW, I + 1 - J), 0));
    std::stable_sort(segs.begin(), segs.end(), [](const int4& a, const int4& b) { return a.z > b.z; });
    const int slots = NWARP * SY_SEG;
    const int npass = std::max(1, ((int)segs.size() + slots - 1) / slots);
    std::vector<int4> tab((size_t)npass * slots, make_int4(0, 0, 0, 0));
    std::vector<int> load(npass * NWARP, 0), used(npass * NWARP, 0);
    for (const int4& sg : segs) {
      int best = -1;
      for (int pw = 0; pw < npass * NWARP; pw++)
        if (used[pw] < SY_SEG && (best < 0 || load[pw] < load[best])) best = pw;
      tab[(size_t)best * SY_SEG + used[best]] = sg;
      used[best]++;
      load[best] += sg.z;
    }
    M.sy_npass = npass;
    int rc2;
    if ((rc2 = upload(e, tab, &M.sy_seg))) return rc2;
  }
  int rc;
  if ((rc = upload(e, Apk, &M.sy_A))) return rc;
  if ((rc = upload(e, Ad, &M.Ad))) return rc;
  if ((rc = upload(e, dcols, &M.dcols))) return rc;
  if ((rc = upload(e, sptr, &M.sing_ptr))) return rc;
  if ((rc = upload(e, scol, &M.sing_col))) return rc;
  if ((rc = upload(e, sw, &M.sing_w))) return rc;
  if ((rc = upload(e, sa, &M.sing_a))) return rc;
  if ((rc = upload(e, colrow, &M.colrow))) return rc;
  if ((rc = upload(e, colval, &M.colval))) return rc;
  return finish_setup(e, max_problems);
}

int pycllp_b200_setup_dense(pycllp_b200_engine* e, int m, int n, const double* A,
                            int max_problems) {
  if (!e || !A || m <= 0 || n <= 0 || max_problems <= 0)
    return fail(e, PYCLLP_B200_ERR_ARG, "setup_dense: bad argument");
  if ((long long)m * (m + 1) / 2 + m > (long long)INT_MAX / 2)
    return fail(e, PYCLLP_B200_ERR_ARG, "setup_dense: m too large for 32-bit factor offsets");
  return guarded(e, [&] { return setup_dense_impl(e, m, n, A, max_problems); });
}

static int setup_sparse_impl(pycllp_b200_engine* e, int m, int n, const int* indptr,
                             const int* indices, const double* data, int max_problems,
                             const int* rperm = nullptr) {
  // indptr comes from the caller: it must be a CSR row pointer before anything is sized from it
  if (indptr[0] != 0) return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: indptr[0] != 0");
  for (int i = 0; i < m; i++)
    if (indptr[i + 1] < indptr[i])
      return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: indptr is not non-decreasing");
  // rperm: analyse and solve with the constraints reordered (internal row i = the caller's row rperm[i])
  std::vector<int> p_indptr, p_indices;
  std::vector<double> p_data;
  if (rperm) {
    p_indptr.assign(m + 1, 0);
    for (int i = 0; i < m; i++) p_indptr[i + 1] = p_indptr[i] + (indptr[rperm[i] + 1] - indptr[rperm[i]]);
    p_indices.resize(p_indptr[m]);
    p_data.resize(p_indptr[m]);
    for (int i = 0; i < m; i++) {
      std::copy(indices + indptr[rperm[i]], indices + indptr[rperm[i] + 1], p_indices.begin() + p_indptr[i]);
      std::copy(data + indptr[rperm[i]], data + indptr[rperm[i] + 1], p_data.begin() + p_indptr[i]);
    }
    indptr = p_indptr.data(); indices = p_indices.data(); data = p_data.data();
  }
  DeviceGuard guard(e->device);
  free_matrix(e);
  Matrix& M = e->A;
  M = Matrix{};
  M.m = m; M.n = n; M.sparse = 1;
  default_params(e->p, true);
  const int nnz = indptr[m];
  for (int k = 0; k < nnz; k++)
    if (indices[k] < 0 || indices[k] >= n)
      return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: column index out of range");
  std::vector<int> Ap(indptr, indptr + m + 1), Ai(indices, indices + nnz);
  std::vector<double> Ax(data, data + nnz);
  // CSR of A' (= CSC of A), rows ascending inside each column
  std::vector<int> Tp(n + 1, 0), Ti(nnz);
  std::vector<double> Tx(nnz);
  for (int k = 0; k < nnz; k++) Tp[Ai[k] + 1]++;
  for (int j = 0; j < n; j++) Tp[j + 1] += Tp[j];
  {
    std::vector<int> pos(Tp.begin(), Tp.end() - 1);
    for (int i = 0; i < m; i++)
      for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        int d = pos[Ai[k]]++;
        Ti[d] = i;
        Tx[d] = Ax[k];
      }
  }
  // symbolic analysis of M = A diag(d) A' (shared pattern): triples (i >= j, k, A_ik A_jk)
  struct Tr { int i, j, k; double w; };
  std::vector<Tr> tr;
  {
    size_t ntr = 0;                                  // sum over columns of c_k (c_k + 1) / 2
    for (int k = 0; k < n; k++) {
      const size_t ck = (size_t)(Tp[k + 1] - Tp[k]);
      ntr += ck * (ck + 1) / 2;
    }
    if (ntr > (size_t)INT_MAX)
      return fail(e, PYCLLP_B200_ERR_ARG,
                  "setup_sparse: the pattern of A A' has more than 2^31 terms (A too dense for the "
                  "sparse path: use setup_dense)");
    tr.reserve(ntr);
  }
  for (int k = 0; k < n; k++)
    for (int a = Tp[k]; a < Tp[k + 1]; a++)
      for (int b2 = Tp[k]; b2 <= a; b2++)
        tr.push_back({Ti[a], Ti[b2], k, Tx[a] * Tx[b2]});
  // entries in the order of the packed column-major factor storage (column j, then row i): the
  // threads of a warp then write neighbouring addresses in form_M_sparse; terms by ascending k
  std::sort(tr.begin(), tr.end(), [](const Tr& u, const Tr& v) {
    if (u.j != v.j) return u.j < v.j;
    if (u.i != v.i) return u.i < v.i;
    return u.k < v.k;
  });
  std::vector<int> me_ptr, me_i, me_j, mt_k;
  std::vector<double> mt_w;
  for (size_t t = 0; t < tr.size(); t++) {
    if (t == 0 || tr[t].i != tr[t - 1].i || tr[t].j != tr[t - 1].j) {
      me_ptr.push_back((int)t);
      me_i.push_back(tr[t].i);
      me_j.push_back(tr[t].j);
    }
    mt_k.push_back(tr[t].k);
    mt_w.push_back(tr[t].w);
  }
  me_ptr.push_back((int)tr.size());
  M.nme = (int)me_i.size();
  int rc;
  // numeric factor: the tiles of the symbolic pattern when L is genuinely sparse, else the dense
  // packed kernels (which also serve every shape whose factor fits in shared memory)
  e->tile_pairs = 0; e->tile_fill = 0.0;
  if (e->sparse_factor_mode != 2) {
    const int nbk = (m + 7) / 8;
    const size_t full = (size_t)nbk * (nbk + 1) / 2;
    TileSym ts;
    const size_t cap = e->sparse_factor_mode == 1 ? (size_t)INT_MAX / 64 : (size_t)(0.4 * (double)full) + 1;
    const bool fits = tile_symbolic(m, me_i, me_j, cap, ts);
    // a band-reducing reordering of the constraints (RCM) when it gives fewer tiles than the order
    // the constraints came in (auto), or on request
    if (!rperm && e->sparse_order_mode != 1 && (e->sparse_factor_mode == 1 || m > 512)) {
      std::vector<int> perm = rcm_ordering(m, me_i, me_j), inv(m);
      for (int i = 0; i < m; i++) inv[perm[i]] = i;
      std::vector<int> qi(me_i.size()), qj(me_i.size());
      for (size_t q = 0; q < me_i.size(); q++) {
        const int a = inv[me_i[q]], b = inv[me_j[q]];
        qi[q] = std::max(a, b); qj[q] = std::min(a, b);
      }
      TileSym tr2;
      const bool fits2 = tile_symbolic(m, qi, qj, cap, tr2);
      if (fits2 && (e->sparse_order_mode == 2 || !fits || tr2.ntiles < ts.ntiles))
        return setup_sparse_impl(e, m, n, indptr, indices, data, max_problems, perm.data());
    }
    if (!fits && e->sparse_factor_mode == 1)
      return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: tile structure exceeds 32-bit positions");
    if (fits && (e->sparse_factor_mode == 1 || m > 512)) {
      M.tiles = 1; M.nbk = ts.nbk; M.ntiles = ts.ntiles;
      e->tile_pairs = (long long)ts.pairs;
      e->tile_fill = (double)ts.ntiles / (double)full;
      if ((rc = upload(e, ts.colptr, &M.tl_colptr))) return rc;
      if ((rc = upload(e, ts.row, &M.tl_row))) return rc;
      if ((rc = upload(e, ts.col, &M.tl_col))) return rc;
      if ((rc = upload(e, ts.updptr, &M.tl_updptr))) return rc;
      if ((rc = upload(e, ts.upda, &M.tl_upda))) return rc;
      if ((rc = upload(e, ts.updb, &M.tl_updb))) return rc;
      {
        std::vector<int> updk(ts.upda.size());
        for (size_t q = 0; q < updk.size(); q++) updk[q] = ts.col[ts.upda[q]];
        if ((rc = upload(e, updk, &M.tl_updk))) return rc;
      }
      if ((rc = upload(e, ts.me_pos, &M.me_pos))) return rc;
    }
  }
  if (rperm) {                                       // (the matrix below IS the reordered one)
    std::vector<int> pv(rperm, rperm + m);
    if ((rc = upload(e, pv, &M.rperm))) return rc;
  }
  if ((rc = upload(e, Ap, &M.Ap))) return rc;
  if ((rc = upload(e, Ai, &M.Ai))) return rc;
  if ((rc = upload(e, Ax, &M.Ax))) return rc;
  if ((rc = upload(e, Tp, &M.Tp))) return rc;
  if ((rc = upload(e, Ti, &M.Ti))) return rc;
  if ((rc = upload(e, Tx, &M.Tx))) return rc;
  if ((rc = upload(e, me_ptr, &M.me_ptr))) return rc;
  if ((rc = upload(e, me_i, &M.me_i))) return rc;
  if ((rc = upload(e, me_j, &M.me_j))) return rc;
  if ((rc = upload(e, mt_k, &M.mt_k))) return rc;
  if ((rc = upload(e, mt_w, &M.mt_w))) return rc;
  return finish_setup(e, max_problems);
}

int pycllp_b200_setup_sparse(pycllp_b200_engine* e, int m, int n, const int* indptr,
                             const int* indices, const double* data, int max_problems) {
  if (!e || !indptr || !indices || !data || m <= 0 || n <= 0 || max_problems <= 0)
    return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: bad argument");
  if ((long long)m * (m + 1) / 2 + m > (long long)INT_MAX / 2)
    return fail(e, PYCLLP_B200_ERR_ARG, "setup_sparse: m too large for 32-bit factor offsets");
  return guarded(e, [&] { return setup_sparse_impl(e, m, n, indptr, indices, data, max_problems); });
}

int pycllp_b200_set_params(pycllp_b200_engine* e, const pycllp_b200_params* p) {
  if (!e || !p) return fail(e, PYCLLP_B200_ERR_ARG, "set_params: null argument");
  if (p->max_iter < 0 || p->max_refine < 0 || !(p->r > 0))
    return fail(e, PYCLLP_B200_ERR_ARG, "set_params: invalid value");
  if (e->ready && e->A.tiles && p->max_refine > 0)
    return fail(e, PYCLLP_B200_ERR_ARG,
                "set_params: the tile-sparse factor keeps no copy of M, so it has no refinement "
                "(like the reference's sparse path, ldl.cl:698-711); use max_refine = 0");
  e->p.eps = p->eps; e->p.delta = p->delta; e->p.r = p->r; e->p.ldl_delta = p->ldl_delta;
  e->p.refine_tol = p->refine_tol; e->p.max_iter = p->max_iter; e->p.max_refine = p->max_refine;
  e->p.nan_guard = p->nan_guard != 0; e->p.carry_v = p->carry_v < 0 ? 0 : p->carry_v;
  e->p.mu_mode = p->mu_mode != 0; e->p.refine_mode = p->refine_mode != 0;
  e->p.theta_floor = p->theta_floor != 0; e->p.dz_mode = p->dz_mode != 0;
  e->p.warm_floor = p->warm_floor > 0 ? p->warm_floor : 0.0;
  return 0;
}

int pycllp_b200_set_preset(pycllp_b200_engine* e, const char* name) {
  if (!e || !name) return fail(e, PYCLLP_B200_ERR_ARG, "set_preset: null argument");
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "set_preset: call setup_dense/setup_sparse first");
  if (!strcmp(name, "cl")) default_params(e->p, e->A.sparse != 0);
  else if (!strcmp(name, "py")) {
    if (e->A.tiles) return fail(e, PYCLLP_B200_ERR_ARG, "set_preset: 'py' needs refinement; not with the tile-sparse factor");
    python_params(e->p, e->A.sparse != 0);
  }
  else return fail(e, PYCLLP_B200_ERR_ARG, "set_preset: unknown preset (cl, py)");
  return 0;
}

int pycllp_b200_get_params(const pycllp_b200_engine* e, pycllp_b200_params* p) {
  if (!e || !p) return PYCLLP_B200_ERR_ARG;
  p->eps = e->p.eps; p->delta = e->p.delta; p->r = e->p.r; p->ldl_delta = e->p.ldl_delta;
  p->refine_tol = e->p.refine_tol; p->max_iter = e->p.max_iter; p->max_refine = e->p.max_refine;
  p->nan_guard = e->p.nan_guard; p->carry_v = e->p.carry_v; p->mu_mode = e->p.mu_mode;
  p->refine_mode = e->p.refine_mode; p->theta_floor = e->p.theta_floor; p->dz_mode = e->p.dz_mode;
  p->warm_floor = e->p.warm_floor;
  return 0;
}

int pycllp_b200_solve_device_ex(pycllp_b200_engine* e, int N, const double* d_b, const double* d_c,
                                const double* d_x0, const double* d_z0, const double* d_y0,
                                double* d_x, double* d_y, double* d_z, int* d_status, int* d_iters,
                                double* d_trace, int trace_iters, void* stream) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "solve: call setup_dense/setup_sparse first");
  if (N < 0 || !d_b || !d_c) return fail(e, PYCLLP_B200_ERR_ARG, "solve: bad argument");
  const int nstart = (d_x0 != nullptr) + (d_z0 != nullptr) + (d_y0 != nullptr);
  if (nstart != 0 && nstart != 3)
    return fail(e, PYCLLP_B200_ERR_ARG, "solve: a warm start needs x0, z0 and y0");
  if (d_trace && trace_iters <= 0) return fail(e, PYCLLP_B200_ERR_ARG, "solve: trace_iters <= 0");
  if (N == 0) return 0;
  DeviceGuard guard(e->device);
  Batch B{};
  B.N = N; B.b = d_b; B.c = d_c; B.x = d_x; B.y = d_y; B.z = d_z;
  B.status = d_status; B.iters = d_iters;
  B.warm = nstart == 3; B.x0 = d_x0; B.z0 = d_z0; B.y0 = d_y0;
  B.trace = d_trace; B.trace_iters = d_trace ? trace_iters : 0;
  return run(e, B, (cudaStream_t)stream);
}

int pycllp_b200_solve_device_packed(pycllp_b200_engine* e, int N, const double* d_b,
                                    const double* d_c, double* d_rec, int warm_start, void* stream) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "solve: call setup_dense/setup_sparse first");
  if (N < 0 || !d_b || !d_c || !d_rec) return fail(e, PYCLLP_B200_ERR_ARG, "solve: bad argument");
  if (N == 0) return 0;
  DeviceGuard guard(e->device);
  const size_t m = e->A.m, n = e->A.n, ld = 2 * n + m + 1;
  Batch B{};
  B.N = N; B.b = d_b; B.c = d_c;
  B.x = d_rec; B.y = d_rec + n; B.z = d_rec + n + m;
  B.status = reinterpret_cast<int*>(d_rec + 2 * n + m);
  B.iters = B.status + 1;
  B.ld_x = B.ld_y = B.ld_z = ld;
  B.ld_s = (int)(2 * ld);
  if (warm_start) { B.warm = 1; B.x0 = B.x; B.y0 = B.y; B.z0 = B.z; B.ld_0 = ld; }
  return run(e, B, (cudaStream_t)stream);
}

int pycllp_b200_solve_device(pycllp_b200_engine* e, int N, const double* d_b, const double* d_c,
                             double* d_x, double* d_y, double* d_z, int* d_status, int* d_iters,
                             void* stream) {
  return pycllp_b200_solve_device_ex(e, N, d_b, d_c, nullptr, nullptr, nullptr, d_x, d_y, d_z, d_status,
                                     d_iters, nullptr, 0, stream);
}

int pycllp_b200_solve_host_ex(pycllp_b200_engine* e, int N, const double* b, const double* c,
                              int warm_start, double* x, double* y, double* z, int* status,
                              int* iters, double* trace, int trace_iters) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "solve: call setup_dense/setup_sparse first");
  if (N < 0 || N > e->max_problems || !b || !c)
    return fail(e, PYCLLP_B200_ERR_ARG, "solve: bad argument (N > max_problems?)");
  if (warm_start && e->resident != N)
    return fail(e, PYCLLP_B200_ERR_STATE,
                "solve: warm start needs the x, y, z of a previous solve of the same N problems");
  if (trace && trace_iters <= 0) return fail(e, PYCLLP_B200_ERR_ARG, "solve: trace_iters <= 0");
  if (N == 0) return 0;
  DeviceGuard guard(e->device);
  const size_t m = e->A.m, n = e->A.n;
  cudaStream_t s = e->stream;
  const size_t tdoubles = trace ? (size_t)N * trace_iters * 3 : 0;
  if (tdoubles > e->trace_cap) {
    if (e->d_trace) CU(cudaFree(e->d_trace));
    e->d_trace = nullptr; e->trace_cap = 0;
    CU(cudaMalloc(&e->d_trace, tdoubles * sizeof(double)));
    e->trace_cap = tdoubles;
  }
  if (trace) CU(cudaMemsetAsync(e->d_trace, 0xff, tdoubles * sizeof(double), s));   // all-ones = NaN
  CU(cudaMemcpyAsync(e->d_b, b, N * m * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->d_c, c, N * n * sizeof(double), cudaMemcpyHostToDevice, s));
  Batch B{};
  B.N = N; B.b = e->d_b; B.c = e->d_c; B.x = e->d_x; B.y = e->d_y; B.z = e->d_z;
  B.status = e->d_status; B.iters = e->d_iters;
  if (warm_start) { B.warm = 1; B.x0 = e->d_x; B.z0 = e->d_z; B.y0 = e->d_y; }
  if (trace) { B.trace = e->d_trace; B.trace_iters = trace_iters; }
  e->resident = 0;
  int rc = run(e, B, s);
  if (rc) return rc;
  if (x) CU(cudaMemcpyAsync(x, e->d_x, N * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (y) CU(cudaMemcpyAsync(y, e->d_y, N * m * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (z) CU(cudaMemcpyAsync(z, e->d_z, N * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (status) CU(cudaMemcpyAsync(status, e->d_status, N * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (iters) CU(cudaMemcpyAsync(iters, e->d_iters, N * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (trace) CU(cudaMemcpyAsync(trace, e->d_trace, tdoubles * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  e->resident = N;
  return 0;
}

int pycllp_b200_solve_host(pycllp_b200_engine* e, int N, const double* b, const double* c,
                           double* x, double* y, double* z, int* status, int* iters) {
  return pycllp_b200_solve_host_ex(e, N, b, c, 0, x, y, z, status, iters, nullptr, 0);
}

int pycllp_b200_solve_primal_normal(pycllp_b200_engine* e, int N, const double* x,
                                    const double* z, const double* y, const double* b,
                                    const double* c, double mu, double* dy) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "solve_primal_normal: call setup first");
  if (N <= 0 || N > e->max_problems || !x || !z || !y || !b || !c || !dy)
    return fail(e, PYCLLP_B200_ERR_ARG, "solve_primal_normal: bad argument");
  DeviceGuard guard(e->device);
  const size_t m = e->A.m, n = e->A.n;
  cudaStream_t s = e->stream;
  e->resident = 0;                                   // (the staging buffers are reused below)
  if ((size_t)N * m > e->dy_cap) {
    if (e->d_dy) CU(cudaFree(e->d_dy));
    e->d_dy = nullptr; e->dy_cap = 0;
    CU(cudaMalloc(&e->d_dy, (size_t)N * m * sizeof(double)));
    e->dy_cap = (size_t)N * m;
  }
  double* d_dy = e->d_dy;
  cudaError_t err = cudaSuccess;
  auto cp = [&](void* d, const void* h, size_t bytes) {
    if (err == cudaSuccess) err = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
  };
  cp(e->d_b, b, N * m * sizeof(double));
  cp(e->d_c, c, N * n * sizeof(double));
  cp(e->d_x, x, N * n * sizeof(double));
  cp(e->d_z, z, N * n * sizeof(double));
  cp(e->d_y, y, N * m * sizeof(double));
  int rc = 0;
  if (err == cudaSuccess) {
    Batch B{};
    B.N = N; B.b = e->d_b; B.c = e->d_c; B.hook = 1;
    B.x0 = e->d_x; B.z0 = e->d_z; B.y0 = e->d_y; B.dy_out = d_dy; B.mu = mu;
    rc = run(e, B, s);
    if (!rc) err = cudaMemcpyAsync(dy, d_dy, N * m * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (!rc && err == cudaSuccess) err = cudaStreamSynchronize(s);
  }
  cudaStreamSynchronize(s);
  if (rc) return rc;
  if (err != cudaSuccess) return fail(e, PYCLLP_B200_ERR_CUDA, cudaGetErrorString(err));
  return 0;
}

int pycllp_b200_ldl(pycllp_b200_engine* e, int N, int m, const double* AA, double* L, double* D,
                    int modified, double beta, double delta) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (N <= 0 || m <= 0 || !AA || !L || !D) return fail(e, PYCLLP_B200_ERR_ARG, "ldl: bad argument");
  DeviceGuard guard(e->device);
  const size_t lsz = (size_t)m * (m + 1) / 2;
  const size_t slot = al16((size_t)2 * m * NB + 512 + packed_doubles(m) + m);
  const int grid = std::min(N, e->num_sms);
  double *d_AA = nullptr, *d_L = nullptr, *d_D = nullptr, *d_s = nullptr;
  cudaError_t err = cudaMalloc(&d_AA, (size_t)N * m * m * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_L, N * lsz * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_D, (size_t)N * m * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_s, slot * grid * sizeof(double));
  cudaStream_t s = e->stream;
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d_AA, AA, (size_t)N * m * m * sizeof(double), cudaMemcpyHostToDevice, s);
  if (err == cudaSuccess) {
    err = launch_ldl_hook(N, m, d_AA, d_L, d_D, modified, beta, delta, d_s, slot, grid, s);
    e->launches += 1;
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(L, d_L, N * lsz * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaMemcpyAsync(D, d_D, (size_t)N * m * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  cudaStreamSynchronize(s);
  cudaFree(d_AA); cudaFree(d_L); cudaFree(d_D); cudaFree(d_s);
  if (err != cudaSuccess) return fail(e, PYCLLP_B200_ERR_CUDA, cudaGetErrorString(err));
  return 0;
}

static int sparse_ldl_impl(pycllp_b200_engine* e, int N, int m, const int* indptr, const int* indices,
                           const double* AA, double* Ldata, double* D, double beta, double delta) {
  if (indptr[0] != 0) return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: indptr[0] != 0");
  for (int i = 0; i < m; i++)
    if (indptr[i + 1] <= indptr[i])
      return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: every row needs its diagonal entry");
  const int nnz = indptr[m];
  std::vector<int> pi(nnz), pj(nnz);
  for (int i = 0; i < m; i++) {
    for (int k = indptr[i]; k < indptr[i + 1]; k++) {
      if (indices[k] < 0 || indices[k] > i)
        return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: pattern must be lower triangular");
      pi[k] = i; pj[k] = indices[k];
    }
    if (indices[indptr[i + 1] - 1] != i)
      return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: the diagonal must be the last entry of its row");
  }
  TileSym ts;
  if (!tile_symbolic(m, pi, pj, (size_t)INT_MAX / 64, ts))
    return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: tile structure exceeds 32-bit positions");
  DeviceGuard guard(e->device);
  Matrix M{};
  M.m = m; M.tiles = 1; M.nbk = ts.nbk; M.ntiles = ts.ntiles;
  std::vector<int> updk(ts.upda.size());
  for (size_t q = 0; q < updk.size(); q++) updk[q] = ts.col[ts.upda[q]];
  std::vector<void*> allocs;
  auto up = [&](const std::vector<int>& h, const int** out) -> cudaError_t {
    int* d = nullptr;
    cudaError_t er = cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(int));
    if (er != cudaSuccess) return er;
    allocs.push_back(d);
    *out = d;
    return h.empty() ? cudaSuccess : cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice);
  };
  const int *d_pi = nullptr, *d_pj = nullptr;
  cudaError_t err = up(ts.colptr, &M.tl_colptr);
  if (err == cudaSuccess) err = up(ts.row, &M.tl_row);
  if (err == cudaSuccess) err = up(ts.col, &M.tl_col);
  if (err == cudaSuccess) err = up(ts.updptr, &M.tl_updptr);
  if (err == cudaSuccess) err = up(ts.upda, &M.tl_upda);
  if (err == cudaSuccess) err = up(ts.updb, &M.tl_updb);
  if (err == cudaSuccess) err = up(updk, &M.tl_updk);
  if (err == cudaSuccess) err = up(ts.me_pos, &M.me_pos);
  if (err == cudaSuccess) err = up(pi, &d_pi);
  if (err == cudaSuccess) err = up(pj, &d_pj);
  const size_t slot = al16((size_t)ts.ntiles * 64 + 8 * (size_t)ts.nbk);
  const int grid = std::min(N, e->num_sms);
  double *d_AA = nullptr, *d_L = nullptr, *d_D = nullptr, *d_s = nullptr;
  if (err == cudaSuccess) err = cudaMalloc(&d_AA, (size_t)N * m * m * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_L, (size_t)N * nnz * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_D, (size_t)N * m * sizeof(double));
  if (err == cudaSuccess) err = cudaMalloc(&d_s, slot * grid * sizeof(double));
  cudaStream_t s = e->stream;
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d_AA, AA, (size_t)N * m * m * sizeof(double), cudaMemcpyHostToDevice, s);
  if (err == cudaSuccess) {
    err = launch_tiles_hook(M, N, nnz, d_pi, d_pj, d_AA, d_L, d_D, beta, delta, d_s, slot, grid, s);
    e->launches += 1;
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(Ldata, d_L, (size_t)N * nnz * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaMemcpyAsync(D, d_D, (size_t)N * m * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  cudaStreamSynchronize(s);
  cudaFree(d_AA); cudaFree(d_L); cudaFree(d_D); cudaFree(d_s);
  for (void* p : allocs) cudaFree(p);
  if (err != cudaSuccess) return fail(e, PYCLLP_B200_ERR_CUDA, cudaGetErrorString(err));
  return 0;
}

int pycllp_b200_sparse_ldl(pycllp_b200_engine* e, int N, int m, const int* Lindptr, const int* Lindices,
                           const double* AA, double* Ldata, double* D, double beta, double delta) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (N <= 0 || m <= 0 || !Lindptr || !Lindices || !AA || !Ldata || !D)
    return fail(e, PYCLLP_B200_ERR_ARG, "sparse_ldl: bad argument");
  return guarded(e, [&] { return sparse_ldl_impl(e, N, m, Lindptr, Lindices, AA, Ldata, D, beta, delta); });
}

int pycllp_b200_phase_profile(pycllp_b200_engine* e, int enable, unsigned long long* out16) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return fail(e, PYCLLP_B200_ERR_STATE, "phase_profile: call setup first");
  DeviceGuard guard(e->device);
  const int pgrid = std::max(e->grid, e->tiny_grid);
  const size_t bytes = (size_t)pgrid * 16 * sizeof(unsigned long long);
  if (out16) {
    for (int k = 0; k < 16; k++) out16[k] = 0;
    if (e->d_prof) {
      std::vector<unsigned long long> h((size_t)pgrid * 16);
      CU(cudaDeviceSynchronize());
      CU(cudaMemcpy(h.data(), e->d_prof, bytes, cudaMemcpyDeviceToHost));
      for (int g = 0; g < pgrid; g++)
        for (int k = 0; k < 16; k++) out16[k] += h[(size_t)g * 16 + k];
    }
  }
  if (enable && !e->d_prof) {
    CU(cudaMalloc(&e->d_prof, bytes));
    e->matrix_allocs.push_back(e->d_prof);
  }
  if (e->d_prof) CU(cudaMemset(e->d_prof, 0, bytes));
  e->sc.prof = enable ? e->d_prof : nullptr;
  return 0;
}

// Page-locked host memory for the plugin layer: with pageable buffers every cudaMemcpyAsync of
// solve_host goes through the driver's staging copy (the reference's pyopencl path has the same
// cost, cl.py:99-121); pinned in place the copies are plain DMA.
int pycllp_b200_host_alloc(pycllp_b200_engine* e, size_t bytes, void** out) {
  if (!e || !out) return PYCLLP_B200_ERR_ARG;
  DeviceGuard guard(e->device);
  *out = nullptr;
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
int pycllp_b200_host_free(pycllp_b200_engine* e, void* p) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  DeviceGuard guard(e->device);
  if (p) CU(cudaFreeHost(p));
  return 0;
}
int pycllp_b200_host_register(pycllp_b200_engine* e, void* p, size_t bytes) {
  if (!e || !p || !bytes) return PYCLLP_B200_ERR_ARG;
  DeviceGuard guard(e->device);
  CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return 0;
}
int pycllp_b200_host_unregister(pycllp_b200_engine* e, void* p) {
  if (!e || !p) return PYCLLP_B200_ERR_ARG;
  DeviceGuard guard(e->device);
  CU(cudaHostUnregister(p));
  return 0;
}

int pycllp_b200_fp64_probe(pycllp_b200_engine* e, double* dmma_tflops) {
  if (!e || !dmma_tflops) return PYCLLP_B200_ERR_ARG;
  DeviceGuard guard(e->device);
  const int blocks = e->num_sms, iters = 20000;
  double* out = nullptr;
  CU(cudaMalloc(&out, sizeof(double) * (size_t)blocks * NT));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t err = cudaEventCreate(&e0);
  if (err == cudaSuccess) err = cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) {          // rep 0 = warm-up
    err = cudaEventRecord(e0, e->stream);
    if (err == cudaSuccess) err = launch_fp64_probe(out, blocks, iters, e->stream);
    if (err == cudaSuccess) err = cudaEventRecord(e1, e->stream);
    if (err == cudaSuccess) err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return fail(e, PYCLLP_B200_ERR_CUDA, cudaGetErrorString(err));
  // 8 tiles x (8 x 8 x 4 MACs) x 2 flops per warp and iteration
  const double flops = 2.0 * 256 * 8 * (double)iters * (NT / 32) * blocks;
  *dmma_tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}

long long pycllp_b200_launch_count(const pycllp_b200_engine* e) { return e ? e->launches : 0; }

int pycllp_b200_info(const pycllp_b200_engine* e, int* num_sms, int* grid, int* block,
                     size_t* smem_bytes, size_t* scratch_bytes, int* factor_in_smem) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (num_sms) *num_sms = e->num_sms;
  const bool tiny = e->tiny_grid > 0 && params_are_cl(e->p) &&
                    (e->small_mode >= 3 || e->max_problems > e->num_sms + e->num_sms / 2);
  if (grid) *grid = tiny ? e->tiny_grid : e->grid;
  if (block) *block = tiny ? 128 : NT;
  if (smem_bytes) *smem_bytes = tiny ? e->tiny_smem : e->smem_bytes;
  if (scratch_bytes) *scratch_bytes = e->sc.slot * sizeof(double) * (size_t)e->grid;
  if (factor_in_smem) *factor_in_smem = e->sc.L_in_smem;
  return 0;
}

// lower pattern (i >= j, diagonal included where a row is non-empty) of A A' from the CSR pattern of A
static bool aat_pattern(int m, int n, const int* indptr, const int* indices, std::vector<int>& me_i,
                        std::vector<int>& me_j) {
  if (indptr[0] != 0) return false;
  for (int i = 0; i < m; i++)
    if (indptr[i + 1] < indptr[i]) return false;
  std::vector<std::vector<int>> colrows(n);
  for (int i = 0; i < m; i++)
    for (int k = indptr[i]; k < indptr[i + 1]; k++) {
      if (indices[k] < 0 || indices[k] >= n) return false;
      colrows[indices[k]].push_back(i);
    }
  std::vector<std::pair<int, int>> ent;
  for (int k = 0; k < n; k++)
    for (size_t a = 0; a < colrows[k].size(); a++)
      for (size_t b = 0; b <= a; b++) ent.push_back({colrows[k][a], colrows[k][b]});
  std::sort(ent.begin(), ent.end());
  ent.erase(std::unique(ent.begin(), ent.end()), ent.end());
  me_i.resize(ent.size());
  me_j.resize(ent.size());
  for (size_t e = 0; e < ent.size(); e++) { me_i[e] = ent[e].first; me_j[e] = ent[e].second; }
  return true;
}

// Host-only: the RCM ordering setup_sparse would consider for this pattern; perm[i] = the caller's
// row placed at position i.
int pycllp_b200_rcm_ordering(int m, int n, const int* indptr, const int* indices, int* perm) {
  if (m <= 0 || n <= 0 || !indptr || !indices || !perm) return PYCLLP_B200_ERR_ARG;
  try {
    std::vector<int> me_i, me_j;
    if (!aat_pattern(m, n, indptr, indices, me_i, me_j)) return PYCLLP_B200_ERR_ARG;
    std::vector<int> p = rcm_ordering(m, me_i, me_j);
    std::copy(p.begin(), p.end(), perm);
    return 0;
  } catch (...) {
    return PYCLLP_B200_ERR_ARG;
  }
}

// Host-only: the symbolic analysis of the tile-sparse factor for a CSR pattern (no device, no
// engine) -- what setup_sparse computes, exposed so that the CPU test-suite can check it.
int pycllp_b200_tile_analysis(int m, int n, const int* indptr, const int* indices, int* nbk, int* ntiles,
                              long long* pairs, int* colptr, int* row, int* col, int* updptr, int* upda,
                              int* updb) {
  if (m <= 0 || n <= 0 || !indptr || !indices) return PYCLLP_B200_ERR_ARG;
  try {
    std::vector<int> me_i, me_j;
    if (!aat_pattern(m, n, indptr, indices, me_i, me_j)) return PYCLLP_B200_ERR_ARG;
    TileSym ts;
    if (!tile_symbolic(m, me_i, me_j, (size_t)INT_MAX / 64, ts)) return PYCLLP_B200_ERR_ARG;
    if (nbk) *nbk = ts.nbk;
    if (ntiles) *ntiles = ts.ntiles;
    if (pairs) *pairs = (long long)ts.pairs;
    if (colptr) std::copy(ts.colptr.begin(), ts.colptr.end(), colptr);
    if (row) std::copy(ts.row.begin(), ts.row.end(), row);
    if (col) std::copy(ts.col.begin(), ts.col.end(), col);
    if (updptr) std::copy(ts.updptr.begin(), ts.updptr.end(), updptr);
    if (upda) std::copy(ts.upda.begin(), ts.upda.end(), upda);
    if (updb) std::copy(ts.updb.begin(), ts.updb.end(), updb);
    return 0;
  } catch (...) {
    return PYCLLP_B200_ERR_ARG;
  }
}

int pycllp_b200_set_small_kernels(pycllp_b200_engine* e, int mode) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (mode < 0 || mode > 4) return fail(e, PYCLLP_B200_ERR_ARG, "set_small_kernels: mode must be 0 .. 4");
  e->small_mode = mode;
  return 0;
}

int pycllp_b200_set_sparse_factor(pycllp_b200_engine* e, int mode) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (mode < 0 || mode > 2) return fail(e, PYCLLP_B200_ERR_ARG, "set_sparse_factor: mode must be 0 (auto), 1 (tiles) or 2 (dense)");
  e->sparse_factor_mode = mode;
  return 0;
}

int pycllp_b200_set_sparse_ordering(pycllp_b200_engine* e, int mode) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (mode < 0 || mode > 2) return fail(e, PYCLLP_B200_ERR_ARG, "set_sparse_ordering: mode must be 0 (auto), 1 (natural) or 2 (rcm)");
  e->sparse_order_mode = mode;
  return 0;
}

int pycllp_b200_sparse_reordered(const pycllp_b200_engine* e) { return (e && e->ready && e->A.rperm) ? 1 : 0; }

int pycllp_b200_sparse_info(const pycllp_b200_engine* e, int* tiles_mode, long long* factor_doubles,
                            long long* dense_factor_doubles, long long* update_pairs, double* tile_fill) {
  if (!e) return PYCLLP_B200_ERR_ARG;
  if (!e->ready) return PYCLLP_B200_ERR_STATE;
  const long long m = e->A.m;
  if (tiles_mode) *tiles_mode = e->A.tiles;
  if (factor_doubles) *factor_doubles = e->A.tiles ? (long long)e->A.ntiles * 64 : (long long)packed_doubles((int)m);
  if (dense_factor_doubles) *dense_factor_doubles = m * (m + 1) / 2;
  if (update_pairs) *update_pairs = e->tile_pairs;
  if (tile_fill) *tile_fill = e->tile_fill;
  return 0;
}

}  // extern "C"
