#!/usr/bin/env python
"""bench.py -- batched shared-A LP solves/sec on 1..8 B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU kernels

A "step" is ONE solve of the whole batch: every LP of the workload taken from the cold
start x = z = y = 1 to its terminal status (cl.py:108-112).  Workloads (all in equality
form, synthetic, generator = examples/random_problem.py):

    cfg3 (default)  dense random LP m=200 n=400, 4096 problems PER GPU (BASELINE.json
                    configs[2], "dense LDL' path on 1 B200"); weak scaling over ranks
    cfg5            dense random LP m=500 n=1000, 65536/8 = 8192 problems per GPU
    cfg1            examples/random_problem.py m=50 n=100, 64 problems (parity-test size)
    cfg4            sparse random LP m=2000 n=5000 (1 % density + slacks), 1024 problems per GPU,
                    sparse solver path (CSR mat-vecs, pattern-based M, super-panel factor)

Timed quantities of the main line
    value   solves/s with b, c already resident in HBM: pycllp_b200_solve_device_packed writes one
            packed record [x | y | z | status, iterations] per LP and, with N > 1 ranks, ONE NCCL
            all_gather_into_tensor of the records straight from that buffer follows inside the
            timed step (the path's only exchange).  CUDA events on the stream the
            kernel is launched on, one event pair per step, L2 flushed between steps, max over ranks.
    e2e     the same through the plugin API a pycllp user calls -- lp.solve(solver) with
            solver_registry['cl_dense_primal_normal' | 'cl_sparse_primal_normal'] and ordinary
            numpy lp.b / lp.c: H2D of b, c and D2H of x, y, z, status, iterations (and the
            all-gather for N > 1) inside the timed region.  e2e_cabi_pinned: the bare C-ABI call
            pycllp_b200_solve_host with caller-pinned buffers, for comparison.
A secondary record (key "secondary") runs BASELINE.json configs[4] as a STRONG-scaling point:
cfg5 with a fixed total of 65 536 LPs split over the ranks, one timed step, gather included.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(m=200, n0=200, density=1.0, batch=4096),
    "cfg5": dict(m=500, n0=500, density=1.0, batch=8192),
    "cfg1": dict(m=50, n0=50, density=0.1, batch=64),
    "cfg4": dict(m=2000, n0=3000, density=0.01, batch=1024, sparse=True),
    # not a BASELINE.json config: a genuinely sparse LP (staircase structure, L < 2 % dense) for the
    # tile-sparse numeric factor; the dense kernels would need 144 MB of factor per LP
    "stair": dict(m=6000, n0=9000, density=4.0 / 48, batch=1024, sparse=True, band=48, per_col=4),
}
METRIC = "batched_lp_solves_per_sec"
STRONG_TOTAL = 65536        # BASELINE.json configs[4]: "batch 65536 sharded across 1/2/4/8 B200"


def make_problem(name, seed_offset, batch=None):
    """A (shared, seed 0) and a slice of b, c (seeded by seed_offset: disjoint problems per rank)."""
    from scipy.sparse import rand
    w = WORKLOADS[name]
    m, n0 = w["m"], w["n0"]
    N = batch or w["batch"]
    if "band" in w:
        from pycllp_b200.problems import staircase_equality_arrays
        A = staircase_equality_arrays(m, n0, w["band"], w["per_col"], 1, seed=0)[0]
        rng = np.random.RandomState(1000 + seed_offset)
        b = 0.5 + rng.rand(N, m)
        c = np.concatenate([0.5 + rng.rand(N, n0), np.zeros((N, m))], axis=1)
        return A, b, c
    np.random.seed(0)
    A0 = rand(m, n0, density=w["density"]).toarray()
    rng = np.random.RandomState(1000 + seed_offset)
    b = 0.5 + rng.rand(N, m)
    c = np.concatenate([0.5 + rng.rand(N, n0), np.zeros((N, m))], axis=1)
    A = np.concatenate([A0, np.eye(m)], axis=1)
    return A, b, c


def flops_per_iteration(m, n, A=None, tile_info=None):
    """M = A D A' lower triangle (m^2 n) + LDL' (m^3/3) + one forward/back
    solve (2 m^2) + the four mat-vecs with A (8 m n).  Sparse A: M costs 2 sum_k c_k(c_k+1)/2
    (c_k = non-zeros of column k), the mat-vecs 8 nnz(A); the factor of A A' is ~dense at the
    config-4 shape and is counted as m^3/3."""
    if A is not None:
        ck = np.asarray((A != 0).sum(axis=0)).ravel().astype(np.float64)
        nnz = float(ck.sum())
        if tile_info and tile_info["factor"] == "tiles":
            # tile-sparse factor: 2 x 8^3 per tile-pair update, 2 x 8 per entry of L in the panel
            # elimination and 4 per entry in the two triangular solves
            fac = 1024.0 * tile_info["update_pairs"] + 20.0 * tile_info["factor_doubles"]
            return float((ck * (ck + 1)).sum()) + fac + 8.0 * nnz
        return float((ck * (ck + 1)).sum()) + m ** 3 / 3.0 + 2.0 * m * m + 8.0 * nnz
    return m * m * n + m ** 3 / 3.0 + 2.0 * m * m + 8.0 * m * n


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([s.strip() for s in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [v for v in sm if v > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(A, b, c, nproblems, threads, sparse=False, force_port=False):
    """Reference CPU path = the reference's own kernels compiled as C (oracle/_ref), else the
    oracle port; process all `nproblems` with `threads` host threads, return (solves/s, kind)."""
    from oracle.bindings import Reference, Oracle
    if Reference.available() and not force_port:
        impl, kind = Reference(), "reference"
    else:
        impl, kind = Oracle(), "port"
    if hasattr(A, "toarray"):
        A = A.toarray()
    t0 = time.perf_counter()
    solve = impl.solve_sparse if sparse else impl.solve_dense
    r = solve(A, b[:nproblems], c[:nproblems], nthreads=threads)
    dt = time.perf_counter() - t0
    return nproblems / dt, kind, dt, r


def measured_traffic(workload):
    """DRAM bytes per LP from the committed ncu capture of this workload, else None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        d = d.get(workload, d if workload == "cfg3" and "dram_bytes_per_lp" in d else None)
        return float(d["dram_bytes_per_lp"]) if d else None
    except Exception:
        return None


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    A, b, c = make_problem(args.workload, 0, batch=max(os.cpu_count() or 1, 1) * 2)
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    sparse = bool(wl.get("sparse"))
    sample = cores * (1 if wl["m"] >= 200 else 2)
    if wl["m"] >= 1000:
        sample = 1        # minutes per LP on one core at config 4: one LP per step is the bounded sample
    for _ in range(min(args.warmup, 1)):
        if wl["m"] < 1000:
            cpu_reference_rate(A, b, c, min(sample, 2), cores, sparse)
    times = []
    for _ in range(args.steps):
        rate, kind, dt, _ = cpu_reference_rate(A, b, c, sample, cores, sparse)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    w = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "m": w["m"], "n": w["m"] + w["n0"],
                   "sample_problems_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": kind,
                         "sample": "%d LPs of the %s workload per step, %d host threads"
                                   % (sample, args.workload, cores)},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class DeviceRun(object):
    """One workload set up on this rank's GPU: engine, device buffers, the timed step."""

    def __init__(self, name, N, rank, world, local, torch, dist):
        from pycllp_b200._cabi import Engine
        self.torch, self.dist, self.world, self.N, self.name = torch, dist, world, N, name
        w = WORKLOADS[name]
        self.m, self.n = w["m"], w["m"] + w["n0"]
        self.sparse = bool(w.get("sparse"))
        self.A, self.b, self.c = make_problem(name, rank, batch=N)
        self.dev = torch.device("cuda", local)
        self.eng = Engine(local)
        t0 = time.perf_counter()
        self.tile_info = None
        if self.sparse:
            from scipy.sparse import csr_matrix
            self.eng.setup_sparse(csr_matrix(self.A), N)
            self.tile_info = self.eng.sparse_info()
        else:
            self.eng.setup_dense(self.A, N)
        self.setup_s = time.perf_counter() - t0
        self.info = self.eng.info()
        f64 = torch.float64
        self.d_b = torch.from_numpy(self.b).to(self.dev)
        self.d_c = torch.from_numpy(self.c).to(self.dev)
        self.width = self.eng.record_width
        self.rec = torch.zeros((N, self.width), dtype=f64, device=self.dev)
        self.gathered = torch.empty((world * N, self.width), dtype=f64, device=self.dev) if world > 1 else None
        self.stream = torch.cuda.current_stream(self.dev)

    def step(self):
        self.eng.solve_device_packed(self.N, self.d_b.data_ptr(), self.d_c.data_ptr(), self.rec.data_ptr(),
                                     self.stream.cuda_stream)
        if self.world > 1:   # the one exchange of the path: every rank's records to every rank
            self.dist.all_gather_into_tensor(self.gathered, self.rec)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, steps, warmup, flush):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.barrier()
        launches0 = self.eng.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for k in range(steps):
            flush.fill_(float(k))            # evict L2 between timed steps (outside the event pair)
            ev[k][0].record(self.stream)
            self.step()
            ev[k][1].record(self.stream)
        self.barrier()
        step_ms = [a.elapsed_time(b_) for a, b_ in ev]
        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(total_ms, op=self.dist.ReduceOp.MAX)
        tail = self.rec[:, self.width - 1:].contiguous().view(torch.int32)     # (N, 2): status, iterations
        agg = torch.stack([tail[:, 1].to(torch.float64).sum(), (tail[:, 0] == 0).sum().to(torch.float64)])
        if self.world > 1:
            self.dist.all_reduce(agg)
        return dict(total_s=float(total_ms.item()) / 1e3, step_ms=step_ms,
                    launches=self.eng.launch_count - launches0,
                    total_iters=float(agg[0].item()), total_ok=int(agg[1].item()))

    def roofline(self, timing, peak, peak_src):
        """achieved = algorithmic flops of one launch / its CUDA-event duration (this rank's launches:
        sum of Newton steps over the world / world ranks)."""
        flops = flops_per_iteration(self.m, self.n, self.A if self.sparse else None, self.tile_info) * \
            (timing["total_iters"] / self.world)
        ms_kernel = float(np.mean(timing["step_ms"]))
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        per_lp = measured_traffic(self.name)
        alg = self.N * ((self.m + self.n) * 8 + (2 * self.n + self.m) * 8 + 8)
        return {"bound": "tensor", "kernel": "ipm_solve_kernel", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": per_lp * self.N if per_lp else None,
                "traffic_note": "DRAM bytes per launch (ncu dram__bytes_read+write per LP x LPs, "
                                "profiles/ncu_traffic.json); algorithmic bytes per launch = %d" % alg,
                "peak_source": peak_src,
                "flops_model": "sum_p (m^2 n + m^3/3 + 2 m^2 + 8 m n) * newton_steps_p (SURVEY.md 8(d)); "
                               "sparse: 2 sum_k c_k(c_k+1)/2 + m^3/3 + 2 m^2 + 8 nnz"}


def fp64_peaks(eng, torch, dev):
    """The roofline denominator, measured NOW on this box: the library's DMMA loop (FP64 tensor
    cores) and, as a cross-check, float64 torch.matmul 8192^3 (cuBLAS)."""
    dmma = eng.fp64_probe()
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    gemm = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    return dmma, gemm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the cfg5 strong-scaling record (65536 LPs over all ranks)")
    ap.add_argument("--strong-total", type=int, default=STRONG_TOTAL)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pycllp_b200.lp import EqualityLP
    from pycllp_b200.solvers import solver_registry

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    N = args.batch or w["batch"]
    run = DeviceRun(args.workload, N, rank, world, local, torch, dist)
    m, n = run.m, run.n
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > L2

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    timing = run.timed(args.steps, args.warmup, flush)

    # ---- e2e: the plugin call a pycllp user makes -- lp.init(solver); lp.solve(solver) ----
    from scipy.sparse import csr_matrix
    b_all = np.concatenate([make_problem(args.workload, r, batch=N)[1] for r in range(world)]) if world > 1 else run.b
    c_all = np.concatenate([make_problem(args.workload, r, batch=N)[2] for r in range(world)]) if world > 1 else run.c
    # the container path a pycllp user takes: StandardLP(A0, b, c0) -> to_equality_form()
    # -> A.tocsr() hand-off to the solver; timed on rank 0, outside the solve timing
    containers = None
    if rank == 0:
        from pycllp_b200.lp import StandardLP, SparseMatrix
        from scipy.sparse import coo_matrix
        n0 = w["n0"]
        t0 = time.perf_counter()
        A0 = csr_matrix(run.A)[:, :n0].tocoo()
        std = StandardLP(SparseMatrix(matrix=A0), run.b, run.c[:, :n0], np.zeros(N))
        t1 = time.perf_counter()
        eq = std.to_equality_form()
        t2 = time.perf_counter()
        Acsr = eq.A.tocsr()
        t3 = time.perf_counter()
        assert eq.ncols == n and eq.nrows == m and Acsr.nnz == csr_matrix(run.A).nnz
        containers = {"build_StandardLP_s": t1 - t0, "to_equality_form_s": t2 - t1, "tocsr_s": t3 - t2,
                      "problems": N, "nnz": int(Acsr.nnz)}
        del std, eq, Acsr, A0
    lp = EqualityLP(csr_matrix(run.A), b_all, c_all, np.zeros(world * N))
    solver = solver_registry["cl_sparse_primal_normal" if run.sparse else "cl_dense_primal_normal"](
        local, group=True if world > 1 else None)
    lp.init(solver)
    lp.solve(solver)                                   # warm-up (page-locks lp.b / lp.c in place)
    run.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lp.solve(solver)
    run.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    plugin_status0 = float((solver.status == 0).mean())

    # ---- the bare C-ABI host call with caller-pinned buffers, for comparison ----
    f64 = torch.float64
    h = {k: torch.empty(s, dtype=f64).pin_memory() for k, s in
         (("b", (N, m)), ("c", (N, n)), ("x", (N, n)), ("y", (N, m)), ("z", (N, n)))}
    h["st"] = torch.empty(N, dtype=torch.int32).pin_memory()
    h["it"] = torch.empty(N, dtype=torch.int32).pin_memory()
    h["b"].copy_(torch.from_numpy(run.b)); h["c"].copy_(torch.from_numpy(run.c))

    def step_host():
        run.eng.solve_host_into(N, h["b"].data_ptr(), h["c"].data_ptr(), h["x"].data_ptr(),
                                h["y"].data_ptr(), h["z"].data_ptr(), h["st"].data_ptr(), h["it"].data_ptr())

    step_host()
    run.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    run.barrier()
    cabi_s = torch.tensor([time.perf_counter() - t0], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(cabi_s, op=dist.ReduceOp.MAX)
    cabi_s = float(cabi_s.item())
    clocks = sampler.stop() if sampler else None

    dmma_peak, gemm_peak = fp64_peaks(run.eng, torch, dev)
    peak_src = ("FP64 DMMA loop of libpycllp_b200.so measured in this run: %.2f TFLOP/s; cross-check float64 "
                "torch.matmul 8192^3 (cuBLAS) in this run: %.2f TFLOP/s" % (dmma_peak, gemm_peak))

    # ---- secondary: cfg5, fixed total split over the ranks (strong scaling), gather included ----
    secondary = None
    if not args.no_secondary:
        del lp, solver, h
        run.eng.close()
        torch.cuda.empty_cache()
        total = args.strong_total
        Ns = total // world
        srun = DeviceRun("cfg5", Ns, rank, world, local, torch, dist)
        srun.eng.solve_device_packed(min(Ns, 296), srun.d_b.data_ptr(), srun.d_c.data_ptr(), srun.rec.data_ptr(),
                                     srun.stream.cuda_stream)          # warm-up: two waves
        st = srun.timed(1, 0, flush)
        hb = torch.from_numpy(srun.b).pin_memory(); hc = torch.from_numpy(srun.c).pin_memory()
        hx = torch.empty((Ns, srun.n), dtype=f64).pin_memory(); hy = torch.empty((Ns, srun.m), dtype=f64).pin_memory()
        hz = torch.empty((Ns, srun.n), dtype=f64).pin_memory()
        hs = torch.empty(Ns, dtype=torch.int32).pin_memory(); hi = torch.empty(Ns, dtype=torch.int32).pin_memory()
        srun.barrier()
        t0 = time.perf_counter()
        srun.eng.solve_host_into(Ns, hb.data_ptr(), hc.data_ptr(), hx.data_ptr(), hy.data_ptr(), hz.data_ptr(),
                                 hs.data_ptr(), hi.data_ptr())
        srun.barrier()
        s_e2e = torch.tensor([time.perf_counter() - t0], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(s_e2e, op=dist.ReduceOp.MAX)
        secondary = {
            "metric": METRIC, "workload": "cfg5", "scaling": "strong", "m": srun.m, "n": srun.n,
            "total_problems": Ns * world, "problems_per_gpu": Ns, "steps": 1, "warmup": "296 LPs",
            "value": Ns * world / st["total_s"], "unit": "solves/s", "ms_per_step": 1e3 * st["total_s"],
            "e2e": {"value": Ns * world / float(s_e2e.item()), "unit": "solves/s",
                    "through": "pycllp_b200_solve_host, pinned buffers, per-rank slice",
                    "h2d_bytes_per_step": Ns * (srun.m + srun.n) * 8,
                    "d2h_bytes_per_step": Ns * (2 * srun.n + srun.m) * 8 + Ns * 8},
            "gather_bytes_per_rank": (world * Ns * srun.width * 8) if world > 1 else 0,
            "status0_fraction": st["total_ok"] / float(world * Ns),
            "mean_newton_steps": st["total_iters"] / float(world * Ns),
            "roofline": srun.roofline(st, dmma_peak, peak_src),
        }

    if rank == 0:
        value = world * N * args.steps / timing["total_s"]
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * timing["total_s"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "m": m, "n": n, "problems_per_gpu": N,
                       "density": w["density"], "l2": "flushed (256 MiB write) between timed steps",
                       "grid": run.info["grid"], "block": run.info["block"], "smem_bytes": run.info["smem_bytes"],
                       "factor_in_smem": run.info["factor_in_smem"],
                       "status0_fraction": timing["total_ok"] / float(world * N),
                       "mean_newton_steps": timing["total_iters"] / float(world * N),
                       "exchange": ("one all_gather_into_tensor of %d B records per LP inside the timed step"
                                    % (run.width * 8)) if world > 1 else "none (1 rank)",
                       "gather_bytes_per_rank": (world * N * run.width * 8) if world > 1 else 0,
                       "setup_s": run.setup_s, "lp_containers": containers,
                       **({"sparse_factor": run.tile_info} if run.tile_info else {})},
            "clocks": clocks,
            "e2e": {"value": world * N * args.steps / e2e_s, "unit": "solves/s",
                    "through": "lp.solve(solver_registry[...]) with numpy lp.b / lp.c (plugin API)",
                    "h2d_bytes_per_step": N * (m + n) * 8,
                    "d2h_bytes_per_step": (world if world > 1 else 1) * N * ((2 * n + m) * 8 + 8),
                    "status0_fraction": plugin_status0},
            "e2e_cabi_pinned": {"value": world * N * args.steps / cabi_s, "unit": "solves/s",
                                "through": "pycllp_b200_solve_host, caller-pinned buffers"},
            "gpu_launches": int(timing["launches"]),
            "roofline": run.roofline(timing, dmma_peak, peak_src),
        }
        if secondary is not None:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            big = m >= 1000
            sample = min(cores, 8) if big else cores * (2 if m >= 200 else 8)
            # config 4: the reference's kernels need ~7 minutes per LP on a core (the per-entry row
            # search of ldl.cl:435-445); its restatement, bit-identical, ~80 s: that one is timed
            rate, kind, dt, _ = cpu_reference_rate(run.A, run.b, run.c, min(sample, N), cores, run.sparse,
                                                   force_port=big)
            line["cpu_baseline"] = {
                "value": rate, "unit": "solves/s", "cores": cores, "kind": kind,
                "sample": "first %d LPs of the workload, %d host threads, %.1f s wall"
                          % (min(sample, N), cores, dt)}
            if args.workload == "cfg1":
                # BASELINE.json configs[0] names "reference CPU solver": the Python/Cython
                # DensePrimalNormalSolver (Oracle A), single-threaded like the reference's loop
                from oracle import oracle_a
                if oracle_a.available():
                    k = 8
                    t0 = time.perf_counter()
                    ra = oracle_a.solve(run.A, run.b[:k], run.c[:k])
                    dt = time.perf_counter() - t0
                    line["cpu_baseline_python"] = {
                        "value": k / dt, "unit": "solves/s", "cores": 1, "kind": "reference",
                        "what": "solvers/normal_eqns.py (restated) + pycllp/_ldl.pyx compiled as is (oracle/_ref)",
                        "sample": "first %d LPs, %.1f s wall; statuses %s (noise-driven, SURVEY fact 1)"
                                  % (k, dt, np.bincount(ra["status"], minlength=6).tolist())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
