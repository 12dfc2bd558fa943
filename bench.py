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
                    sparse solver path (CSR mat-vecs, pattern-based M, dense packed factor)

Timed quantities
    value   solves/s with b, c already resident in HBM (device-pointer C-ABI entry), CUDA
            events on torch's current stream, one event pair per step, L2 flushed between
            steps, max over ranks.
    e2e     the same through pycllp_b200_solve_host with pinned HOST buffers: H2D of b, c
            and D2H of x, y, z, status, iterations inside the timed region.
One process per GPU (torchrun); ranks solve disjoint slices, no collective in the solve;
the final status/iteration gather (NCCL all_gather) is inside the timed step when N > 1.
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
}
METRIC = "batched_lp_solves_per_sec"


def make_problem(name, rank, batch=None):
    """A (shared, seed 0) and this rank's slice of b, c (seeded by rank: disjoint problems)."""
    from scipy.sparse import rand
    w = WORKLOADS[name]
    m, n0 = w["m"], w["n0"]
    N = batch or w["batch"]
    np.random.seed(0)
    A0 = rand(m, n0, density=w["density"]).toarray()
    rng = np.random.RandomState(1000 + rank)
    b = 0.5 + rng.rand(N, m)
    c = np.concatenate([0.5 + rng.rand(N, n0), np.zeros((N, m))], axis=1)
    A = np.concatenate([A0, np.eye(m)], axis=1)
    return A, b, c


def flops_per_iteration(m, n, A=None):
    """M = A D A' lower triangle (m^2 n) + LDL' (m^3/3) + one forward/back
    solve (2 m^2) + the four mat-vecs with A (8 m n).  Sparse A: M costs 2 sum_k c_k(c_k+1)/2
    (c_k = non-zeros of column k), the mat-vecs 8 nnz(A); the factor of A A' is ~dense at the
    config-4 shape and is counted as m^3/3."""
    if A is not None:
        ck = (A != 0).sum(axis=0).astype(np.float64)
        return float((ck * (ck + 1)).sum()) + m ** 3 / 3.0 + 2.0 * m * m + 8.0 * float((A != 0).sum())
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


def cpu_reference_rate(A, b, c, nproblems, threads, sparse=False):
    """Reference CPU path = the reference's own kernels compiled as C (oracle/_ref), else the
    oracle port; process all `nproblems` with `threads` host threads, return (solves/s, kind)."""
    from oracle.bindings import Reference, Oracle
    if Reference.available():
        impl, kind = Reference(), "reference"
    else:
        impl, kind = Oracle(), "port"
    t0 = time.perf_counter()
    solve = impl.solve_sparse if sparse else impl.solve_dense
    r = solve(A, b[:nproblems], c[:nproblems], nthreads=threads)
    dt = time.perf_counter() - t0
    return nproblems / dt, kind, dt, r


def fp64_peak():
    """FP64 tensor (DMMA) peak measured on this pool (tools/fp64_probe.cu); MEASURED_PEAKS.json
    holds only HBM and bf16."""
    path = os.path.join(ROOT, "profiles", "fp64_peak.json")
    try:
        d = json.load(open(path))
        return float(d["dmma_tflops"]), "measured (tools/fp64_probe.cu, profiles/fp64_peak.json)"
    except Exception:
        return 37.0, "fallback (nominal B200 FP64 ~37 TFLOP/s)"


def measured_traffic(workload, nproblems):
    """DRAM bytes per launch from the committed ncu capture (config 3 only), else None."""
    if workload != "cfg3":
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(d["dram_bytes_per_lp"]) * nproblems
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pycllp_b200._cabi import Engine

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    m, n = w["m"], w["m"] + w["n0"]
    N = args.batch or w["batch"]
    A, b, c = make_problem(args.workload, rank, batch=N)

    eng = Engine(local)
    if w.get("sparse"):
        from scipy.sparse import csr_matrix
        eng.setup_sparse(csr_matrix(A), N)
    else:
        eng.setup_dense(A, N)
    info = eng.info()

    f64 = torch.float64
    d_b = torch.from_numpy(b).to(dev)
    d_c = torch.from_numpy(c).to(dev)
    d_x = torch.empty(N, n, dtype=f64, device=dev)
    d_y = torch.empty(N, m, dtype=f64, device=dev)
    d_z = torch.empty(N, n, dtype=f64, device=dev)
    d_st = torch.empty(N, dtype=torch.int32, device=dev)
    d_it = torch.empty(N, dtype=torch.int32, device=dev)
    gathered = torch.empty(world * N, 2, dtype=torch.int32, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > L2
    stream = torch.cuda.current_stream(dev)

    def step_device():
        eng.solve_device(N, d_b.data_ptr(), d_c.data_ptr(), d_x.data_ptr(), d_y.data_ptr(),
                         d_z.data_ptr(), d_st.data_ptr(), d_it.data_ptr(), stream.cuda_stream)
        if world > 1:   # the one exchange of the path: collect status / iteration counts
            dist.all_gather_into_tensor(gathered, torch.stack([d_st, d_it], dim=1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_device()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))            # evict L2 between timed steps (outside the event pair)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    launches = eng.launch_count - launches0
    step_ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    iters_local = d_it.to(torch.float64).sum()
    status_ok = (d_st == 0).sum().to(torch.float64)
    agg = torch.stack([iters_local, status_ok])
    if world > 1:
        dist.all_reduce(agg)
    total_iters, total_ok = float(agg[0].item()), int(agg[1].item())

    # ---- e2e: host buffers through the C ABI (pinned), H2D + solve + D2H per step ----
    h = {k: torch.empty(s, dtype=f64).pin_memory() for k, s in
         (("b", (N, m)), ("c", (N, n)), ("x", (N, n)), ("y", (N, m)), ("z", (N, n)))}
    h["st"] = torch.empty(N, dtype=torch.int32).pin_memory()
    h["it"] = torch.empty(N, dtype=torch.int32).pin_memory()
    h["b"].copy_(torch.from_numpy(b)); h["c"].copy_(torch.from_numpy(c))

    def step_host():
        eng.solve_host_into(N, h["b"].data_ptr(), h["c"].data_ptr(), h["x"].data_ptr(),
                            h["y"].data_ptr(), h["z"].data_ptr(), h["st"].data_ptr(), h["it"].data_ptr())

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    launches += 0  # e2e launches are outside the device-timed region
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        value = world * N * args.steps / total_s
        peak, peak_src = fp64_peak()
        flops_per_step = flops_per_iteration(m, n, A if w.get("sparse") else None) * (total_iters / world)
        ms_kernel = float(np.mean(step_ms))
        achieved = flops_per_step / (ms_kernel * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "m": m, "n": n, "problems_per_gpu": N,
                       "density": w["density"], "l2": "flushed (256 MiB write) between timed steps",
                       "grid": info["grid"], "block": info["block"], "smem_bytes": info["smem_bytes"],
                       "factor_in_smem": info["factor_in_smem"],
                       "status0_fraction": total_ok / float(world * N),
                       "mean_newton_steps": total_iters / float(world * N)},
            "clocks": clocks,
            "e2e": {"value": world * N * args.steps / e2e_s, "unit": "solves/s",
                    "h2d_bytes_per_step": N * (m + n) * 8,
                    "d2h_bytes_per_step": N * (2 * n + m) * 8 + N * 8},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "ipm_solve_kernel", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, N),
                         "traffic_note": "DRAM bytes per launch (ncu dram__bytes_read+write per LP x LPs); "
                                         "algorithmic bytes per launch = %d" % (N * ((m + n) * 8 + (2 * n + m) * 8 + 8)),
                         "peak_source": "FP64 DMMA " + peak_src,
                         "flops_model": "sum_p (m^2 n + m^3/3 + 2 m^2 + 8 m n) * newton_steps_p"},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample = cores * (2 if m >= 200 else 8)
            if m >= 1000:
                sample = 0      # config 4: minutes per LP on a core; see `--impl reference`
            if sample:
                rate, kind, dt, _ = cpu_reference_rate(A, b, c, min(sample, N), cores, bool(w.get("sparse")))
            else:
                rate, kind, dt = None, "reference", 0.0
            line["cpu_baseline"] = {
                "value": rate, "unit": "solves/s", "cores": cores, "kind": kind,
                "sample": "first %d LPs of the workload, %d host threads, %.1f s wall"
                          % (min(sample, N), cores, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
