"""Text summary of an ncu report for profiles/: selected raw metrics, warp-stall totals, per-line stalls.

usage: ncu_summary.py <report.ncu-rep> <lib.so> "<header comment>" > profiles/ncu_summary_XXX.txt
"""
import csv, io, os, subprocess, sys, collections

rep, lib, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
KEEP = """dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum launch__block_size launch__grid_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic lts__t_sector_hit_rate.pct sm__cycles_elapsed.avg sm__icc_request_hit_rate.pct
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum""".split()
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {n: i for i, n in enumerate(hdr)}
print("#", note)
print("#", vals[col["Kernel Name"]])
for k in KEEP:
    if k in col:
        print("%s [%s] = %s" % (k, units[col[k]], vals[col[k]]))
stall = collections.Counter()
for n, i in col.items():
    if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
        try:
            stall[n[len("smsp__pcsamp_warps_issue_stalled_"):]] += float(vals[i].replace(",", ""))
        except ValueError:
            pass
tot = sum(stall.values()) or 1.0
print("\n# warp stall samples by reason")
for k, v in stall.most_common(10):
    print("  %-28s %10d %5.1f%%" % (k, v, 100 * v / tot))
print("\n# per-source-line stall samples (tools/ncu_lines.py on the same report)")
sys.stdout.flush()
here = os.path.dirname(os.path.abspath(__file__))
subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, lib, sys.argv[4] if len(sys.argv) > 4 else "ipm_solve_kernel", "30"])
