#!/bin/bash
# Round evidence batch (run under gpurun):  bash tools/evidence.sh TAG
# tests, bench records of every workload, ncu launch list + one --set full capture of the top kernel.
TAG=${1:-r02}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
python bench.py > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc $?"
python bench.py --workload cfg4 --steps 2 --warmup 1 --no-secondary > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc $?"
python bench.py --workload cfg5 --steps 2 --warmup 1 --no-secondary > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc $?"
python bench.py --workload cfg1 --steps 20 --warmup 3 --no-secondary > $O/bench_cfg1_64.json 2> $O/bench_cfg1_64.err; echo "cfg1 rc $?"
python bench.py --workload cfg1 --batch 4096 --steps 20 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_cfg1_4096.json 2> $O/bench_cfg1_4096.err; echo "cfg1b rc $?"
python bench.py --workload cfg1 --batch 65536 --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_cfg1_65536.json 2> $O/bench_cfg1_65536.err; echo "cfg1c rc $?"
python bench.py --workload stair --steps 2 --warmup 1 --no-secondary > $O/bench_stair.json 2> $O/bench_stair.err; echo "stair rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_cfg3_reference_arm.json 2> $O/ref.err; echo "ref rc $?"
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu list rc $?"
# one full capture of the solve kernel (1184 LPs = 8 waves) + DRAM bytes / tensor pipe of the bench-sized launches
ncu --set full --clock-control none --import-source on -k regex:ipm_solve -c 1 -o $O/prof_cfg3 -f \
  python tools/prof_run.py cfg3 1184 > $O/ncu_full.log 2>&1; echo "ncu full rc $?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct
ncu --metrics $M --clock-control none -k regex:ipm_solve -c 1 --csv --log-file $O/traffic_cfg3_4096.csv python tools/prof_run.py cfg3 4096 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:ipm_solve -c 1 --csv --log-file $O/traffic_cfg4_148.csv python tools/prof_run.py cfg4 148 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:ipm_solve -c 1 --csv --log-file $O/traffic_cfg5_296.csv python tools/prof_run.py cfg5 296 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:ipm_small -c 1 --csv --log-file $O/traffic_cfg1_2368.csv python tools/tiny_prof.py 2368 > /dev/null 2>&1
python tests/tools/gpu_check.py prof3 > $O/phase_cfg3.txt 2>&1
python tests/tools/gpu_check.py prof5 > $O/phase_cfg5.txt 2>&1
python tests/tools/gpu_check.py prof4 > $O/phase_cfg4.txt 2>&1
python tests/tools/gpu_check.py proft > $O/phase_stair.txt 2>&1
PB200_SMALL=3 python tests/tools/gpu_check.py prof1 prof1s > $O/phase_cfg1.txt 2>&1
# (read the report HERE, not on the box:  python tools/ncu_summary.py $O/prof_cfg3.ncu-rep pycllp_b200/libpycllp_b200.so \
#    "<note>" ipm_solve_kernelILb1ELb1ELb1ELi1 > profiles/ncu_summary_$TAG.txt ;  python tools/sass_hist.py > profiles/sass_r02.txt)
ls -la $O
