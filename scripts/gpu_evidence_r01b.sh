# ncu evidence for the current kernels (run under gpurun from the repo root):
#  1. launch list of the bench command (per-launch gpu time, serialised, cold cache)
#  2. one full capture of the dominant kernel on the 1200-forest probe
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r01_launches_bench_c2_100k_v7.csv \
  python bench.py --workload c2_100k --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/bench_under_ncu_v7.json 2> gpurun_out/bench_under_ncu_v7.err
wc -l gpurun_out/r01_launches_bench_c2_100k_v7.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pb2_xi_auto_diag -c 1 \
  -o gpurun_out/r01_xi_auto_diag_v7_full python scripts/perf_probe.py --n 1200 --side 7.6 --brute 0 --reps 1 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
