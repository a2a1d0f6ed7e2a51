set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-pb2_xi_auto} -c 1 -o gpurun_out/${OUT:-prof} python scripts/perf_probe.py --n 1200 --side 7.6 --brute 0 --reps 1 2>&1 | tail -3
