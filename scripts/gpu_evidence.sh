set -x
timeout 600 python -m pytest tests/test_edge_cases_gpu.py -x -q 2>&1 | tail -4
timeout 900 python scripts/perf_dmat_xcf.py > gpurun_out/dmat_xcf_probe.log 2>&1; grep -E "dmat rep|xcf rep|xdmat rep|Error|error" gpurun_out/dmat_xcf_probe.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --workload c2_100k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
tail -c 400 gpurun_out/bench_under_ncu.json; wc -l gpurun_out/launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pb2_xi_auto_diag -c 1 -o gpurun_out/prof_diag_v3 python scripts/perf_probe.py --n 1200 --side 7.6 --brute 0 --reps 1 2>&1 | tail -3
