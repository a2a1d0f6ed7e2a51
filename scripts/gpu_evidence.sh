set -x
timeout 900 python scripts/perf_dmat_xcf.py > gpurun_out/dmat_xcf_probe.log 2>&1; grep -E "dmat rep|xcf rep|xdmat rep|Error|error" gpurun_out/dmat_xcf_probe.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --workload c2_100k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
tail -c 400 gpurun_out/bench_under_ncu.json; wc -l gpurun_out/launches_bench.csv
