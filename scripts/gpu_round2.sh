# Round-2 evidence run (under gpurun): GPU tests, default bench, reference arm, launch list.
set -x
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_gputest.log
tail -4 gpurun_out/r02_gputest.log
python bench.py > gpurun_out/r02_bench_c2_100k.json 2> gpurun_out/r02_bench_c2_100k.err
tail -c 600 gpurun_out/r02_bench_c2_100k.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
tail -c 300 gpurun_out/r02_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_bench_c2_100k.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench.log 2>&1
tail -c 300 gpurun_out/r02_launches_bench.log
