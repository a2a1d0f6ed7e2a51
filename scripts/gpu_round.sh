# one GPU session: parity tests, smoke, a bench line (run under gpurun from the repo root)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -6
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python bench.py --workload ${WORKLOAD:-c2_100k} --steps ${STEPS:-2} --warmup ${WARMUP:-3} > gpurun_out/bench_${TAG:-r01}.json 2> gpurun_out/bench_${TAG:-r01}.err
tail -c 2500 gpurun_out/bench_${TAG:-r01}.json; tail -3 gpurun_out/bench_${TAG:-r01}.err
