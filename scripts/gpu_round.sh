# one GPU session: parity tests, smoke, a bench line (run under gpurun from the repo root)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -15
if [ -z "$SKIP_BENCH" ]; then
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1200 python bench.py --workload ${WORKLOAD:-c2_20k} --steps ${STEPS:-2} --warmup ${WARMUP:-3} 2>&1 | tail -3
fi
