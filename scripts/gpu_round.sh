# one GPU session: parity tests, smoke, a bench line (run under gpurun from the repo root)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1200 python bench.py --workload ${WORKLOAD:-c2_20k} --steps ${STEPS:-2} --warmup ${WARMUP:-3} 2>&1 | tail -3
