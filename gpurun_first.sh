set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_cf_gpu.py -x -q 2>&1 | tail -30
timeout 600 python scripts/perf_probe.py --n 3000 --side 12 2>&1 | tail -20
