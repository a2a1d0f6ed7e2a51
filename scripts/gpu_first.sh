set -x
timeout 600 python -m pytest tests/test_cf_gpu.py -x -q 2>&1 | tail -5
timeout 600 python scripts/perf_probe.py --n 3000 --side 12 2>&1 | grep -v computing | tail -12
