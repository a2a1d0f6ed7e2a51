set -x
timeout 900 python -m pytest tests/test_cf_gpu.py tests/test_golden_gpu.py tests/test_edge_cases_gpu.py -x -q 2>&1 | tail -15
timeout 600 python scripts/perf_probe.py --n 3000 --side 12 --brute ${BRUTE:-1} 2>&1 | grep -v computing | tail -14
