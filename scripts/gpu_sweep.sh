if [ -z "$NOTEST" ]; then timeout 600 python -m pytest tests/test_cf_gpu.py tests/test_golden_gpu.py tests/test_edge_cases_gpu.py -x -q 2>&1 | tail -5; fi
echo "== default"; timeout 300 python scripts/perf_probe.py --n 3000 --side 12 --brute 0 --reps 2 2>&1 | grep "rep 1"
for f in $(ls gpurun_variants/*.so 2>/dev/null); do echo "== $f"; PICCA_B200_LIB=$PWD/$f timeout 300 python scripts/perf_probe.py --n 3000 --side 12 --brute 0 --reps 2 2>&1 | grep "rep 1\|Error\|error" | tail -3; done
