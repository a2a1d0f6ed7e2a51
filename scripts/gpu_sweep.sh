timeout 300 python -m pytest tests/test_cf_gpu.py -x -q 2>&1 | tail -2
echo "== default"; timeout 300 python scripts/perf_probe.py --n 3000 --side 12 --brute 0 --reps 2 2>&1 | grep "rep 1"
for f in gpurun_variants/*.so; do echo "== $f"; PICCA_B200_LIB=$PWD/$f timeout 300 python scripts/perf_probe.py --n 3000 --side 12 --brute 0 --reps 2 2>&1 | grep "rep 1"; done
