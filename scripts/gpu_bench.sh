# full-size bench line on 1 GPU (+ ncu launch list), run under gpurun from the repo root
set -x
timeout 1500 python bench.py --workload ${WORKLOAD:-c2_100k} --steps ${STEPS:-2} --warmup ${WARMUP:-3} > gpurun_out/bench_${TAG:-r01}.json 2> gpurun_out/bench_${TAG:-r01}.err
tail -c 3000 gpurun_out/bench_${TAG:-r01}.json; tail -3 gpurun_out/bench_${TAG:-r01}.err
