# DRAM traffic of ONE launch of the dominant kernel on the bench workload (run under gpurun).
# Writes gpurun_out/r02_traffic_raw.csv and gpurun_out/r02_traffic.json (copy into profiles/).
set -x
WORKLOAD=${WORKLOAD:-c2_100k}
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:pb2_xi_auto_diag -c 1 --csv --log-file gpurun_out/r02_traffic_raw.csv \
  python bench.py --workload $WORKLOAD --steps 1 --warmup 0 --no-cpu-baseline --no-dmat \
  > gpurun_out/traffic_bench.log 2>&1
python scripts/traffic_json.py gpurun_out/r02_traffic_raw.csv $WORKLOAD > gpurun_out/r02_traffic.json
cat gpurun_out/r02_traffic.json
