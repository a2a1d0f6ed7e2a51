# Executed-FP64 and DRAM counters of the distortion-matrix and cross-correlation kernels on the
# bench workload (run under gpurun; ncu replays every profiled launch a few times).
# Writes gpurun_out/r02_dmat_counters.json, r02_xcf_counters.json (copy into profiles/).
set -x
WORKLOAD=${WORKLOAD:-c2_100k}
SEG=${SEG:-8}
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
timeout 1500 ncu --metrics $M --clock-control none -k regex:pb2_dmat_auto -c $SEG --csv \
  --log-file gpurun_out/r02_dmat_counters_raw.csv \
  python bench.py --workload $WORKLOAD --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e \
  --no-xcf --dmat-steps 1 --dmat-segments $SEG > gpurun_out/counters_dmat_bench.log 2>&1
python scripts/counters_json.py gpurun_out/r02_dmat_counters_raw.csv $WORKLOAD pb2_dmat_auto $SEG \
  > gpurun_out/r02_dmat_counters.json
cat gpurun_out/r02_dmat_counters.json
timeout 900 ncu --metrics $M --clock-control none -k regex:pb2_xi_cross_chunk -c 1 --csv \
  --log-file gpurun_out/r02_xcf_counters_raw.csv \
  python bench.py --workload $WORKLOAD --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e \
  --no-dmat --xcf-steps 1 > gpurun_out/counters_xcf_bench.log 2>&1
python scripts/counters_json.py gpurun_out/r02_xcf_counters_raw.csv $WORKLOAD pb2_xi_cross_chunk 1 \
  > gpurun_out/r02_xcf_counters.json
cat gpurun_out/r02_xcf_counters.json
