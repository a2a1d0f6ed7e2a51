"""Turn an `ncu --csv --metrics ...` log of the launches of ONE bench step into the small JSON
bench.py reads: DRAM bytes and executed FP64 thread instructions (DADD + DMUL + DFMA, one op
each -- the unit of the FP64 issue peak), summed over the launches of the matching kernels.
usage: counters_json.py raw.csv workload kernel_regex launches_per_step > out.json"""
import csv
import json
import re
import sys

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
head = rows[0]
k_id, k_name, k_metric, k_unit, k_val = (head.index(x) for x in
                                         ("ID", "Kernel Name", "Metric Name", "Metric Unit",
                                          "Metric Value"))
pat = re.compile(sys.argv[3])
per_step = int(sys.argv[4])
launches = {}
for r in rows[1:]:
    if not pat.search(r[k_name]):
        continue
    d = launches.setdefault(r[k_id], {"kernel": r[k_name].split("(")[0]})
    val = float(r[k_val].replace(",", ""))
    if r[k_metric].startswith("dram__bytes"):
        val *= UNIT.get(r[k_unit], 1.)
    d[r[k_metric]] = val
ids = sorted(launches, key=int)[:per_step]
tot = lambda key: sum(launches[i].get(key, 0.) for i in ids)
fp64 = sum(tot("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op)
           for op in ("dadd", "dmul", "dfma"))
out = {"workload": sys.argv[2], "kernels": sorted({launches[i]["kernel"] for i in ids}),
       "launches": len(ids),
       "source": "ncu --clock-control none, metrics summed over the launches of one bench step",
       "dram_bytes_read": tot("dram__bytes_read.sum"), "dram_bytes_write": tot("dram__bytes_write.sum"),
       "dram_bytes_per_step": tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum"),
       "dram_bytes_per_launch": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) /
       max(1, len(ids)),
       "fp64_thread_ops_per_step": fp64,
       "fp64_dfma": tot("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"),
       "fp64_dmul": tot("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
       "fp64_dadd": tot("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
       "gpu_time_under_ncu_ms": tot("gpu__time_duration.sum") / 1e6
       if tot("gpu__time_duration.sum") > 1e5 else tot("gpu__time_duration.sum")}
print(json.dumps(out))
