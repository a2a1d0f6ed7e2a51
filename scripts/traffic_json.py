"""Turn an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` log of one kernel
launch into the small JSON bench.py reads for `roofline.traffic`."""
import csv
import json
import sys

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
head = rows[0]
k_name, k_metric, k_unit, k_val = (head.index(x) for x in
                                   ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
out = {"workload": sys.argv[2], "kernel": None, "source": "ncu --metrics dram__bytes_read.sum,"
       "dram__bytes_write.sum --clock-control none, one launch"}
tot = 0.
for r in rows[1:]:
    out["kernel"] = r[k_name].split("(")[0]
    val = float(r[k_val].replace(",", ""))
    if r[k_metric].startswith("dram__bytes"):
        b = val * UNIT.get(r[k_unit], 1.)
        out[r[k_metric]] = b
        tot += b
    elif r[k_metric] == "gpu__time_duration.sum":
        out["gpu_time_under_ncu_" + r[k_unit]] = val
out["dram_bytes_per_launch"] = tot
print(json.dumps(out))
