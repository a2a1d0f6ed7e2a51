"""Shares of the kernels in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/launch_summary.py profiles/r02_launches_bench_c2_100k.csv "<command>" > ..._summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity"
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[mv].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1., "s": 1e3}.get(r[mu], 1e-6)
    name = re.sub(r"\(.*", "", r[kn]).strip()
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("launch list of `%s` under" % cmd)
print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: shares, not absolutes)")
print("total %.1f ms over %d launches\n" % (total, sum(cnt.values())))
for name, ms in tot.most_common():
    print("%6.2f %% %11.2f ms %6d x  %s" % (100. * ms / total, ms, cnt[name], name[:110]))
