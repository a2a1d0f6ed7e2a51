"""Summarise an .ncu-rep (read on the CPU box): key metrics, opcode mix and the hot loop.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--hot 2.5e8]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
hot_thr = float(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for h, u, v in zip(hdr, units, val):
    if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            if h.startswith("smsp__average") and float(v) < 0.15:
                continue
        except ValueError:
            pass
        print("%-80s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
H = rows[hi]
ie, sc, sm = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
ops, tot, data = collections.Counter(), 0, []
for r in rows[hi + 1:]:
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[sc])
    op = (m.group(2) if m else r[sc][:8]).split(".")[0]
    ops[op] += n
    tot += n
    data.append((n, int(r[sm] or 0), r[sc]))
print("total warp instructions %.4e" % tot)
print("  ".join("%s %.1f%%" % (k, 100. * v / tot) for k, v in ops.most_common(28)))
if hot_thr:
    hot = [d for d in data if d[0] >= hot_thr]
    print(len(hot), "instructions executed >= %.2e times; sum %.3e" % (hot_thr, sum(d[0] for d in hot)))
    for d in hot:
        print("%.2e %6d %s" % (d[0], d[1], d[2][:90]))
