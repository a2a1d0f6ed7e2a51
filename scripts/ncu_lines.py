"""Per source line totals of an .ncu-rep captured with -lineinfo / --import-source on:
instructions executed and stall samples, top lines first.
usage: python scripts/ncu_lines.py x.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
        lines.append((int(r[ie] or 0), int(r[ns] or 0), fname, int(r[0]), r[1].strip()))
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print("total warp instructions %.4e, samples %d" % (tot_i, tot_s))
for key, name in ((1, "stall samples"), (0, "instructions")):
    print("---- top lines by", name)
    for l in sorted(lines, key=lambda x: -x[key])[:top]:
        print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100. * l[0] / tot_i, 100. * l[1] / tot_s,
                                                        l[2], l[3], l[4][:100]))
