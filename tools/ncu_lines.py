"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`: instructions executed, stall samples,
local loads/stores.  Usage: ncu_lines.py export.csv [divisor] [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fname, hdr = None, None
agg = defaultdict(lambda: [0.0, 0.0, 0.0, ""])
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_exec = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        agg[cur][3] = r[1].strip()[:90]
        continue
    if cur is None or len(r) <= i_exec or r[2] in ("...", "-"):
        continue
    try:
        n = float(r[i_exec] or 0)
        s = float(r[i_samp] or 0)
    except ValueError:
        continue
    agg[cur][0] += n
    agg[cur][1] += s
    sass = r[3]
    if "LDL" in sass or "STL" in sass:
        agg[cur][2] += n
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()); tl = sum(v[2] for v in agg.values())
print("total: %.4g instr (%.1f per unit), %d samples, %.4g local ld/st (%.1f per unit)" % (ti, ti / div, ts, tl, tl / div))
print("%-22s %9s %7s %7s  %s" % ("file:line", "instr/u", "stall%", "local/u", "source"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-22s %9.1f %6.1f%% %7.1f  %s" % ("%s:%d" % k, v[0] / div, 100 * v[1] / max(ts, 1), v[2] / div, v[3]))
