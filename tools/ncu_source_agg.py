"""Aggregate an `ncu --page source --csv` export (SASS view): executed instructions and stall samples by opcode class,
and the hottest SASS stretches.  Usage: ncu_source_agg.py src.csv [warp_ticks]  (warp_ticks: divisor for per-tick figures)"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ix = {h: i for i, h in enumerate(hdr)}
ops, stalls = Counter(), Counter()
tot_i = tot_s = 0
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
    op = m.group(2) if m else "?"
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    ops[op] += n
    stalls[op] += s
    tot_i += n
    tot_s += s
    recs.append((r[ix["Address"]], sass, n, s))
print("total instructions executed %.4g (%.1f per unit), samples %d" % (tot_i, tot_i / div, tot_s))
print("%-12s %12s %8s %8s" % ("opcode", "executed", "per unit", "stall %"))
for op, n in ops.most_common(28):
    print("%-12s %12.4g %8.1f %7.1f%%" % (op, n, n / div, 100.0 * stalls[op] / max(tot_s, 1)))
