"""Stall samples of an `ncu --page source --csv --print-source cuda,sass` export by code region and stall reason.
Usage: ncu_regions.py export.csv warp_ticks"""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0


def region(f, l):
    # (line ranges of the sources at the end of round 2)
    if f == "ekf_core.cuh":
        for hi, name in ((59, "core:scalar wrappers (fma/sqrt..)"), (341, "core:P ld/st + 3x3 helpers"), (503, "core:quat/attitude"), (541, "core:init_state"), (606, "core:pred_nominal"), (782, "core:pred_cov"), (848, "core:sym6inv"), (954, "core:corr_front"), (1163, "core:correction")):
            if l <= hi:
                return name
        return "core:gate"
    if f == "ekf_synth.cuh":
        return "synth:nees" if l >= 243 else "synth:noise"
    if f == "ekf_kernels.cuh":
        for hi, name in ((229, "k:inputs"), (300, "k:stats_sample"), (343, "k:load/store filter"), (359, "k:correction_call"), (444, "k:votes"), (480, "k:SmemInt"), (690, "k:run_filter (single rate)"), (760, "k:checkpoint/advance_call"), (1004, "k:run_filter_mr (ring)"), (1034, "k:advance_synth_call"), (1283, "k:run_filter_mrs (no ring)")):
            if l <= hi:
                return name
        return "k:kernels"
    return f


fname = hdr = cur = st = None
reg = defaultdict(Counter)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None:
        continue
    if r[0] != "":
        try:
            cur = (fname, int(r[0]))
        except ValueError:
            cur = None
        continue
    if cur is None or len(r) < len(hdr):
        continue
    g = region(*cur)
    try:
        reg[g]["instr"] += float(r[ix["Instructions Executed"]] or 0)
        for h in st:
            reg[g][h] += float(r[ix[h]] or 0)
        if "LDL" in r[3] or "STL" in r[3]:
            reg[g]["local"] += float(r[ix["Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        pass
tot = sum(sum(v[h] for h in st) for v in reg.values())
cols = ["stall_wait", "stall_long_sb", "stall_selected", "stall_barrier", "stall_math", "stall_short_sb", "stall_not_selected", "stall_no_inst", "stall_lg"]
print("(per-line attribution: inlined code is counted at every line of its inline stack, so instruction columns overlap)")
print("%-36s %8s %7s %6s | %s other" % ("region", "instr/u", "local/u", "stall%", " ".join(c[6:11].rjust(5) for c in cols)))
for g, v in sorted(reg.items(), key=lambda kv: -sum(kv[1][h] for h in st)):
    s = sum(v[h] for h in st)
    if s < tot * 0.003:
        continue
    f = [100 * v[h] / tot for h in cols]
    print("%-36s %8.1f %7.1f %5.1f%% | %s %5.1f" % (g, v["instr"] / div, v["local"] / div, 100 * s / tot, " ".join("%5.1f" % x for x in f),
                                                  100 * s / tot - sum(f)))
allc = Counter()
for v in reg.values():
    for h in st:
        allc[h] += v[h]
print("total by reason: " + ", ".join("%s %.1f%%" % (h[6:], 100 * n / tot) for h, n in allc.most_common(10)))
