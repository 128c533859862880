#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed-instruction mix and the
top stall sites of the hottest loop.  usage: tools/ncu_hot.py source.csv [n_top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[h]
data = []
for r in rows[h + 1:]:  # first kernel instance only
    if "Source" in r or (r and r[0] == "Kernel Name"):
        break
    if len(r) > 10:
        data.append(r)
c = {k: hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "stall_wait", "stall_long_sb", "stall_math",
                               "stall_not_selected", "stall_short_sb", "stall_branch_resolving", "stall_dispatch", "stall_selected", "stall_no_inst", "stall_lg")}
f = lambda r, k: float(r[c[k]] or 0)
op = lambda s: re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
mix = collections.Counter()
for r in data:
    mix[op(r[c["Source"]])] += f(r, "Instructions Executed")
tot = sum(mix.values())
print("executed warp-instructions: %.3e" % tot)
print("  " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in mix.most_common(18)))
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print("  FP64-pipe share of issued instructions: %.1f%%" % (100 * fp64 / tot))
samples = sum(f(r, "# Samples") for r in data)
print("stall samples by reason:", {k[6:]: "%.1f%%" % (100 * sum(f(r, k) for r in data) / samples) for k in c if k.startswith("stall_")})
print("top sample sites:")
for i, r in sorted(enumerate(data), key=lambda ir: -f(ir[1], "# Samples"))[:top]:
    print("  %5d %6.0f  wait=%4.0f long=%4.0f math=%4.0f nsel=%4.0f  %s" % (i, f(r, "# Samples"), f(r, "stall_wait"), f(r, "stall_long_sb"),
                                                                      f(r, "stall_math"), f(r, "stall_not_selected"), r[c["Source"]].strip()[:70]))
