#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break  # next captured launch
    if len(r) == len(hdr):
        body.append(r)
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("stall mix:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body_idx = list(enumerate(body))
for i, r in sorted(body_idx, key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  #%-4d exec=%-8s thr=%-5s %-60s %s" % (s, 100.0 * s / max(tot, 1), i, r[col["Instructions Executed"]],
          r[col["Avg. Threads Executed"]], r[col["Source"]].strip()[:60], ",".join("%s:%d" % (k, v) for v, k in top if v)))
