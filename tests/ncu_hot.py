#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples / executed count."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        samples = int(r[idx["# Samples"]] or 0)
        execd = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    stalls = {h: int(r[idx[h]] or 0) for h in stall_cols}
    data.append((samples, execd, r[idx["Source"]], stalls))
tot = sum(d[0] for d in data)
tot_exec = sum(d[1] for d in data)
print("total samples", tot, "total warp-instructions", tot_exec)
print("-- top by samples")
for s, e, src, st in sorted(data, key=lambda d: -d[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/tot:5.1f}% exec={e:9d} {src[:70]:70s} {top}")
print("-- top by executed")
for s, e, src, st in sorted(data, key=lambda d: -d[1])[:12]:
    print(f"{100*e/tot_exec:5.1f}% samples={100*s/tot:4.1f}% {src[:80]}")
agg = {}
for s, e, src, st in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print("-- stall totals", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
