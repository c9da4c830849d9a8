#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`: top CUDA source lines by stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file, hdr, out = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(range(len(hdr)), r))
        i_s = hdr.index("# Samples")
        i_e = hdr.index("Instructions Executed")
        stalls = {hdr[i]: int(r[i] or 0) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i] and r[i].isdigit()}
        out.append((int(r[i_s] or 0), int(r[i_e] or 0), cur_file, r[0], r[1].strip()[:90], stalls))
tot = sum(o[0] for o in out) or 1
tote = sum(o[1] for o in out) or 1
print("samples", tot, "warp-instr", tote)
for s, e, f, ln, src, st in sorted(out, key=lambda o: -o[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100*s/tot:5.1f}% {100*e/tote:5.1f}%i {f}:{ln:>4s} {src:90s} [{top}]")
