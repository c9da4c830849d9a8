#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into the launch list + per-kernel summary under profiles/.

    python tests/ncu_launches.py gpurun_out/launches.csv profiles/rNN_launches_<what>.csv profiles/rNN_launches_<what>_summary.txt "<command>"
"""
import collections
import csv
import re
import sys

src, out_csv, out_txt, cmd = sys.argv[1:5]
rows = list(csv.reader(open(src, errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[h]
kn, gs, bs, mv = H.index("Kernel Name"), H.index("Grid Size"), H.index("Block Size"), H.index("Metric Value")


def short(name):
    name = re.sub(r"\(.*", "", name)                       # drop the argument list
    name = re.sub(r"lcm::|<?unnamed>::|\(anonymous namespace\)::|void ", "", name)
    return name.strip()


launches = [(short(r[kn]), r[gs], r[bs], int(float(r[mv].replace(",", "")))) for r in rows[h + 1:] if len(r) > mv]
with open(out_csv, "w") as f:
    f.write(f"# {cmd}\n# cold-cache, serialised launches: compare shares, not absolute times\nid,kernel,grid,block,duration_ns\n")
    for i, (k, g, b, d) in enumerate(launches):
        f.write(f'{i},"{k}","{g}","{b}",{d}\n')
tot = sum(d for *_, d in launches)
agg = collections.OrderedDict()
for k, _, _, d in launches:
    a = agg.setdefault(k, [0, 0]); a[0] += 1; a[1] += d
with open(out_txt, "w") as f:
    f.write(f"ncu launch list of `{cmd}` ({len(launches)} launches), share of device time per kernel\n"
            "(per-launch times under ncu are cold-cache and serialised; bench.py's own CUDA-event shares are in BENCH json `roofline.per_kernel`)\n\n")
    for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:48s} n={n:4d} {d / 1e3:10.1f} us {100.0 * d / tot:5.1f}%\n")
print(open(out_txt).read())
