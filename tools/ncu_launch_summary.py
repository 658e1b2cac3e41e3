"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv
python bench.py ...`): python tools/ncu_launch_summary.py x.csv "<command line>" > profiles/<name>.txt
ncu times every launch alone and cold, so compare SHARES with bench.py's in-step `kernel_shares`, not absolute times."""
import collections
import csv
import re
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = open(path, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.reader(lines[start:]))
hdr = {h: i for i, h in enumerate(rows[0])}
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= hdr["Metric Value"] or r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    u = r[hdr["Metric Unit"]]
    us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
    name = re.sub(r"\(.*$", "", r[hdr["Kernel Name"]])[:100]
    t = tot.setdefault(name, [0.0, 0])
    t[0] += us
    t[1] += 1
total = sum(t[0] for t in tot.values()) or 1.0
print(cmd)
print("total %.1f ms over %d launches (each launch timed alone by ncu: cold caches, serialised)" % (total / 1e3, sum(t[1] for t in tot.values())))
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%9.2f ms %5.1f%% n=%6d avg=%8.1f us  %s" % (us / 1e3, 100 * us / total, n, us / n, name))
