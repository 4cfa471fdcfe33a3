"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_summary.py launches.csv > summary.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
tot = 0.0
for r in rd:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    a = agg[name]
    a[0] += 1
    a[1] += v
    a[2][(r[ix["Grid Size"]], r[ix["Block Size"]])] += 1
    tot += v
print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of serialized kernel time")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    shapes = ", ".join(f"{g}x{b}:{n}" for (g, b), n in a[2].most_common(3))
    print(f"{k[:64]:64s} n={a[0]:6d} total={a[1] / 1e3:8.2f} ms {100 * a[1] / tot:5.1f} %  avg={a[1] / a[0]:7.1f} us   [{shapes}]")
