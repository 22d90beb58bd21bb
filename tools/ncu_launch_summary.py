"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum ... --csv --log-file X.csv): time per kernel type.
    python tools/ncu_launch_summary.py X.csv ["title line"] > profiles/rNN_<what>_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ix = {c: j for j, c in enumerate(hdr)}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v = v / 1000.0 if r[ix["Metric Unit"]] in ("ns", "nsecond") else (v * 1000.0 if r[ix["Metric Unit"]] in ("ms", "msecond") else v)
    name = r[ix["Kernel Name"]]
    name = name[:name.index("(")] if "(" in name else name
    key = f"{name[:84]} grid={r[ix['Grid Size']]} blk={r[ix['Block Size']]}"
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
if len(sys.argv) > 2:
    print(sys.argv[2])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% n={n:5d} avg={t / n:8.2f} us  {k}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
