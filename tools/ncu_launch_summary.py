"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: device time per kernel of this library
and each kernel's share of the captured launches (cold-cache, serialised: compare SHARES, not absolutes).

    python tools/ncu_launch_summary.py profiles/r2_ncu_launches.csv > profiles/r2_ncu_launches_summary.txt
"""
import collections, csv, sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    n = r["Kernel Name"]
    if not any(t in n for t in ("fm::", "mom::", "seg::", "segb::", "segq::", "p1::", "p1f::", "d128::", "prologue::", "exact::", "dec::")):
        continue
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(n.split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches of this library, {tot:.2f} ms of device time")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} launches {a[0]:3d}  total {a[1]:8.3f} ms  share {100 * a[1] / tot:5.1f} %  per launch {a[1] / a[0]:.3f} ms")
