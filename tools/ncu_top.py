"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ix["# Samples"]])
    except: continue
    st = {h: int(r[ix[h]] or 0) for h in stall_cols}
    data.append((n, r[ix["Address"]][-5:], r[ix["Source"]].strip(), st, r[ix["Instructions Executed"]]))
tot = sum(d[0] for d in data)
print("total samples", tot)
agg = {}
for d in data:
    for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for n, addr, src, st, ex in sorted(data, key=lambda d: -d[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{n:7d} {100*n/tot:5.1f}% {addr} ex={ex:>8} {src[:70]:70s} {top}")
