"""Summarise an `ncu --page source --csv` dump: stall-reason totals, hottest SASS lines, op mix.
    ncu -i rep.ncu-rep --page source --csv --kernel-id ::<name>:<n> > src.csv ; python tools/ncu_src_summary.py src.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[ix["# Samples"]]) for r in data)
tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("total samples", total)
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:24s} {v:8d} {v / max(total, 1):.3f}")
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    best = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Address"]][-5:],
          r[ix["Source"]][:64].ljust(64), best)
mix = collections.Counter()
for r in data:
    toks = r[ix["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    mix[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
t = sum(mix.values())
print("op mix:", [(k, round(v / t, 3)) for k, v in mix.most_common(20)])
