"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares.
    python tools/summarize_launches.py gpurun_out/launches.csv "<comment>" > profiles/rN_launches_summary.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none ; per-launch times are cold-cache/serialised: compare SHARES")
print("kernel,launches,total_us,share")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k},{cnt[k]},{v:.1f},{v / s:.4f}")
