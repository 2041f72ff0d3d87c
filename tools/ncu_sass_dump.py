"""Prints the SASS lines of an `ncu --page source --csv` dump that carry samples (or are sync / TMA /
MMA / TMEM instructions), in address order, with per-execution stall estimates.
    python tools/ncu_sass_dump.py src.csv [min_samples]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 20
seen = set()
tot = 0
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    tot += s
    src = r[ix["Source"]]
    if s >= mins or any(k in src for k in ("SYNCS", "UTC", "UTMA", "LDTM", "STTM", "BAR", "STG", "LDG", "STS", "LDS")):
        st = {h: int(r[ix[h]] or 0) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        best = max(st.items(), key=lambda x: x[1]) if st else ("", 0)
        print(r[0][-5:], str(s).rjust(6), r[ix["Instructions Executed"]].rjust(9), src[:86].ljust(86), best[0][6:], best[1])
print("total samples", tot)
