"""Counts the Blackwell-specific SASS opcodes per kernel of the built library (cuobjdump -sass; no GPU needed):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), plus the
pipes the softmax cares about (MUFU.EX2, F2FP, FFMA2, DFMA).
    python tools/sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gw_whisper_b200", "libgww_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "MUFU.EX2", "F2FP", "FFMA2",
        "FADD2", "DFMA", "LDGSTS", "SYNCS"]
kern = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op.startswith(p):
            counts[kern][p] += 1
            break
print(f"# SASS opcode counts per kernel of {os.path.basename(lib)} (cuobjdump -sass, sm_100a)")
print("# " + " ".join(pats))
tot = collections.Counter()
for k, c in counts.items():
    if not c:
        continue
    tot.update(c)
    print(f"{k}\n    " + "  ".join(f"{p}={c[p]}" for p in pats if c[p]))
print("TOTAL\n    " + "  ".join(f"{p}={tot[p]}" for p in pats if tot[p]))
