"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (evidence that the hot path runs
on tcgen05 / TMEM / the bulk-copy engine):  python tools/sass_summary.py > profiles/r2_sass_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "oprl_b200", "liboprl_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "ACQBULK", "ATOM", "RED",
             "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "BAR", "NANOSLEEP", "STL", "LDL"]
cur, counts, lines = None, collections.OrderedDict(), {}
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        lines[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if cur and m:
        lines[cur] += 1
        op = m.group(1)
        for k in MNEMONICS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction counts per kernel (static, not dynamic)")
print(f"# {'kernel':58s} {'instrs':>7s}  " + " ".join(f"{k:>7s}" for k in MNEMONICS if any(c[k] for c in counts.values())))
for name, c in counts.items():
    print(f"  {name[:58]:58s} {lines[name]:7d}  " + " ".join(f"{c[k]:7d}" for k in MNEMONICS if any(cc[k] for cc in counts.values())))
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (tensor memory), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk")
print("# (1-D bulk copies: operands are pre-tiled CT32, so no tensor maps -> no UTMALDG), SYNCS = mbarrier ops, ELECT = elect.sync")
