#!/usr/bin/env python
"""SASS opcode histogram of libflappy_b200.so, per kernel family (cuobjdump -sass): the evidence that the hot kernels are
tcgen05 / TMA code (UTCHMMA, UTMALDG, UBLKCP, LDTM) and what the env kernel issues.  Writes profiles/r02_sass_histogram.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dqnflappybird_b200", "libflappy_b200.so")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
fam = collections.defaultdict(collections.Counter)
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = re.sub(r"\(.*", "", name)
        cur = re.sub(r"^void ", "", cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        fam[cur][m.group(1)] += 1
key = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "STG", "LDG", "STS", "LDS", "IMAD", "LOP3", "SHF", "PRMT"]
lines = ["SASS opcode counts per kernel of libflappy_b200.so (cuobjdump -sass, sm_100a); columns: " + " ".join(key) + " | total", ""]
tot = collections.Counter()
for k in sorted(fam):
    c = fam[k]
    tot.update(c)
    lines.append(f"{k[:86]:86s} " + " ".join(f"{c.get(o, 0):5d}" for o in key) + f" | {sum(c.values()):6d}")
lines.append("")
lines.append(f"{'ALL KERNELS':86s} " + " ".join(f"{tot.get(o, 0):5d}" for o in key) + f" | {sum(tot.values()):6d}")
path = os.path.join(ROOT, "profiles", "r02_sass_histogram.txt")
open(path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
print("wrote", path)
