"""Per-kernel counts of the Blackwell-native SASS mnemonics in diffusiondrive_b200/_ddh.so
(tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, tcgen05.commit -> UTCBAR,
legacy mma.sync -> HMMA, cp.async -> LDGSTS).  Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "diffusiondrive_b200", "_ddh.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTMAPF", "HMMA", "LDGSTS", "LDSM", "SYNCS")
cur, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if cur:
        for k in MN:
            if re.search(r"\b" + k + r"[\.\s]", line):
                counts[cur][k] += 1
print(f"# {os.path.relpath(so, ROOT)}  arch {arch}  (cuobjdump -sass; instruction counts per kernel)")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in MN))
tot = collections.Counter()
for name, c in counts.items():
    if sum(c[k] for k in MN[:8]) == 0 and c["LDGSTS"] == 0:
        continue
    print(f"{name[-70:]:70s} " + " ".join(f"{c[k]:8d}" for k in MN))
    tot.update(c)
print(f"{'TOTAL':70s} " + " ".join(f"{tot[k]:8d}" for k in MN))
