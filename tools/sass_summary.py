"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UBLKCP = bulk async copy (TMA unit), UTCBAR = tcgen05.commit, HMMA = legacy mma.sync) from `cuobjdump -sass libfdpt.so`.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "framedipt_b200", "libfdpt.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTC*MMA (other)": r"\bUTC(?!HMMA|BAR|ATOM)[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UBLKCP": r"\bUBLKCP",
       "UTCBAR": r"\bUTCBAR", "SYNCS (mbarrier)": r"\bSYNCS", "HMMA (legacy)": r"\bHMMA", "LDGSTS (cp.async)": r"\bLDGSTS"}
cur, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# cuobjdump -sass framedipt_b200/libfdpt.so  (arch {arch}); instruction counts per kernel")
cols = list(PAT)
print(f"{'kernel':70s} " + " ".join(f"{c:>18s}" for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(f"{k[:70]:70s} " + " ".join(f"{c.get(x, 0):18d}" for x in cols))
    tot.update(c)
print(f"{'TOTAL':70s} " + " ".join(f"{tot.get(x, 0):18d}" for x in cols))
print(f"# kernels without any of these (plain SIMT): {sum(1 for c in counts.values() if sum(c.values()) == 0)}")
