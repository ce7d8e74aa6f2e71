"""Per-kernel tally of the SASS mnemonics that prove which hardware paths the shipped library uses
(tcgen05 int8 MMA, TMA, TMEM loads, FP64 tensor-core MMA, 1-D bulk copies, mbarriers):
    python tools/sass_tally.py > profiles/r02_sass_tally.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dbslmm_b200", "libdbslmm_b200.so")
KEYS = ["UTCIMMA", "UTMALDG", "LDTM", "DMMA", "UBLKCP", "UTMAPF", "SYNCS", "BAR.SYNC", "LDGSTS", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
cur, tab, arch = None, collections.OrderedDict(), set()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("dbslmm::", "")
        tab[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    if cur:
        for k in KEYS:
            if re.search(r"\b" + re.escape(k), ln):
                tab[cur][k] += 1
print(f"# cuobjdump -sass dbslmm_b200/libdbslmm_b200.so ({len(tab)} kernels, arch {sorted(arch)}): occurrences per kernel")
print(f"{'kernel':44s} " + " ".join(f"{k:>8s}" for k in KEYS))
for k, c in tab.items():
    print(f"{k[:44]:44s} " + " ".join(f"{c[x]:8d}" for x in KEYS))
