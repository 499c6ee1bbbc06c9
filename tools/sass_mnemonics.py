"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel takes (cuobjdump -sass of the
built library; runs on the CPU box).  usage: python tools/sass_mnemonics.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rrt_mil_b200", "librrt_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "LDSM",
        "LDGSTS", "MUFU.EX2", "REDG", "RED", "ATOMG", "USETMAXREG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
lines = ["# SASS mnemonics per kernel of librrt_b200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma, UTCBAR = "
         "tcgen05.commit, LDTM / STTM = tcgen05.ld / st,",
         "# UTMALDG / UTMASTG = TMA tile load / store, UTMAPF = prefetch.tensormap, SYNCS = mbarrier, HMMA = mma.sync "
         "(legacy tensor path), LDSM = ldmatrix,",
         "# LDGSTS = cp.async, MUFU.EX2 = ex2.approx, RED / REDG / ATOMG = global reductions.  Kernels with none of "
         "these are omitted."]
for k, c in counts.items():
    if c:
        lines.append(k[:150])
        lines.append("    " + " ".join(f"{w}={c[w]}" for w in WANT if c[w]))
text = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text[:3000])
