"""SASS instruction count / code bytes per kernel of the built library (cuobjdump; runs on the CPU box).
usage: python tools/sass_sizes.py [min_instructions]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rrt_mil_b200", "librrt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
rows, cur, n = [], None, 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur: rows.append((n, cur))
        cur, n = m.group(1), 0
    elif re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        n += 1
if cur: rows.append((n, cur))
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (n, _), nm in sorted(zip(rows, names), reverse=True):
    if n >= lo:
        print(f"{n:6d} instr {n * 16 // 1024:4d} KB  {nm[:120]}")
