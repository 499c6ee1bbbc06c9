"""Top stall lines per kernel from an .ncu-rep (source page, SASS view).  Tooling only.
usage: python tools/ncu_hot.py report.ncu-rep [kernel-substring] [top_n]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = row
        elif len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
seen = set()
for b in blocks:
    if want not in b["name"] or b["name"] in seen:
        continue
    seen.add(b["name"])
    h = b["hdr"]
    si, ai = h.index("Source"), h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[ai] or 0) for r in b["rows"]) or 1
    print(f"=== {b['name'][:110]}  (samples {tot}, {len(b['rows'])} SASS lines)")
    rows = sorted(b["rows"], key=lambda r: -int(r[ai] or 0))[:top]
    for r in rows:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"  {int(r[ai])/tot*100:5.1f}%  {r[si].strip()[:70]:70s} {st}")
