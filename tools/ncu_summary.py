"""Per-kernel summary table from an .ncu-rep (raw page).  Tooling only.
usage: python tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']
idx = [hdr.index(w) for w in want if w in hdr]
names = [hdr[i] for i in idx]
short = [n.replace('.avg.pct_of_peak_sustained_', '%').replace('launch__', '').replace('.sum', '') for n in names]
table = [short, [units[i] for i in idx]] + [[r[i] for i in idx] for r in rows[2:]]
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w")).writerows(table)
seen = set()
for r in table[2:]:
    key = r[0][:80]
    if key in seen:
        continue
    seen.add(key)
    print(r[0][:90])
    print("    " + " | ".join(f"{s}={v}" for s, v in zip(short[1:], r[1:])))
