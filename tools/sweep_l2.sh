#!/usr/bin/env bash
# GPU-box experiment (NOT a bench number): how much of the per-bag time is HBM traffic?  Fewer distinct bags per step
# keep inputs / intermediates in the 126 MB L2.
for b in 16 8 4 2 1; do for l in 8 2 1; do
 if [ $l -le $b ]; then echo -n "bags=$b lanes=$l : "; timeout 200 python bench.py --no-train --no-workloads --no-cpu-baseline --steps 10 --bags $b --lanes $l --reps $((64 / b)) 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['us_per_bag'],2), 'us/bag')"; fi; done; done
