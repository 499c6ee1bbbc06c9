#!/usr/bin/env bash
# GPU-box tool: headline us/bag under a few knob settings (attention SM cap, lanes, attention kernel)
run() { echo -n "$* : "; env "$@" timeout 200 python bench.py --no-train --no-workloads --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['us_per_bag'],2), 'us/bag', round(d['value']/1e6,1), 'M patches/s')"; }
run RRT_ATTN=auto
run RRT_ATTN=mma
run RRT_ATTN_SMS=96
run RRT_ATTN_SMS=64
run RRT_ATTN_SMS=43
for l in 2 4 6; do echo -n "lanes=$l : "; timeout 200 python bench.py --no-train --no-workloads --no-cpu-baseline --steps 10 --lanes $l 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['us_per_bag'],2), 'us/bag')"; done
