#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
for occ in 1 2; do
RRT_CRB_OCC=$occ timeout 600 python bench.py --no-cpu-baseline --no-workloads --steps 3 > $O/c15_occ$occ.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/c15_occ$occ.json').read().strip().splitlines()[-1])
print('occ', $occ, d['train_step']['stages_us_per_step']['bwd_crmsa'], d['train_step']['us_per_bag_fwd_bwd'])
PY
done
RRT_CRB_OCC=2 timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -2
