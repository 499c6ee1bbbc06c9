#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
T=r02g
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/${T}_tests.log
timeout 600 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/${T}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-workloads > $O/${T}_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none -s 14 -c 7 -o $O/${T}_fwd_full -f python tools/one_bag.py 3 > $O/${T}_ncu_full.log 2>&1; tail -1 $O/${T}_ncu_full.log
python tools/ncu_summary.py $O/${T}_fwd_full.ncu-rep $O/${T}_ncu_full_summary.csv > /dev/null 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['us_per_bag'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['encoder']['frac'], d['clocks'])
print(d['stages_us_per_launch'])
print(d['train_step']['us_per_bag_fwd_bwd'], d['workloads']['train_configs4']['cuda_graph']['us_per_step'], d['workloads']['mil_configs2']['value'])
PY
