#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${TAG:-c14}
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_mil_head.py tests/test_graph.py tests/test_gpu_range.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c14_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['us_per_bag'])
print(json.dumps(d.get('train_step'))[:900])
print(json.dumps(d.get('workloads'))[:1500])
PY
