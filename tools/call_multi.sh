#!/usr/bin/env bash
# scratch driver: N-GPU bench (+ the NCCL tests at N = 2)
cd "$(dirname "$0")/.."
O=gpurun_out
N=${N:-2}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline > $O/r02f_bench_${N}gpu.json 2> $O/r02f_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 1500 $O/r02f_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['us_per_bag'], d['per_rank_ms'], d['e2e']['value'], d.get('gather'))
w=d.get('workloads') or {}
print({k:(v.get('value'), v.get('cuda_graph',{}).get('us_per_step') if isinstance(v,dict) else None) for k,v in w.items()})
PY
