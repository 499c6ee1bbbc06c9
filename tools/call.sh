#!/usr/bin/env bash
cd "$(dirname "$0")/.."
run() { echo -n "$* : "; env "$@" timeout 120 python tools/lanes_sweep.py 4 2>&1 | tail -1; }
echo -n "default: "; timeout 120 python tools/lanes_sweep.py 1 8 2>&1 | tail -1
run RRT_ATTN_SMS_4=0
run RRT_ATTN_SMS_4=64
run RRT_ATTN_SMS_4=86
run RRT_ATTN_SMS_4=103
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_graph.py -x -q -m gpu 2>&1 | tail -2
