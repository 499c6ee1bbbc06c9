#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/lanes_sweep.py 1 4 8 2>&1 | tail -1
timeout 120 python tools/lanes_sweep.py 8 2>&1 | tail -1
RRT_GEMM_SMS=37 RRT_GEMM_PAIR_CAPPED=0 timeout 120 python tools/gemm_trace.py 2>&1 | tail -2
timeout 120 python tools/stage_probe.py 2>&1 | tail -2
