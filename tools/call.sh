#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/stage_probe.py 2>&1 | tail -2
timeout 200 python tools/graph_probe.py 2>&1 | tail -5
