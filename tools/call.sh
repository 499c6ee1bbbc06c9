#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${TAG:-c11}
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/${T}_tests.log
timeout 120 python tools/stage_probe.py 2>&1 | tail -2
timeout 120 python tools/lanes_sweep.py 1 4 8 2>&1 | tail -1
timeout 120 python tools/gemm_trace.py 2>&1 | tail -4
timeout 120 python tools/attn_probe.py --quick --trace 2>&1 | tail -16
