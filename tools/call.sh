#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${TAG:-c3}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_mil_head.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
timeout 120 python tools/stage_probe.py 2>&1 | tail -2
timeout 120 python tools/gemm_trace.py 2>&1 | tail -2
timeout 120 python tools/lanes_sweep.py 1 4 8 2>&1 | tail -1
for c in 24 48 64 96; do echo -n "resid cap $c: "; RRT_GEMM_SMS_RESID=$c timeout 120 python tools/lanes_sweep.py 8 2>&1 | tail -1; done
for c in 48 64 ; do echo -n "all-gemm cap $c: "; RRT_GEMM_SMS=$c timeout 120 python tools/lanes_sweep.py 8 2>&1 | tail -1; done
