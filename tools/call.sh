#!/usr/bin/env bash
# scratch driver for one gpurun call (edited per call)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${TAG:-c12}
timeout 300 python tools/attn_probe.py --trace 2>&1 | tail -32
timeout 120 python tools/lanes_sweep.py 1 8 2>&1 | tail -1
