#!/usr/bin/env bash
# First GPU call of the next round: verify and measure the code written after round 1's GPU budget was spent
# (fused LN+QKV GEMM single-CTA / CTA-pair, cluster CR-MSA front end, CUDA-graph replay)
# before anything is built on it.  Run from the repo root on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'   (typically ~6 min; every step has its own timeout)
# Every step runs under its own timeout (a pipeline bug traps after ~2 s of SM clocks, it does not hang).
set -u
mkdir -p gpurun_out
# 1. the default path still green (fast subset) -- the library was rebuilt with the new kernel in it
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2a_tests_default.log 2>&1
echo "default parity rc=$?"
# 2. parity of the experimental kernel (block level vs the default kernels, goldens, lanes == serial)
RRT_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k experimental_fused > gpurun_out/r2a_tests_fused.log 2>&1
echo "fused parity rc=$?"
RRT_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k experimental_cluster > gpurun_out/r2a_tests_cluster_front.log 2>&1
echo "cluster CR-MSA front parity rc=$?"
RRT_CRMSA_FRONT=cluster timeout 200 python tools/stage_probe.py > gpurun_out/r2a_stage_probe_cluster_front.log 2>&1
timeout 200 python tools/stage_probe.py > gpurun_out/r2a_stage_probe_default.log 2>&1
# 2b. CUDA-graph replay of the forward (rrt_mil_b200/graph.py)
RRT_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_graph.py -m gpu -x -q > gpurun_out/r2a_tests_graph.log 2>&1
echo "graph replay rc=$?"
timeout 200 python tools/graph_probe.py > gpurun_out/r2a_graph_probe.log 2>&1
echo "graph probe rc=$?"
# 3. A/B: stage times, us/bag per lane count, phase trace
timeout 300 python tools/fused_qkv_probe.py --trace > gpurun_out/r2a_fused_probe.log 2>&1
echo "probe rc=$?"
timeout 300 python tools/fused_qkv_probe.py --pair --trace > gpurun_out/r2a_fused_probe_pair.log 2>&1
echo "probe (pair) rc=$?"
# 4. bench line with the fused kernel on (compare with profiles/r01c_bench_default.json)
RRT_QKV_FUSED_LN=1 timeout 400 python bench.py > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err
echo "bench rc=$?"
# 5. one ncu --set full capture of the fused kernel (1 GPU; never a bench number)
RRT_QKV_FUSED_LN=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:gemm_lnqkv -c 2 -o gpurun_out/r2a_fused_qkv -f python tools/stage_probe.py \
  > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
tail -n 5 gpurun_out/r2a_tests_fused.log gpurun_out/r2a_fused_probe.log gpurun_out/r2a_fused_probe_pair.log
