#!/usr/bin/env python
"""Benchmark of the RRTEncoder hot path: patches/sec through RRTEncoder at N=9000, D=512
(BASELINE.json metric; workload = configs[1]'s shape, forward, eval mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = ``--reps`` passes of the encoder over ``--bags`` distinct synthetic bags (4 x 16 bags of
9000 x 512 fp32 by default: 295 MB of distinct inputs and as much output, > the 126 MB L2, so an iteration never
finds its inputs in L2; 64 bags per step keep the timed region near 100 ms).  Each rank (one process per GPU)
owns its own bags; bags are independent, so the data path needs no collective and scaling is weak.

Printed JSON (rank 0, one line): the base contract's keys plus
  roofline      the stage with the LARGEST TIME SHARE against its own bound (measured peaks), the whole encoder
                against the tensor peak (``roofline.encoder``) and every stage's fraction (``roofline.stages``)
  per_rank_ms   the timed region of every rank (the headline takes the max)
  gather        (N > 1) the same step followed by ONE NCCL all-gather of the ragged outputs (north star's
                "single NCCL gather of outputs"): patches/s including it, and the collective's own time
  e2e           host buffers in, host buffers out, copies inside the timed region; ``copy_ceiling`` = what the
                host<->device path carries with all ranks copying at once (the bound of e2e)
  workloads     BASELINE configs[2] (RRTMIL over 8 ragged R50-shaped bags, only logits leave the GPU) and
                configs[4] (full RRTMIL train step, data-parallel, gradient all-reduce overlapped with backward)
  cpu_baseline  the reference's CPU forward on this box's host cores (N = 1 only)

``--impl reference`` times the reference's own CPU implementation (the unmodified ``modules/rrt.py`` staged under
the git-ignored ``oracle/_ref/`` by ``oracle/build_ref.py``; the oracle's reference-order port if it is absent)
on all host threads, one bag per step, and prints the same line with ``"impl": "reference"``.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "patches/sec through RRTEncoder at N=9000 D=512"
UNIT = "patches/s"
N_TOKENS, DIM = 9000, 512
ENC_KW = dict(mlp_dim=DIM, region_num=8, epeg_k=15, crmsa_k=3, n_layers=2, n_heads=8, crmsa_heads=8)
NVLINK_PEER_GBS = 770.0   # measured peer copy per direction on this pool (B200_PROFILING.md)


# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


def _grids(L):
    H = math.isqrt(L - 1) + 1
    Hr = H + (-H) % ENC_KW["region_num"]
    Hc = H + (-H) % 8
    return Hr * Hr, (Hr // ENC_KW["region_num"]) ** 2, Hc * Hc


def encoder_flops(L):
    """Algorithmic forward FLOPs of one bag."""
    D, ke, kc = DIM, ENC_KW["epeg_k"], ENC_KW["crmsa_k"]
    np_r, p_r, np_c = _grids(L)
    T = kc * 64
    return np_r * (8 * D * D + 4 * p_r * D + 2 * ke * D) + T * (8 * D * D + 4 * 64 * D) + 6 * np_c * D * kc


def stage_algorithmic(stage, L):
    """(bound, algorithmic flops, algorithmic bytes) of one launch of a stage at N=9000/D=512.
    Bytes: fp32 residual stream (4 B), f16 internal activations / weight shadows (2 B); DESIGN.md 4-5."""
    D = DIM
    np_r, p_r, _ = _grids(L)
    if stage == "qkv_gemm":
        return "tensor", 2.0 * np_r * 3 * D * D, 2.0 * (np_r * D + 3 * D * D + np_r * 3 * D)
    if stage == "proj_gemm_residual":   # 4.8 GFLOP but 46.8 MB of residual / output traffic: HBM binds (7.1 vs 2.9 us)
        return "hbm", 2.0 * np_r * D * D, 2.0 * (np_r * D + D * D) + 4.0 * 2 * L * D
    if stage == "rmsa_attention":       # 2.7 GFLOP, 37.7 MB: HBM binds (5.8 vs 1.6 us)
        return "hbm", 4.0 * np_r * p_r * D, 2.0 * (np_r * 3 * D + np_r * D)
    if stage == "ln_partition":
        return "hbm", 0.0, 4.0 * L * D + 2.0 * np_r * D
    if stage == "crmsa_landmarks":
        return "hbm", 0.0, 4.0 * L * D
    if stage == "crmsa_dispatch_final_ln":
        return "hbm", 0.0, 4.0 * 2 * L * D
    return "latency", 0.0, 0.0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms.  The process is started BEFORE the warm-up steps
    (its NVML start-up takes driver locks for ~100 ms and stalled the launches of the first timed steps when every
    rank started one at the top of the timed region) and keeps sampling through the timed region; `mark()` sets
    the window whose samples are summarised (falls back to every sample under load when the window is shorter
    than a sampling period)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def mark(self, start):
        if start:
            self.t0 = time.monotonic()
        else:
            self.t1 = time.monotonic()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1 + 0.05]
        window = "timed region"
        if not inside:      # region shorter than a sampling period: every sample since the warm-up started
            inside, window = [r for _, r in self.rows], "warm-up + timed region"
        for r in inside:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
# reference side (CPU baseline leg and --impl reference: the only places bench.py executes oracle/)
def reference_forward_factory(threads, device="cpu", autocast=False):
    """(callable running one N=9000 forward, kind).  kind = "reference": the UNMODIFIED modules/rrt.py staged
    under oracle/_ref/ (or /root/reference in the build container); "port": the oracle's reference-order
    restatement, when no reference tree is available."""
    from oracle import rrt_oracle as O
    from oracle import _reference_shim as shim
    torch.set_num_threads(threads)
    torch.manual_seed(2021)                      # reference default seed (main.py:645)
    if shim.available():
        ref = shim.import_reference_rrt()
        m = ref.RRTEncoder(need_init=True, **ENC_KW).to(device).eval()
        x = torch.randn(1, N_TOKENS, DIM).to(device)

        def run():
            with torch.no_grad():
                if autocast:
                    with torch.autocast("cuda", dtype=torch.float16):
                        return m(x)
                return m(x)
        return run, "reference"
    cfg = O.EncoderConfig(**{k: v for k, v in ENC_KW.items()})
    w = {k: v.to(device) for k, v in O.make_weights(cfg, 2021, dtype=torch.float32, randomize_bias=False).items()}
    x = O.make_bag(N_TOKENS, DIM, 7, dtype=torch.float32).to(device)

    def run():
        with torch.no_grad():
            return O.encoder_forward(x, w, cfg, "reference")
    return run, "port"


def time_cpu_baseline(budget_s=12.0, min_bags=5):
    threads = os.cpu_count() or 1
    run, kind = reference_forward_factory(threads)
    for _ in range(2):
        run()
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < min_bags or time.perf_counter() < t_end:
        t0 = time.perf_counter(); run(); times.append(time.perf_counter() - t0)
        if len(times) >= 200:
            break
    med = statistics.median(times)
    what = ("the unmodified reference modules/rrt.py RRTEncoder (oracle/_ref)" if kind == "reference"
            else "oracle/rrt_oracle.py reference operator order")
    return {"value": N_TOKENS / med, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{len(times)} forwards of one N={N_TOKENS} D={DIM} bag, fp32, eval, median "
                      f"{med * 1e3:.1f} ms/bag ({what})"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    run, kind = reference_forward_factory(threads)
    for _ in range(max(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    val = N_TOKENS * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"RRTEncoder forward, one bag N={N_TOKENS} D={DIM} region_num=8 "
                               "epeg_k=15 crmsa_k=3 per step, CPU", "bags_per_step": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steps x 1 bag of N={N_TOKENS}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    # Informational extras ("beside the reference GPU-eager number"; the north star's ">= 10x the
    # reference 1-GPU PyTorch forward"): the same module run eagerly on cuda:0 with the reference's defaults
    # (fp32, allow_tf32 off) and under fp16 autocast as its --amp flag does (main.py:101-102,439).
    if torch.cuda.is_available():
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        for key, ac in (("gpu_eager_fp32", False), ("gpu_eager_fp16_autocast", True)):
            try:
                grun, gkind = reference_forward_factory(threads, device="cuda", autocast=ac)
                for _ in range(3):
                    grun()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    grun()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                line[key] = {"value": N_TOKENS / (ms * 1e-3), "unit": UNIT, "ms_per_bag": ms, "kind": gkind,
                             "what": "the reference's forward on cuda:0, eager, CUDA events, 10 bags back to back"}
            except Exception as e:  # informational only
                line[key] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def bind_rank_to_cores(local, n_local):
    """Give every rank of a node its own slice of the host cores (launch threads and pinned-buffer first touch
    stay off each other; the box is one NUMA node, so this is all the placement there is to do)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(n_local, 1)
        if n_local > 1 and per >= 1:
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]))
            return per
    except (AttributeError, OSError):
        pass
    return None


def copy_ceiling(dev, world, barrier, dist):
    """Host<->device GB/s each way with BOTH directions busy on every rank at once (pinned 64 MB buffers)."""
    n = 16 << 20
    hx, hy = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    hx.fill_(1.0); hy.fill_(0.0)
    dx, dy = torch.empty(n, device=dev), torch.ones(n, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(reps):
        for _ in range(reps):
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        torch.cuda.synchronize(dev)
    run(2)
    barrier()
    t0 = time.perf_counter()
    run(10)
    dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * n * 4 / dt / 1e9], device=dev, dtype=torch.float64)
    lo = gbs.clone()
    if world > 1:
        dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return {"gbs_each_way_sum_over_ranks": float(gbs.item()), "gbs_each_way_slowest_rank": float(lo.item()),
            "how": "every rank copies 64 MB pinned buffers H2D and D2H concurrently on two streams, 10 rounds, wall clock"}


def run_b200_arm(args):
    import torch.distributed as dist
    from rrt_mil_b200 import RRTEncoder, RRTMIL, cabi, parallel
    from rrt_mil_b200.optim import Adam
    from rrt_mil_b200.pipeline import HostPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product arm)")
    cores_per_rank = bind_rank_to_cores(local, n_local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.manual_seed(2021)  # reference default seed (main.py:645)
    enc = RRTEncoder(need_init=True, **ENC_KW).to(dev).eval()
    B, reps = args.bags, max(1, args.reps)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    bags = [torch.randn(N_TOKENS, DIM, device=dev, generator=gen) for _ in range(B)]
    outs = [torch.empty_like(b) for b in bags]
    patches_per_step = reps * B * N_TOKENS

    def step():
        with torch.no_grad():
            for _ in range(reps):
                enc.forward_bags(bags, outs, lanes=args.lanes)

    with ClockSampler(local) as clk:
        time.sleep(0.3)                      # nvidia-smi is up and sampling before anything is timed
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()

        # ---- timed region: device-resident inputs -------------------------------------------------
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = cabi.launch_count()
        barrier(); torch.cuda.synchronize()
        clk.mark(True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        clk.mark(False)
    barrier()
    launches = cabi.launch_count() - n0
    my_ms = e0.elapsed_time(e1)
    per_rank = torch.tensor([my_ms], device=dev, dtype=torch.float64)
    if world > 1:
        allms = torch.empty(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allms, per_rank)
        per_rank = allms
    per_rank_ms = [round(float(v), 4) for v in per_rank.tolist()]
    ms_total = max(per_rank_ms)
    value = world * patches_per_step * args.steps / (ms_total * 1e-3)

    # ---- the same step followed by ONE NCCL all-gather of the ragged outputs (N > 1) -------------
    gather = None
    if world > 1:
        g_steps = max(2, min(args.steps, 5))
        with torch.no_grad():
            enc.forward_bags(bags, outs, lanes=args.lanes)
            parallel.gather_ragged(outs, B * world)          # warm-up (NCCL channels, buffers)
        torch.cuda.synchronize(); barrier()
        ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        coll_ms = 0.0
        ga.record()
        for _ in range(g_steps):
            with torch.no_grad():
                enc.forward_bags(bags, outs, lanes=args.lanes)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            allouts = parallel.gather_ragged(outs, B * world)
            c1.record()
            torch.cuda.synchronize()
            coll_ms += c0.elapsed_time(c1)
            del allouts
        gb.record()
        torch.cuda.synchronize()
        g_ms = max_over_ranks(ga.elapsed_time(gb))
        c_ms = max_over_ranks(coll_ms) / g_steps
        recv_bytes = (world - 1) * B * N_TOKENS * DIM * 4
        gather = {"value": world * B * N_TOKENS * g_steps / (g_ms * 1e-3), "unit": UNIT, "steps": g_steps,
                  "bags_per_step_per_gpu": B, "ms_per_step": g_ms / g_steps, "collective_ms_per_step": c_ms,
                  "bytes_received_per_rank_per_step": recv_bytes,
                  "achieved_gbs_per_rank": recv_bytes / (c_ms * 1e-3) / 1e9, "nvlink_peer_gbs": NVLINK_PEER_GBS,
                  "what": "forward_bags + parallel.gather_ragged: one all_gather_into_tensor of the padded "
                          "outputs (every rank ends up with every bag's [N, D] output) + one tiny all-gather of "
                          "the sizes; outputs are as large as inputs, so this line is NVLink-bound by design"}

    # ---- per-stage CUDA-event timing inside the library (same workload; bags back to back on one
    # stream so that every interval brackets exactly one kernel running alone) --------------------
    cabi.stage_timing(True)
    for _ in range(min(args.steps, 5)):
        with torch.no_grad():
            enc.forward_bags(bags, outs, lanes=1)
    torch.cuda.synchronize()
    stages = cabi.read_stage_timing()
    cabi.stage_timing(False)
    stage_avg_us = {k: v[0] / v[1] * 1e3 for k, v in stages.items()}
    tot = sum(v[0] for v in stages.values()) or 1.0
    stage_share = {k: v[0] / tot for k, v in stages.items()}
    peaks = measured_peaks()

    def stage_roof(name):
        bound, fl, by = stage_algorithmic(name, N_TOKENS)
        dur_s = stage_avg_us[name] * 1e-6
        if bound == "tensor":
            a = fl / dur_s / 1e12
            return {"bound": "tensor", "achieved": a, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": a / peaks["bf16_tflops"], "flops_per_launch": fl}
        if bound == "hbm":
            a = by / dur_s / 1e9
            return {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": a / peaks["hbm_gbs"], "bytes_per_launch": by}
        return {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None}

    dom = max(stages, key=lambda k: stages[k][0])     # the stage with the largest time share
    roof = {"kernel": dom, **stage_roof(dom), "traffic": None, "avg_launch_us": stage_avg_us[dom],
            "time_share": stage_share[dom], "peak_source": peaks["source"],
            "how": "CUDA events recorded by the library around the kernel on its launching stream, bags back to "
                   "back (lanes=1, every GEMM on all SMs); the interval includes the launch gap.  In the timed "
                   "region the bag-sized GEMMs run under an SM cap (37 SMs with 8 bags in flight; the QKV GEMM as CTA pairs) "
                   "and the attention kernel on a 64-SM persistent grid, so the "
                   "honest whole-step figure is roofline.encoder"}
    tr_file = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes/launch from ncu --set full
    if os.path.isfile(tr_file):
        roof["traffic"] = json.load(open(tr_file)).get(dom)
    fl_bag = encoder_flops(N_TOKENS)
    enc_tflops = fl_bag * reps * B * args.steps / (ms_total * 1e-3) / 1e12 / world * world  # per GPU = whole / N * N
    enc_tflops_per_gpu = fl_bag * reps * B * args.steps / (ms_total * 1e-3) / 1e12
    roof["encoder"] = {"bound": "tensor", "achieved": enc_tflops_per_gpu, "peak": peaks["bf16_tflops"],
                       "unit": "TFLOP/s per GPU", "frac": enc_tflops_per_gpu / peaks["bf16_tflops"],
                       "flops_per_bag": fl_bag,
                       "what": "algorithmic FLOPs of the whole forward / the timed region (binding roofline of the "
                               "encoder is the tensor pipe: 13.7 us per bag at the measured peak)"}
    roof["stages"] = {k: {"us": round(stage_avg_us[k], 2), "share": round(stage_share[k], 4),
                          "bound": stage_roof(k)["bound"],
                          "frac": (round(stage_roof(k)["frac"], 4) if stage_roof(k)["frac"] is not None else None)}
                      for k in stages}

    # ---- e2e: pinned host bags in, pinned host results out, copies inside the timed region ------
    pipe = HostPipeline(enc, n_streams=args.e2e_streams, device=dev)
    hx = [torch.randn(N_TOKENS, DIM).pin_memory() for _ in range(B)]
    hy = [torch.empty(N_TOKENS, DIM).pin_memory() for _ in range(B)]
    hy2 = [torch.empty(N_TOKENS, DIM).pin_memory() for _ in range(B)]
    for _ in range(2):
        pipe.run(hx, hy)
        pipe.run(hx, hy2)
    e2e_steps = max(3, min(args.steps, 10))
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        # steps are streamed: step i+1's uploads start while step i's results are still going back
        # (alternating host result buffers); the clock stops when the LAST result is in host memory
        pipe.run(hx, hy if i % 2 == 0 else hy2, sync=False)
    pipe.wait()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * B * N_TOKENS * e2e_steps / t_e2e
    bytes_step = B * N_TOKENS * DIM * 4
    ceiling = copy_ceiling(dev, world, barrier, dist)
    del hx, hy, hy2, pipe

    # ---- forward + backward of the encoder alone (BASELINE configs[1] names fwd+bwd on 1 GPU) -----
    train = None
    if not args.no_train and world == 1:
        enc_t = RRTEncoder(need_init=True, **ENC_KW).to(dev).train()
        xt = bags[0].clone().requires_grad_()
        gout = torch.randn_like(xt)
        tparams = list(enc_t.parameters())

        def train_step():
            for q in tparams:      # optimizer.zero_grad(set_to_none=True)
                q.grad = None
            xt.grad = None
            y = enc_t(xt)
            y.backward(gout)
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
        n_tr = max(5, min(args.steps, 20))
        l0 = cabi.launch_count()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        th0 = time.perf_counter()
        t0e.record()
        for _ in range(n_tr):
            train_step()
        t1e.record()
        loop_us = (time.perf_counter() - th0) * 1e6 / n_tr   # wall time of the enqueue loop (throttled by the GPU
        torch.cuda.synchronize()                            # once the launch queue fills: NOT the host cost)
        us = t0e.elapsed_time(t1e) * 1e3 / n_tr
        host = []
        for _ in range(7):   # host cost of ONE step enqueued into an empty queue
            torch.cuda.synchronize()
            h0 = time.perf_counter()
            train_step()
            host.append((time.perf_counter() - h0) * 1e6)
        torch.cuda.synchronize()
        host_us = statistics.median(host)
        cabi.stage_timing(True)
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
        tr_stages = {k: round(v[0] / 3 * 1e3, 1) for k, v in cabi.read_stage_timing().items()}
        cabi.stage_timing(False)
        train = {"us_per_bag_fwd_bwd": us, "patches_per_s": N_TOKENS / (us * 1e-6), "drop_out": 0.1,
                 "host_enqueue_us_per_step": host_us, "enqueue_loop_us_per_step": loop_us,
                 "stages_us_per_step": tr_stages,
                 "launches_per_step": (cabi.launch_count() - l0) // n_tr,
                 "what": "RRTEncoder.train() forward (tape + proj dropout) + backward of one N=9000 bag "
                         "through torch.autograd, all parameter gradients, one stream, CUDA events"}
        del enc_t, xt, gout

    # ---- BASELINE configs[2] and configs[4] as extra workloads -------------------------------------
    workloads = {}
    if not args.no_workloads:
        # configs[2]: 8 ragged R50-shaped bags (input_dim 1024) through RRTMIL, bag-parallel, only logits leave
        g8 = torch.Generator().manual_seed(3)
        lens = [int(v) for v in torch.randint(8000, 10001, (8,), generator=g8)]
        gdev = torch.Generator(device=dev).manual_seed(77)       # same bags on every rank
        mil_bags = [torch.randn(n, 1024, device=dev, generator=gdev) for n in lens]
        torch.manual_seed(2021)
        mil = RRTMIL(input_dim=1024, n_classes=2).to(dev).eval()

        def mil_step():
            with torch.no_grad():
                if world > 1:
                    return parallel.classify_bags_parallel(mil, mil_bags)
                return torch.cat([mil(b.unsqueeze(0)) for b in mil_bags])
        for _ in range(2):
            mil_step()
        torch.cuda.synchronize(); barrier()
        m_steps = max(3, min(args.steps, 10))
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(m_steps):
            logits = mil_step()
        m1.record()
        torch.cuda.synchronize()
        m_ms = max_over_ranks(m0.elapsed_time(m1))
        workloads["mil_configs2"] = {
            "value": sum(lens) * m_steps / (m_ms * 1e-3), "unit": UNIT, "ms_per_step": m_ms / m_steps,
            "bags": len(lens), "patches_per_step": sum(lens), "logits_shape": list(logits.shape),
            "d2d_collective_bytes_per_step": (0 if world == 1 else world * math.ceil(8 / world) * 2 * 4),
            "what": "RRTMIL(input_dim=1024): patch_to_emb + RRTEncoder + DAttention pooling + predictor over 8 "
                    "ragged bags N~U[8000,10000], bag i on rank i mod W, one all-gather of the logits"}
        del mil, mil_bags

        # configs[4]: full RRTMIL train step (TCGA-NSCLC-R50 encoder shape), data-parallel: one bag per rank per
        # step, gradients averaged by all-reduces launched from autograd hooks while backward still runs
        torch.manual_seed(2021)
        tm = RRTMIL(input_dim=1024, n_classes=2, epeg_k=21, crmsa_k=5, n_layers=2).to(dev).train()
        opt = Adam(tm.parameters(), lr=2e-4, weight_decay=1e-5)   # main.py:224-233
        red = parallel.GradReducer(parallel.rrtmil_grad_groups(tm)) if world > 1 else None
        tb = torch.randn(1, N_TOKENS, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank))
        label = torch.tensor([rank % 2], device=dev)
        n_coll = 0

        def rrtmil_step():
            nonlocal n_coll
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(tm(tb), label)
            loss.backward()
            if red is not None:
                n_coll = red.finish()
            opt.step()
            return loss
        for _ in range(3):
            rrtmil_step()
        torch.cuda.synchronize(); barrier()
        t_steps = max(5, min(args.steps, 20))
        l0 = cabi.launch_count()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(t_steps):
            loss = rrtmil_step()
        q1.record()
        torch.cuda.synchronize()
        t_ms = max_over_ranks(q0.elapsed_time(q1))
        t_launches = (cabi.launch_count() - l0) // t_steps
        graph_us, graph_err = None, None
        # the same step captured in a CUDA graph (rrt_mil_b200.graph.GraphedTrainStep); with W > 1 the gradient
        # all-reduces launched by the reducer's hooks are part of the graph
        try:
            from rrt_mil_b200.graph import GraphedTrainStep
            loss = float(loss.detach())   # the eager steps' autograd graph must be gone before the capture
            gstep = GraphedTrainStep(tm, opt, N_TOKENS, 1024, reducer=red)
            for _ in range(3):
                gstep(tb, label)
            torch.cuda.synchronize(); barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(t_steps):
                gloss = gstep(tb, label)
            g1.record()
            torch.cuda.synchronize()
            graph_us = max_over_ranks(g0.elapsed_time(g1)) / t_steps * 1e3
            gh = []
            for _ in range(5):
                torch.cuda.synchronize(); barrier()
                h0 = time.perf_counter()
                gstep(tb, label)
                gh.append((time.perf_counter() - h0) * 1e6)
            torch.cuda.synchronize()
            graph_host_us = statistics.median(gh)
            graph_loss = float(gloss.detach())
            del gloss
            gstep.close()     # a graph holding NCCL collectives must not outlive the process group
            del gstep
        except Exception as exc:   # optional leg: never take the bench line down with it
            graph_us, graph_err = None, f"{type(exc).__name__}: {str(exc).splitlines()[0][:200]}"
        host = []
        for _ in range(5):   # host cost of ONE step enqueued into an empty queue (all ranks in step: collectives)
            torch.cuda.synchronize(); barrier()
            h0 = time.perf_counter()
            rrtmil_step()
            host.append((time.perf_counter() - h0) * 1e6)
        torch.cuda.synchronize()
        host_us = statistics.median(host)
        workloads["train_configs4"] = {
            "value": world * N_TOKENS * t_steps / (t_ms * 1e-3), "unit": UNIT, "us_per_step": t_ms / t_steps * 1e3,
            "bags_per_step": world, "host_enqueue_us_per_step": host_us,
            "launches_per_step": t_launches, "collectives_per_step": n_coll,
            "collectives_in_place": (red.last_in_place if red is not None else 0),
            "gradient_bytes": sum(q.numel() for q in tm.parameters()) * 4,
            "final_loss": (loss if isinstance(loss, float) else float(loss.detach())),
            "cuda_graph": ({"error": graph_err} if graph_us is None else {
                "us_per_step": graph_us, "value": world * N_TOKENS / (graph_us * 1e-6),
                "host_enqueue_us_per_step": graph_host_us,
                "final_loss": graph_loss,
                "what": "the same step (zero_grad, forward, CE, backward, Adam) captured once in a CUDA graph and "
                        "replayed; dropout seed and Adam bias corrections come from a 16-byte device buffer "
                        "updated before every replay"}),
            "what": "RRTMIL(input_dim=1024, epeg_k=21, crmsa_k=5): forward + cross entropy + backward + gradient "
                    "all-reduce (3 buckets launched from autograd hooks, overlapping backward) + fused Adam; one "
                    "N=9000 bag per rank per step (= batch-W SGD)"}
        if red is not None:
            red.remove()
        del tm, opt, tb

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"RRTEncoder forward (eval), bags of N={N_TOKENS} D={DIM}, region_num=8 "
                                   "epeg_k=15 crmsa_k=3 n_layers=2 (BASELINE configs[1] shape), fp32 I/O and residual "
                                   "stream, f16 tensor-core operands (10-bit mantissa, as tf32), fp32 accumulate",
                       "bags_per_step_per_gpu": reps * B,
                       "l2_policy": f"inputs larger than L2: {B} distinct bags ({bytes_step / 1e6:.0f} MB in, same "
                                    f"out) walked {reps}x per step",
                       "bags_in_flight_per_gpu": args.lanes,
                       "parallelism": f"bag-parallel x{world}: no collective inside the headline region; the "
                                      "single NCCL gather of the outputs is timed in `gather`",
                       "host_cores_per_rank": cores_per_rank},
            "us_per_bag": ms_total / args.steps / (reps * B) * 1e3,
            "per_rank_ms": per_rank_ms,
            "encoder_tflops": enc_tflops_per_gpu * world,
            "roofline": roof,
            "stages_us_per_launch": {k: round(v, 2) for k, v in stage_avg_us.items()},
            "stages_share": {k: round(v, 4) for k, v in stage_share.items()},
            "gather": gather,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": bytes_step,
                    "d2h_bytes_per_step": bytes_step, "steps": e2e_steps,
                    "gbs_each_way_sum_over_ranks": world * bytes_step * e2e_steps / t_e2e / 1e9,
                    "copy_ceiling": ceiling,
                    "how": f"HostPipeline.run over pinned host bags: H2D stream -> compute stream -> D2H "
                           f"stream over a ring of {args.e2e_streams} device buffers, steps streamed back to "
                           "back, wall clock until the last result is in host memory"},
            "train_step": train,
            "workloads": workloads,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        # non-default kernel variants selected through the environment make the line self-describing
        knobs = {k: v for k, v in os.environ.items() if k.startswith("RRT_")}
        if knobs:
            line["config"]["tuning_knobs"] = knobs
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = time_cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        import gc
        # the line is out; a communicator teardown that hangs must not hold the job (seen once with a CUDA graph that
        # still referenced NCCL work): leave after 45 s at the latest
        sys.stdout.flush()
        threading.Timer(45.0, lambda: os._exit(0)).start()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bags", type=int, default=16, help="distinct bags per GPU")
    ap.add_argument("--reps", type=int, default=4, help="passes over the bags per step (64 bags per step by default)")
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--lanes", type=int, default=8, help="bags in flight per GPU (internal streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the encoder forward+backward measurement")
    ap.add_argument("--no-workloads", action="store_true", help="skip the configs[2] / configs[4] workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
