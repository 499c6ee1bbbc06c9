"""CUDA-graph replay of the encoder forward for fixed bag geometries (measured on a B200, round 2:
N=512 one bag 63.2 -> 43.3 us, N=2000 55.1 -> 51.3 us, N=9000 94.8 -> 92.5 us; replay == eager bit for bit).

A bag of a few hundred patches is launch-latency bound (10 kernels per bag: N=512 takes 77 us on a B200, of
which the kernels themselves are a fraction).  The library call is capture-safe by construction -- no host
synchronisation, caller-owned workspace, internal lanes forked from / joined into the caller's stream with
events -- so the whole kernel chain of ``forward_bags`` for a FIXED list of bag lengths can be recorded once
into a ``torch.cuda.CUDAGraph`` and replayed with a single launch.

The graph bakes in the parameter pointers and, in eval mode, the fp16 weight shadows that existed at capture
time; ``__call__`` re-captures when a parameter's storage or version counter changed since then (same test
as ``RRTEncoder._weights``), so results always equal the eager path's.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import cabi


class GraphedForward:
    """``g = GraphedForward(encoder, [N_0, N_1, ...]); outs = g(bags)`` -- inference only (eval mode).

    ``bags[i]`` must be float32 CUDA ``[N_i, D]``; they are copied into the graph's static input buffers
    (device-to-device, on the current stream) and the returned tensors are the graph's static outputs:
    they are overwritten by the next call.  ``g.inputs`` / ``g.outputs`` are exposed so that a producer can
    write the static inputs directly and call ``g.replay()``."""

    def __init__(self, encoder, lengths: Sequence[int], lanes: int = cabi.RRT_MAX_LANES, device=None):
        if encoder.training:
            raise RuntimeError("GraphedForward is inference-only: call .eval() first")
        if not lengths or any(int(n) < 1 for n in lengths):
            raise ValueError("lengths must be a non-empty list of positive bag sizes")
        self.enc = encoder
        self.lengths = [int(n) for n in lengths]
        self.lanes = max(1, min(int(lanes), cabi.RRT_MAX_LANES))
        self.device = (torch.device(device) if device is not None
                       else next(encoder.parameters()).device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphedForward needs the encoder on a CUDA device; there is no CPU fallback")
        D = encoder.final_dim
        self.inputs: List[torch.Tensor] = [torch.zeros(n, D, device=self.device) for n in self.lengths]
        self.outputs: List[torch.Tensor] = [torch.empty(n, D, device=self.device) for n in self.lengths]
        self._graph = None
        self._key = None
        self.captures = 0

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.enc._named_param_cache()[1])

    @torch.no_grad()
    def _capture(self) -> None:
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # one-time work (function attributes, lane streams, TMA descriptor cache, shadows)
                self.enc.forward_bags(self.inputs, self.outputs, lanes=self.lanes)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            # the workspace is allocated inside the capture: it lives in the graph's private pool
            self.enc.forward_bags(self.inputs, self.outputs, lanes=self.lanes)
        self._graph, self._key = graph, self._weights_key()
        self.captures += 1

    def replay(self) -> List[torch.Tensor]:
        if self.enc.training:
            raise RuntimeError("GraphedForward is inference-only: call .eval() first")
        if self._graph is None or self._key != self._weights_key():
            self._capture()
        self._graph.replay()
        return self.outputs

    @torch.no_grad()
    def __call__(self, bags: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        if len(bags) != len(self.lengths):
            raise ValueError(f"expected {len(self.lengths)} bags, got {len(bags)}")
        for x, buf in zip(bags, self.inputs):
            if x.shape != buf.shape or x.dtype != torch.float32 or x.device != self.device:
                raise ValueError(f"expected a float32 {tuple(buf.shape)} bag on {self.device}, "
                                 f"got {x.dtype} {tuple(x.shape)} on {x.device}")
        for x, buf in zip(bags, self.inputs):
            if x.data_ptr() != buf.data_ptr():
                buf.copy_(x, non_blocking=True)
        return self.replay()


class GraphedTrainStep:
    """One training step of a fixed-geometry bag -- ``zero_grad``, forward, loss, ``backward``, optimizer step --
    recorded once into a ``torch.cuda.CUDAGraph`` and replayed with a single launch (the eager
    step is host-bound, ~70 kernel launches + autograd + Python per step against ~0.9 ms of GPU work).

    What changes from step to step and would otherwise be frozen into the graph lives in a 16-byte DEVICE buffer
    (``rrt_set_step_state``): the dropout seed (added to every mask's seed when the mask is evaluated, forward and
    backward alike) and Adam's bias corrections.  ``__call__(bag, label)`` copies the inputs into the static
    buffers, uploads the step state and replays; it returns the static loss tensor (overwritten by the next call).

    Requirements: ``model(bag) -> logits [1, C]`` built from this package's modules (their C calls are capture
    safe: no host synchronisation, caller-owned workspaces), ``optimizer`` = ``rrt_mil_b200.optim.Adam`` with one
    parameter group, one CUDA device per process; ``reducer`` = an optional ``parallel.GradReducer`` (data-parallel
    step: its NCCL all-reduces become part of the graph).  The optimizer's Python-side ``state[p]["step"]`` is
    brought up to date by ``sync_optimizer_state()``.

    Drop every tensor that still hangs on to an EARLIER eager step's autograd graph (e.g. its loss) before the
    first call: such a graph keeps the parameters' AccumulateGrad nodes alive on the stream that step ran on
    (usually the legacy default stream), and torch then has to synchronise that stream with the capturing one,
    which CUDA forbids ("would make the legacy stream depend on a capturing blocking stream")."""

    _RING = 64

    def __init__(self, model, optimizer, n_patches: int, in_dim: int, loss_fn=None, seed: int = 0, device=None,
                 reducer=None):
        from .optim import Adam
        if not isinstance(optimizer, Adam) or len(optimizer.param_groups) != 1:
            raise TypeError("GraphedTrainStep needs rrt_mil_b200.optim.Adam with a single parameter group")
        if any(getattr(mod, "drop_path_rate", 0.0) > 0.0 for mod in model.modules()):
            raise NotImplementedError("GraphedTrainStep: drop_path > 0 changes the kernel sequence from step to step "
                                      "(a dropped block is skipped); train those models eagerly")
        self.model, self.opt = model, optimizer
        # data-parallel: a parallel.GradReducer whose hooks launch the gradient all-reduces during backward;
        # finish() is part of the recorded step (NCCL collectives are capturable), so a replay is the whole
        # batch-W step of this rank
        self.reducer = reducer
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device; there is no CPU fallback")
        self.loss_fn = loss_fn if loss_fn is not None else torch.nn.functional.cross_entropy
        self.bag = torch.zeros(1, int(n_patches), int(in_dim), device=self.device)
        self.label = torch.zeros(1, dtype=torch.long, device=self.device)
        self.loss = None
        self._state = torch.zeros(2, dtype=torch.int64, device=self.device)        # {u64 seed | f32 bc1, f32 bc2}
        self._host = torch.zeros(self._RING, 2, dtype=torch.int64).pin_memory()
        self._host_f32 = self._host.view(torch.float32)                            # [RING, 4]: floats 2, 3 of a row
        self._events = [None] * self._RING
        self._seed = int(seed)
        self._graph = None
        self.capture_error_mode = "thread_local"
        st = [optimizer.state[p].get("step", 0) for p in optimizer.param_groups[0]["params"] if p in optimizer.state]
        self.t = int(max(st)) if st else 0     # optimizer steps taken so far
        self.replays = 0

    # -- per-step device state ------------------------------------------------------------------------------
    def _upload_state(self, t: int) -> None:
        k = t % self._RING
        if self._events[k] is not None:
            self._events[k].synchronize()       # the copy that last read this pinned slot (64 steps ago)
        b1, b2 = self.opt.param_groups[0]["betas"]
        self._host[k, 0] = (self._seed + t * 0x9E3779B97F4A7C15) % (1 << 63)
        self._host_f32[k, 2] = 1.0 - b1 ** t
        self._host_f32[k, 3] = 1.0 / (1.0 - b2 ** t) ** 0.5
        self._state.copy_(self._host[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events[k] = ev

    def _step_eager(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.bag), self.label)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss

    def _capture(self) -> None:
        lib = cabi.lib()
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep captures a TRAINING step: call .train() first")
        cabi.check(lib.rrt_set_step_state(self._state.data_ptr()), "rrt_set_step_state")
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):      # optimizer state, function attributes, TMA descriptor cache, allocator warm-up
                    self.t += 1
                    self._upload_state(self.t)
                    self._step_eager()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            import gc
            gc.collect()            # dead autograd graphs of earlier eager steps release their AccumulateGrad nodes
            graph = torch.cuda.CUDAGraph()
            # thread_local: torch runs the backward on its autograd thread and other host threads of the process
            # (pinned-memory bookkeeping, NCCL watchdogs, monitoring) may call CUDA while the step is being recorded
            with torch.cuda.graph(graph, stream=side, capture_error_mode=self.capture_error_mode):
                self.loss = self._step_eager()
            self._graph = graph
        finally:
            cabi.check(lib.rrt_set_step_state(None), "rrt_set_step_state")
        # the capture itself did not run the kernels; Adam.step counted one step on the Python side
        self.sync_optimizer_state()

    def close(self) -> None:
        """Drops the recorded graph (and its private memory pool).  A graph that contains NCCL collectives must be
        gone before ``torch.distributed.destroy_process_group()``, or the teardown of the communicator hangs."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph = None
            self.loss = None
        self.sync_optimizer_state()

    def sync_optimizer_state(self) -> None:
        for p in self.opt.param_groups[0]["params"]:
            if p in self.opt.state and self.opt.state[p]:
                self.opt.state[p]["step"] = self.t

    def __call__(self, bag: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        if bag.dim() == 2:
            bag = bag.unsqueeze(0)
        if bag.shape != self.bag.shape:
            raise ValueError(f"expected a bag of shape {tuple(self.bag.shape)}, got {tuple(bag.shape)}")
        if self._graph is None:
            self._capture()
        if bag.data_ptr() != self.bag.data_ptr():
            self.bag.copy_(bag, non_blocking=True)
        self.label.copy_(label.reshape(1), non_blocking=True)
        self.t += 1
        self._upload_state(self.t)
        self._graph.replay()
        self.replays += 1
        # the graph wrote the parameters through raw pointers: caches keyed on version counters must see it
        for p in self.opt.param_groups[0]["params"]:
            torch.autograd.graph.increment_version(p)
        return self.loss
