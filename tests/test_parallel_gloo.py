"""Bag-parallel host logic over world_size-2 gloo on CPU: sharding and the single ragged gather.
(The encoder itself is CUDA-only; here the per-rank "encoder" is a deterministic CPU stand-in so
that the collective plumbing is what is tested.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrt_mil_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StandIn(torch.nn.Module):
    """forward_bags(x) = 2*x + rank-independent bias: any rank must produce the same result."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    def forward_bags(self, bags):
        return [2.0 * b + 1.0 for b in bags]


def _worker(rank, world, port, lens, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        bags = [torch.randn(n, 8, generator=g) for n in lens]
        outs = parallel.encode_bags_parallel(_StandIn(), bags)
        ok = len(outs) == len(bags) and all(torch.equal(o, 2.0 * b + 1.0) for o, b in zip(outs, bags))
        only0 = parallel.encode_bags_parallel(_StandIn(), bags, dst=0)
        ok = ok and ((only0 is None) == (rank != 0))
        local = parallel.encode_bags_parallel(_StandIn(), bags, gather=False)
        ok = ok and len(local) == len(parallel.shard_indices(len(bags), world, rank))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lens", [[5, 3, 9, 1, 7], [4], [6, 6], [2, 0 + 1, 3]])
def test_sharded_encode_and_single_gather_world2(lens):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, lens, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_shard_indices_partition_the_bags():
    for n in (0, 1, 7, 8, 9):
        for w in (1, 2, 4, 8):
            seen = sorted(i for r in range(w) for i in parallel.shard_indices(n, w, r))
            assert seen == list(range(n))


def _grad_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
        frozen = torch.nn.Parameter(torch.ones(3))            # never receives a gradient on any rank
        params = list(model.parameters()) + [frozen]
        g = torch.Generator().manual_seed(1)
        bags = [torch.randn(10 + 3 * i, 16, generator=g) for i in range(world)]   # one bag per rank
        model(bags[rank]).square().mean().backward()
        n_coll = parallel.allreduce_gradients(params, bucket_bytes=1024)           # several buckets
        # serial oracle: average of the per-bag gradients
        ref = [torch.zeros_like(p) for p in params]
        for b in bags:
            m2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
            m2.load_state_dict(model.state_dict())
            m2(b).square().mean().backward()
            for r, p in zip(ref, m2.parameters()):
                r += p.grad / world
        ok = n_coll >= 2 and all(torch.allclose(p.grad, r, rtol=1e-6, atol=1e-7) for p, r in zip(params, ref))
        ok = ok and float(frozen.grad.abs().sum()) == 0.0
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_equals_serial_average_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_grad_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


class _StandInMil(torch.nn.Module):
    """logits = (mean over tokens) @ W: deterministic, rank independent."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.w = torch.nn.Parameter(torch.randn(8, 3, generator=g))

    def forward(self, x):
        return x[0].mean(0, keepdim=True) @ self.w


def _logit_worker(rank, world, port, lens, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        bags = [torch.randn(n, 8, generator=g) for n in lens]
        m = _StandInMil()
        out = parallel.classify_bags_parallel(m, bags)
        ref = torch.cat([m(b.unsqueeze(0)) for b in bags]).detach()
        ok = out.shape == (len(bags), 3) and torch.allclose(out, ref, atol=1e-6)
        only0 = parallel.classify_bags_parallel(m, bags, dst=0)
        results[rank] = bool(ok and ((only0 is None) == (rank != 0)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lens", [[5, 3, 9, 1, 7], [4]])
def test_logits_gather_world2(lens):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_logit_worker, args=(world, port, lens, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def _reducer_worker(rank, world, port, results):
    """GradReducer: hooks launch one asynchronous all-reduce per group while backward is still running; the
    result must equal the serial average, step after step (hooks re-arm), also when a group gets no gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        front, back = torch.nn.Linear(16, 32), torch.nn.Linear(32, 4)
        unused = torch.nn.Linear(4, 4)                         # a group that never takes part in the loss
        groups = [list(back.parameters()), list(front.parameters()), list(unused.parameters())]
        red = parallel.GradReducer(groups)
        g = torch.Generator().manual_seed(1)
        ok = True
        for step in range(3):
            bags = [torch.randn(10 + 3 * i + step, 16, generator=g) for i in range(world)]
            for q in [p for grp in groups for p in grp]:
                q.grad = None
            back(torch.tanh(front(bags[rank]))).square().mean().backward()
            n = red.finish()
            ok = ok and n == 3
            ok = ok and all(q.grad is not None and float(q.grad.abs().sum()) == 0.0 for q in unused.parameters())
            mine = [q.grad.clone() for q in list(back.parameters()) + list(front.parameters())]
            ref = [torch.zeros_like(q) for q in mine]
            for b in bags:
                for q in [p for grp in groups for p in grp]:
                    q.grad = None
                back(torch.tanh(front(b))).square().mean().backward()
                for r, q in zip(ref, list(back.parameters()) + list(front.parameters())):
                    r += q.grad / world
            red._pending = [len(grp) for grp in red.groups]     # the oracle passes fired the hooks too: re-arm
            for w in red._work:
                w[2].wait()
            red._work, red._launched = [], [False] * len(red.groups)
            ok = ok and all(torch.allclose(a, b, atol=1e-6) for a, b in zip(mine, ref))
        red.remove()
        # gradients that are views of ONE buffer (what the library's backward Functions hand out) are reduced in
        # place: same storage afterwards, padding between the views untouched by anything but the average
        ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
        flat = torch.full((64 + 7,), float(rank + 1))
        ps[0].grad, ps[1].grad = flat[:15].view(5, 3), flat[64:71]
        red2 = parallel.GradReducer([ps])
        ok = ok and red2._shared_span([q.grad for q in ps]) is not None
        ok = ok and red2.finish() == 1
        ok = ok and ps[0].grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
        ok = ok and bool(torch.allclose(flat, torch.full_like(flat, (1 + world) / 2.0)))
        red2.remove()
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_reducer_overlapped_allreduce_equals_serial_average_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_reducer_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def _empty_rank_worker(rank, world, port, results):
    """world_size > n_bags: the rank without a bag must still join both collectives."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bags = [torch.arange(12.0).view(3, 4)]
        outs = parallel.encode_bags_parallel(_StandIn(), bags)
        results[rank] = len(outs) == 1 and torch.equal(outs[0], 2.0 * bags[0] + 1.0)
    finally:
        dist.destroy_process_group()


def test_gather_with_a_rank_that_holds_no_bag_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_empty_rank_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
