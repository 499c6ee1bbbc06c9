"""Bag-parallel execution across GPUs: one process per GPU, bags sharded round-robin, no collective
on the data path (bags are independent forwards); ONE gather at the end brings the
per-bag results to every rank (or to rank 0).

The reference has no distributed code at all (single device, batch 1: main.py:103,639).  The host
logic here (sharding, ragged gather) is backend-agnostic: NCCL over NVLink on the B200 box, gloo in
the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_bags: int, world_size: int, rank: int) -> List[int]:
    """Bags of rank ``rank``: i with i % world_size == rank (bag i -> GPU i mod W)."""
    return list(range(rank, n_bags, world_size))


def _default_device(group=None) -> torch.device:
    """Device of the collectives' buffers for a rank that holds no tensor to take it from: the current CUDA
    device under NCCL (every rank must hand CUDA tensors to the same collective), the CPU otherwise."""
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_ragged(local: Sequence[torch.Tensor], n_bags: int, group=None,
                  dst: Optional[int] = None, device: Optional[torch.device] = None,
                  dtype: torch.dtype = torch.float32) -> Optional[List[torch.Tensor]]:
    """All ranks hold the outputs of their ``shard_indices`` bags (``[N_i, D]`` each, ragged N_i).
    Returns the ``n_bags`` outputs in bag order on every rank (``dst=None``) or on ``dst`` only.

    One collective for the sizes (tiny) and ONE for the payload: the local bags are packed into a
    single padded ``[rows_max, D]`` buffer and all-gathered; views are then cut per bag.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_indices(n_bags, world, rank)
    if len(mine) != len(local):
        raise ValueError(f"rank {rank} holds {len(local)} bags, expected {len(mine)}")
    # a rank without bags (world_size > n_bags) must still join the collectives with tensors on the same
    # kind of device as everybody else: `device` from the caller, else the backend's default
    if local:
        device, dtype = local[0].device, local[0].dtype
    elif device is None:
        device = _default_device(group)
    width = local[0].shape[1] if local else 0
    per_rank = (n_bags + world - 1) // world
    # sizes: [world, per_rank + 1] (last column = feature width, so empty ranks learn it too)
    sizes = torch.zeros(per_rank + 1, dtype=torch.int64, device=device)
    for j, t in enumerate(local):
        sizes[j] = t.shape[0]
    sizes[per_rank] = width
    all_sizes = torch.empty(world * (per_rank + 1), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, per_rank + 1).cpu()
    width = int(all_sizes[:, per_rank].max())
    rows = all_sizes[:, :per_rank].sum(1)
    rows_max = int(rows.max())
    send = torch.zeros(rows_max, width, dtype=dtype, device=device)
    off = 0
    for t in local:
        send[off:off + t.shape[0]] = t
        off += t.shape[0]
    recv = torch.empty(world * rows_max, width, dtype=dtype, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)  # the single payload collective
    recv = recv.view(world, rows_max, width)
    if dst is not None and rank != dst:
        return None
    out: List[Optional[torch.Tensor]] = [None] * n_bags
    for r in range(world):
        off = 0
        for j, i in enumerate(shard_indices(n_bags, world, r)):
            n = int(all_sizes[r, j])
            out[i] = recv[r, off:off + n]
            off += n
    return out  # type: ignore[return-value]


@torch.no_grad()
def encode_bags_parallel(encoder, bags: Sequence[torch.Tensor], group=None, gather: bool = True,
                         dst: Optional[int] = None):
    """Every rank receives the same list of bags (host or device tensors), encodes its shard on its
    own GPU with ``encoder.forward_bags`` and, if ``gather``, returns all outputs in bag order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_indices(len(bags), world, rank)
    dev = next(encoder.parameters()).device
    local_in = [bags[i].to(dev, non_blocking=True).contiguous() for i in mine]
    local_out = encoder.forward_bags(local_in) if local_in else []
    if not gather:
        return local_out
    return gather_ragged(local_out, len(bags), group=group, dst=dst, device=dev)


@torch.no_grad()
def classify_bags_parallel(model, bags: Sequence[torch.Tensor], group=None, dst: Optional[int] = None):
    """BASELINE configs[2]: a pooling head follows the encoder, so only the ``[n_classes]`` logits of each bag
    leave a GPU.  Every rank receives the same list of bags ``[N_i, C_in]``, runs ``model`` (e.g. ``RRTMIL``) on
    its shard and ONE all-gather of ``ceil(n_bags / world) * n_classes`` floats returns the ``[n_bags, n_classes]``
    logits in bag order on every rank (on ``dst`` only if given)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_indices(len(bags), world, rank)
    dev = next(model.parameters()).device
    local = [model(bags[i].to(dev, non_blocking=True).unsqueeze(0)).reshape(-1) for i in mine]
    per_rank = (len(bags) + world - 1) // world
    ncls = torch.tensor([local[0].numel() if local else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ncls, op=dist.ReduceOp.MAX, group=group)      # ranks without a bag learn the width
    ncls = int(ncls.item())
    send = torch.zeros(per_rank, ncls, dtype=torch.float32, device=dev)
    for j, t in enumerate(local):
        send[j] = t.float()
    recv = torch.empty(world * per_rank, ncls, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)           # the single payload collective
    if dst is not None and rank != dst:
        return None
    recv = recv.view(world, per_rank, ncls)
    out = torch.empty(len(bags), ncls, dtype=torch.float32, device=dev)
    for r in range(world):
        for j, i in enumerate(shard_indices(len(bags), world, r)):
            out[i] = recv[r, j]
    return out


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Data-parallel training step glue (training): every rank has run forward +
    backward on ITS bag; sum (mean) the gradients of ``params`` over the ranks.  Gradients are packed
    into flat buckets of at most ``bucket_bytes`` (the encoder's 8.4 MB of gradients is ONE bucket, i.e.
    one all-reduce per step), reduced asynchronously, and unpacked.  Parameters without a gradient on
    some rank contribute zeros (every rank must call with the same parameter list).  Returns the number
    of collectives issued.  Equals batch-``world`` SGD: compare with the serial average in the tests."""
    world = dist.get_world_size(group)
    plist = [p for p in params]
    if world == 1 or not plist:
        return 0
    for p in plist:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    buckets, cur, size = [], [], 0
    for p in plist:
        nb = p.grad.numel() * p.grad.element_size()
        if cur and size + nb > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nb
    if cur:
        buckets.append(cur)
    work = []
    for b in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in b])
        work.append((b, flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)))
    for b, flat, w in work:
        w.wait()
        if average:
            flat.div_(world)
        off = 0
        for p in b:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return len(buckets)


class GradReducer:
    """Data-parallel gradient averaging OVERLAPPED with the backward pass (BASELINE configs[4]; the reference has
    no distributed training at all, main.py:103).

    ``groups`` is a list of parameter lists in the order their gradients become available during backward (for
    ``RRTMIL``: pooling head + predictor, encoder, ``patch_to_emb``).  A post-accumulate hook on every parameter
    counts the group down; when the last gradient of a group has been written its all-reduce is launched
    asynchronously -- IN PLACE on the buffer the group's gradients are views of (the library's backward Functions
    allocate the gradients of one Function from one buffer), or on a packed copy when they are not -- so the
    collective of the head runs under the encoder's backward kernels, the encoder's under ``patch_to_emb``'s.  ``finish()`` (call it after ``loss.backward()``,
    before the optimizer step) waits for the collectives, averages and scatters the buckets back into ``.grad``.
    Parameters that received no gradient in a step contribute zeros, so every rank reduces the same layout.
    """

    def __init__(self, groups, group=None, average: bool = True):
        self.pg = group
        self.average = average
        self.world = dist.get_world_size(group)
        self.groups = [[p for p in g if p.requires_grad] for g in groups]
        self.groups = [g for g in self.groups if g]
        self._pending = [len(g) for g in self.groups]
        self._work = []          # (group index, flat, handle, in_place, divide_after)
        self.last_in_place = 0
        self._launched = [False] * len(self.groups)
        self._hooks = []
        if self.world > 1:
            for gi, g in enumerate(self.groups):
                for p in g:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(gi)))

    def _make_hook(self, gi):
        def hook(_param):
            self._pending[gi] -= 1
            if self._pending[gi] == 0:
                self._launch(gi)
        return hook

    @staticmethod
    def _shared_span(grads):
        """One flat tensor over the storage range the gradients of a group occupy, when they are contiguous
        views of ONE buffer (the library's backward Functions allocate theirs that way); else None."""
        st = grads[0].untyped_storage()
        base = st.data_ptr()
        lo, hi = None, 0
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base:
                return None
            a = g.storage_offset()
            lo = a if lo is None else min(lo, a)
            hi = max(hi, a + g.numel())
        if (hi - lo) > 2 * sum(g.numel() for g in grads) + 4096:   # mostly foreign data in between: do not touch it
            return None
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, lo, (hi - lo,))

    def _launch(self, gi):
        g = self.groups[gi]
        grads = [p.grad for p in g]
        flat = self._shared_span(grads) if all(x is not None for x in grads) else None
        in_place = flat is not None
        if not in_place:
            flat = torch.cat([(x if x is not None else torch.zeros_like(p)).reshape(-1) for x, p in zip(grads, g)])
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide in finish()
        avg_op = self.average and dist.get_backend(self.pg) == "nccl"
        op = dist.ReduceOp.AVG if avg_op else dist.ReduceOp.SUM
        self._work.append((gi, flat, dist.all_reduce(flat, op=op, group=self.pg, async_op=True), in_place,
                           self.average and not avg_op))
        self._launched[gi] = True

    def finish(self) -> int:
        """Waits for the launched collectives (launching those of groups whose hooks did not all fire) and, for
        groups that had to be packed, writes the averaged gradients back; returns the number of collectives of
        this step and re-arms the hooks.  Groups whose gradients are views of one buffer were reduced in place:
        nothing is copied."""
        if self.world == 1:
            return 0
        for gi in range(len(self.groups)):
            if not self._launched[gi]:
                self._launch(gi)
        n = len(self._work)
        self.last_in_place = sum(1 for w in self._work if w[3])   # buckets reduced without packing (introspection)
        for gi, flat, handle, in_place, divide in self._work:
            handle.wait()
            if divide:
                flat.div_(self.world)
            if in_place:
                continue
            off = 0
            for p in self.groups[gi]:
                k = p.numel()
                if p.grad is None:
                    p.grad = flat[off:off + k].view_as(p).clone()
                else:
                    p.grad.copy_(flat[off:off + k].view_as(p.grad))
                off += k
        self._work = []
        self._pending = [len(g) for g in self.groups]
        self._launched = [False] * len(self.groups)
        return n

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def rrtmil_grad_groups(model):
    """The three groups of an ``RRTMIL`` in backward order: head (pooling + predictor), encoder, patch_to_emb."""
    head = list(model.pool_fn.parameters()) + list(model.predictor.parameters())
    return [head, list(model.online_encoder.parameters()), list(model.patch_to_emb.parameters())]
