"""Bag loader of the training / inference harness: the reference reads one
``<root>/pt/<slide>.pt`` feature tensor per step with ``torch.load`` inside a ``Dataset`` and moves it
with a blocking ``.to(device)`` (dataloader.py:162-203, main.py:425-434).  Here a background thread reads
ahead into PINNED host buffers, so that the host->device copy of bag i+1 (asynchronous from pinned
memory) and the file read of bag i+2 overlap the kernels of bag i.

    for bag, label in PinnedBagLoader(names, labels, root, device="cuda", prefetch=2):
        logits = model(bag)          # bag: [1, N, C] float32 on the device

Iteration order is the order of ``names`` (shuffle the lists, as the reference's sampler does).  With
``persistence=True`` every file is read once and kept (pinned) in host memory, like the reference's
``persistence`` flag.  Host-side logic only; no kernel is involved.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

import torch


class PinnedBagLoader:
    def __init__(self, names: Sequence[str], labels: Sequence[int], root: str, device=None, prefetch: int = 2,
                 persistence: bool = False, subdir: str = "pt", pin: Optional[bool] = None):
        if len(names) != len(labels):
            raise ValueError("names and labels differ in length")
        self.names, self.labels = list(names), [int(v) for v in labels]
        self.dir = os.path.join(root, subdir)
        self.device = torch.device(device) if device is not None else None
        self.prefetch = max(1, int(prefetch))
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self.persistence = persistence
        self._cache: List[Optional[torch.Tensor]] = [None] * len(self.names)

    def __len__(self) -> int:
        return len(self.names)

    def _read(self, i: int) -> torch.Tensor:
        if self._cache[i] is not None:
            return self._cache[i]
        t = torch.load(os.path.join(self.dir, self.names[i] + ".pt"), map_location="cpu")
        if not isinstance(t, torch.Tensor) or t.dim() != 2:
            raise ValueError(f"{self.names[i]}.pt: expected a [N, C] feature tensor")
        t = t.float().contiguous()
        if self.pin:
            t = t.pin_memory()
        if self.persistence:
            self._cache[i] = t
        return t

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, int]]:
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def producer():
            try:
                for i in range(len(self.names)):
                    if stop.is_set():
                        return
                    q.put((self._read(i), self.labels[i]))
                q.put(None)
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                host, label = item
                if self.device is not None and self.device.type == "cuda":
                    # asynchronous from pinned memory; torch's pinned-memory allocator records the copy's
                    # stream event, so the host block is not reused before the copy has drained
                    dev = host.to(self.device, non_blocking=True)
                    yield dev.unsqueeze(0), label
                else:
                    yield host.unsqueeze(0), label
        finally:
            stop.set()
            while not q.empty():   # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
