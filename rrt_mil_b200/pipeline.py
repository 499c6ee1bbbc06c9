"""Host-buffer front end: bags that live in (pinned) host memory go through the encoder with the
host->device copy, the kernels and the device->host copy of consecutive bags overlapped on a few
CUDA streams.  This is the end-to-end path ``bench.py`` reports as ``e2e``.

The reference moves one bag at a time with a blocking ``.to(device)`` (main.py:434); here the PCIe
copies (18.4 MB each way for a 9000 x 512 bag) dominate the per-bag time, so hiding them behind the
previous bag's kernels is what the host side can contribute.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class HostPipeline:
    """``run(bags_host, outs_host)``: host bags through the encoder with one stream per DMA direction.

    A dedicated host->device stream, a compute stream and a dedicated device->host stream are chained
    with events over a ring of ``n_streams`` device buffer pairs, so both copy engines always have the
    next transfer queued (measured on the B200 box, tools/pcie_probe.py: 54.7 GB/s H2D alone, 56.0 GB/s
    D2H alone, 47.5 GB/s each way when both run -- the bound of this path)."""

    def __init__(self, encoder, n_streams: int = 3, device=None):
        self.enc = encoder
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.depth = max(2, int(n_streams))
        self.h2d = torch.cuda.Stream(device=self.device)
        self.compute = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        ev = lambda: [torch.cuda.Event() for _ in range(self.depth)]  # noqa: E731
        self._copied, self._computed, self._returned = ev(), ev(), ev()
        self._x = [None] * self.depth
        self._y = [None] * self.depth
        self._n = 0   # bags submitted so far (ring position; carries over between run() calls)

    def _slot(self, i: int, n: int, d: int):
        if self._x[i] is None or self._x[i].shape[0] < n or self._x[i].shape[1] != d:
            if self._x[i] is not None:   # growing a slot: its old buffers may still be in flight on the
                torch.cuda.synchronize(self.device)  # copy streams; drain before they go back to the pool
            self._x[i] = torch.empty(n, d, device=self.device)
            self._y[i] = torch.empty(n, d, device=self.device)
        return self._x[i][:n], self._y[i][:n]

    @torch.no_grad()
    def run(self, bags_host: Sequence[torch.Tensor], outs_host: Sequence[torch.Tensor],
            sync: bool = True) -> List[torch.Tensor]:
        """``outs_host[i] = encoder(bags_host[i])`` for host tensors ``[N_i, D]`` (pinned memory makes
        the copies asynchronous).  Returns after every result has landed in host memory; with
        ``sync=False`` it returns once everything is enqueued, so that the next call's first uploads
        overlap this call's last downloads (a stream of batches pays the pipeline fill / drain -- one
        un-overlapped copy each way -- once instead of once per call); ``wait()`` then blocks until the
        results of every call so far are in host memory."""
        if len(bags_host) != len(outs_host):
            raise ValueError("bags_host and outs_host differ in length")
        cur = torch.cuda.current_stream(self.device)
        for s in (self.h2d, self.compute, self.d2h):
            s.wait_stream(cur)
        for hx, hy in zip(bags_host, outs_host):
            k = self._n % self.depth
            reuse = self._n >= self.depth
            self._n += 1
            # allocate (first use / larger bag) on the compute stream's side of the allocator
            with torch.cuda.stream(self.compute):
                x, y = self._slot(k, hx.shape[0], hx.shape[1])
            with torch.cuda.stream(self.h2d):
                if reuse:
                    self.h2d.wait_event(self._computed[k])   # the bag that used x[k] has been encoded
                x.copy_(hx, non_blocking=True)
                self._copied[k].record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self._copied[k])
                if reuse:
                    self.compute.wait_event(self._returned[k])  # y[k] has left for the host
                self.enc.forward_bags([x], [y])
                self._computed[k].record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self._computed[k])
                hy.copy_(y, non_blocking=True)
                self._returned[k].record(self.d2h)
        if sync:
            self.wait()
        return list(outs_host)

    def wait(self) -> None:
        """Block until every result submitted so far has landed in host memory."""
        self.d2h.synchronize()
        torch.cuda.current_stream(self.device).wait_stream(self.compute)
