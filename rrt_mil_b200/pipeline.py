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
    def __init__(self, encoder, n_streams: int = 3, device=None):
        self.enc = encoder
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, n_streams))]
        self._x = [None] * len(self.streams)
        self._y = [None] * len(self.streams)

    def _slot(self, i: int, n: int, d: int):
        if self._x[i] is None or self._x[i].shape[0] < n or self._x[i].shape[1] != d:
            self._x[i] = torch.empty(n, d, device=self.device)
            self._y[i] = torch.empty(n, d, device=self.device)
        return self._x[i][:n], self._y[i][:n]

    @torch.no_grad()
    def run(self, bags_host: Sequence[torch.Tensor], outs_host: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """``outs_host[i] = encoder(bags_host[i])`` for host tensors ``[N_i, D]`` (pinned memory makes
        the copies asynchronous).  Returns after every result has landed in host memory."""
        if len(bags_host) != len(outs_host):
            raise ValueError("bags_host and outs_host differ in length")
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for i, (hx, hy) in enumerate(zip(bags_host, outs_host)):
            k = i % len(self.streams)
            s = self.streams[k]
            with torch.cuda.stream(s):
                x, y = self._slot(k, hx.shape[0], hx.shape[1])
                x.copy_(hx, non_blocking=True)
                self.enc.forward_bags([x], [y])
                hy.copy_(y, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return list(outs_host)
