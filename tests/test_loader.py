"""Host-side bag loader: order, labels, read-ahead, persistence, error surfacing.  CPU only."""
import os

import pytest
import torch

from rrt_mil_b200.loader import PinnedBagLoader


def _make(tmp_path, n=7):
    os.makedirs(tmp_path / "pt")
    g = torch.Generator().manual_seed(0)
    names, bags = [], []
    for i in range(n):
        t = torch.randn(10 + 3 * i, 16, generator=g)
        torch.save(t, tmp_path / "pt" / f"slide_{i}.pt")
        names.append(f"slide_{i}")
        bags.append(t)
    return names, bags


@pytest.mark.parametrize("prefetch,persistence", [(1, False), (3, False), (2, True)])
def test_loader_yields_bags_in_order(tmp_path, prefetch, persistence):
    names, bags = _make(tmp_path)
    labels = [i % 2 for i in range(len(names))]
    ld = PinnedBagLoader(names, labels, str(tmp_path), prefetch=prefetch, persistence=persistence, pin=False)
    for _ in range(2):   # two epochs (the second one served from memory with persistence)
        got = list(ld)
        assert len(got) == len(ld) == len(names)
        for (b, y), ref, lab in zip(got, bags, labels):
            assert b.shape == (1,) + tuple(ref.shape) and torch.equal(b[0], ref) and y == lab
    if persistence:
        os.remove(tmp_path / "pt" / "slide_0.pt")
        assert torch.equal(next(iter(ld))[0][0], bags[0])


def test_loader_surfaces_read_errors_and_early_exit(tmp_path):
    names, bags = _make(tmp_path, 4)
    ld = PinnedBagLoader(names + ["missing"], [0] * 5, str(tmp_path), pin=False)
    with pytest.raises(FileNotFoundError):
        list(ld)
    it = iter(PinnedBagLoader(names, [0] * 4, str(tmp_path), prefetch=1, pin=False))
    next(it)
    it.close()           # abandoning the iterator must not leave the reader thread blocked
    torch.save({"not": "a tensor"}, tmp_path / "pt" / "bad.pt")
    with pytest.raises(ValueError):
        list(PinnedBagLoader(["bad"], [0], str(tmp_path), pin=False))
    with pytest.raises(ValueError):
        PinnedBagLoader(names, [0], str(tmp_path))
