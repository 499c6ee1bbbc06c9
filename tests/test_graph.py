"""GraphedForward (rrt_mil_b200/graph.py): CUDA-graph replay of the encoder forward for fixed bag lengths.
CPU: argument validation.  GPU: replay == eager, bit for bit."""
import pytest
import torch

from rrt_mil_b200 import RRTEncoder
from rrt_mil_b200.graph import GraphedForward


def test_graphed_forward_validates_its_arguments():
    m = RRTEncoder(mlp_dim=128, n_heads=4, crmsa_heads=4)
    with pytest.raises(RuntimeError, match="inference-only"):
        GraphedForward(m, [100])                       # training mode
    m.eval()
    with pytest.raises(ValueError):
        GraphedForward(m, [])
    with pytest.raises(ValueError):
        GraphedForward(m, [100, 0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GraphedForward(m, [100])                       # parameters on the CPU


@pytest.mark.gpu
def test_graph_replay_equals_eager_forward():
    torch.manual_seed(0)
    m = RRTEncoder(need_init=True).cuda().eval()
    for lengths, lanes in [([512], 1), ([9000], 1), ([3000, 777, 1, 4097], 3), ([2000] * 9, 8)]:
        g = GraphedForward(m, lengths, lanes=lanes)
        for trial in range(2):
            bags = [torch.randn(n, 512, device="cuda") for n in lengths]
            with torch.no_grad():
                want = [m(b) for b in bags]
            got = g(bags)
            torch.cuda.synchronize()
            for a, b in zip(want, got):
                assert torch.equal(a, b), (lengths, lanes, trial)
        assert g.captures == 1
        # a weight update (version counter bump) must trigger a re-capture with fresh fp16 shadows
        with torch.no_grad():
            m.layers[0].attn.attn.qkv.weight.mul_(1.01)
            want = [m(b) for b in bags]
        got = g(bags)
        torch.cuda.synchronize()
        assert g.captures == 2
        for a, b in zip(want, got):
            assert torch.equal(a, b)
        with pytest.raises(ValueError):
            g(bags[:-1] + [bags[-1][:-1]] if bags[-1].shape[0] > 1 else bags + bags)
