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


def _mil(seed, **kw):
    from rrt_mil_b200 import RRTMIL
    torch.manual_seed(seed)
    return RRTMIL(input_dim=512, n_classes=2, epeg_k=9, crmsa_k=3, **kw).cuda().train()


@pytest.mark.gpu
def test_graphed_train_step_follows_the_eager_training_trajectory():
    """Captured step == eager step: same bags, dropout off (no masks), 8 steps of Adam.  The backward sums with fp32
    atomics, so trajectories agree to rounding, not bit for bit: losses within 2e-3 relative, parameters within a
    small fraction of what the optimizer moved them."""
    from rrt_mil_b200.graph import GraphedTrainStep
    from rrt_mil_b200.optim import Adam
    g = torch.Generator(device="cuda").manual_seed(5)
    bags = [torch.randn(1, 700, 512, device="cuda", generator=g) for _ in range(8)]
    labels = [torch.tensor([i % 2], device="cuda") for i in range(8)]
    ma, mb = _mil(3, dropout=0.0, trans_dropout=0.0), _mil(3, dropout=0.0, trans_dropout=0.0)
    start = [p.detach().clone() for p in ma.parameters()]
    oa, ob = Adam(ma.parameters(), lr=2e-4, weight_decay=1e-5), Adam(mb.parameters(), lr=2e-4, weight_decay=1e-5)
    step = GraphedTrainStep(mb, ob, 700, 512)
    la, lb = [], []
    # the capture runs three warm-up steps on its static (zero) bag: give the eager model the same three
    for _ in range(3):
        oa.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(ma(torch.zeros(1, 700, 512, device="cuda")),
                                          torch.zeros(1, dtype=torch.long, device="cuda")).backward()
        oa.step()
    for x, y in zip(bags, labels):
        oa.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(ma(x), y)
        loss.backward()
        oa.step()
        la.append(float(loss))
        lb.append(float(step(x, y)))
    torch.cuda.synchronize()
    assert step.replays == 8 and step.t == 11
    assert all(abs(a - b) <= 2e-3 * max(1.0, abs(a)) for a, b in zip(la, lb)), (la, lb)
    moved = sum(float((p - s).abs().sum()) for p, s in zip(ma.parameters(), start))
    apart = sum(float((p - q).abs().sum()) for p, q in zip(ma.parameters(), mb.parameters()))
    assert apart <= 0.05 * moved, (apart, moved)
    step.sync_optimizer_state()
    assert all(ob.state[p]["step"] == 11 for p in mb.parameters())


@pytest.mark.gpu
def test_graphed_train_step_draws_a_new_dropout_mask_per_replay_and_learns():
    from rrt_mil_b200.graph import GraphedTrainStep
    from rrt_mil_b200.optim import Adam
    m = _mil(4, dropout=0.25, trans_dropout=0.1)
    opt = Adam(m.parameters(), lr=0.0)                         # lr 0: only the masks differ between replays
    step = GraphedTrainStep(m, opt, 600, 512, seed=123)
    x = torch.randn(1, 600, 512, device="cuda")
    y = torch.tensor([1], device="cuda")
    losses = [float(step(x, y)) for _ in range(4)]
    assert len({round(v, 6) for v in losses}) == 4, losses     # four replays, four masks
    m2 = _mil(4, dropout=0.25, trans_dropout=0.1)
    # an eager step on the default stream first (a loop that switches to the graph after a few steps): its
    # autograd graph must be released before the capture, see the class docstring
    eager_loss = torch.nn.functional.cross_entropy(m2(x), y)
    eager_loss.backward()
    m2.zero_grad(set_to_none=True)
    del eager_loss
    step2 = GraphedTrainStep(m2, Adam(m2.parameters(), lr=2e-4), 600, 512, seed=7)
    first = sum(float(step2(x, y)) for _ in range(5)) / 5
    for _ in range(40):
        step2(x, y)
    last = sum(float(step2(x, y)) for _ in range(5)) / 5
    assert last < 0.5 * first, (first, last)
    # eval after graphed training sees the updated weights (version counters were bumped)
    with torch.no_grad():
        logits = m2.eval()(x)
    assert int(logits.argmax()) == 1
