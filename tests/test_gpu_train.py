"""Training harness pieces on the GPU: the multi-tensor Adam / AdamW step against
torch.optim, and an end-to-end optimisation loop through the CUDA forward / backward."""
import pytest
import torch

from oracle import rrt_oracle as O
import gpu_util as G
from rrt_mil_b200 import RRTEncoder
from rrt_mil_b200.optim import Adam, AdamW

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decoupled,wd", [(False, 1e-5), (False, 0.0), (True, 1e-2)])
def test_adam_matches_torch_optim(decoupled, wd):
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(1536, 512), (512,), (3,), (8, 1, 15, 1), (1,), (4097,), (512, 3)] + [(17, 5)] * 50   # > 48 tensors
    ours = [torch.randn(s, device="cuda", generator=g).requires_grad_() for s in shapes]
    ref = [p.detach().clone().requires_grad_() for p in ours]
    kw = dict(lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    opt = (AdamW if decoupled else Adam)(ours, **kw)
    topt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, **kw)
    for step in range(5):
        for p, q in zip(ours, ref):
            gr = torch.randn(p.shape, device="cuda", generator=g) * (10.0 ** (step - 2))
            p.grad, q.grad = gr.clone(), gr.clone()
        if step == 3:
            ours[2].grad = None      # a parameter that skips a step keeps its own step count
            ref[2].grad = None
        opt.step()
        topt.step()
    for p, q in zip(ours, ref):
        assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (p.shape, (p - q).abs().max())
    sd = opt.state_dict()["state"]
    assert set(sd[0]) == {"step", "exp_avg", "exp_avg_sq"} and sd[0]["step"] == 5 and sd[2]["step"] == 4
    with pytest.raises(RuntimeError):      # CPU parameters: no fallback
        _cpu_param_step()


def _cpu_param_step():
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    Adam([p]).step()


def test_encoder_training_loop_reduces_the_loss():
    """A few Adam steps on a regression target through the CUDA forward (training mode, dropout 0.1) and
    backward: the loss must fall, and the same seeds must reproduce the run (the dropout masks exactly;
    the parameters up to the order of the fp32 atomic gradient reductions)."""
    def run():
        torch.manual_seed(11)
        cfg = O.EncoderConfig()
        m = G.make_encoder(cfg, O.make_weights(cfg, 3)).train()
        opt = Adam(m.parameters(), lr=2e-4, weight_decay=1e-5)      # the reference's optimiser settings
        x = O.make_bag(600, 512, 4, kind="relu").float().cuda()
        target = O.make_bag(600, 512, 5).float().cuda() * 0.1
        losses = []
        for _ in range(8):
            opt.zero_grad(set_to_none=True)
            loss = (m(x) - target).square().mean()
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return losses, torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    l1, p1 = run()
    l2, p2 = run()
    assert l1[-1] < 0.97 * l1[0] and all(b < a for a, b in zip(l1, l1[1:])), l1   # measured: 1.034 -> 0.945
    assert torch.allclose(torch.tensor(l1), torch.tensor(l2), rtol=1e-4)
    # Adam divides by sqrt(v): an element whose gradient is ~0 can flip the sign of its (lr-sized) update
    # when the fp32 atomic reductions of the backward sum in another order, so compare in the mean and
    # bound the worst case by the distance 8 steps of size lr can cover
    d = (p1 - p2).abs()
    assert float(d.mean()) < 1e-5 and float(d.max()) <= 2 * 8 * 2e-4


def test_eval_after_fused_adam_uses_the_updated_weights():
    """The fused step writes parameters through raw pointers; the eval-mode fp16 weight shadows (cached by
    storage + version counter) must notice: eval output after training == a fresh module with the same state."""
    cfg = O.EncoderConfig()
    m = G.make_encoder(cfg, O.make_weights(cfg, 3))
    x = O.make_bag(500, 512, 4).float().cuda()
    with torch.no_grad():
        y_before = m(x)                     # builds the shadows
    m.train()
    opt = Adam(m.parameters(), lr=1e-2)
    v0 = m.norm.weight._version
    m(x).square().mean().backward()
    opt.step()
    assert m.norm.weight._version > v0
    m.eval()
    fresh = RRTEncoder(**cfg.to_dict()).cuda().eval()
    fresh.load_state_dict(m.state_dict(), strict=True)
    with torch.no_grad():
        y_after, y_fresh = m(x), fresh(x)
    assert torch.equal(y_after, y_fresh) and not torch.equal(y_after, y_before)
