"""Survival head / NLL-surv loss: plain torch on the [1, n_bins] logits.  The CPU tests
pin the restatement against the reference's own loss function (when /root/reference is present) and a
hand-computed case; the GPU test runs a survival train step through the CUDA path."""
import importlib.util
import os

import pytest
import torch

from rrt_mil_b200.survival import hazards_and_survival, nll_surv_loss

REF = "/root/reference/Survival/utils/loss.py"


def test_hazards_survival_and_loss_by_hand():
    logits = torch.tensor([[0.0, 1.0, -1.0, 2.0]], dtype=torch.float64)
    h, S = hazards_and_survival(logits)
    assert torch.allclose(h[0, 0], torch.tensor(0.5, dtype=torch.float64))
    assert torch.allclose(S[0], torch.cumprod(1 - h[0], 0))
    # event observed in bin 2: -(log S(1) + log h(2));   censored in bin 2: -log S(2)
    l_event = nll_surv_loss(h, S, torch.tensor([2]), torch.tensor([0]))
    l_cens = nll_surv_loss(h, S, torch.tensor([2]), torch.tensor([1]))
    assert torch.allclose(l_event, -(torch.log(S[0, 1]) + torch.log(h[0, 2])))
    assert torch.allclose(l_cens, -torch.log(S[0, 2]))


@pytest.mark.skipif(not os.path.isfile(REF), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("alpha", [0.0, 0.4])
def test_loss_matches_reference_function(alpha):
    spec = importlib.util.spec_from_file_location("ref_surv_loss", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(5, 4, generator=g, dtype=torch.float64)
    h, S = hazards_and_survival(logits)
    Y = torch.randint(0, 4, (5,), generator=g)
    c = torch.randint(0, 2, (5,), generator=g)
    assert torch.allclose(nll_surv_loss(h, S, Y, c, alpha), ref.nll_loss(h, S, Y, c, alpha=alpha), atol=1e-14)
    assert torch.allclose(nll_surv_loss(h, None, Y, c, alpha), ref.nll_loss(h, None, Y, c, alpha=alpha), atol=1e-14)


@pytest.mark.gpu
def test_survival_train_step_on_the_cuda_path():
    from rrt_mil_b200.survival import SurvivalRRTMIL
    from rrt_mil_b200.optim import Adam
    torch.manual_seed(3)
    m = SurvivalRRTMIL(input_dim=512, n_classes=4).cuda().train()
    opt = Adam(m.parameters(), lr=2e-4, weight_decay=1e-5)
    x = torch.randn(1, 900, 512, device="cuda")
    Y, c = torch.tensor([2], device="cuda"), torch.tensor([0], device="cuda")
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        h, S = m(x)
        assert h.shape == (1, 4) and S.shape == (1, 4)
        loss = nll_surv_loss(h, S, Y, c)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    with torch.no_grad():
        h, S = m.eval()(x)
    assert bool((S[0, 1:] <= S[0, :-1]).all()) and bool(((h > 0) & (h < 1)).all())
