"""Range safety of the fp16 tensor-core operand path and half-precision inputs.

The internal activations that feed the tensor cores (LayerNorm output, q/k/v, attention output, landmarks) are
fp16: 10 mantissa bits like tf32 but a 5-bit exponent.  LayerNorm protects the first of them from the scale of the
features; the others depend on the weights.  These tests feed features scaled by 1e3 and 1e-4 and weights with row
norms far above the Xavier scale, forward and backward, against the fp64 oracle at the north star's 1e-3 / the
backward's 3e-3 bar; and fp16 / bf16 bags as an autocast host hands them over (main.py:101-102,439)."""
import pytest
import torch

from oracle import rrt_oracle as O
import gpu_util as G
from rrt_mil_b200 import RRTEncoder

pytestmark = pytest.mark.gpu


def _scaled_weights(cfg, seed, w_scale):
    w = O.make_weights(cfg, seed)
    if w_scale != 1.0:
        for k in w:
            if k.endswith(("qkv.weight", "proj.weight")) or k.endswith("phi"):
                w[k] = w[k] * w_scale
    return w


def _operand_noise_sensitivity(x, w, cfg, ref):
    """How far the fp64 forward moves when its INPUT is perturbed by the rounding of a 10-bit mantissa (2^-11
    relative, random sign): a lower bound on what any tf32 / fp16-operand implementation can promise for these
    weights.  Xavier-scale weights: 3e-4; attention weights x2: 1.4e-3, x3: 3.8e-3, x6: 2.8e-2 (peaked softmax)."""
    g = torch.Generator().manual_seed(0)
    xp = x * (1 + (torch.rand(x.shape, generator=g, dtype=torch.float64) - 0.5) * 2.0 ** -10)
    return O.rel_err(O.encoder_forward(xp, w, cfg, "spec"), ref)


@pytest.mark.parametrize("x_scale,w_scale", [(1e3, 1.0), (1e-2, 1.0), (1.0, 2.0), (1.0, 6.0), (30.0, 4.0), (1e3, 0.2)])
def test_forward_parity_holds_for_scaled_features_and_large_weights(x_scale, w_scale):
    """Exponent range is not the limit of the fp16 operand path (features x1e3 / x1e-2, weights x6: finite, no
    saturation); the limit is the 10-bit mantissa once the weights make the function ill-conditioned.  The bar:
    1e-3 (north star) or three times the sensitivity of the fp64 function to operand-level input rounding,
    whichever is larger.  (Not covered: bags whose per-row variance is far below LayerNorm's eps = 1e-5, e.g.
    features x1e-4 -- there the reference's own fp32 forward is 4.8e-4 away from fp64.)"""
    cfg = O.EncoderConfig(epeg_k=9, crmsa_k=5)
    w = _scaled_weights(cfg, 41, w_scale)
    x = O.make_bag(1500, cfg.mlp_dim, 42, kind="relu") * x_scale
    ref = O.encoder_forward(x, w, cfg, "spec")
    m = G.make_encoder(cfg, w)
    with torch.no_grad():
        y = m(x.float().cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    err, floor = O.rel_err(y.cpu(), ref), _operand_noise_sensitivity(x, w, cfg, ref)
    assert err < max(1e-3, 3 * floor), (x_scale, w_scale, err, floor)


@pytest.mark.parametrize("x_scale,g_scale", [(1e3, 1.0), (1e-2, 1.0), (1.0, 1e4), (1.0, 1e-6)])
def test_backward_parity_holds_for_scaled_features_and_gradients(x_scale, g_scale):
    """dY far outside the fp16 range (1e4 * unit rows, 1e-6 * unit rows): the backward rescales every stage by a
    power of two from an amax probe, so the gradients keep the 3e-3 bar."""
    cfg = O.EncoderConfig()
    w = O.make_weights(cfg, 7)
    x = (O.make_bag(400, cfg.mlp_dim, 8) * x_scale)
    gout = O.make_bag(400, cfg.mlp_dim, 9) * g_scale
    wr = {k: v.clone().requires_grad_() for k, v in w.items()}
    xr = x.clone().requires_grad_()
    (O.encoder_forward(xr, wr, cfg, "spec") * gout).sum().backward()
    m = G.make_encoder(cfg, w).train()
    m.drop_out = 0.0                       # the oracle pass above has no dropout
    xd = x.float().cuda().requires_grad_()
    m(xd).backward(gout.float().cuda())
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double().cpu() - b).norm() / (b.norm() + 1e-300))
    assert rel(xd.grad, xr.grad) < 3e-3
    for name, p in m.named_parameters():
        ref = wr[name].grad
        if ref is None or float(ref.norm()) < 1e-12 * float(abs(g_scale)):   # pe.bias: exactly zero gradient
            continue
        tol = 2e-2 if name.endswith("phi") else 3e-3    # phi: min / max ties of the normaliser
        assert rel(p.grad, ref) < tol, (name, rel(p.grad, ref))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_bags_are_widened_exactly(dtype):
    torch.manual_seed(0)
    m = RRTEncoder(need_init=True).cuda().eval()
    xh = torch.randn(1, 2000, 512, device="cuda").to(dtype)
    with torch.no_grad():
        y = m(xh)
        want = m(xh.float())
        bags = m.forward_bags([xh[0], xh[0, :700].contiguous()])
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and torch.equal(y, want)
    assert torch.equal(bags[0], want[0]) and bags[1].dtype == torch.float32
    with pytest.raises(NotImplementedError):
        m.train()(xh.requires_grad_())


def test_encoder_runs_inside_an_autocast_host():
    """The reference's --amp path: Linear + ReLU under fp16 autocast produce half rows, the encoder follows."""
    torch.manual_seed(1)
    enc = RRTEncoder(need_init=True).cuda().eval()
    host = torch.nn.Sequential(torch.nn.Linear(1024, 512), torch.nn.ReLU(), enc).cuda().eval()
    x = torch.randn(1, 3000, 1024, device="cuda")
    with torch.no_grad():
        full = host(x)
        with torch.autocast("cuda", dtype=torch.float16):
            amp = host(x)
    torch.cuda.synchronize()
    assert amp.dtype == torch.float32 and torch.isfinite(amp).all()
    rel = float((amp.double() - full.double()).norm() / full.double().norm())
    assert rel < 1e-2, rel     # the fp16 Linear in front is the autocast host's, not ours
