"""RRTMIL = patch_to_emb -> RRTEncoder -> DAttention pooling -> predictor.
CPU part: the oracle restatement against goldens generated from the reference's RRTMIL, and the
drop-in surface.  GPU part: the CUDA path against the same goldens / the fp64 oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rrt_oracle as O
from oracle import _reference_shim as shim
from golden_util import GOLDEN_DIR, MANIFEST

MIL = {c["name"]: c for c in MANIFEST.get("mil_cases", [])}
MIL_TRAIN = {c["name"]: c for c in MANIFEST.get("mil_train_cases", [])}


def _head_kw(c):
    """RRTMIL keywords beyond the fixed columns of a fixture (da_gated, da_dropout, act)."""
    return dict(c.get("extra") or {})


def load_mil(name, dtype=torch.float64):
    c = MIL[name]
    cfg = O.EncoderConfig(**c["config"])
    w = O.make_mil_weights(cfg, c["input_dim"], c["n_classes"], c["weight_seed"], da_bias=c["da_bias"], dtype=dtype,
                           **{k: v for k, v in _head_kw(c).items() if k in ("da_gated", "da_dropout")})
    x = O.make_bag(c["L"], c["input_dim"], c["bag_seed"], dtype=dtype)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return c, cfg, w, x, gold


@pytest.mark.parametrize("name", sorted(MIL))
@pytest.mark.parametrize("order", ["reference", "spec"])
def test_oracle_mil_matches_golden(name, order):
    c, cfg, w, x, gold = load_mil(name)
    logits, attn = O.mil_forward(x, w, cfg, c["act"], c["da_act"], order)
    assert np.abs(logits.numpy() - gold["logits"]).max() < 1e-9
    assert np.abs(attn.numpy() - gold["attn"]).max() < 1e-6 * gold["attn"].max() + 1e-9


@pytest.mark.skipif(not shim.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("kw", [dict(), dict(input_dim=512, n_classes=4, act='gelu', da_act='tanh', da_bias=True),
                                dict(epeg_k=21, crmsa_k=5, da_dropout=True),
                                dict(da_gated=True), dict(da_gated=True, da_bias=True, da_dropout=True, da_act='gelu')])
def test_rrtmil_state_dict_interchangeable_with_reference(kw):
    from rrt_mil_b200 import RRTMIL
    ref = shim.import_reference_rrt().RRTMIL(**kw)
    ours = RRTMIL(**kw)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_rrtmil_unsupported_options_raise():
    from rrt_mil_b200 import RRTMIL
    with pytest.raises(NotImplementedError):
        RRTMIL(pool='avg')
    m = RRTMIL().eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(torch.randn(1, 10, 1024))


def load_mil_train(name, dtype=torch.float64):
    c = MIL_TRAIN[name]
    cfg = O.EncoderConfig(**c["config"])
    w = O.make_mil_weights(cfg, c["input_dim"], c["n_classes"], c["weight_seed"], da_bias=c["da_bias"], dtype=dtype,
                           **{k: v for k, v in _head_kw(c).items() if k in ("da_gated", "da_dropout")})
    x = O.make_bag(c["L"], c["input_dim"], c["bag_seed"], dtype=dtype)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return c, cfg, w, x, gold


def grad_errors(grads, gold, floor):
    """rel. Frobenius error over the stored rows + of the whole-tensor norm, per parameter gradient."""
    e = {}
    for k in gold:
        if not k.startswith("g:"):
            continue
        name = k[2:]
        ref, fro = gold[k].astype(np.float64), float(gold["gfro:" + name])
        if fro < floor:          # exactly-zero reference gradient (pe.bias)
            continue
        a = grads[name].detach().double().cpu().numpy()
        a2 = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(1, -1)
        st = max(1, -(-a2.shape[0] // 96))
        e[name] = float(np.linalg.norm(a2[::st] - ref) / max(np.linalg.norm(ref), 1e-300))
        e["|" + name + "|"] = abs(np.linalg.norm(a) - fro) / fro
    return e


@pytest.mark.parametrize("name", sorted(MIL_TRAIN))
def test_oracle_mil_train_step_matches_reference_autograd(name):
    """Oracle RRTMIL in training mode (dp + proj_drop masks) + CrossEntropy + torch autograd vs the fixture
    the reference produced with the same masks installed."""
    c, cfg, w, x, gold = load_mil_train(name)
    w = {k: v.clone().requires_grad_() for k, v in w.items()}
    kw = _head_kw(c)
    logits, _ = O.mil_forward(x, w, cfg, kw.get("act", "relu"), c["da_act"], "spec",
                              drop=(c["dropout"], c["seed"], c["trans_dropout"], c["seed"] + 1),
                              pool_drop=(0.25, c["seed"]) if kw.get("da_dropout") else None)
    loss = torch.nn.functional.cross_entropy(logits[None], torch.tensor([c["label"]]))
    loss.backward()
    assert np.abs(logits.detach().numpy() - gold["logits"]).max() < 1e-9 and abs(float(loss) - float(gold["loss"])) < 1e-9
    e = grad_errors({k: v.grad for k, v in w.items()}, gold, 1e-14)
    assert max(e.values()) < 2e-6, e


# ---- GPU -----------------------------------------------------------------------------------------
TOL = 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MIL_TRAIN))
def test_cuda_rrtmil_train_step_matches_reference_fixture(name):
    """Full RRTMIL train step on the GPU (patch_to_emb + dp, taped encoder with proj dropout, pooling head,
    CrossEntropy, loss.backward() through the three CUDA backward entry points) vs the reference fixture."""
    from rrt_mil_b200 import RRTMIL
    c, cfg, w, x, gold = load_mil_train(name)
    kw = {k: v for k, v in c["config"].items() if k in ("region_num", "n_layers", "epeg_k", "crmsa_k",
                                                          "all_shortcut", "crmsa_heads")}
    hk = _head_kw(c)
    m = RRTMIL(input_dim=c["input_dim"], n_classes=c["n_classes"], da_act=c["da_act"], da_bias=c["da_bias"],
               dropout=c["dropout"], trans_dropout=c["trans_dropout"], **kw, **hk).cuda().train()
    m.load_state_dict({k: v.float() for k, v in w.items()}, strict=True)
    m._dropout_seed, m.online_encoder._dropout_seed = c["seed"], c["seed"] + 1
    logits = m(x.float().cuda().unsqueeze(0))
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([c["label"]], device="cuda"))
    loss.backward()
    torch.cuda.synchronize()
    gl = gold["logits"]
    assert np.abs(logits[0].detach().cpu().numpy() - gl).max() <= TOL * max(1.0, np.abs(gl).max())
    assert abs(float(loss.detach()) - float(gold["loss"])) < 2e-3
    e = grad_errors({n: p.grad for n, p in m.named_parameters()}, gold, 1e-12)
    print(name, {k: f"{v:.1e}" for k, v in e.items()})
    # Tolerance of the COMPOSED step: the softmax-pooling backward forms da_l = a_l (h_l . dpooled - pooled .
    # dpooled), a difference of nearly equal numbers, so the 2e-4 forward error of the fp16-operand encoder is
    # amplified ~50x before it enters the encoder's backward (measured: 2.2e-2 on patch_to_emb.weight, 7.8e-3
    # on norm.weight for miltrain_r50_n800, while each stage alone -- tests below, and
    # test_gpu_backward.py for the encoder -- sits at 1-3e-3 on exact inputs).  Mixed-precision gradient noise,
    # not a defect of a stage: 3e-2 here, the per-stage tests hold the tight bars.
    for k, v in e.items():
        assert v < 3e-2, (k, v)
    # eval mode still runs the inference kernels and ignores both dropouts
    with torch.no_grad():
        le = m.eval()(x.float().cuda().unsqueeze(0))
    ref_eval, _ = O.mil_forward(x, {k: v for k, v in load_mil_train(name)[2].items()}, cfg, hk.get("act", "relu"),
                                c["da_act"], "spec")
    assert np.abs(le[0].cpu().numpy() - ref_eval.numpy()).max() <= TOL * max(1.0, float(ref_eval.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MIL))
def test_cuda_rrtmil_matches_golden(name):
    from rrt_mil_b200 import RRTMIL
    c, cfg, w, x, gold = load_mil(name)
    kw = {k: v for k, v in c["config"].items() if k in ("region_num", "n_layers", "epeg_k", "crmsa_k",
                                                          "all_shortcut", "crmsa_heads", "crmsa_mlp")}
    m = RRTMIL(input_dim=c["input_dim"], n_classes=c["n_classes"], act=c["act"], da_act=c["da_act"],
               da_bias=c["da_bias"], **kw, **_head_kw(c)).cuda().eval()
    m.load_state_dict({k: v.float() for k, v in w.items()}, strict=True)
    with torch.no_grad():
        logits, attn = m(x.float().cuda().unsqueeze(0), return_attn=True)
        logits2 = m(x.float().cuda().unsqueeze(0))
        _, raw = m(x.float().cuda().unsqueeze(0), return_attn=True, no_norm=True)
    torch.cuda.synchronize()
    assert logits.shape == (1, c["n_classes"]) and attn.shape == (1, c["L"])
    assert torch.equal(logits, logits2)
    gl = gold["logits"]
    assert np.abs(logits[0].cpu().numpy() - gl).max() <= TOL * max(1.0, np.abs(gl).max())
    ga = gold["attn"].astype(np.float64)
    a = attn[0].double().cpu().numpy()
    assert abs(a.sum() - 1.0) < 1e-4
    assert np.linalg.norm(a - ga) <= 5 * TOL * np.linalg.norm(ga)   # exp() amplifies score errors
    # raw scores (no_norm) reproduce the normalised map through a softmax
    assert np.allclose(torch.softmax(raw[0].double(), 0).cpu().numpy(), a, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("L,din,act", [(1000, 1024, "relu"), (333, 512, "gelu"), (64, 64, "none")])
def test_cuda_patch_embed_matches_fp64(L, din, act):
    import ctypes as C
    from rrt_mil_b200 import cabi
    g = torch.Generator().manual_seed(L + din)
    x = torch.randn(L, din, generator=g)
    w = torch.randn(512, din, generator=g) / din ** 0.5
    b = torch.randn(512, generator=g)
    ref = O._act(act)(x.double() @ w.double().T + b.double())
    lib = cabi.lib()
    n = C.c_size_t()
    cabi.check(lib.rrt_mil_head_workspace_bytes(L, din, 512, 128, C.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device="cuda")
    out = torch.empty(L, 512, device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    code = {"none": 0, "relu": 1, "gelu": 2}[act]
    cabi.check(lib.rrt_patch_embed_forward(xd.data_ptr(), L, din, 512, wd.data_ptr(), bd.data_ptr(), None, code,
                                           out.data_ptr(), ws.data_ptr(), n.value, 0.0, 0, None,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert O.rel_err(out.cpu(), ref) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("L,da_act,bias,ncls,gated,dadrop", [
    (800, "relu", False, 2, False, False), (600, "tanh", True, 3, False, False), (65, "relu", True, 4, False, False),
    (9000, "relu", False, 2, False, False),
    (700, "gelu", True, 2, False, False), (500, "relu", False, 2, False, True), (640, "tanh", True, 3, False, True),
    (900, "relu", False, 2, True, False), (450, "gelu", True, 3, True, True), (300, "tanh", True, 2, True, True)])
def test_cuda_attn_pool_backward_matches_fp64_autograd(L, da_act, bias, ncls, gated, dadrop):
    """The pooling head + predictor backward alone vs torch autograd in fp64: plain and gated head
    (modules/datten.py:5-83), every activation, with and without the nn.Dropout inside the score MLP.

    The reference evaluates the activation of the score MLP at the pre-activations the fp16-operand forward
    actually produced (straight-through: value of the fp16-rounded product, gradient of the exact one):
    d/dz ReLU is a step function, and ~0.04 % of the pre-activations sit within the forward's rounding
    error of the kink; evaluated at the fp64 values instead, those flips alone are 0.5-2 % of the gradient
    norm (measured 1.6e-2 at L=800, and reproduced to 4 digits by a numpy emulation of the pipeline) --
    a property of running the forward on fp16 tensor-core operands, not of the backward kernels."""
    import torch.nn.functional as F
    from rrt_mil_b200 import RRTMIL
    from rrt_mil_b200.mil import _AttnPoolFunction
    g = torch.Generator().manual_seed(L)
    h = torch.randn(L, 512, generator=g, dtype=torch.float64)
    m = RRTMIL(input_dim=512, n_classes=ncls, da_act=da_act, da_bias=bias, da_gated=gated, da_dropout=dadrop).cuda()
    m.train(dadrop)
    seed = 4242
    m._dropout_seed = seed
    firsts, score, _ = m.pool_fn.attention.head_params()
    pred = m.predictor
    lins = firsts + [score, pred]
    with torch.no_grad():
        for l in lins:
            l.weight.copy_(torch.randn(l.weight.shape, generator=g) * (2.0 / sum(l.weight.shape)) ** 0.5)
            if l.bias is not None:
                l.bias.copy_(0.1 * torch.randn(l.bias.shape, generator=g))
    dbl = lambda t: None if t is None else t.detach().double().cpu().requires_grad_()
    rw, rb = [dbl(l.weight) for l in lins], [dbl(l.bias) for l in lins]
    hr = h.clone().requires_grad_()

    def first_layer(i):   # value of the fp16-operand product, gradient of the exact one
        pre = F.linear(hr, rw[i])
        pre = pre + (F.linear(h.half().double(), rw[i].detach().half().double()) - pre).detach()
        return pre if rb[i] is None else pre + rb[i]

    hid = O._act(da_act)(first_layer(0))
    N = 256 if gated else 128
    mask = O.dropout_mask(L, N, 0.25, seed, O.DROP_STREAM_POOL) if dadrop else torch.ones(L, N, dtype=torch.float64)
    hid = hid * mask[:, :128]
    if gated:
        hid = hid * (torch.sigmoid(first_layer(1)) * mask[:, 128:])
    ns = len(firsts)
    sc = F.linear(hid, rw[ns], rb[ns]).squeeze(-1)
    logits_ref = F.linear(torch.softmax(sc, 0) @ hr, rw[ns + 1], rb[ns + 1])
    label = torch.tensor([ncls - 1])
    F.cross_entropy(logits_ref[None], label).backward()
    hd = h.float().cuda().requires_grad_()
    gate = firsts[1] if gated else None
    logits = _AttnPoolFunction.apply(m, hd, firsts[0].weight, firsts[0].bias, None if gate is None else gate.weight,
                                     None if gate is None else gate.bias, score.weight.view(-1), score.bias,
                                     pred.weight, pred.bias)
    F.cross_entropy(logits[None], label.cuda()).backward()
    torch.cuda.synchronize()
    assert float((logits.detach().cpu().double() - logits_ref.detach()).abs().max()) < \
        2e-3 * max(1.0, float(logits_ref.abs().max()))
    errs = {"dh": O.rel_err(hd.grad.cpu(), hr.grad)}
    names = (["wa", "wb"] if gated else ["w1"]) + ["w2", "pw"]
    for n, l, qw, qb in zip(names, lins, rw, rb):
        errs[n] = O.rel_err(l.weight.grad.cpu(), qw.grad)
        if l.bias is None:
            continue
        if n == "w2":   # softmax is shift invariant: d loss / d b2 == 0 exactly; only rounding noise may remain
            assert float(l.bias.grad.abs().max()) <= 1e-4 * float(l.weight.grad.abs().max())
            continue
        errs[n + ".b"] = O.rel_err(l.bias.grad.cpu(), qb.grad)
    print(L, da_act, gated, dadrop, errs)
    for n, e in errs.items():
        assert e < 2e-3, (n, e, errs)


@pytest.mark.gpu
@pytest.mark.parametrize("L,din,act,p", [(800, 1024, "relu", 0.25), (333, 512, "relu", 0.0), (500, 512, "none", 0.25),
                                         (700, 512, "gelu", 0.25), (260, 1024, "gelu", 0.0)])
def test_cuda_patch_embed_backward_matches_fp64(L, din, act, p):
    """patch_to_emb + dp backward alone: dW, db vs fp64 (the mask of the forward regenerated / read off
    the output), and the training forward itself (dropout applied) vs the oracle with the same mask."""
    import ctypes as C
    from rrt_mil_b200 import cabi
    g = torch.Generator().manual_seed(L + din)
    x = torch.randn(L, din, generator=g)
    w = torch.randn(512, din, generator=g) / din ** 0.5
    b = torch.randn(512, generator=g)
    dout = torch.randn(L, 512, generator=g) * 1e-3
    seed = 1234
    mask = O.dropout_mask(L, 512, p, seed, O.DROP_STREAM_PATCH)
    z = x.double() @ w.double().T + b.double()
    out_ref = O._act(act)(z) * mask
    lib, st = cabi.lib(), torch.cuda.current_stream().cuda_stream
    n = C.c_size_t()
    cabi.check(lib.rrt_mil_head_workspace_bytes(L, din, 512, 128, C.byref(n)))
    tape = torch.empty(n.value, dtype=torch.uint8, device="cuda")
    out = torch.empty(L, 512, device="cuda")
    xd, wd, bd, dd = x.cuda(), w.cuda(), b.cuda(), dout.cuda()
    code = {"none": 0, "relu": 1, "gelu": 2}[act]
    pre = torch.empty(L, 512, device="cuda") if act == "gelu" else None
    pre_p = None if pre is None else pre.data_ptr()
    cabi.check(lib.rrt_patch_embed_forward(xd.data_ptr(), L, din, 512, wd.data_ptr(), bd.data_ptr(), None, code,
                                           out.data_ptr(), tape.data_ptr(), n.value, p, seed, pre_p, st))
    dw, db = torch.empty_like(wd), torch.empty_like(bd)
    nws = 512 + L * 512 * 2 + 256
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    if act == "gelu":   # without the forward's pre-activations the backward refuses
        assert lib.rrt_patch_embed_backward(dd.data_ptr(), out.data_ptr(), None, L, din, 512, code, p, seed,
                                            tape.data_ptr(), n.value, dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                            nws, st) == cabi.RRT_E_INVALID
    cabi.check(lib.rrt_patch_embed_backward(dd.data_ptr(), out.data_ptr(), pre_p, L, din, 512, code, p, seed, tape.data_ptr(),
                                            n.value, dw.data_ptr(), db.data_ptr(), ws.data_ptr(), nws, st))
    torch.cuda.synchronize()
    assert O.rel_err(out.cpu(), out_ref) < TOL
    # the backward of the forward that was actually computed: ReLU's kink makes d/dz a step function, and the
    # fp16-operand forward puts ~0.04 % of the pre-activations on the other side of it than fp64 does (9e-3 of
    # the gradient norm), so the reference takes the kept / dropped pattern from the GPU's own output
    keep = (out.cpu().double() != 0).double() / (1.0 - p) if act == "relu" else mask
    if act == "gelu":   # smooth: the derivative at the fp64 pre-activations is the reference
        assert O.rel_err(pre.cpu(), z) < TOL
        zr = z.clone().requires_grad_()
        torch.nn.functional.gelu(zr).sum().backward()
        keep = mask * zr.grad
    dz = dout.double() * keep
    assert O.rel_err(dw.cpu(), dz.T @ x.double()) < 2e-3 and O.rel_err(db.cpu(), dz.sum(0)) < 2e-3
