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


def load_mil(name, dtype=torch.float64):
    c = MIL[name]
    cfg = O.EncoderConfig(**c["config"])
    w = O.make_mil_weights(cfg, c["input_dim"], c["n_classes"], c["weight_seed"], da_bias=c["da_bias"], dtype=dtype)
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
                                dict(epeg_k=21, crmsa_k=5, da_dropout=True)])
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
        RRTMIL(da_gated=True)
    with pytest.raises(NotImplementedError):
        RRTMIL(pool='avg')
    m = RRTMIL().eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(torch.randn(1, 10, 1024))


# ---- GPU -----------------------------------------------------------------------------------------
TOL = 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MIL))
def test_cuda_rrtmil_matches_golden(name):
    from rrt_mil_b200 import RRTMIL
    c, cfg, w, x, gold = load_mil(name)
    kw = {k: v for k, v in c["config"].items() if k in ("region_num", "n_layers", "epeg_k", "crmsa_k",
                                                          "all_shortcut", "crmsa_heads", "crmsa_mlp")}
    m = RRTMIL(input_dim=c["input_dim"], n_classes=c["n_classes"], act=c["act"], da_act=c["da_act"],
               da_bias=c["da_bias"], **kw).cuda().eval()
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
                                           out.data_ptr(), ws.data_ptr(), n.value,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert O.rel_err(out.cpu(), ref) < TOL
