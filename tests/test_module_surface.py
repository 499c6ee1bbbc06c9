"""Drop-in surface of rrt_mil_b200.RRTEncoder: constructor keywords, state_dict layout,
error behaviour.  CPU only (no kernel is launched)."""
import inspect

import pytest
import torch

from oracle import rrt_oracle as O
from oracle import _reference_shim as shim
from rrt_mil_b200 import RRTEncoder

VARIANTS = [
    dict(),
    dict(epeg_k=21, crmsa_k=5),
    dict(crmsa_mlp=True, crmsa_heads=1, all_shortcut=True, epeg_k=13),
    dict(qkv_bias=False, epeg=False),
    dict(n_layers=3, cr_msa=False),
    dict(mlp_dim=256, region_num=16, epeg_bias=False),
    dict(pos='ppeg', pos_pos=-1),
    dict(pos='peg', peg_k=5, peg_1d=True, peg_bias=False, n_layers=3),
    dict(ffn=True),
    dict(ffn=True, ffn_act='relu', mlp_ratio=2.0, mlp_dim=256, n_layers=3, cr_msa=False),
    dict(epeg_2d=True, epeg_k=5),
    dict(epeg_type='value_bf', epeg_k=7),
    dict(epeg_type='value_af', epeg_2d=True, epeg_k=3, epeg_bias=False),
]


@pytest.mark.parametrize("over", VARIANTS)
def test_state_dict_keys_and_shapes_match_oracle_table(over):
    m = RRTEncoder(**over)
    cfg = O.EncoderConfig(**over)
    want = O.weight_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want


@pytest.mark.skipif(not shim.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("over", VARIANTS)
def test_state_dict_interchangeable_with_reference(over):
    ref = shim.import_reference_rrt().RRTEncoder(**over)
    ours = RRTEncoder(**over)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)


@pytest.mark.skipif(not shim.available(), reason="/root/reference only exists in the build container")
def test_constructor_signature_matches_reference():
    ref = inspect.signature(shim.import_reference_rrt().RRTEncoder.__init__)
    ours = inspect.signature(RRTEncoder.__init__)
    assert list(ref.parameters) == list(ours.parameters)
    for name, p in ref.parameters.items():
        assert ours.parameters[name].default == p.default, name


def test_module_behaves_like_an_nn_module():
    m = RRTEncoder(need_init=True)
    assert m.final_dim == 512
    assert sum(p.numel() for p in m.parameters()) == 2_105_984
    seq = torch.nn.Sequential(torch.nn.Linear(1024, 512), torch.nn.ReLU(), m)
    assert any(isinstance(c, torch.nn.Conv2d) for c in seq.modules())
    assert float(m.layers[0].attn.attn.qkv.bias.detach().abs().sum()) == 0.0
    m.eval(); m.train(); repr(m)


def test_unsupported_options_raise_loudly():
    with pytest.raises(NotImplementedError):
        RRTEncoder(attn='ntrans')
    with pytest.raises(NotImplementedError):
        RRTEncoder(pos='sincos')
    with pytest.raises(ValueError):
        RRTEncoder(pos='ppeg', peg_k=4)
    with pytest.raises(ValueError):
        RRTEncoder(ffn=True, mlp_ratio=0.3)
    with pytest.raises(NotImplementedError):
        RRTEncoder(epeg_type='value_xx')
    with pytest.raises(ValueError):
        RRTEncoder(epeg_k=4)
    with pytest.raises(TypeError):
        RRTEncoder(not_an_option=1)


def test_no_cpu_fallback():
    m = RRTEncoder().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.randn(1, 100, 512))
    with pytest.raises(ValueError):
        with torch.no_grad():
            m(torch.randn(2, 10, 512, device="meta") if False else torch.randn(2, 10, 512))


def test_product_never_imports_the_oracle():
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "rrt_mil_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


@pytest.mark.skipif(not shim.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("host", ["attmil.DAttention", "attmil.AttentionGated", "mean_max.MeanMIL", "mean_max.MaxMIL",
                                  "dsmil.MILNet"])
def test_drops_into_the_reference_mil_hosts(host):
    """The reference's own aggregators take the encoder as ``rrt=<module>`` (main.py:138-155).  Built once with
    the reference encoder and once with ours they must have the same parameter tree, exchange checkpoints with
    strict=True, and the host's own ``initialize_weights`` pass (isinstance checks on nn.Linear / nn.Conv2d /
    nn.LayerNorm, modules/attmil.py:6-25) must reach the parameters our holders own."""
    import importlib
    ref_rrt = shim.import_reference_rrt()
    mod_name, cls_name = host.split(".")
    mod = importlib.import_module("modules." + mod_name)
    cls = getattr(mod, cls_name)
    kw = dict(input_dim=1024, n_classes=2, dropout=True, act="relu")

    def build(enc):
        if cls_name == "MILNet":                                # dsmil: (n_classes, dropout, act, input_dim, rrt)
            return cls(n_classes=2, dropout=True, act="relu", input_dim=1024, rrt=enc)
        try:
            return cls(rrt=enc, **kw)
        except TypeError:
            return cls(rrt=enc, input_dim=1024, act="relu")   # AttentionGated(input_dim, act, bias, dropout, rrt)

    enc_kw = dict(epeg_k=9, crmsa_k=5)
    theirs, ours = build(ref_rrt.RRTEncoder(**enc_kw)), build(RRTEncoder(**enc_kw))
    ks_t = {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
    ks_o = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert ks_t == ks_o
    ours.load_state_dict(theirs.state_dict(), strict=True)
    theirs.load_state_dict(ours.state_dict(), strict=True)
    enc = [m for m in ours.modules() if isinstance(m, RRTEncoder)][0]
    # the host's initialisation pass zeroed our biases and reset our LayerNorms
    assert float(enc.layers[0].attn.attn.qkv.bias.detach().abs().sum()) == 0.0
    assert float((enc.norm.weight.detach() - 1).abs().sum()) == 0.0
    assert sum(p.numel() for p in enc.parameters()) == sum(
        p.numel() for p in [m for m in theirs.modules() if isinstance(m, ref_rrt.RRTEncoder)][0].parameters())


def test_encoder_survives_deepcopy_and_pickle_with_populated_caches():
    """EMA / teacher copies (copy.deepcopy) and torch.save(model) must work after a forward has populated the
    pointer caches (ctypes structs, fp16 shadows, the parameter-walk closure): none of them may travel."""
    import copy
    import io
    m = RRTEncoder(mlp_dim=128, n_heads=4, crmsa_heads=4, need_init=True)
    m._named_param_cache()
    m.__dict__["_w_cache"] = (("cpu", False, ()), object.__new__(type("Opaque", (), {"__reduce__": None})))
    m._shadow[1] = ((0, 0), torch.zeros(1))
    c = copy.deepcopy(m)
    assert "_w_cache" not in c.__dict__ and "_np_cache" not in c.__dict__ and c._shadow == {}
    assert bytes(c._cfg) == bytes(m._cfg) and c._cfg is not m._cfg
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), c.named_parameters()):
        assert n1 == n2 and p1 is not p2 and torch.equal(p1, p2)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert bytes(r._cfg) == bytes(m._cfg) and r.state_dict().keys() == m.state_dict().keys()
    # the copy is a working module: it rebuilds its caches on demand
    assert [n for n in r._named_param_cache()[0]] == [n for n, _ in m.named_parameters()]
