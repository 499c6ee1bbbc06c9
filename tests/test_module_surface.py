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
        RRTEncoder(epeg_2d=True)
    with pytest.raises(NotImplementedError):
        RRTEncoder(epeg_type='value_bf')
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
