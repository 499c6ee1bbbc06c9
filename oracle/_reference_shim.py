"""Read-only import of the upstream reference (``/root/reference``) for pinning the oracle.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used in the build container by ``oracle/make_golden.py`` and by
the ``not gpu`` oracle tests when ``/root/reference`` exists.  ``/root/reference`` itself never travels;
``oracle/build_ref.py`` stages byte-for-byte copies of the hot path's modules under the git-ignored
``oracle/_ref/``, which does travel to the GPU box, where ``bench.py --impl reference`` and
``tests/test_gpu_dropin.py`` import them through this module.  The product never does.

The reference imports ``timm.models.layers.DropPath`` (modules/rrt.py:7) and
``trunc_normal_`` (modules/emb_position.py:4); ``timm`` is not installed, so two stand-in
symbols are registered in ``sys.modules`` before the import.  Nothing under
``/root/reference`` is modified.
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

# /root/reference in the build container; on the GPU box the tree staged by oracle/build_ref.py (oracle/_ref/,
# git-ignored, travels with the gpurun snapshot)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("RRT_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/modules/rrt.py") else _STAGED)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "rrt.py"))


class _StochasticDepth(nn.Module):
    """Stand-in for timm's DropPath; only constructed when drop_path > 0."""

    def __init__(self, p: float = 0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1.0 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def import_reference_rrt():
    """Return the reference's ``modules.rrt`` module object."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _StochasticDepth
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import modules.rrt as ref_rrt  # noqa: E402  (the reference's own package name)

    return ref_rrt


def build_reference_encoder(cfg, weights, dtype=torch.float64):
    """Instantiate the real ``RRTEncoder`` with ``cfg`` and load ``weights`` (strict)."""
    ref = import_reference_rrt()
    m = ref.RRTEncoder(mlp_dim=cfg.mlp_dim, region_num=cfg.region_num, n_layers=cfg.n_layers,
                       n_heads=cfg.n_heads, epeg=cfg.epeg, epeg_k=cfg.epeg_k,
                       region_size=cfg.region_size, min_region_num=cfg.min_region_num,
                       min_region_ratio=cfg.min_region_ratio, qkv_bias=cfg.qkv_bias,
                       cr_msa=cfg.cr_msa, crmsa_k=cfg.crmsa_k, all_shortcut=cfg.all_shortcut,
                       crmsa_mlp=cfg.crmsa_mlp, crmsa_heads=cfg.crmsa_heads,
                       epeg_bias=cfg.epeg_bias, pos=cfg.pos, pos_pos=cfg.pos_pos, peg_k=cfg.peg_k,
                       peg_bias=cfg.peg_bias, peg_1d=cfg.peg_1d, ffn=cfg.ffn, ffn_act=cfg.ffn_act,
                       mlp_ratio=cfg.mlp_ratio, epeg_2d=cfg.epeg_2d, epeg_type=cfg.epeg_type)
    m = m.to(dtype).eval()
    m.load_state_dict({k: v.to(dtype) for k, v in weights.items()}, strict=True)
    return m
