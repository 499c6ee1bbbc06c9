"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference (run in the build container).

    python -m oracle.make_golden            # writes tests/golden/<case>.npz + manifest.json

For every case the real ``modules.rrt.RRTEncoder`` (imported read-only through
``oracle/_reference_shim.py``) runs in float64, eval mode, on a seeded bag with seeded
weights.  Inputs and weights are NOT stored: they are regenerated from the seeds with
``oracle.rrt_oracle.make_weights`` / ``make_bag`` (numpy legacy ``RandomState`` streams, which
are frozen across numpy versions and platforms).  Stored per case: the config, the seeds, the
reference output (all rows for small bags, every ``row_stride``-th row for large ones) and
whole-output checksums (per-row sums and sums of squares, per-channel column sums, the
Frobenius norm) so that a sampled fixture still constrains every row.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import _reference_shim as shim
from . import rrt_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, L, config overrides, bag seed, bag kind.  At most MAX_ROWS evenly strided rows are stored.
CASES = [
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    ("c1_n512_d512", 512, dict(), 7, "randn"),
    # BASELINE.json configs[1] shape (C16-PLIP): N=9000 D=512 region_num=8
    ("c2_n9000_d512", 9000, dict(), 7, "randn"),
    # configs[4] encoder shape: epeg_k=21 crmsa_k=5
    ("c5_n9000_k21_c5", 9000, dict(epeg_k=21, crmsa_k=5), 7, "relu"),
    # configs[3]: long survival bag, region_num=16 (CR-MSA still partitions 8x8, P=784)
    ("c4_n50000_g16", 50000, dict(region_num=16), 7, "randn"),
    # README per-dataset settings (modules/rrt.py:252-258)
    ("plip_k9_shortcut", 1337, dict(epeg_k=9, crmsa_k=3, all_shortcut=True), 7, "relu"),
    ("r50_k15_c1_shortcut", 2000, dict(epeg_k=15, crmsa_k=1, all_shortcut=True), 7, "relu"),
    ("nsclc_plip_mlp_h1", 1500, dict(epeg_k=13, crmsa_k=3, crmsa_heads=1, all_shortcut=True,
                                     crmsa_mlp=True), 7, "randn"),
    ("brca_r50_h1", 1111, dict(epeg_k=17, crmsa_k=3, crmsa_heads=1), 7, "randn"),
    # degenerate / ragged bags
    ("tiny_n1", 1, dict(), 7, "randn"),
    ("tiny_n50", 50, dict(all_shortcut=True), 7, "randn"),
    ("n63", 63, dict(), 7, "randn"),
    ("n64_square", 64, dict(), 7, "randn"),
    ("n65", 65, dict(), 7, "randn"),
    ("n576_exact_grid", 576, dict(), 7, "randn"),
    # option coverage at a smaller width (D=256 -> head_dim 32)
    ("d256_g4", 777, dict(mlp_dim=256, region_num=4, epeg_k=7, crmsa_k=2, crmsa_heads=4), 8, "randn"),
    ("d256_nobias_noepeg", 900, dict(mlp_dim=256, qkv_bias=False, epeg=False), 7, "randn"),
    ("d256_three_layers_nocr", 900, dict(mlp_dim=256, n_layers=3, cr_msa=False), 7, "randn"),
    ("d256_region_size5", 600, dict(mlp_dim=256, region_size=5), 7, "randn"),
    ("d256_min_region_num", 600, dict(mlp_dim=256, min_region_num=700), 7, "randn"),
    ("d256_min_region_ratio", 300, dict(mlp_dim=256, min_region_ratio=5.0, region_num=16), 7, "randn"),
    ("d128_h2_k3", 400, dict(mlp_dim=128, n_heads=2, crmsa_heads=2, epeg_k=3), 7, "randn"),
]

# RRTMIL end-to-end: name, L, input_dim, n_classes, act, da_act, da_bias, encoder overrides
MIL_CASES = [
    ("mil_r50_n3000", 3000, 1024, 2, "relu", "relu", False, dict()),
    ("mil_plip_n1200_gelu_bias", 1200, 512, 4, "gelu", "tanh", True, dict(epeg_k=9, all_shortcut=True)),
    ("mil_n9000_k21", 9000, 1024, 2, "relu", "relu", False, dict(epeg_k=21, crmsa_k=5)),
]

MAX_ROWS = 96
WEIGHT_SEED = 2021  # the reference's default --seed (main.py:645)
MIN_LOGIT_RANGE = 1e-2  # see crmsa_conditioning


def crmsa_conditioning(x, w, cfg):
    """Smallest (max - min) of the CR-MSA logits over the regions that hold a real token and more
    than one slot.  The reference's min-max dispatch weight (modules/rmsa.py:312-314) is a step
    function of the logit ordering when that range collapses (e.g. a region with ONE real token next
    to zero pads: weight = 1 if logit > 0 else 0), so a fixture whose range is ~0 sits on a
    discontinuity of the reference function and pins nothing but rounding noise."""
    if not cfg.cr_msa:
        return float("inf")
    h = x
    for i in range(cfg.n_layers - 1):
        p = f"layers.{i}."
        h = h + O.rmsa_block(O.layer_norm(h, w[p + "norm.weight"], w[p + "norm.bias"]), w,
                             p + "attn.", cfg, "spec")
    L = h.shape[0]
    H, rs, _ = O.grid_geometry(L, 8)
    if rs == 1:
        return float("inf")
    z = O._to_regions(O.layer_norm(h, w["cr_msa.norm.weight"], w["cr_msa.norm.bias"]), L, H, rs)
    if cfg.crmsa_mlp:
        lg = torch.tanh(z @ w["cr_msa.attn.phi.0.weight"].T) @ w["cr_msa.attn.phi.2.weight"].T
    else:
        lg = z @ w["cr_msa.attn.phi"]
    real = (O.region_slot_map(H, rs) < L).view(-1, rs * rs).any(1)
    rng = (lg.max(1).values - lg.min(1).values)[real]
    return float(rng.min())


def generate(name, L, overrides, bag_seed, kind):
    cfg = O.EncoderConfig(**overrides)
    w = O.make_weights(cfg, WEIGHT_SEED)
    x = O.make_bag(L, cfg.mlp_dim, bag_seed, kind=kind)
    cond = crmsa_conditioning(x, w, cfg)
    if cond < MIN_LOGIT_RANGE:
        raise SystemExit(f"{name}: CR-MSA logit range {cond:.2e} < {MIN_LOGIT_RANGE}: fixture sits on a "
                         "discontinuity of the reference's min-max dispatch; pick another bag seed")
    model = shim.build_reference_encoder(cfg, w)
    with torch.no_grad():
        y = model(x.unsqueeze(0))[0]
    y = y.numpy()
    row_stride = max(1, -(-L // MAX_ROWS))
    rows = np.arange(0, L, row_stride)
    np.savez(os.path.join(GOLDEN_DIR, name + ".npz"),
             out_rows=y[rows].astype(np.float32), row_index=rows.astype(np.int64),
             row_sum=y.sum(1).astype(np.float32), row_sqsum=(y * y).sum(1).astype(np.float32),
             col_sum=y.sum(0), fro=np.array(np.linalg.norm(y)))
    return dict(name=name, L=L, config=cfg.to_dict(), row_stride=row_stride, bag_kind=kind,
                weight_seed=WEIGHT_SEED, bag_seed=bag_seed, min_crmsa_logit_range=cond)


def generate_mil(name, L, input_dim, n_classes, act, da_act, da_bias, overrides):
    cfg = O.EncoderConfig(**overrides)
    w = O.make_mil_weights(cfg, input_dim, n_classes, WEIGHT_SEED, da_bias=da_bias)
    x = O.make_bag(L, input_dim, 7, kind="randn")
    ref = shim.import_reference_rrt()
    m = ref.RRTMIL(input_dim=input_dim, n_classes=n_classes, act=act, da_act=da_act, da_bias=da_bias,
                   region_num=cfg.region_num, n_layers=cfg.n_layers, epeg_k=cfg.epeg_k,
                   crmsa_k=cfg.crmsa_k, all_shortcut=cfg.all_shortcut, crmsa_heads=cfg.crmsa_heads,
                   crmsa_mlp=cfg.crmsa_mlp).double().eval()
    m.load_state_dict(w, strict=True)
    with torch.no_grad():
        logits, attn = m(x.unsqueeze(0), return_attn=True)
    np.savez(os.path.join(GOLDEN_DIR, name + ".npz"), logits=logits[0].numpy(),
             attn=attn[0].numpy().astype(np.float32))
    return dict(name=name, L=L, input_dim=input_dim, n_classes=n_classes, act=act, da_act=da_act,
                da_bias=da_bias, config=cfg.to_dict(), weight_seed=WEIGHT_SEED, bag_seed=7)


def main():
    if not shim.available():
        sys.exit("reference tree not present; goldens can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_grad_enabled(False)
    manifest = []
    for case in CASES:
        manifest.append(generate(*case))
        print("golden", case[0], flush=True)
    ref_commit = None
    sub = os.path.join(shim.REFERENCE_ROOT, ".SUBMODULES.json")
    if os.path.isfile(sub):
        ref_commit = json.load(open(sub)).get("commit")
    mil = []
    for case in MIL_CASES:
        mil.append(generate_mil(*case))
        print("golden", case[0], flush=True)
    json.dump(dict(reference_commit=ref_commit, torch=torch.__version__, numpy=np.__version__,
                   cases=manifest, mil_cases=mil),
              open(os.path.join(GOLDEN_DIR, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
