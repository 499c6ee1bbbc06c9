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

# name, L, config overrides, (unused), bag kind.  At most MAX_ROWS evenly strided rows are stored.
CASES = [
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    ("c1_n512_d512", 512, dict(), 1, "randn"),
    # BASELINE.json configs[1] shape (C16-PLIP): N=9000 D=512 region_num=8
    ("c2_n9000_d512", 9000, dict(), 36, "randn"),
    # configs[4] encoder shape: epeg_k=21 crmsa_k=5
    ("c5_n9000_k21_c5", 9000, dict(epeg_k=21, crmsa_k=5), 36, "relu"),
    # configs[3]: long survival bag, region_num=16 (CR-MSA still partitions 8x8, P=784)
    ("c4_n50000_g16", 50000, dict(region_num=16), 200, "randn"),
    # README per-dataset settings (modules/rrt.py:252-258)
    ("plip_k9_shortcut", 1337, dict(epeg_k=9, crmsa_k=3, all_shortcut=True), 1, "relu"),
    ("r50_k15_c1_shortcut", 2000, dict(epeg_k=15, crmsa_k=1, all_shortcut=True), 4, "relu"),
    ("nsclc_plip_mlp_h1", 1500, dict(epeg_k=13, crmsa_k=3, crmsa_heads=1, all_shortcut=True,
                                     crmsa_mlp=True), 3, "randn"),
    ("brca_r50_h1", 1111, dict(epeg_k=17, crmsa_k=3, crmsa_heads=1), 2, "randn"),
    # degenerate / ragged bags
    ("tiny_n1", 1, dict(), 1, "randn"),
    ("tiny_n50", 50, dict(all_shortcut=True), 1, "randn"),
    ("n63", 63, dict(), 1, "randn"),
    ("n64_square", 64, dict(), 1, "randn"),
    ("n65", 65, dict(), 1, "randn"),
    ("n576_exact_grid", 576, dict(), 1, "randn"),
    # option coverage at a smaller width (D=256 -> head_dim 32)
    ("d256_g4", 777, dict(mlp_dim=256, region_num=4, epeg_k=7, crmsa_k=2, crmsa_heads=4), 1, "randn"),
    ("d256_nobias_noepeg", 900, dict(mlp_dim=256, qkv_bias=False, epeg=False), 1, "randn"),
    ("d256_three_layers_nocr", 900, dict(mlp_dim=256, n_layers=3, cr_msa=False), 1, "randn"),
    ("d256_region_size5", 600, dict(mlp_dim=256, region_size=5), 1, "randn"),
    ("d256_min_region_num", 600, dict(mlp_dim=256, min_region_num=700), 1, "randn"),
    ("d256_min_region_ratio", 300, dict(mlp_dim=256, min_region_ratio=5.0, region_num=16), 1, "randn"),
    ("d128_h2_k3", 400, dict(mlp_dim=128, n_heads=2, crmsa_heads=2, epeg_k=3), 1, "randn"),
]

MAX_ROWS = 96
WEIGHT_SEED, BAG_SEED = 2021, 7  # 2021 is the reference's default --seed (main.py:645)


def generate(name, L, overrides, row_stride, kind):
    cfg = O.EncoderConfig(**overrides)
    w = O.make_weights(cfg, WEIGHT_SEED)
    x = O.make_bag(L, cfg.mlp_dim, BAG_SEED, kind=kind)
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
                weight_seed=WEIGHT_SEED, bag_seed=BAG_SEED)


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
    json.dump(dict(reference_commit=ref_commit, torch=torch.__version__, numpy=np.__version__,
                   cases=manifest),
              open(os.path.join(GOLDEN_DIR, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
