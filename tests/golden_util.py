"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load a committed golden
fixture (tests/golden, generated from the unmodified reference by oracle/make_golden.py),
regenerate its seeded inputs, and compare an output against it."""
import json
import os

import numpy as np
import torch

from oracle import rrt_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN_DIR, "manifest.json")))
CASES = {c["name"]: c for c in MANIFEST["cases"]}


def load_case(name, dtype=torch.float64):
    c = CASES[name]
    cfg = O.EncoderConfig(**c["config"])
    w = O.make_weights(cfg, c["weight_seed"], dtype=dtype)
    x = O.make_bag(c["L"], cfg.mlp_dim, c["bag_seed"], dtype=dtype, kind=c["bag_kind"])
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return cfg, w, x, gold


def golden_errors(y: torch.Tensor, gold) -> dict:
    """Relative errors of ``y`` [L,D] against one fixture: the sampled rows (Frobenius-relative
    and max-abs), and the per-row / per-column / whole-tensor checksums that cover every row."""
    y = y.detach().double().cpu().numpy()
    rows = gold["row_index"]
    ref = gold["out_rows"].astype(np.float64)
    got = y[rows]
    e = {}
    e["rows_rel"] = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
    e["rows_maxabs"] = float(np.abs(got - ref).max())
    L, D = y.shape
    # Output rows are LayerNorm rows (rms ~ 1 per element), so a checksum error divided by the
    # number of summed elements is the MEAN error per element of the worst row / column, directly
    # comparable with the relative tolerance.  (tf32 rounding of a weight column gives errors that
    # are correlated down a column, so no sqrt(n) cancellation may be assumed.)  A single wrong row
    # outside the sample moves row_sum by ~sqrt(D)/D = 4e-2 and row_sqsum by O(1).
    e["row_sum"] = float(np.abs(y.sum(1) - gold["row_sum"]).max() / D)
    e["row_sqsum"] = float(np.abs((y * y).sum(1) - gold["row_sqsum"]).max() / D)
    e["col_sum"] = float(np.abs(y.sum(0) - gold["col_sum"]).max() / L)
    e["fro"] = float(abs(np.linalg.norm(y) - float(gold["fro"])) / float(gold["fro"]))
    return e


def assert_matches_golden(y, gold, tol, what=""):
    e = golden_errors(y, gold)
    bad = {k: v for k, v in e.items() if k != "rows_maxabs" and not (v <= tol)}
    assert not bad, f"{what}: golden mismatch beyond {tol:g}: {bad} (all: {e})"
    return e
