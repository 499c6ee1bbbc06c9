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
TRAIN_CASES = {c["name"]: c for c in MANIFEST.get("train_cases", [])}


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


def load_train_case(name, dtype=torch.float64):
    """Training-mode fixture (reference in .train() with the library's dropout masks, + autograd):
    (cfg, weights, x, grad_out, (drop_p, seed), golden arrays)."""
    c = TRAIN_CASES[name]
    cfg = O.EncoderConfig(**c["config"])
    w = O.make_weights(cfg, c["weight_seed"], dtype=dtype)
    x = O.make_bag(c["L"], cfg.mlp_dim, c["bag_seed"], dtype=dtype, kind=c["bag_kind"])
    gout = torch.randn(c["L"], cfg.mlp_dim, generator=torch.Generator().manual_seed(c["grad_seed"]),
                       dtype=torch.float64).to(dtype)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    # (drop_out p, seed, drop_path rate, keep per block) -- the last two None without stochastic depth
    return cfg, w, x, gout, (c["drop_out"], c["dropout_seed"], c.get("drop_path"), c.get("drop_path_keep")), gold


def branch_scales(drop):
    """Per-block branch factors of a train fixture with stochastic depth (None otherwise)."""
    if len(drop) < 4 or drop[2] is None:
        return None
    return [(1.0 / (1.0 - drop[2])) if k else 0.0 for k in drop[3]]


def _rows(a, stride):
    a = a.detach().double().cpu().numpy()
    a2 = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(1, -1)
    return a2[::stride]


def train_errors(y, dx, grads, gold) -> dict:
    """Relative errors (Frobenius over the stored rows, and of the whole-tensor norm) of a training
    forward/backward against a train fixture.  ``grads``: name -> gradient (None = zero)."""
    e = {}
    stride = int(gold["row_stride"])

    def rel(got, ref):
        return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))

    for key, t in (("out", y), ("dx", dx)):
        got = _rows(t, stride)
        e[key] = rel(got, gold[key + "_rows"].astype(np.float64))
        full = t.detach().double().cpu().numpy()
        e[key + "_fro"] = abs(np.linalg.norm(full) - float(gold[key + "_fro"])) / float(gold[key + "_fro"])
        e[key + "_row_sum"] = float(np.abs(full.sum(1) - gold[key + "_row_sum"]).max() /
                                    (float(gold[key + "_fro"]) / np.sqrt(full.shape[0])) / np.sqrt(full.shape[1]))
    for k in gold:
        if not k.startswith("g:"):
            continue
        name = k[2:]
        ref = gold[k].astype(np.float64)
        fro = float(gold["gfro:" + name])
        g = grads.get(name)
        if fro < 1e-12 * max(1.0, float(gold["dx_fro"])):   # exactly-zero reference gradient (pe.bias)
            continue
        assert g is not None, name
        a = g.detach().double().cpu().numpy()
        a2 = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(1, -1)
        st = max(1, -(-a2.shape[0] // 96))
        e["g:" + name] = float(np.linalg.norm(a2[::st] - ref) / max(np.linalg.norm(ref), 1e-300))
        e["gfro:" + name] = abs(np.linalg.norm(a) - fro) / fro
    return e
