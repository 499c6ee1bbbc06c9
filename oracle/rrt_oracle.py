"""CPU restatement of DearCaat/RRT-MIL's ``RRTEncoder.forward`` (eval mode, one bag).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product never imports this.

Parity status: the reference repository holds no tests, golden vectors or known-answer
fixtures for this path.  The oracle is therefore pinned against
OUTPUTS OF THE REFERENCE ITSELF, imported read-only in the build container
(``oracle/_reference_shim.py``): ``oracle/make_golden.py`` runs the real
``modules.rrt.RRTEncoder`` in float64 on seeded inputs and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks both restatements below against
those fixtures (and, when ``/root/reference`` is present, against the live reference).

Two restatements of the same math, both plain PyTorch on CPU, any float dtype:

* ``order="reference"`` follows the reference's operator sequence one-for-one (pad with
  zeros after LayerNorm, gather into regions, materialise the ``[R,h,P,P]`` logit map,
  depthwise ``Conv2d`` (k,1) on that map, the ``[R,k,P,D]`` CR-MSA tensors ...).  This is
  what the CPU baseline in ``bench.py`` times: it issues the same ATen work as the
  reference's own forward.
* ``order="spec"`` is the algebraically reduced form the CUDA kernels implement (EPEG as a
  depthwise 1-D conv on the scaled Q, CR-MSA as two rank-k products).

Weights travel as a flat ``dict`` keyed by the reference's ``state_dict`` names.

Reference lines followed (``/root/reference``):
  modules/rrt.py:108-131   TransLayer.forward_trans      -> ``_residual_block``
  modules/rrt.py:165-202   RRTEncoder.forward            -> ``encoder_forward``
  modules/rmsa.py:28-54    region_partition / reverse    -> ``region_slot_map``
  modules/rmsa.py:91-134   InnerAttention.forward        -> ``inner_attention``
  modules/rmsa.py:175-230  RegionAttntion.padding/forward-> ``grid_geometry``, ``rmsa_block``
  modules/rmsa.py:261-337  CrossRegionAttntion           -> ``crmsa_block``
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5  # nn.LayerNorm default, modules/rrt.py:47,139


@dataclass
class EncoderConfig:
    """Constructor options of ``RRTEncoder`` that reach the hot path (modules/rrt.py:134)."""

    mlp_dim: int = 512
    region_num: int = 8
    n_layers: int = 2
    n_heads: int = 8
    epeg: bool = True
    epeg_k: int = 15
    epeg_2d: bool = False        # ablation: k x k kernel instead of (k, 1)  (modules/rmsa.py:76-87)
    epeg_type: str = "attn"      # ablation: 'attn' | 'value_bf' | 'value_af'  (modules/rmsa.py:104-129)
    region_size: int = 0
    min_region_num: int = 0
    min_region_ratio: float = 0.0
    qkv_bias: bool = True
    epeg_bias: bool = True
    cr_msa: bool = True
    crmsa_k: int = 3
    all_shortcut: bool = False
    crmsa_mlp: bool = False
    crmsa_heads: int = 8
    # ablation positional encoding (modules/rrt.py:150-160)
    pos: str = "none"
    pos_pos: int = 0
    peg_k: int = 7
    peg_bias: bool = True
    peg_1d: bool = False
    # ablation FFN (modules/rrt.py:25-41,106,128-129)
    ffn: bool = False
    ffn_act: str = "gelu"
    mlp_ratio: float = 4.0

    def to_dict(self):
        return asdict(self)


# ----------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------
def _ceil_sqrt(n: int) -> int:
    return 0 if n <= 0 else math.isqrt(n - 1) + 1


def grid_geometry(L: int, region_num: int, region_size: int = 0, min_region_num: int = 0,
                  min_region_ratio: float = 0.0) -> Tuple[int, int, int]:
    """(H, rs, add_length): side of the padded square grid, side of one region, pad tokens.

    modules/rmsa.py:175-198 (identical code at :261-284 for CR-MSA).
    """
    H = _ceil_sqrt(L)
    if region_size and region_size > 0:
        H += (-H) % region_size
        rs = region_size
    else:
        H += (-H) % region_num
        rs = H // region_num
    add = H * H - L
    # "if padding much, give up region attention" escape (never fires with the 0/0 defaults)
    if add > L / (min_region_ratio + 1e-8) or L < min_region_num:
        H = _ceil_sqrt(L)
        H += (-H) % 2
        add = H * H - L
        rs = H
    return H, rs, add


def region_slot_map(H: int, rs: int) -> torch.Tensor:
    """slot -> padded token index.  Slot ``rho*P + p`` (region-major) holds grid cell
    ``(row, col)`` with ``rho = (row//rs)*(H//rs) + col//rs`` and ``p = (row%rs)*rs + col%rs``
    (modules/rmsa.py:37-38)."""
    g = H // rs
    slot = torch.arange(H * H)
    rho, p = slot // (rs * rs), slot % (rs * rs)
    row = (rho // g) * rs + p // rs
    col = (rho % g) * rs + p % rs
    return row * H + col


# ----------------------------------------------------------------------------------------
# training-mode proj_drop (modules/rmsa.py:70,132)
# ----------------------------------------------------------------------------------------
DROP_STREAM_CRMSA = 64  # == RRT_DROP_STREAM_CRMSA (include/rrt_b200.h); R-MSA layer i uses stream i
DROP_STREAM_PATCH = 65  # == RRT_DROP_STREAM_PATCH: RRTMIL.dp behind patch_to_emb
DROP_STREAM_POOL = 66   # == RRT_DROP_STREAM_POOL: nn.Dropout inside the pooling head's score MLP (da_dropout)
_M64 = (1 << 64) - 1


def dropout_mask(rows: int, D: int, p: float, seed: int, stream: int, dtype=torch.float64) -> torch.Tensor:
    """keep/(1-p) factors [rows, D] of the library's counter-based dropout (csrc/common.cuh):
    one splitmix64 hash per 4 consecutive elements (flat index), 16 bits per element, dropped when
    the 16-bit field is < round(p * 65536).  It is ``nn.Dropout(p)`` in distribution; the reference
    draws its masks from torch's Philox stream, which no other implementation can reproduce, so
    parity is checked by giving the reference / the oracle THIS mask (oracle/make_golden.py)."""
    assert (rows * D) % 4 == 0
    p32 = np.float32(p)
    if not p32 > 0:
        return torch.ones(rows, D, dtype=dtype)
    thresh = min(int(float(p32) * 65536.0 + 0.5), 65535)
    scale = float(np.float32(1.0) / (np.float32(1.0) - p32))
    key = np.uint64((seed + (stream + 1) * 0x9E3779B97F4A7C15) & _M64)
    with np.errstate(over="ignore"):
        z = np.arange(rows * D // 4, dtype=np.uint64) + key
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    fields = np.stack([(z >> np.uint64(16 * j)) & np.uint64(0xFFFF) for j in range(4)], 1).reshape(rows, D)
    return torch.from_numpy(np.where(fields >= thresh, scale, 0.0)).to(dtype)


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def layer_norm(x, w, b):
    # the same ATen operator nn.LayerNorm dispatches to (biased variance, eps inside the sqrt)
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def _to_regions(z: torch.Tensor, L: int, H: int, rs: int) -> torch.Tensor:
    """[L,D] -> zero-pad to H*H tokens -> [R,P,D] (modules/rmsa.py:199-215)."""
    D, g = z.shape[-1], H // rs
    zp = torch.cat([z, torch.zeros(H * H - L, D, dtype=z.dtype, device=z.device)]) if H * H > L else z
    # grid (row, col) -> (region row, row in region, region col, col in region); the slot order of
    # ``region_slot_map`` is exactly this axis swap
    return zp.view(g, rs, g, rs, D).transpose(1, 2).reshape(g * g, rs * rs, D)


def _from_regions(y: torch.Tensor, L: int, H: int, rs: int) -> torch.Tensor:
    """inverse of ``_to_regions`` followed by dropping the pad tail (modules/rmsa.py:221-228)."""
    D, g = y.shape[-1], H // rs
    return y.reshape(g, g, rs, rs, D).transpose(1, 2).reshape(H * H, D)[:L]


def inner_attention(xr: torch.Tensor, w: Dict[str, torch.Tensor], prefix: str, heads: int,
                    order: str, epeg_type: str = "attn") -> torch.Tensor:
    """Multi-head attention over each row-block of ``xr`` [B_,S,D] (modules/rmsa.py:91-134).

    ``prefix`` points at the InnerAttention (``...attn.attn.``).  EPEG is applied when
    ``prefix + 'pe.weight'`` exists; its kernel is (k, 1) or k x k (``epeg_2d``), read off the weight's shape.
    ``epeg_type``: 'attn' = conv on the logit map; 'value_bf' / 'value_af' = depthwise conv on V folded into the
    region's sqrt(S) x sqrt(S) grid, added to V before resp. to the output after the attention
    (modules/rmsa.py:114-129, including the reference's channel reinterpretation: the conv sees channel
    d*heads + h, its result is read back as channel h'*d_head + d').  Eval mode: both dropouts are identities.
    """
    B_, S, D = xr.shape
    d = D // heads
    qkv = F.linear(xr, w[prefix + "qkv.weight"], w.get(prefix + "qkv.bias"))
    qkv = qkv.view(B_, S, 3, heads, d).permute(2, 0, 3, 1, 4)  # [3,B_,h,S,d]
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
    pe_w = w.get(prefix + "pe.weight")
    pe_b = w.get(prefix + "pe.bias")
    on_logits = pe_w is not None and epeg_type == "attn"
    two_d = pe_w is not None and pe_w.shape[3] > 1

    def value_pe():
        side = _ceil_sqrt(S)
        kk = pe_w.shape[2]
        img = v.permute(0, 3, 1, 2).reshape(B_, D, side, side)
        return F.conv2d(img, pe_w, pe_b, padding=(kk // 2, kk // 2 if two_d else 0), groups=D)

    if order == "reference" or (on_logits and two_d):
        logits = q @ k.transpose(-1, -2)  # [B_,h,S,S]
        if on_logits:
            kk = pe_w.shape[2]
            logits = logits + F.conv2d(logits, pe_w, pe_b, padding=(kk // 2, kk // 2 if two_d else 0),
                                       groups=heads)
    else:
        if on_logits:
            kk = pe_w.shape[2]
            # depthwise conv along the token axis of q, one tap vector per head; the conv
            # bias is constant along the key axis and vanishes in the softmax.
            taps = pe_w.view(heads, 1, kk).repeat_interleave(d, 0)  # [h*d,1,kk]
            qc = q.permute(0, 1, 3, 2).reshape(B_, heads * d, S)
            qc = F.conv1d(qc, taps, padding=kk // 2, groups=heads * d)
            q = q + qc.view(B_, heads, d, S).permute(0, 1, 3, 2)
        logits = q @ k.transpose(-1, -2)
    attn = torch.softmax(logits, -1)
    if pe_w is not None and epeg_type == "value_bf":
        v = v + value_pe().reshape(B_, heads, d, S).permute(0, 1, 3, 2)
    o = (attn @ v).transpose(1, 2).reshape(B_, S, D)  # [B_,S,D]
    if pe_w is not None and epeg_type == "value_af":
        o = o + value_pe().reshape(B_, D, S).transpose(-1, -2)
    return F.linear(o, w[prefix + "proj.weight"], w[prefix + "proj.bias"])


def rmsa_block(z, w, prefix, cfg: EncoderConfig, order, mask=None):
    """RegionAttntion.forward on an already-normalised [L,D] (modules/rmsa.py:204-230).
    ``mask`` [L,D]: training-mode proj_drop factors in token order (pad rows are dropped afterwards,
    so their mask is irrelevant)."""
    L = z.shape[0]
    H, rs, _ = grid_geometry(L, cfg.region_num, cfg.region_size, cfg.min_region_num,
                             cfg.min_region_ratio)
    y = inner_attention(_to_regions(z, L, H, rs), w, prefix + "attn.", cfg.n_heads, order, cfg.epeg_type)
    y = _from_regions(y, L, H, rs)
    return y if mask is None else y * mask


def crmsa_block(z, w, prefix, cfg: EncoderConfig, order, mask=None):
    """CrossRegionAttntion.forward on an already-normalised [L,D] (modules/rmsa.py:290-337).

    The CR-MSA TransLayer is built without ``n_region`` / ``region_size`` / ``min_region_*``
    (modules/rrt.py:148), so it always partitions with the TransLayer defaults
    (``n_region=8, region_size=0, min_region_num=0, min_region_ratio=0``; modules/rrt.py:44).
    """
    L, D = z.shape
    H, rs, _ = grid_geometry(L, 8, 0, 0, 0.0)
    xr = _to_regions(z, L, H, rs)  # [R,P,D]
    if cfg.crmsa_mlp:
        hid = torch.tanh(xr @ w[prefix + "phi.0.weight"].T)
        logits = (hid @ w[prefix + "phi.2.weight"].T).transpose(1, 2)  # [R,k,P]
    else:
        logits = (xr @ w[prefix + "phi"]).transpose(1, 2)  # [R,k,P]
    combine = logits.softmax(-1)
    dispatch = logits.softmax(1)
    lo = logits.min(-1, keepdim=True).values
    hi = logits.max(-1, keepdim=True).values
    dispatch_mm = (logits - lo) / (hi - lo + 1e-8)
    if order == "reference":
        lm = (xr.unsqueeze(1) * combine.unsqueeze(-1)).sum(-2)  # [R,k,P,D] -> [R,k,D]
    else:
        lm = combine @ xr  # [R,k,D]
    lm = inner_attention(lm.transpose(0, 1), w, prefix + "attn.", cfg.crmsa_heads, order)  # [k,R,D]
    if mask is not None:  # proj_drop of the landmark MHA: factors [k*R, D], row = n*R + rho
        lm = lm * mask.view(lm.shape)
    lm = lm.transpose(0, 1)  # [R,k,D]
    if order == "reference":
        y = lm.unsqueeze(2) * dispatch_mm.unsqueeze(-1)  # [R,k,P,D]
        y = (y * dispatch.unsqueeze(-1)).sum(1)
    else:
        y = (dispatch_mm * dispatch).transpose(1, 2) @ lm  # [R,P,D]
    return _from_regions(y, L, H, rs)


def pos_embedding(x: torch.Tensor, w: Dict[str, torch.Tensor], cfg: EncoderConfig) -> torch.Tensor:
    """PEG / PPEG on one bag [L,D] (modules/emb_position.py:36-58, 66-82): fold the tokens row-major
    into a ceil(sqrt(L)) square, filling the tail with the FIRST tokens; PPEG zero-extends grids smaller
    than 7x7; depthwise "same" convs (k; PPEG also 5 and 3) plus the identity; drop the fill."""
    L, D = x.shape
    H = _ceil_sqrt(L)
    add = H * H - L
    g = torch.cat([x, x[:add]]) if add > 0 else x
    if cfg.pos == "ppeg" and H < 7:
        g = torch.cat([g, torch.zeros(49 - H * H, D, dtype=x.dtype, device=x.device)])
        H = 7
    feat = g.t().reshape(1, D, H, H)
    out = feat
    for name, k in (("proj", cfg.peg_k),) + ((("proj1", 5), ("proj2", 3)) if cfg.pos == "ppeg" else ()):
        pad = (k // 2, 0) if cfg.peg_1d else k // 2
        out = out + F.conv2d(feat, w[f"pos_embedding.{name}.weight"], w.get(f"pos_embedding.{name}.bias"),
                             padding=pad, groups=D)
    return out.reshape(D, H * H).t()[:L]


def ffn_block(h: torch.Tensor, w: Dict[str, torch.Tensor], prefix: str, cfg: EncoderConfig) -> torch.Tensor:
    """``x + mlp(norm2(x))`` of one TransLayer (modules/rrt.py:128-129, 35-41), eval mode."""
    z = layer_norm(h, w[prefix + "norm2.weight"], w[prefix + "norm2.bias"])
    z = F.linear(z, w[prefix + "mlp.fc1.weight"], w[prefix + "mlp.fc1.bias"])
    z = F.gelu(z) if cfg.ffn_act == "gelu" else torch.relu(z)
    return h + F.linear(z, w[prefix + "mlp.fc2.weight"], w[prefix + "mlp.fc2.bias"])


def encoder_forward(x: torch.Tensor, w: Dict[str, torch.Tensor], cfg: EncoderConfig,
                    order: str = "reference", drop=None, branch_scale=None) -> torch.Tensor:
    """``RRTEncoder.forward`` for one bag ``x`` [L,D] -> [L,D] (modules/rrt.py:165-202).
    ``drop=(p, seed)``: training mode with ``drop_out=p`` (proj_drop masks from ``dropout_mask``).
    ``branch_scale``: one factor per block (R-MSA layers, then CR-MSA) on its residual branch -- stochastic depth
    (``drop_path``, modules/rrt.py:102,125) for a batch of one: 0 = dropped, 1 / keep otherwise."""
    assert order in ("reference", "spec")
    assert x.dim() == 2 and x.shape[1] == cfg.mlp_dim
    L, D = x.shape

    def mask(rows, stream):
        return None if drop is None else dropout_mask(rows, D, drop[0], drop[1], stream, x.dtype)

    has_pos = cfg.pos in ("peg", "ppeg")
    h = x
    if has_pos and cfg.pos_pos == -1:                      # modules/rrt.py:181-182
        h = pos_embedding(h, w, cfg)
    for i in range(cfg.n_layers - 1):
        if i == 1 and has_pos and cfg.pos_pos == 0:        # modules/rrt.py:186-187
            h = pos_embedding(h, w, cfg)
        p = f"layers.{i}."
        bs = 1.0 if branch_scale is None else float(branch_scale[i])
        if bs != 0.0:
            h = h + bs * rmsa_block(layer_norm(h, w[p + "norm.weight"], w[p + "norm.bias"]), w,
                                    p + "attn.", cfg, order, mask(L, i))
        if cfg.ffn:
            h = ffn_block(h, w, p, cfg)
    if cfg.cr_msa:
        p = "cr_msa."
        bs = 1.0 if branch_scale is None else float(branch_scale[cfg.n_layers - 1])
        if bs != 0.0:
            h = h + bs * crmsa_block(layer_norm(h, w[p + "norm.weight"], w[p + "norm.bias"]), w,
                                     p + "attn.", cfg, order, mask(cfg.crmsa_k * 64, DROP_STREAM_CRMSA))
        if cfg.ffn:
            h = ffn_block(h, w, p, cfg)
    if cfg.all_shortcut:
        h = h + x
    return layer_norm(h, w["norm.weight"], w["norm.bias"])


# ----------------------------------------------------------------------------------------
# RRTMIL = patch_to_emb -> encoder -> DAttention pooling -> predictor
# ----------------------------------------------------------------------------------------
def _act(name):
    return {"relu": torch.relu, "gelu": F.gelu, "tanh": torch.tanh}.get(name, lambda t: t)


def mil_forward(x: torch.Tensor, w: Dict[str, torch.Tensor], cfg: EncoderConfig, act: str = "relu",
                da_act: str = "relu", order: str = "reference", drop=None, pool_drop=None):
    """``RRTMIL.forward`` (eval) for one bag ``x`` [L, input_dim] -> (logits [C], attention [L])
    (modules/rrt.py:227-246, modules/datten.py:28-38,94-101).  Encoder weights carry the reference's
    ``online_encoder.`` prefix."""
    h = _act(act)(F.linear(x, w["patch_to_emb.0.weight"], w["patch_to_emb.0.bias"]))
    enc_drop = None
    if drop is not None:   # training mode: drop = (dp p, dp seed, trans_dropout p, encoder seed)
        h = h * dropout_mask(h.shape[0], h.shape[1], drop[0], drop[1], DROP_STREAM_PATCH, h.dtype)
        enc_drop = (drop[2], drop[3]) if drop[2] > 0 else None
    enc = {k[len("online_encoder."):]: v for k, v in w.items() if k.startswith("online_encoder.")}
    h = encoder_forward(h, enc, cfg, order, drop=enc_drop)
    # pool_drop = (p, seed): the nn.Dropout(0.25) inside the score MLP (da_dropout=True) in training mode, the
    # library's counter-based mask over the hidden buffer [L, 128] (gated: [L, 256] = act branch | gate branch)
    pa = "pool_fn.attention."
    if pa + "attention_c.weight" in w:                       # AttentionGated (modules/datten.py:66-83)
        ga = _act(da_act)(F.linear(h, w[pa + "attention_a.0.weight"], w.get(pa + "attention_a.0.bias")))
        gb = torch.sigmoid(F.linear(h, w[pa + "attention_b.0.weight"], w.get(pa + "attention_b.0.bias")))
        if pool_drop is not None and pool_drop[0] > 0:
            m = dropout_mask(h.shape[0], 2 * ga.shape[1], pool_drop[0], pool_drop[1], DROP_STREAM_POOL, h.dtype)
            ga, gb = ga * m[:, :ga.shape[1]], gb * m[:, ga.shape[1]:]
        a = F.linear(ga * gb, w[pa + "attention_c.weight"], w.get(pa + "attention_c.bias")).squeeze(-1)
    else:                                                    # Attention (modules/datten.py:28-38)
        keys = sorted(k for k in w if k.startswith(pa + "attention.") and k.endswith("weight"))
        k0, k1 = keys[0], keys[-1]
        a = _act(da_act)(F.linear(h, w[k0], w.get(k0[:-6] + "bias")))
        if pool_drop is not None and pool_drop[0] > 0:
            a = a * dropout_mask(h.shape[0], a.shape[1], pool_drop[0], pool_drop[1], DROP_STREAM_POOL, h.dtype)
        a = F.linear(a, w[k1], w.get(k1[:-6] + "bias")).squeeze(-1)  # [L]
    attn = torch.softmax(a, 0)
    pooled = attn @ h
    logits = F.linear(pooled, w["predictor.weight"], w["predictor.bias"])
    return logits, attn


def make_mil_weights(cfg: EncoderConfig, input_dim: int, n_classes: int, seed: int, da_bias: bool = False,
                     dtype=torch.float64, da_gated: bool = False, da_dropout: bool = False) -> Dict[str, torch.Tensor]:
    enc = make_weights(cfg, seed, dtype)
    rs = np.random.RandomState(seed + 1)

    def lin(o, i):
        return torch.from_numpy(rs.standard_normal((o, i)) * math.sqrt(2.0 / (o + i))).to(dtype)

    def vec(n):
        return torch.from_numpy(0.1 * rs.standard_normal(n)).to(dtype)

    w = {"online_encoder." + k: v for k, v in enc.items()}
    w["patch_to_emb.0.weight"], w["patch_to_emb.0.bias"] = lin(512, input_dim), vec(512)
    if da_gated:      # key names of modules/datten.py:47-65
        pa = "pool_fn.attention."
        w[pa + "attention_a.0.weight"], w[pa + "attention_b.0.weight"] = lin(128, cfg.mlp_dim), lin(128, cfg.mlp_dim)
        w[pa + "attention_c.weight"] = lin(1, 128)
        if da_bias:
            w[pa + "attention_a.0.bias"], w[pa + "attention_b.0.bias"], w[pa + "attention_c.bias"] = \
                vec(128), vec(128), vec(1)
    else:
        last = 3 if da_dropout else 2     # the nn.Dropout shifts the index of the score Linear in the Sequential
        w["pool_fn.attention.attention.0.weight"] = lin(128, cfg.mlp_dim)
        w[f"pool_fn.attention.attention.{last}.weight"] = lin(1, 128)
        if da_bias:
            w["pool_fn.attention.attention.0.bias"], w[f"pool_fn.attention.attention.{last}.bias"] = vec(128), vec(1)
    w["predictor.weight"], w["predictor.bias"] = lin(n_classes, cfg.mlp_dim), vec(n_classes)
    return w


# ----------------------------------------------------------------------------------------
# seeded synthetic weights / inputs (platform-stable: numpy legacy RandomState streams)
# ----------------------------------------------------------------------------------------
def weight_shapes(cfg: EncoderConfig) -> Dict[str, Tuple[int, ...]]:
    D = cfg.mlp_dim
    shp: Dict[str, Tuple[int, ...]] = {"norm.weight": (D,), "norm.bias": (D,)}

    def attn(prefix, epeg):
        shp[prefix + "qkv.weight"] = (3 * D, D)
        if cfg.qkv_bias:
            shp[prefix + "qkv.bias"] = (3 * D,)
        shp[prefix + "proj.weight"] = (D, D)
        shp[prefix + "proj.bias"] = (D,)
        if epeg:
            ch = cfg.n_heads if cfg.epeg_type == "attn" else D     # modules/rmsa.py:76-87
            shp[prefix + "pe.weight"] = (ch, 1, cfg.epeg_k, cfg.epeg_k if cfg.epeg_2d else 1)
            if cfg.epeg_bias:
                shp[prefix + "pe.bias"] = (ch,)

    for i in range(cfg.n_layers - 1):
        shp[f"layers.{i}.norm.weight"] = (D,)
        shp[f"layers.{i}.norm.bias"] = (D,)
        attn(f"layers.{i}.attn.attn.", cfg.epeg)
    if cfg.cr_msa:
        shp["cr_msa.norm.weight"] = (D,)
        shp["cr_msa.norm.bias"] = (D,)
        if cfg.crmsa_mlp:
            shp["cr_msa.attn.phi.0.weight"] = (D // 4, D)
            shp["cr_msa.attn.phi.2.weight"] = (cfg.crmsa_k, D // 4)
        else:
            shp["cr_msa.attn.phi"] = (D, cfg.crmsa_k)
        attn("cr_msa.attn.attn.", False)
    # last, so that the seeded streams of the entries above do not move when pos is switched on
    if cfg.pos in ("peg", "ppeg"):
        for name, k in (("proj", cfg.peg_k),) + ((("proj1", 5), ("proj2", 3)) if cfg.pos == "ppeg" else ()):
            shp[f"pos_embedding.{name}.weight"] = (D, 1, k, 1 if cfg.peg_1d else k)
            if cfg.peg_bias:
                shp[f"pos_embedding.{name}.bias"] = (D,)
    if cfg.ffn:
        hid = int(D * cfg.mlp_ratio)
        for pre in [f"layers.{i}." for i in range(cfg.n_layers - 1)] + (["cr_msa."] if cfg.cr_msa else []):
            shp[pre + "norm2.weight"], shp[pre + "norm2.bias"] = (D,), (D,)
            shp[pre + "mlp.fc1.weight"], shp[pre + "mlp.fc1.bias"] = (hid, D), (hid,)
            shp[pre + "mlp.fc2.weight"], shp[pre + "mlp.fc2.bias"] = (D, hid), (D,)
    return shp


def make_weights(cfg: EncoderConfig, seed: int, dtype=torch.float64,
                 randomize_bias: bool = True) -> Dict[str, torch.Tensor]:
    """Xavier-normal matrices as ``initialize_weights`` draws them (modules/rrt.py:9-23); biases
    and LayerNorm affines are re-drawn N(0,0.1^2) / 1+N(0,0.1^2) so that bias paths are
    exercised (the reference's zero-initialised biases would hide bias bugs)."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in weight_shapes(cfg).items():
        if name.endswith("norm.weight") or name.endswith("norm2.weight"):
            a = 1.0 + 0.1 * rs.standard_normal(shape) if randomize_bias else np.ones(shape)
        elif name.endswith("bias"):
            a = 0.1 * rs.standard_normal(shape) if randomize_bias else np.zeros(shape)
        elif name.startswith("pos_embedding.") and name.endswith("weight"):
            a = rs.standard_normal(shape) * (0.5 / math.sqrt(shape[2] * shape[3]))
        elif name.endswith("pe.weight"):
            fan = shape[2] * shape[3]  # Conv2d(h,h,(k,1) | (k,k),groups=h): fan_in = fan_out = taps per group
            a = rs.standard_normal(shape) * math.sqrt(2.0 / (fan + fan))
        elif name.endswith("phi"):
            bound = 1.0 / math.sqrt(shape[1])  # kaiming_uniform(a=sqrt5) on [D,k]: fan_in=k
            a = rs.uniform(-bound, bound, shape)
        else:
            a = rs.standard_normal(shape) * math.sqrt(2.0 / (shape[0] + shape[1]))
        out[name] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return out


def make_bag(L: int, D: int, seed: int, dtype=torch.float64, kind: str = "randn") -> torch.Tensor:
    rs = np.random.RandomState(seed)
    a = rs.standard_normal((L, D))
    if kind == "relu":  # look-alike of Linear+ReLU+Dropout(0.25) output, modules/rrt.py:208-217
        a = np.maximum(a, 0.0) * (rs.uniform(size=(L, D)) > 0.25) / 0.75
    return torch.from_numpy(a).to(dtype)


def rel_err(y: torch.Tensor, ref: torch.Tensor) -> float:
    y, ref = y.double(), ref.double()
    return float((y - ref).norm() / ref.norm().clamp_min(1e-300))
