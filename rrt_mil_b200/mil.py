"""``RRTMIL`` -- the reference's end-to-end model (modules/rrt.py:204-246) on the B200 kernels:
``patch_to_emb`` (Linear + act) -> ``RRTEncoder`` -> ``DAttention`` pooling -> ``predictor``.
SURVEY.md 8(f) rows f1 (pooling head) and f2 (patch_to_emb front end).

Same constructor keywords, parameter tree and ``state_dict`` keys as the reference, so reference
checkpoints load with ``strict=True``.  Eval mode / ``torch.no_grad()`` run the inference kernels; grad
mode or ``.train()`` run the training path: ``patch_to_emb`` + ``dp`` dropout, the taped
encoder and the pooling head each behind a ``torch.autograd.Function`` whose backward is CUDA
(``rrt_patch_embed_backward``, ``rrt_encoder_backward``, ``rrt_attn_pool_backward``), so
``loss.backward()`` fills the gradient of every parameter.  Options the kernels do not cover raise.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import cabi
from .encoder import RRTEncoder, initialize_weights

_ACT = {"relu": (nn.ReLU, cabi.RRT_ACT_RELU), "gelu": (nn.GELU, cabi.RRT_ACT_GELU),
        "tanh": (nn.Tanh, cabi.RRT_ACT_TANH)}


class Attention(nn.Module):
    """Parameters of modules/datten.py:5-26 (``attention`` = Linear(L,128)-act-[Dropout]-Linear(128,1))."""

    def __init__(self, input_dim=512, act='relu', bias=False, dropout=False):
        super().__init__()
        self.L, self.D, self.K = input_dim, 128, 1
        layers = [nn.Linear(self.L, self.D, bias=bias)]
        self.act_code = cabi.RRT_ACT_NONE
        if act in _ACT:
            layers.append(_ACT[act][0]())
            self.act_code = _ACT[act][1]
        self.drop_p = 0.25 if dropout else 0.0
        if dropout:
            layers.append(nn.Dropout(0.25))
        layers.append(nn.Linear(self.D, self.K, bias=bias))
        self.attention = nn.Sequential(*layers)

    def head_params(self):
        """(first-layer Linear modules, score Linear, act code for the C entry)."""
        return [self.attention[0]], self.attention[-1], self.act_code


class AttentionGated(nn.Module):
    """Parameters of modules/datten.py:40-65: ``attention_a`` = Linear-act-[Dropout], ``attention_b`` =
    Linear-Sigmoid-[Dropout], ``attention_c`` = Linear(128,1); scores = attention_c(a * b)."""

    def __init__(self, input_dim=512, act='relu', bias=False, dropout=False):
        super().__init__()
        self.L, self.D, self.K = input_dim, 128, 1
        a = [nn.Linear(self.L, self.D, bias=bias)]
        self.act_code = cabi.RRT_ACT_NONE
        if act in _ACT:
            a.append(_ACT[act][0]())
            self.act_code = _ACT[act][1]
        b = [nn.Linear(self.L, self.D, bias=bias), nn.Sigmoid()]
        self.drop_p = 0.25 if dropout else 0.0
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(self.D, self.K, bias=bias)

    def head_params(self):
        return [self.attention_a[0], self.attention_b[0]], self.attention_c, self.act_code | cabi.RRT_ACT_GATED


class DAttention(nn.Module):
    """modules/datten.py:85-101."""

    def __init__(self, input_dim=512, act='relu', gated=False, bias=False, dropout=False):
        super().__init__()
        self.gated = gated
        self.attention = (AttentionGated if gated else Attention)(input_dim, act, bias, dropout)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _flat_views(shapes, device):
    """Gradient tensors of one backward Function as views of ONE buffer (256-byte aligned pieces; ``None`` stays
    ``None``): a data-parallel reducer can then all-reduce the whole group in place (parallel.GradReducer)."""
    offs, total = [], 0
    for shp in shapes:
        offs.append(total)
        if shp is not None:
            total += (int(torch.Size(shp).numel()) + 63) // 64 * 64
    flat = torch.empty(max(total, 1), dtype=torch.float32, device=device)
    return [None if shp is None else flat[o:o + torch.Size(shp).numel()].view(shp) for o, shp in zip(offs, shapes)]


class _PatchEmbedFunction(torch.autograd.Function):
    """h0 = dp(act(x W^T + b)) (modules/rrt.py:228-229); no gradient for the bag's features."""

    @staticmethod
    def forward(ctx, mil, bag, weight, bias):
        lib, dev = cabi.lib(), bag.device
        L, in_dim = bag.shape
        dim = weight.shape[0]
        drop_p, seed = mil._train_dropout()
        with torch.cuda.device(dev):
            n = C.c_size_t()
            cabi.check(lib.rrt_mil_head_workspace_bytes(L, max(in_dim, dim), dim, 128, C.byref(n)), "workspace")
            tape = torch.empty(n.value, dtype=torch.uint8, device=dev)
            out = torch.empty(L, dim, device=dev)
            # nn.GELU: the backward needs the pre-activations (gelu' is not a function of the output)
            pre = torch.empty(L, dim, device=dev) if mil._fc_act == cabi.RRT_ACT_GELU else None
            cabi.check(lib.rrt_patch_embed_forward(bag.data_ptr(), L, in_dim, dim, weight.data_ptr(),
                                                   RRTMIL._p(bias), None, mil._fc_act, out.data_ptr(),
                                                   tape.data_ptr(), n.value, drop_p, seed, RRTMIL._p(pre),
                                                   _stream(dev)),
                       "rrt_patch_embed_forward")
        ctx.mil, ctx.drop, ctx.shape = mil, (drop_p, seed), (L, in_dim, dim)
        ctx.save_for_backward(out, tape, weight, bias, pre)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, tape, weight, bias, pre = ctx.saved_tensors
        lib, dev = cabi.lib(), out.device
        L, in_dim, dim = ctx.shape
        dout = dout.contiguous().float()
        with torch.cuda.device(dev):
            dw, db = _flat_views([weight.shape, None if bias is None else bias.shape], dev)
            nws = 512 + L * dim * 2 + 256
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
            cabi.check(lib.rrt_patch_embed_backward(dout.data_ptr(), out.data_ptr(), RRTMIL._p(pre), L, in_dim, dim,
                                                    ctx.mil._fc_act, ctx.drop[0], ctx.drop[1], tape.data_ptr(),
                                                    tape.numel(), dw.data_ptr(), RRTMIL._p(db), ws.data_ptr(),
                                                    nws, _stream(dev)), "rrt_patch_embed_backward")
        return None, None, dw, db


class _AttnPoolFunction(torch.autograd.Function):
    """logits = predictor(DAttention(h)) (modules/datten.py:28-38,66-83, modules/rrt.py:241).  ``wb`` / ``bb`` are
    the gate branch of AttentionGated (None for the plain head): the C entry takes [W_a; W_b] as one matrix."""

    @staticmethod
    def forward(ctx, mil, h, wa, ba, wb, bb, w2, b2, pw, pb):
        lib, dev = cabi.lib(), h.device
        L, dim = h.shape
        gated = wb is not None
        w1 = torch.cat([wa, wb]) if gated else wa
        b1 = (torch.cat([ba, bb]) if gated else ba) if ba is not None else None
        hid, n1, ncls = wa.shape[0], w1.shape[0], pw.shape[0]
        act = mil.pool_fn.attention.head_params()[2]
        drop_p, seed = mil._pool_dropout()
        with torch.cuda.device(dev):
            n = C.c_size_t()
            cabi.check(lib.rrt_mil_head_workspace_bytes(L, dim, dim, n1, C.byref(n)), "workspace")
            tape = torch.empty(n.value, dtype=torch.uint8, device=dev)
            pooled, logits = torch.empty(dim, device=dev), torch.empty(ncls, device=dev)
            pre = torch.empty(L, hid, device=dev) if (act & 0xff) == cabi.RRT_ACT_GELU else None
            cabi.check(lib.rrt_attn_pool_forward(h.data_ptr(), L, dim, hid, w1.data_ptr(), RRTMIL._p(b1), None,
                                                 act, w2.data_ptr(), RRTMIL._p(b2),
                                                 pw.data_ptr(), RRTMIL._p(pb), ncls, pooled.data_ptr(),
                                                 logits.data_ptr(), None, 0, drop_p, seed, RRTMIL._p(pre),
                                                 tape.data_ptr(), n.value, _stream(dev)), "rrt_attn_pool_forward")
        ctx.mil, ctx.act, ctx.drop, ctx.gated, ctx.has_b1 = mil, act, (drop_p, seed), gated, b1 is not None
        ctx.save_for_backward(h, tape, pooled, w1, w2, b2, pw, pb, pre)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        h, tape, pooled, w1, w2, b2, pw, pb, pre = ctx.saved_tensors
        lib, dev = cabi.lib(), h.device
        L, dim = h.shape
        n1, ncls = w1.shape[0], pw.shape[0]
        hid = n1 // 2 if ctx.gated else n1
        dlogits = dlogits.contiguous().float()
        with torch.cuda.device(dev):
            n = C.c_size_t()
            cabi.check(lib.rrt_mil_head_backward_workspace_bytes(L, dim, n1, C.byref(n)), "workspace")
            ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
            dh = torch.empty_like(h)
            dw1, db1, dw2, db2, dpw, dpb = _flat_views(
                [w1.shape, (n1,) if ctx.has_b1 else None, w2.shape, None if b2 is None else b2.shape, pw.shape,
                 None if pb is None else pb.shape], dev)
            cabi.check(lib.rrt_attn_pool_backward(h.data_ptr(), L, dim, hid, w1.data_ptr(), ctx.act,
                                                  w2.data_ptr(), pw.data_ptr(), ncls, pooled.data_ptr(),
                                                  dlogits.data_ptr(), ctx.drop[0], ctx.drop[1], RRTMIL._p(pre),
                                                  tape.data_ptr(), tape.numel(), dh.data_ptr(), dw1.data_ptr(),
                                                  RRTMIL._p(db1), dw2.data_ptr(), RRTMIL._p(db2), dpw.data_ptr(),
                                                  RRTMIL._p(dpb), ws.data_ptr(), n.value, _stream(dev)),
                       "rrt_attn_pool_backward")
        if ctx.gated:
            dwa, dwb = dw1[:hid], dw1[hid:]
            dba, dbb = (db1[:hid], db1[hid:]) if db1 is not None else (None, None)
        else:
            dwa, dwb, dba, dbb = dw1, None, db1, None
        return None, dh, dwa, dba, dwb, dbb, dw2, db2, dpw, dpb


class RRTMIL(nn.Module):
    def __init__(self, input_dim=1024, mlp_dim=512, act='relu', n_classes=2, dropout=0.25, pos_pos=0,
                 pos='none', peg_k=7, attn='rmsa', pool='attn', region_num=8, n_layers=2, n_heads=8,
                 drop_path=0., da_act='relu', trans_dropout=0.1, ffn=False, ffn_act='gelu', mlp_ratio=4.,
                 da_gated=False, da_bias=False, da_dropout=False, trans_dim=64, epeg=True,
                 min_region_num=0, qkv_bias=True, **kwargs):
        super().__init__()
        if pool != 'attn':
            raise NotImplementedError("pool != 'attn' is not built (and is shape-broken in the reference)")
        if mlp_dim != 512:
            raise ValueError("the reference's patch_to_emb always emits 512 channels (modules/rrt.py:208)")
        layers = [nn.Linear(input_dim, 512)]
        self._fc_act = cabi.RRT_ACT_NONE
        if act.lower() in ("relu", "gelu"):
            layers.append(_ACT[act.lower()][0]())
            self._fc_act = _ACT[act.lower()][1]
        self.dp = nn.Dropout(dropout) if dropout > 0. else nn.Identity()
        self.patch_to_emb = nn.Sequential(*layers)
        self.online_encoder = RRTEncoder(mlp_dim=mlp_dim, pos_pos=pos_pos, pos=pos, peg_k=peg_k, attn=attn,
                                         region_num=region_num, n_layers=n_layers, n_heads=n_heads,
                                         drop_path=drop_path, drop_out=trans_dropout, ffn=ffn,
                                         ffn_act=ffn_act, mlp_ratio=mlp_ratio, trans_dim=trans_dim,
                                         epeg=epeg, min_region_num=min_region_num, qkv_bias=qkv_bias,
                                         **kwargs)
        self.pool_fn = DAttention(self.online_encoder.final_dim, da_act, gated=da_gated, bias=da_bias,
                                  dropout=da_dropout)
        self.predictor = nn.Linear(self.online_encoder.final_dim, n_classes)
        self.apply(initialize_weights)
        self._shadow = {}
        self.dropout_p = float(dropout)
        self._dropout_seed = None   # tests: pin the dp seed (the encoder has its own ``_dropout_seed``)

    # ------------------------------------------------------------------------------------------
    def _f16(self, param):
        key = id(param)
        ent = self._shadow.get(key)
        if ent is None or ent[0] != (param.data_ptr(), param._version):
            buf = torch.empty(param.shape, dtype=torch.float16, device=param.device)
            cabi.check(cabi.lib().rrt_convert_f16(param.data_ptr(), buf.data_ptr(), param.numel(),
                                                  torch.cuda.current_stream(param.device).cuda_stream),
                       "rrt_convert_f16")
            ent = ((param.data_ptr(), param._version), buf)
            self._shadow[key] = ent
        return ent[1].data_ptr()

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def _train_dropout(self):
        """(p, seed) of ``dp`` for this forward; active in training mode only."""
        if not self.training or self.dropout_p <= 0.0:
            return 0.0, 0
        seed = self._dropout_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        return self.dropout_p, seed

    def _pool_dropout(self):
        """(p, seed) of the nn.Dropout(0.25) inside the pooling head's score MLP (da_dropout=True); training only.
        Same step seed as ``dp``: the two masks come from different counter streams."""
        p = self.pool_fn.attention.drop_p
        if not self.training or p <= 0.0:
            return 0.0, 0
        if self._dropout_seed is not None:
            return p, self._dropout_seed
        return p, int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    def _head_weights(self):
        """Inference operands of the pooling head: (w1 fp32, b1, w1 fp16 shadow, score Linear, act code, hid).
        The gated head's two first layers are concatenated once per parameter version."""
        firsts, score, act = self.pool_fn.attention.head_params()
        if len(firsts) == 1:
            a0 = firsts[0]
            return a0.weight, a0.bias, self._f16(a0.weight), score, act, a0.out_features
        ps = [q for l in firsts for q in (l.weight, l.bias) if q is not None]
        key = tuple((q.data_ptr(), q._version) for q in ps)
        ent = self._shadow.get("gated")
        if ent is None or ent[0] != key:
            with torch.no_grad():
                w1 = torch.cat([l.weight for l in firsts]).contiguous()
                b1 = torch.cat([l.bias for l in firsts]).contiguous() if firsts[0].bias is not None else None
                w16 = torch.empty(w1.shape, dtype=torch.float16, device=w1.device)
                cabi.check(cabi.lib().rrt_convert_f16(w1.data_ptr(), w16.data_ptr(), w1.numel(), _stream(w1.device)),
                           "rrt_convert_f16")
            ent = (key, w1, b1, w16)
            self._shadow["gated"] = ent
        return ent[1], ent[2], ent[3].data_ptr(), score, act, firsts[0].out_features

    def _forward_train(self, bag):
        """Autograd / training path: three CUDA-backed autograd Functions in a row."""
        fc, pred = self.patch_to_emb[0], self.predictor
        firsts, score, _ = self.pool_fn.attention.head_params()
        if fc.out_features % 128 or firsts[0].out_features != 128:
            raise NotImplementedError("training needs a 128-aligned patch_to_emb width and the reference's "
                                      "128-wide pooling MLP")
        h0 = _PatchEmbedFunction.apply(self, bag, fc.weight, fc.bias)
        h1 = self.online_encoder.forward_bag(h0)
        gate = firsts[1] if len(firsts) == 2 else None
        logits = _AttnPoolFunction.apply(self, h1, firsts[0].weight, firsts[0].bias,
                                         None if gate is None else gate.weight,
                                         None if gate is None else gate.bias,
                                         score.weight.view(-1), score.bias, pred.weight, pred.bias)
        return logits.unsqueeze(0)

    def forward(self, x, return_attn=False, no_norm=False):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.dim() != 3 or x.shape[0] != 1:
            raise ValueError("RRTMIL processes one bag [1, N, input_dim] per call")
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("RRTMIL (rrt_mil_b200) needs float32 CUDA input; there is no CPU fallback")
        bag = x[0].contiguous()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.training and return_attn:
            raise NotImplementedError("return_attn is an inference option: call .eval() first")
        if (needs_grad or self.training) and not return_attn:
            return self._forward_train(bag)
        # inference kernels (also for eval-mode attention maps requested outside torch.no_grad(): the map and
        # the logits returned with it carry no autograd graph)
        L, in_dim = bag.shape
        dev = bag.device
        lib = cabi.lib()
        fc, pred = self.patch_to_emb[0], self.predictor
        w1, b1, w1_f16, score, act, hid = self._head_weights()
        dim, ncls = self.online_encoder.final_dim, pred.out_features
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            n = C.c_size_t()
            cabi.check(lib.rrt_mil_head_workspace_bytes(L, max(in_dim, dim), dim, w1.shape[0], C.byref(n)),
                       "workspace")
            ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
            h0 = torch.empty(L, dim, device=dev)
            cabi.check(lib.rrt_patch_embed_forward(bag.data_ptr(), L, in_dim, dim, fc.weight.data_ptr(),
                                                   self._p(fc.bias), self._f16(fc.weight), self._fc_act,
                                                   h0.data_ptr(), ws.data_ptr(), n.value, 0.0, 0, None, st),
                       "rrt_patch_embed_forward")
            h1 = self.online_encoder.forward_bag(h0)
            pooled = torch.empty(dim, device=dev)
            logits = torch.empty(ncls, device=dev)
            attn = torch.empty(L, device=dev) if return_attn else None
            cabi.check(lib.rrt_attn_pool_forward(h1.data_ptr(), L, dim, hid, w1.data_ptr(),
                                                 self._p(b1), w1_f16, act,
                                                 score.weight.data_ptr(), self._p(score.bias),
                                                 pred.weight.data_ptr(), self._p(pred.bias), ncls,
                                                 pooled.data_ptr(), logits.data_ptr(), self._p(attn),
                                                 int(bool(no_norm)), 0.0, 0, None, ws.data_ptr(), n.value, st),
                       "rrt_attn_pool_forward")
        if return_attn:
            return logits.unsqueeze(0), attn.unsqueeze(0)
        return logits.unsqueeze(0)
