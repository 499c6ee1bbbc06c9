"""``RRTEncoder`` -- the reference's plug-in surface over the B200 C-ABI library.

Mirrors ``/root/reference/modules/rrt.py:133-202`` (constructor keywords and defaults, sub-module
tree, ``state_dict`` keys, ``final_dim``, 2-D / 3-D / 4-D input handling) so that it drops into the
reference's MIL aggregators (``rrt=<module>``) and loads reference checkpoints with ``strict=True``.
The sub-modules below only OWN parameters (real ``nn.Linear`` / ``nn.Conv2d`` / ``nn.LayerNorm``
objects, so host-side ``initialize_weights`` passes keep working); all arithmetic of the forward
happens in ``librrt_b200.so`` (hand-written sm_100a CUDA) in one C call per bag.

No fallback: CPU tensors, missing library, or options the kernels do not cover raise.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import cabi


def initialize_weights(module: nn.Module) -> None:
    """Same initialisation the reference applies with ``need_init=True`` (modules/rrt.py:9-23):
    Xavier-normal Linear/Conv2d weights, zero biases, LayerNorm (1, 0)."""
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class _ParamHolder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} only owns parameters; call RRTEncoder.forward")


class InnerAttention(_ParamHolder):
    """Parameters of modules/rmsa.py:56-89: ``qkv``, ``proj`` and the EPEG conv ``pe``."""

    def __init__(self, dim, num_heads, qkv_bias, epeg, epeg_k, epeg_bias, epeg_2d=False, epeg_type='attn'):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if epeg:
            # modules/rmsa.py:76-87: per-head conv on the logit map, or per-channel conv on V (value_bf / value_af)
            ch = num_heads if epeg_type == 'attn' else dim
            ks, pad = (epeg_k, epeg_k // 2) if epeg_2d else ((epeg_k, 1), (epeg_k // 2, 0))
            self.pe = nn.Conv2d(ch, ch, ks, padding=pad, groups=ch, bias=epeg_bias)
        else:
            self.pe = None


class RegionAttention(_ParamHolder):
    """modules/rmsa.py:152-173 (``RegionAttntion``)."""

    def __init__(self, dim, num_heads, qkv_bias, epeg, epeg_k, epeg_bias, epeg_2d=False, epeg_type='attn'):
        super().__init__()
        self.attn = InnerAttention(dim, num_heads, qkv_bias, epeg, epeg_k, epeg_bias, epeg_2d, epeg_type)


class CrossRegionAttention(_ParamHolder):
    """modules/rmsa.py:232-259 (``CrossRegionAttntion``): ``phi`` + an InnerAttention without EPEG."""

    def __init__(self, dim, num_heads, qkv_bias, crmsa_k, crmsa_mlp):
        super().__init__()
        self.attn = InnerAttention(dim, num_heads, qkv_bias, False, 0, False)
        if crmsa_mlp:
            self.phi = nn.Sequential(nn.Linear(dim, dim // 4, bias=False), nn.Tanh(),
                                     nn.Linear(dim // 4, crmsa_k, bias=False))
        else:
            self.phi = nn.Parameter(torch.empty(dim, crmsa_k))
            nn.init.kaiming_uniform_(self.phi, a=math.sqrt(5))


class Mlp(_ParamHolder):
    """Parameters of modules/rrt.py:25-41 (the FFN ablation)."""

    def __init__(self, in_features, hidden_features, act_layer, drop):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop = nn.Dropout(drop)


class TransLayer(_ParamHolder):
    """modules/rrt.py:43-106: pre-LayerNorm + attention (+ residual, done in the kernels); with
    ``ffn`` also ``norm2`` + ``mlp``."""

    def __init__(self, dim, attn_module, ffn=False, ffn_act='gelu', mlp_ratio=4., drop_out=0.1):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim) if ffn else nn.Identity()
        self.attn = attn_module
        self.ffn = ffn
        self.mlp = (Mlp(dim, int(dim * mlp_ratio), nn.GELU if ffn_act == 'gelu' else nn.ReLU, drop_out)
                    if ffn else nn.Identity())


class PEG(_ParamHolder):
    """Parameters of modules/emb_position.py:60-82: one depthwise Conv2d ``proj``."""

    def __init__(self, dim=512, k=7, bias=True, conv_1d=False):
        super().__init__()
        ks, pad = ((k, 1), (k // 2, 0)) if conv_1d else (k, k // 2)
        self.proj = nn.Conv2d(dim, dim, ks, 1, pad, groups=dim, bias=bias)


class PPEG(_ParamHolder):
    """Parameters of modules/emb_position.py:24-58: depthwise Conv2d ``proj`` (k), ``proj1`` (5), ``proj2`` (3)."""

    def __init__(self, dim=512, k=7, conv_1d=False, bias=True):
        super().__init__()

        def conv(kk):
            ks, pad = ((kk, 1), (kk // 2, 0)) if conv_1d else (kk, kk // 2)
            return nn.Conv2d(dim, dim, ks, 1, pad, groups=dim, bias=bias)

        self.proj, self.proj1, self.proj2 = conv(k), conv(5), conv(3)


class _EncoderFunction(torch.autograd.Function):
    """Autograd bridge: forward with a tape (``rrt_encoder_forward_train``), backward through
    ``rrt_encoder_backward``.  Parameters travel as explicit inputs so that autograd routes their
    gradients; all arithmetic stays in the library."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        lib, cfg, N = cabi.lib(), enc._cfg, x.shape[0]
        drop_p, seed = enc._train_dropout()
        scales = enc._branch_scales()       # stochastic depth of this step (drop_path), or None
        with torch.cuda.device(x.device):
            n = C.c_size_t()
            cabi.check(lib.rrt_train_tape_bytes(C.byref(cfg), N, C.byref(n)), "rrt_train_tape_bytes")
            tape = torch.empty(n.value, dtype=torch.uint8, device=x.device)
            out = torch.empty_like(x)
            w = enc._weights(x.device)
            rc = lib.rrt_encoder_forward_train(C.byref(cfg), C.byref(w), x.data_ptr(), out.data_ptr(), N,
                                               tape.data_ptr(), n.value, drop_p, seed, scales,
                                               torch.cuda.current_stream(x.device).cuda_stream)
        cabi.check(rc, "rrt_encoder_forward_train")
        ctx.enc = enc
        ctx.drop = (drop_p, seed)
        ctx.scales = scales
        ctx.save_for_backward(x, tape, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        enc = ctx.enc
        x, tape, *params = ctx.saved_tensors
        lib, cfg, N = cabi.lib(), enc._cfg, x.shape[0]
        dout = dout.contiguous().float()
        names = enc._named_param_cache()[0]
        # one zero-initialised flat buffer for every parameter gradient (64-float aligned views)
        offs, total = [], 0
        for p_ in params:
            offs.append(total)
            total += (p_.numel() + 63) // 64 * 64
        with torch.cuda.device(x.device):
            flat = torch.zeros(total, dtype=torch.float32, device=x.device)
            views = [flat[o:o + p_.numel()].view(p_.shape) for o, p_ in zip(offs, params)]
            g = enc._grads(dict(zip(names, views)))
            n = C.c_size_t()
            cabi.check(lib.rrt_backward_workspace_bytes(C.byref(cfg), N, C.byref(n)),
                       "rrt_backward_workspace_bytes")
            ws = torch.empty(n.value, dtype=torch.uint8, device=x.device)
            dx = torch.empty_like(x)
            w = enc._weights(x.device)
            rc = lib.rrt_encoder_backward(C.byref(cfg), C.byref(w), x.data_ptr(), dout.data_ptr(), N,
                                          tape.data_ptr(), tape.numel(), C.byref(g), dx.data_ptr(),
                                          ws.data_ptr(), n.value, ctx.drop[0], ctx.drop[1], ctx.scales,
                                          torch.cuda.current_stream(x.device).cuda_stream)
        cabi.check(rc, "rrt_encoder_backward")
        grads = [v if p_.requires_grad else None for v, p_ in zip(views, params)]
        return (None, dx if ctx.needs_input_grad[1] else None, *grads)


class RRTEncoder(nn.Module):
    def __init__(self, mlp_dim=512, pos_pos=0, pos='none', peg_k=7, attn='rmsa', region_num=8,
                 drop_out=0.1, n_layers=2, n_heads=8, drop_path=0., ffn=False, ffn_act='gelu',
                 mlp_ratio=4., trans_dim=64, epeg=True, epeg_k=15, region_size=0, min_region_num=0,
                 min_region_ratio=0, qkv_bias=True, peg_bias=True, peg_1d=False, cr_msa=True,
                 crmsa_k=3, all_shortcut=False, crmsa_mlp=False, crmsa_heads=8, need_init=False,
                 **kwargs):
        super().__init__()
        # ---- options the kernels do not cover raise instead of silently computing something else
        if attn != 'rmsa':
            raise NotImplementedError(f"attn={attn!r}: only 'rmsa' is built (the reference also "
                                      "raises for unknown values, modules/rrt.py:89-90)")
        if pos not in ('none', None, 'peg', 'ppeg'):
            # 'sincos' cannot be constructed in the reference either with numpy >= 1.24 (np.float,
            # modules/emb_position.py:97); any other string is nn.Identity there
            raise NotImplementedError(f"pos={pos!r}: only 'none', 'peg' and 'ppeg' are built")
        if pos in ('peg', 'ppeg') and (peg_k % 2 == 0 or peg_k > 31 or pos_pos not in (-1, 0)):
            raise ValueError("peg_k must be odd and <= 31, pos_pos -1 or 0")
        if ffn and (int(mlp_dim * mlp_ratio) % 64 or int(mlp_dim * mlp_ratio) < 64):
            raise ValueError("ffn: int(mlp_dim * mlp_ratio) must be a multiple of 64")
        epeg_2d = kwargs.pop('epeg_2d', False)
        epeg_type = kwargs.pop('epeg_type', 'attn')
        epeg_bias = kwargs.pop('epeg_bias', True)
        region_attn = kwargs.pop('region_attn', 'native')
        if epeg_type not in ('attn', 'value_bf', 'value_af'):
            raise NotImplementedError(f"epeg_type={epeg_type!r}: 'attn', 'value_bf' and 'value_af' are built")
        if region_attn != 'native':
            raise NotImplementedError("region_attn != 'native' (ablation) is not built")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        if n_layers < 1 or n_layers - 1 > cabi.RRT_MAX_RMSA_LAYERS:
            raise ValueError("n_layers out of range")
        if epeg and n_layers > 1 and (epeg_k % 2 == 0 or epeg_k > cabi.RRT_MAX_EPEG_K):
            raise ValueError("epeg_k must be odd (the reference's forward fails on even kernels: "
                             "Conv2d padding k//2 yields P+1 rows, modules/rmsa.py:83,108) and <= 63")

        self.final_dim = mlp_dim
        self.all_shortcut = all_shortcut
        self.drop_out = float(drop_out)
        self.drop_path_rate = float(drop_path)
        self.pos_pos = pos_pos
        self.norm = nn.LayerNorm(mlp_dim)
        self.layers = nn.Sequential(*[
            TransLayer(mlp_dim, RegionAttention(mlp_dim, n_heads, qkv_bias, epeg, epeg_k, epeg_bias,
                                                bool(epeg_2d), epeg_type),
                       ffn, ffn_act, mlp_ratio, drop_out)
            for _ in range(n_layers - 1)])
        self.cr_msa = (TransLayer(mlp_dim, CrossRegionAttention(mlp_dim, crmsa_heads, qkv_bias,
                                                                crmsa_k, crmsa_mlp),
                                  ffn, ffn_act, mlp_ratio, drop_out)
                       if cr_msa else nn.Identity())
        if pos == 'ppeg':
            self.pos_embedding = PPEG(dim=mlp_dim, k=peg_k, bias=peg_bias, conv_1d=peg_1d)
        elif pos == 'peg':
            self.pos_embedding = PEG(mlp_dim, k=peg_k, bias=peg_bias, conv_1d=peg_1d)
        else:
            self.pos_embedding = nn.Identity()

        cfg = cabi.RrtConfig()
        cfg.dim, cfg.n_rmsa_layers, cfg.n_heads = mlp_dim, n_layers - 1, n_heads
        cfg.region_num, cfg.region_size = region_num, int(region_size or 0)
        cfg.min_region_num, cfg.min_region_ratio = int(min_region_num), float(min_region_ratio)
        cfg.epeg, cfg.epeg_k, cfg.qkv_bias = int(bool(epeg)), int(epeg_k), int(bool(qkv_bias))
        cfg.cr_msa, cfg.crmsa_k, cfg.crmsa_heads = int(bool(cr_msa)), int(crmsa_k), int(crmsa_heads)
        cfg.crmsa_mlp, cfg.all_shortcut = int(bool(crmsa_mlp)), int(bool(all_shortcut))
        cfg.math_mode = cabi.RRT_MATH_F16
        cfg.pos = {'peg': cabi.RRT_POS_PEG, 'ppeg': cabi.RRT_POS_PPEG}.get(pos, cabi.RRT_POS_NONE)
        cfg.pos_pos, cfg.peg_k, cfg.peg_1d = int(pos_pos), int(peg_k), int(bool(peg_1d))
        cfg.ffn, cfg.ffn_hidden = int(bool(ffn)), int(mlp_dim * mlp_ratio) if ffn else 0
        cfg.ffn_act = cabi.RRT_ACT_GELU if ffn_act == 'gelu' else cabi.RRT_ACT_RELU
        cfg.epeg_2d = int(bool(epeg_2d))
        cfg.epeg_type = {'attn': cabi.RRT_EPEG_ATTN, 'value_bf': cabi.RRT_EPEG_VALUE_BF,
                         'value_af': cabi.RRT_EPEG_VALUE_AF}[epeg_type]
        self._cfg = cfg
        self._crmsa_mlp = bool(crmsa_mlp)
        self._shadow = {}
        self._dropout_seed = None       # tests: pin the dropout seed of the next training forwards
        self.last_dropout_seed = None   # seed the last training forward used

        if need_init:
            self.apply(initialize_weights)

    # ------------------------------------------------------------------------------------------
    def extra_repr(self) -> str:
        c = self._cfg
        return (f"dim={c.dim}, n_layers={c.n_rmsa_layers + 1}, region_num={c.region_num}, "
                f"epeg_k={c.epeg_k if c.epeg else None}, crmsa_k={c.crmsa_k if c.cr_msa else None}, "
                f"all_shortcut={bool(c.all_shortcut)}")

    @staticmethod
    def _ptr(t, device):
        if t is None:
            return None
        if t.device != device or t.dtype != torch.float32:
            raise RuntimeError("all RRTEncoder parameters must be float32 on the input's device")
        if not t.is_contiguous():
            raise RuntimeError("RRTEncoder parameters must be contiguous")
        return t.data_ptr()

    def _f16_shadow(self, param: torch.Tensor):
        """fp16 copy of a GEMM weight (the form the tcgen05 kernels consume), cached in eval mode and
        refreshed when the parameter's storage or version counter changes.  In training mode
        (weights change every step) no shadow is passed and the library converts per call.
        In-place edits through ``.data`` bypass the version counter: call
        ``invalidate_weight_cache()`` after such edits."""
        if self.training:
            return None
        key = id(param)
        ent = self._shadow.get(key)
        if ent is None or ent[0] != (param.data_ptr(), param._version):
            buf = torch.empty(param.shape, dtype=torch.float16, device=param.device)
            rc = cabi.lib().rrt_convert_f16(param.data_ptr(), buf.data_ptr(), param.numel(),
                                            torch.cuda.current_stream(param.device).cuda_stream)
            cabi.check(rc, "rrt_convert_f16")
            ent = ((param.data_ptr(), param._version), buf)
            self._shadow[key] = ent
        return ent[1].data_ptr()

    def invalidate_weight_cache(self) -> None:
        self._shadow.clear()
        self.__dict__.pop("_w_cache", None)

    # The caches hold ctypes structs full of device pointers, fp16 shadow tensors keyed by id() and a closure:
    # none of it may travel with a copy or a pickle (copy.deepcopy for EMA / teacher models, torch.save(model)).
    _TRANSIENT = ("_w_cache", "_np_cache")

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in self._TRANSIENT:
            state.pop(k, None)
        state["_shadow"] = {}
        state["_cfg"] = bytes(self._cfg)          # ctypes.Structure -> plain bytes
        return state

    def __setstate__(self, state):
        cfg = state.pop("_cfg")
        self.__dict__.update(state)
        self._cfg = cabi.RrtConfig.from_buffer_copy(cfg) if isinstance(cfg, (bytes, bytearray)) else cfg

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        state = self.__getstate__()
        cfg = state.pop("_cfg")
        new.__dict__.update(copy.deepcopy(state, memo))
        new._cfg = cabi.RrtConfig.from_buffer_copy(cfg)
        return new

    def _attn_weights(self, inner: InnerAttention, dst: cabi.RrtAttnWeights, device, shadows=False):
        p = self._ptr
        dst.qkv_w, dst.qkv_b = p(inner.qkv.weight, device), p(inner.qkv.bias, device)
        dst.proj_w, dst.proj_b = p(inner.proj.weight, device), p(inner.proj.bias, device)
        dst.pe_w = p(inner.pe.weight, device) if inner.pe is not None else None
        dst.pe_b = p(inner.pe.bias, device) if inner.pe is not None and inner.pe.bias is not None else None
        if shadows:
            dst.qkv_w_f16 = self._f16_shadow(inner.qkv.weight)
            dst.proj_w_f16 = self._f16_shadow(inner.proj.weight)

    def _weights(self, device) -> cabi.RrtWeights:
        """``rrt_weights`` over the parameters (and, in eval mode, their fp16 shadows).  Building it walks
        ~60 pointers with device / dtype / layout checks (~100 us); it is rebuilt only when a parameter's
        storage or version counter, the device or the train/eval mode changed."""
        params = self._named_param_cache()[1]
        key = (device, self.training, tuple((q.data_ptr(), q._version) for q in params))
        c = self.__dict__.get("_w_cache")
        if c is not None and c[0] == key:
            return c[1]
        w = self._build_weights(device)
        self.__dict__["_w_cache"] = (key, w)
        return w

    def _build_weights(self, device) -> cabi.RrtWeights:
        w, p = cabi.RrtWeights(), self._ptr
        w.norm_w, w.norm_b = p(self.norm.weight, device), p(self.norm.bias, device)
        for i, layer in enumerate(self.layers):
            w.layer_norm_w[i], w.layer_norm_b[i] = p(layer.norm.weight, device), p(layer.norm.bias, device)
            self._attn_weights(layer.attn.attn, w.layer_attn[i], device, shadows=True)
        if self._cfg.cr_msa:
            cr = self.cr_msa
            w.cr_norm_w, w.cr_norm_b = p(cr.norm.weight, device), p(cr.norm.bias, device)
            if self._crmsa_mlp:
                w.cr_phi_w1, w.cr_phi_w2 = p(cr.attn.phi[0].weight, device), p(cr.attn.phi[2].weight, device)
                w.cr_phi_w1_f16 = self._f16_shadow(cr.attn.phi[0].weight)
            else:
                w.cr_phi = p(cr.attn.phi, device)
            self._attn_weights(cr.attn.attn, w.cr_attn, device, shadows=True)
        if self._cfg.ffn:
            def ffn_w(layer, dst):
                dst.norm_w, dst.norm_b = p(layer.norm2.weight, device), p(layer.norm2.bias, device)
                dst.fc1_w, dst.fc1_b = p(layer.mlp.fc1.weight, device), p(layer.mlp.fc1.bias, device)
                dst.fc2_w, dst.fc2_b = p(layer.mlp.fc2.weight, device), p(layer.mlp.fc2.bias, device)
                dst.fc1_w_f16 = self._f16_shadow(layer.mlp.fc1.weight)
                dst.fc2_w_f16 = self._f16_shadow(layer.mlp.fc2.weight)
            for i, layer in enumerate(self.layers):
                ffn_w(layer, w.layer_ffn[i])
            if self._cfg.cr_msa:
                ffn_w(self.cr_msa, w.cr_ffn)
        if self._cfg.pos != cabi.RRT_POS_NONE:
            pe = self.pos_embedding
            for j, conv in enumerate([pe.proj] + ([pe.proj1, pe.proj2] if self._cfg.pos == cabi.RRT_POS_PPEG else [])):
                w.pos_w[j], w.pos_b[j] = p(conv.weight, device), p(conv.bias, device)
        return w

    def _grads(self, by_name) -> cabi.RrtGrads:
        """``rrt_grads`` over gradient buffers keyed by parameter name (``named_parameters``)."""
        g = cabi.RrtGrads()

        def ptr(name):
            t = by_name.get(name)
            return t.data_ptr() if t is not None else None

        def attn(prefix, dst):
            dst.qkv_w, dst.qkv_b = ptr(prefix + "qkv.weight"), ptr(prefix + "qkv.bias")
            dst.proj_w, dst.proj_b = ptr(prefix + "proj.weight"), ptr(prefix + "proj.bias")
            dst.pe_w = ptr(prefix + "pe.weight")   # pe.bias: exactly zero gradient, never written

        g.norm_w, g.norm_b = ptr("norm.weight"), ptr("norm.bias")
        for i in range(len(self.layers)):
            g.layer_norm_w[i], g.layer_norm_b[i] = ptr(f"layers.{i}.norm.weight"), ptr(f"layers.{i}.norm.bias")
            attn(f"layers.{i}.attn.attn.", g.layer_attn[i])
        if self._cfg.cr_msa:
            g.cr_norm_w, g.cr_norm_b = ptr("cr_msa.norm.weight"), ptr("cr_msa.norm.bias")
            g.cr_phi = ptr("cr_msa.attn.phi")
            g.cr_phi_w1, g.cr_phi_w2 = ptr("cr_msa.attn.phi.0.weight"), ptr("cr_msa.attn.phi.2.weight")
            attn("cr_msa.attn.attn.", g.cr_attn)
        if self._cfg.pos != cabi.RRT_POS_NONE:
            for j, conv in enumerate(("proj", "proj1", "proj2")):
                g.pos_w[j] = ptr(f"pos_embedding.{conv}.weight")
                g.pos_b[j] = ptr(f"pos_embedding.{conv}.bias")
        return g

    def _named_param_cache(self):
        """(names, parameters) in ``named_parameters()`` order.  Walking the module tree costs ~150 us
        per training step; the tree of an encoder is fixed after construction, so the walk is cached
        and redone only if a parameter object was replaced (``module.weight = nn.Parameter(...)``)."""
        c = self.__dict__.get("_np_cache")
        if c is None or any(self_p is not q for self_p, q in zip(c[1], c[2]())):
            named = list(self.named_parameters())
            holders = [(self.get_submodule(n.rpartition(".")[0]) if "." in n else self, n.rpartition(".")[2])
                       for n, _ in named]
            getter = lambda: [h._parameters[k] for h, k in holders]  # noqa: E731
            c = ([n for n, _ in named], [p for _, p in named], getter)
            self.__dict__["_np_cache"] = c
        return c

    def _train_dropout(self):
        """(p, seed) of this forward: ``drop_out`` is active in training mode only, like ``nn.Dropout``.
        The seed comes from torch's CPU generator, so ``torch.manual_seed`` makes steps reproducible
        (and no device synchronisation is needed to draw it); ``_dropout_seed`` pins it (tests)."""
        if not self.training or self.drop_out <= 0.0:
            return 0.0, 0
        seed = self._dropout_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        self.last_dropout_seed = seed
        return float(self.drop_out), seed

    def _branch_scales(self):
        """Stochastic depth (``drop_path``, modules/rrt.py:102,125) of this training forward: one Bernoulli(keep)
        per block (R-MSA layers, then CR-MSA) from torch's CPU generator -- the bag is the whole batch, so a dropped
        block is skipped outright and a kept one has its branch scaled by 1 / keep (timm's DropPath).  Returns a
        ctypes float array for the C entries, or None.  ``_drop_path_keep`` (list of bools) pins the draw (tests)."""
        p = self.drop_path_rate
        if not self.training or p <= 0.0:
            return None
        nb = self._cfg.n_rmsa_layers + (1 if self._cfg.cr_msa else 0)
        keep = self.__dict__.get("_drop_path_keep")
        if keep is None:
            keep = (torch.rand(nb) >= p).tolist()
        self.last_drop_path_keep = [bool(k) for k in keep]
        vals = [(1.0 / (1.0 - p)) if k else 0.0 for k in keep[:nb]]
        while len(vals) < self._cfg.n_rmsa_layers + 1:      # the C array always has n_rmsa_layers + 1 entries
            vals.append(1.0)
        return (C.c_float * len(vals))(*vals)

    def _needs_grad(self, x) -> bool:
        return torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self._named_param_cache()[1]))

    def _check_mode(self, x, allow_grad=False):
        if not x.is_cuda:
            raise RuntimeError("RRTEncoder (rrt_mil_b200) runs on CUDA only; there is no CPU fallback")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise NotImplementedError(f"input dtype {x.dtype}: float32 bags (or float16 / bfloat16 rows from an "
                                      "autocast host, widened to float32) are supported")
        if x.dtype != torch.float32 and self._needs_grad(x):
            raise NotImplementedError("half-precision inputs are an inference convenience (autocast hosts); "
                                      "train with float32 bags")
        if self._needs_grad(x):
            if not allow_grad:
                raise NotImplementedError("forward_bags is inference-only: call it under "
                                          "torch.no_grad(), or use forward() for autograd")
            if self._cfg.ffn:
                raise NotImplementedError("backward through the FFN ablation is not built")
            # limits that depend on the BAG, not only on the configuration (region size, head_dim): checked here,
            # before the taped forward, so that a training loop fails at the call and not inside loss.backward()
            N = x.shape[-2] if x.dim() >= 2 else 0
            lib = cabi.lib()
            if N >= 1 and lib.rrt_backward_supported(C.byref(self._cfg), N) != cabi.RRT_OK:
                why = lib.rrt_last_error().decode("utf-8", "replace")
                raise NotImplementedError(
                    f"training is not covered for this bag / configuration (N={N}, region_num={self._cfg.region_num}, "
                    f"R-MSA head_dim={self._cfg.dim // max(self._cfg.n_heads, 1)}, CR-MSA head_dim="
                    f"{self._cfg.dim // max(self._cfg.crmsa_heads, 1)}): {why}")
        if self.training and self._cfg.ffn and self.drop_out > 0:
            # the reference's Mlp applies nn.Dropout(drop_out) twice in training mode (modules/rrt.py:25-41);
            # the FFN kernels have no dropout, so computing on would silently differ from the reference
            raise NotImplementedError("training-mode dropout inside the FFN ablation is not built: "
                                      "use .eval() or drop_out=0")
        if self.training and self.drop_path_rate > 0 and self._cfg.ffn:
            raise NotImplementedError("training-mode drop_path with the FFN ablation is not built")
        if self.training and (self.drop_out > 0 or self.drop_path_rate > 0) and not allow_grad:
            raise NotImplementedError("forward_bags is inference-only (no dropout / drop_path): call .eval() first")

    @staticmethod
    def _as_f32(x: torch.Tensor) -> torch.Tensor:
        """float16 / bfloat16 rows (a host under autocast: the reference's ``--amp``, main.py:101-102,439) are
        widened to float32 by the library; like the reference under autocast (its LayerNorms run in float32),
        the result is float32."""
        if x.dtype == torch.float32:
            return x
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = cabi.lib().rrt_widen_f32(x.data_ptr(), int(x.dtype == torch.bfloat16), out.data_ptr(), x.numel(),
                                          torch.cuda.current_stream(x.device).cuda_stream)
        cabi.check(rc, "rrt_widen_f32")
        return out

    def forward_bag(self, x: torch.Tensor) -> torch.Tensor:
        """One bag ``[N, D]`` float32 CUDA -> ``[N, D]``; enqueues on the current stream."""
        self._check_mode(x, allow_grad=True)
        if x.dim() != 2 or x.shape[1] != self.final_dim:
            raise ValueError(f"expected a [N, {self.final_dim}] bag, got {tuple(x.shape)}")
        N = x.shape[0]
        if N < 1:
            raise ValueError("empty bag")
        x = self._as_f32(x.contiguous())
        if self._needs_grad(x) or (self.training and (self.drop_out > 0 or self.drop_path_rate > 0)):
            # autograd / training path: forward with a tape (+ proj dropout), backward kernels
            return _EncoderFunction.apply(self, x, *self._named_param_cache()[1])
        lib, cfg = cabi.lib(), self._cfg
        with torch.cuda.device(x.device):
            nbytes = cabi.workspace_bytes(cfg, N)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            out = torch.empty_like(x)
            w = self._weights(x.device)
            rc = lib.rrt_encoder_forward(C.byref(cfg), C.byref(w), x.data_ptr(), out.data_ptr(), N,
                                         ws.data_ptr(), nbytes,
                                         torch.cuda.current_stream(x.device).cuda_stream)
        cabi.check(rc, "rrt_encoder_forward")
        return out

    def forward_bags(self, bags, outs=None, lanes: int = cabi.RRT_MAX_LANES):
        """Independent bags ``[N_i, D]`` with ONE C call; results equal per-bag ``forward``.
        Up to ``lanes`` bags run concurrently on the library's internal streams, forked from and
        joined back into the current stream (bags are independent; the small latency-bound kernels
        of one bag fill the SMs another bag leaves idle).  ``lanes=1`` runs them back to back."""
        if not bags:
            return []
        for x in bags:
            self._check_mode(x)
            if x.dim() != 2 or x.shape[1] != self.final_dim or x.shape[0] < 1 or not x.is_contiguous():
                raise ValueError(f"expected contiguous [N, {self.final_dim}] bags")
            if x.device != bags[0].device:
                raise ValueError("all bags of one call must live on the same device")
        device = bags[0].device
        bags = [self._as_f32(x) for x in bags]
        if outs is None:
            outs = [torch.empty_like(x) for x in bags]
        n = len(bags)
        lib, cfg = cabi.lib(), self._cfg
        with torch.cuda.device(device):
            per_bag = (cabi.workspace_bytes(cfg, max(x.shape[0] for x in bags)) + 255) // 256 * 256
            nbytes = per_bag * max(1, min(int(lanes), n, cabi.RRT_MAX_LANES))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            w = self._weights(device)
            xs = (C.c_void_p * n)(*[x.data_ptr() for x in bags])
            os_ = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
            ls = (C.c_int64 * n)(*[x.shape[0] for x in bags])
            rc = lib.rrt_encoder_forward_batch(C.byref(cfg), C.byref(w), xs, os_, ls, n, ws.data_ptr(),
                                               nbytes, torch.cuda.current_stream(device).cuda_stream)
        cabi.check(rc, "rrt_encoder_forward_batch")
        return outs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape_len = x.dim()
        if shape_len == 2:          # [N, C]  (clam / dsmil hosts)
            x3 = x.unsqueeze(0)
        elif shape_len == 4:        # [B, C, H, W]
            x3 = x.reshape(x.size(0), x.size(1), -1).transpose(1, 2)
        elif shape_len == 3:
            x3 = x
        else:
            raise ValueError(f"expected a 2-D, 3-D or 4-D input, got {tuple(x.shape)}")
        batch, num_patches, channels = x3.shape
        if batch != 1:
            # the reference silently mixes the bags of a batch inside CR-MSA;
            # every host calls with one bag, which is the contract kept here
            raise ValueError("RRTEncoder processes one bag per call (batch dimension must be 1)")
        y = self.forward_bag(x3[0]).unsqueeze(0)
        if shape_len == 2:
            y = y.squeeze(0)
        elif shape_len == 4:
            side = int(num_patches ** 0.5)
            y = y.transpose(1, 2).reshape(batch, channels, side, side)
        return y
