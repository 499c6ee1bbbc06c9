"""Helpers for the GPU parity tests: build the product module from an oracle config + weight dict,
and call the block-level C-ABI entry points on torch device buffers."""
import ctypes as C

import torch

from rrt_mil_b200 import RRTEncoder, cabi


def make_encoder(cfg, weights, device="cuda"):
    m = RRTEncoder(**cfg.to_dict()).to(device).eval()
    m.load_state_dict({k: v.float() for k, v in weights.items()}, strict=True)
    return m


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def workspace(cfg_struct, L, device="cuda"):
    n = cabi.workspace_bytes(cfg_struct, L)
    return torch.empty(n, dtype=torch.uint8, device=device), n


def rmsa_block(m: RRTEncoder, layer_idx: int, x: torch.Tensor) -> torch.Tensor:
    layer = m.layers[layer_idx]
    aw = cabi.RrtAttnWeights()
    m._attn_weights(layer.attn.attn, aw, x.device)
    ws, n = workspace(m._cfg, x.shape[0])
    out = torch.empty_like(x)
    rc = cabi.lib().rrt_rmsa_block_forward(C.byref(m._cfg), layer.norm.weight.data_ptr(),
                                           layer.norm.bias.data_ptr(), C.byref(aw), x.data_ptr(),
                                           out.data_ptr(), x.shape[0], ws.data_ptr(), n, stream_ptr())
    cabi.check(rc, "rrt_rmsa_block_forward")
    return out


def crmsa_block(m: RRTEncoder, x1: torch.Tensor, x0, final_norm: bool) -> torch.Tensor:
    w = m._weights(x1.device)
    ws, n = workspace(m._cfg, x1.shape[0])
    out = torch.empty_like(x1)
    rc = cabi.lib().rrt_crmsa_block_forward(C.byref(m._cfg), C.byref(w), x1.data_ptr(),
                                            x0.data_ptr() if x0 is not None else None,
                                            out.data_ptr(), x1.shape[0], int(final_norm),
                                            ws.data_ptr(), n, stream_ptr())
    cabi.check(rc, "rrt_crmsa_block_forward")
    return out


def linear(a, w, b):
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty(M, N, device=a.device, dtype=torch.float32)
    rc = cabi.lib().rrt_linear_forward(a.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                       c.data_ptr(), M, N, K, stream_ptr())
    cabi.check(rc, "rrt_linear_forward")
    return c


def convert_f16(x):
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    cabi.check(cabi.lib().rrt_convert_f16(x.data_ptr(), out.data_ptr(), x.numel(), stream_ptr()), "rrt_convert_f16")
    return out


def linear_f16(a, w, b):
    """The tcgen05 GEMM: fp16 operands (converted here with the library's own kernel), fp32 out."""
    M, K = a.shape
    N = w.shape[0]
    a, w = convert_f16(a.contiguous()), convert_f16(w.contiguous())
    c = torch.empty(M, N, device=a.device, dtype=torch.float32)
    rc = cabi.lib().rrt_linear_f16_forward(a.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                           c.data_ptr(), M, N, K, stream_ptr())
    cabi.check(rc, "rrt_linear_f16_forward")
    return c


def layernorm(x, g, b):
    out = torch.empty_like(x)
    rc = cabi.lib().rrt_layernorm_forward(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(),
                                          x.shape[0], x.shape[1], stream_ptr())
    cabi.check(rc, "rrt_layernorm_forward")
    return out
