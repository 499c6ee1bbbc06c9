"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden
vectors.  Tolerance: the north-star's fp32 bar, rel <= 1e-3 (tf32 tensor-core operands, fp32
accumulation / softmax / LayerNorm); the HBM-bound fp32 kernels are held to 1e-5."""
import os

import pytest
import torch

from oracle import rrt_oracle as O
from golden_util import CASES, load_case, assert_matches_golden
import gpu_util as G

pytestmark = pytest.mark.gpu

TOL_TF32 = 1e-3   # BASELINE.json north_star: "within 1e-3 rel fp32"
TOL_FP32 = 1e-5   # kernels with no tensor-core contraction


def dev(w):
    return {k: v.float().cuda() for k, v in w.items()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_encoder_matches_golden(name):
    cfg, w, x, gold = load_case(name)
    m = G.make_encoder(cfg, w)
    with torch.no_grad():
        y = m(x.float().cuda().unsqueeze(0))[0]
    torch.cuda.synchronize()
    assert y.shape == x.shape and torch.isfinite(y).all()
    e = assert_matches_golden(y, gold, TOL_TF32, f"cuda {name}")
    print(name, e)


@pytest.mark.parametrize("L,over", [
    (512, dict()),
    (300, dict(mlp_dim=256, region_num=4, epeg_k=5, crmsa_k=4, crmsa_heads=2, all_shortcut=True)),
    (1000, dict(crmsa_mlp=True, crmsa_heads=1, crmsa_k=5)),
    (2500, dict(region_num=16, n_layers=3, epeg_k=21)),
    (97, dict(mlp_dim=128, n_heads=4, crmsa_heads=4, epeg_k=3)),
    (700, dict(mlp_dim=1024, n_heads=8, crmsa_heads=8)),
])
def test_encoder_matches_oracle_fp64(L, over):
    cfg = O.EncoderConfig(**over)
    w = O.make_weights(cfg, 31)
    x = O.make_bag(L, cfg.mlp_dim, 32, kind="relu")
    ref = O.encoder_forward(x, w, cfg, "spec")
    m = G.make_encoder(cfg, w)
    with torch.no_grad():
        y = m(x.float().cuda())          # 2-D input path (clam / dsmil hosts)
    assert y.dim() == 2
    assert O.rel_err(y.cpu(), ref) < TOL_TF32


@pytest.mark.parametrize("L,over", [(512, dict()), (1300, dict(region_num=4, epeg_k=9)),
                                    (200, dict(mlp_dim=256, epeg=False, qkv_bias=False))])
def test_rmsa_block_matches_oracle(L, over):
    cfg = O.EncoderConfig(**over)
    w = O.make_weights(cfg, 5)
    x = O.make_bag(L, cfg.mlp_dim, 6)
    p = "layers.0."
    ref = x + O.rmsa_block(O.layer_norm(x, w[p + "norm.weight"], w[p + "norm.bias"]), w,
                           p + "attn.", cfg, "spec")
    m = G.make_encoder(cfg, w)
    y = G.rmsa_block(m, 0, x.float().cuda())
    assert O.rel_err(y.cpu(), ref) < TOL_TF32
    # the residual branch alone (what the kernels compute) must also be within tolerance
    assert O.rel_err(y.cpu().double() - x, ref - x) < 3 * TOL_TF32


@pytest.mark.parametrize("L,over,final", [
    (512, dict(), True), (512, dict(all_shortcut=True, crmsa_k=1), True),
    (3000, dict(crmsa_k=5, crmsa_heads=1), False), (900, dict(crmsa_mlp=True, mlp_dim=256), True)])
def test_crmsa_block_matches_oracle(L, over, final):
    cfg = O.EncoderConfig(**over)
    w = O.make_weights(cfg, 8)
    x1 = O.make_bag(L, cfg.mlp_dim, 9)
    x0 = O.make_bag(L, cfg.mlp_dim, 10)
    p = "cr_msa."
    ref = x1 + O.crmsa_block(O.layer_norm(x1, w[p + "norm.weight"], w[p + "norm.bias"]), w,
                             p + "attn.", cfg, "spec")
    if cfg.all_shortcut:
        ref = ref + x0
    if final:
        ref = O.layer_norm(ref, w["norm.weight"], w["norm.bias"])
    m = G.make_encoder(cfg, w)
    y = G.crmsa_block(m, x1.float().cuda(), x0.float().cuda(), final)
    assert O.rel_err(y.cpu(), ref) < TOL_TF32


@pytest.mark.parametrize("M,N,K", [(1, 2, 32), (192, 1536, 512), (333, 130, 96), (9216, 512, 512),
                                   (2000, 1536, 256)])
def test_linear_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g, dtype=torch.float64)
    w = torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5
    b = torch.randn(N, generator=g, dtype=torch.float64)
    ref = a @ w.T + b
    y = G.linear(a.float().cuda(), w.float().cuda(), b.float().cuda())
    assert O.rel_err(y.cpu(), ref) < TOL_TF32
    y0 = G.linear(a.float().cuda(), w.float().cuda(), None)
    assert O.rel_err(y0.cpu(), ref - b) < TOL_TF32


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (9216, 1536, 512), (9216, 512, 512), (576, 1536, 512),
                                   (100, 384, 128), (64, 256, 256), (192, 1536, 512), (1, 4, 64),
                                   (50176, 1536, 512), (3000, 3072, 1024)])
def test_tcgen05_linear_matches_fp64(M, N, K):
    """The TMA + tcgen05 + TMEM GEMM (every linear layer of the path): M / N tails, single tile,
    narrow-tile variant, many waves."""
    g = torch.Generator().manual_seed(M * 13 + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    y = G.linear_f16(a.cuda(), w.cuda(), b.cuda())
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T + b.double()
    assert O.rel_err(y.cpu(), ref) < TOL_TF32
    # against the SAME fp16 operands the error is fp32-accumulation only
    ah, wh = G.convert_f16(a.cuda()).cpu().double(), G.convert_f16(w.cuda()).cpu().double()
    assert O.rel_err(y.cpu(), ah @ wh.T + b.double()) < 2e-6
    y2 = G.linear_f16(a.cuda(), w.cuda(), b.cuda())
    assert torch.equal(y, y2)


def test_convert_f16_rounds_to_nearest_and_saturates():
    x = torch.randn(4096, generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(G.convert_f16(x), x.half())
    big = torch.tensor([1e6, -1e6, 65504.0, 3.0], device="cuda")
    assert G.convert_f16(big).float().tolist() == [65504.0, -65504.0, 65504.0, 3.0]


def test_linear_is_linear():
    g = torch.Generator().manual_seed(3)
    a1, a2 = torch.randn(500, 512, generator=g).cuda(), torch.randn(500, 512, generator=g).cuda()
    w = (torch.randn(384, 512, generator=g) / 22).cuda()
    lhs = G.linear(a1 + a2, w, None)
    rhs = G.linear(a1, w, None) + G.linear(a2, w, None)
    assert O.rel_err(lhs.cpu(), rhs.cpu()) < TOL_TF32


@pytest.mark.parametrize("L,D", [(1, 128), (1000, 512), (77, 1024), (4096, 256)])
def test_layernorm_matches_fp64(L, D):
    g = torch.Generator().manual_seed(L + D)
    x = torch.randn(L, D, generator=g, dtype=torch.float64) * 3 + 1
    gm = torch.randn(D, generator=g, dtype=torch.float64)
    bt = torch.randn(D, generator=g, dtype=torch.float64)
    y = G.layernorm(x.float().cuda(), gm.float().cuda(), bt.float().cuda())
    assert O.rel_err(y.cpu(), O.layer_norm(x, gm, bt)) < TOL_FP32


# ---- size-independent properties at BASELINE.json's full sizes -------------------------------
def _default_encoder(**over):
    cfg = O.EncoderConfig(**over)
    return cfg, G.make_encoder(cfg, O.make_weights(cfg, 2021, randomize_bias=False))


@pytest.mark.parametrize("L,over", [(9000, dict()), (50000, dict(region_num=16))])
def test_full_size_output_rows_are_normalised_and_deterministic(L, over):
    cfg, m = _default_encoder(**over)
    x = O.make_bag(L, 512, 1, dtype=torch.float32).cuda()
    with torch.no_grad():
        y1 = m(x.unsqueeze(0))[0]
        y2 = m(x.unsqueeze(0))[0]
    assert torch.isfinite(y1).all()
    assert torch.equal(y1, y2)                      # idempotent / deterministic
    # final LayerNorm with (1,0) affine: every row has mean 0, variance 1
    assert y1.mean(1).abs().max() < 1e-4
    assert (y1.var(1, unbiased=False) - 1).abs().max() < 1e-3


def test_full_size_rmsa_regions_are_independent():
    """Changing the tokens of one region must leave every other region's rows bit-identical
    (R-MSA attends within a region only)."""
    cfg, m = _default_encoder()
    L = 9000
    x = O.make_bag(L, 512, 4, dtype=torch.float32).cuda()
    H, rs, _ = O.grid_geometry(L, 8)
    slot_to_tok = O.region_slot_map(H, rs).cuda()
    P = rs * rs
    toks = slot_to_tok[5 * P:6 * P]
    toks = toks[toks < L]
    x2 = x.clone()
    x2[toks] += 1.0
    y1, y2 = G.rmsa_block(m, 0, x), G.rmsa_block(m, 0, x2)
    mask = torch.ones(L, dtype=torch.bool, device="cuda")
    mask[toks] = False
    assert torch.equal(y1[mask], y2[mask])
    assert not torch.equal(y1[toks], y2[toks])


def test_concurrent_lanes_equal_serial_per_bag_forward():
    """forward_bags (several bags in flight on internal streams) == one forward per bag, bit for bit,
    for ragged bag lengths."""
    cfg, m = _default_encoder()
    g = torch.Generator().manual_seed(5)
    lens = [9000, 777, 8123, 1, 9216, 5000, 4097]
    bags = [torch.randn(n, 512, generator=g).cuda() for n in lens]
    with torch.no_grad():
        serial = [m(b) for b in bags]
        for lanes in (1, 2, 4):
            outs = m.forward_bags(bags, lanes=lanes)
            torch.cuda.synchronize()
            for a, b in zip(serial, outs):
                assert torch.equal(a, b)


@pytest.mark.parametrize("knob,value,restore", [
    ("rrt_debug_set_attention_kernel", 2, 1),   # tcgen05 attention core wherever supported (auto: regions > 128)
    ("rrt_debug_set_attention_kernel", 0, 1),   # mma.sync attention core everywhere
    ("rrt_debug_set_gemm_cluster", 2, 11),      # cta_group::2 GEMM (CTA pairs, M=256 tiles)
    ("rrt_debug_set_gemm_cluster", 21, 11),     # 2x1 cluster, TMA multicast of the W tile
    ("rrt_debug_set_gemm_cluster", 22, 11),     # 2x2 cluster, multicast of both operand tiles
])
def test_alternative_kernel_variants_keep_parity(knob, value, restore):
    """The selectable kernel variants (kept for tuning) must meet the same parity bar
    as the default kernels on golden cases with one and with two 128-row attention blocks."""
    from rrt_mil_b200 import cabi
    lib = cabi.lib()
    getattr(lib, knob)(value)
    try:
        for name in ("c2_n9000_d512", "plip_k9_shortcut", "c4_n50000_g16", "c1_n512_d512", "n65", "d256_g4"):
            cfg, w, x, gold = load_case(name)
            m = G.make_encoder(cfg, w)
            with torch.no_grad():
                y = m(x.float().cuda())
            torch.cuda.synchronize()
            assert_matches_golden(y, gold, TOL_TF32, f"{knob}={value} {name}")
    finally:
        getattr(lib, knob)(restore)


def test_tiny_bag_crmsa_contributes_nothing():
    """N < 64: every token is its own region, min-max normalised dispatch weight is 0/(0+1e-8)=0
     -> the CR-MSA block is the identity on x1."""
    cfg, m = _default_encoder()
    x1 = O.make_bag(50, 512, 4, dtype=torch.float32).cuda()
    y = G.crmsa_block(m, x1, None, False)
    assert torch.equal(y, x1)


def test_input_rank_handling_and_errors():
    cfg, m = _default_encoder()
    x = O.make_bag(400, 512, 2, dtype=torch.float32).cuda()
    with torch.no_grad():
        y3 = m(x.unsqueeze(0))
        y2 = m(x)
        x4 = x.t().reshape(1, 512, 20, 20).contiguous()
        y4 = m(x4)
    assert y3.shape == (1, 400, 512) and y2.shape == (400, 512) and y4.shape == (1, 512, 20, 20)
    assert torch.equal(y3[0], y2)
    assert torch.equal(y4.reshape(1, 512, 400).transpose(1, 2)[0], y2)
    xg = x.unsqueeze(0).clone().requires_grad_()    # grad mode runs the taped forward + CUDA backward
    yg = m(xg)
    assert torch.allclose(yg.detach(), y3, rtol=0, atol=1e-5)
    yg.square().sum().backward()
    assert xg.grad is not None and xg.grad.shape == xg.shape and torch.isfinite(xg.grad).all()
    with torch.no_grad():   # half rows (a host under autocast) are widened by the library; the result is fp32
        yh = m(x.half())
        yf = m(x.half().float())
    assert yh.dtype == torch.float32 and torch.equal(yh, yf)


def test_host_pipeline_equals_per_bag_forward():
    """Host buffers in / out through the three-stream pipeline (ragged bags, more bags than ring slots,
    a slot that has to grow) == one forward per bag, bit for bit."""
    from rrt_mil_b200.pipeline import HostPipeline
    cfg, m = _default_encoder()
    lens = [700, 300, 1500, 64, 2000, 999, 1500, 31]
    hx = [O.make_bag(n, 512, 100 + i, dtype=torch.float32).pin_memory() for i, n in enumerate(lens)]
    hy = [torch.empty(n, 512).pin_memory() for n in lens]
    pipe = HostPipeline(m, n_streams=3)
    hy2 = [torch.empty(n, 512).pin_memory() for n in lens]
    pipe.run(hx, hy)
    pipe.run(hx, hy2, sync=False)      # streamed calls: no wait in between
    pipe.run(hx[::-1], hy[::-1], sync=False)
    pipe.wait()
    with torch.no_grad():
        for x, y, y2 in zip(hx, hy, hy2):
            ref = m(x.cuda()).cpu()
            assert torch.equal(ref, y) and torch.equal(ref, y2)


@pytest.mark.parametrize("L,D,kind,k,one_d,bias", [
    (1000, 512, "ppeg", 7, False, True), (900, 256, "peg", 5, False, False), (30, 128, "ppeg", 3, True, True),
    (1, 128, "ppeg", 7, False, True), (49, 128, "peg", 7, False, True), (2, 128, "peg", 3, False, True),
    (5000, 384, "ppeg", 9, True, False)])
def test_peg_ppeg_match_oracle(L, D, kind, k, one_d, bias):
    """PEG / PPEG alone through rrt_peg_forward: wrap-around fill with the first tokens, PPEG's zero
    extension below 7x7, 2-D and (k,1) kernels, with and without bias; fp32 streaming kernel -> 1e-5."""
    import ctypes as C
    from rrt_mil_b200 import cabi
    cfg = O.EncoderConfig(mlp_dim=D, pos=kind, pos_pos=-1, peg_k=k, peg_1d=one_d, peg_bias=bias,
                          n_heads=4, crmsa_heads=4)
    w = O.make_weights(cfg, 17)
    x = O.make_bag(L, D, 18)
    ref = O.pos_embedding(x, w, cfg)
    names = ["proj"] + (["proj1", "proj2"] if kind == "ppeg" else [])
    wd = [w[f"pos_embedding.{n}.weight"].float().cuda().contiguous() for n in names]
    bd = [w[f"pos_embedding.{n}.bias"].float().cuda() for n in names] if bias else []
    wp = (C.c_void_p * 3)(*([t.data_ptr() for t in wd] + [None] * (3 - len(wd))))
    bp = (C.c_void_p * 3)(*([t.data_ptr() for t in bd] + [None] * (3 - len(bd))))
    xd = x.float().cuda()
    out = torch.empty_like(xd)
    rc = cabi.lib().rrt_peg_forward(xd.data_ptr(), out.data_ptr(), L, D, k, int(kind == "ppeg"), int(one_d),
                                    wp, bp, G.stream_ptr())
    cabi.check(rc, "rrt_peg_forward")
    torch.cuda.synchronize()
    assert O.rel_err(out.cpu(), ref) < TOL_FP32


def test_peg_encoder_autograd_limits():
    """PEG / PPEG train (tests/test_gpu_backward.py); what does not is PPEG's zero-extended 7x7 grid of bags below
    37 tokens, which raises at the call."""
    cfg = O.EncoderConfig(pos="ppeg", pos_pos=-1)
    m = G.make_encoder(cfg, O.make_weights(cfg, 3))
    x = O.make_bag(100, 512, 1).float().cuda().requires_grad_()
    m(x).square().mean().backward()
    assert torch.isfinite(x.grad).all() and float(m.pos_embedding.proj1.weight.grad.abs().sum()) > 0
    with pytest.raises(NotImplementedError):
        m(O.make_bag(30, 512, 1).float().cuda().requires_grad_())
