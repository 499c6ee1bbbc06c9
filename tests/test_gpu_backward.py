"""Backward pass (through the C ABI / autograd bridge) against torch autograd over the fp64 oracle.

Tolerances: the backward GEMMs and the attention backward use fp16 tensor-core operands (10 mantissa
bits, fp32 accumulation) with automatic power-of-two scaling, like the forward; gradients go through
about twice as many rounding stages as activations, so the bar is rel <= 3e-3 per gradient tensor
(Frobenius), 1e-5 for the fp32 streaming kernels."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import rrt_oracle as O
import gpu_util as G
from golden_util import TRAIN_CASES, load_train_case, train_errors
from rrt_mil_b200 import RRTEncoder, cabi

pytestmark = pytest.mark.gpu

TOL_GRAD = 3e-3
TOL_FP32 = 1e-5


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize("L,D", [(300, 512), (1, 128), (1000, 1024), (77, 256)])
def test_layernorm_backward(L, D):
    g = torch.Generator().manual_seed(L + D)
    x = torch.randn(L, D, generator=g, dtype=torch.float64) * 2 + 0.3
    gam = (1 + 0.1 * torch.randn(D, generator=g, dtype=torch.float64))
    bet = 0.1 * torch.randn(D, generator=g, dtype=torch.float64)
    dy = torch.randn(L, D, generator=g, dtype=torch.float64)
    xr, gr, br = x.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
    (F.layer_norm(xr, (D,), gr, br) * dy).sum().backward()
    xd, gd, dyd = x.float().cuda(), gam.float().cuda(), dy.float().cuda()
    dx = torch.empty_like(xd)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    rc = cabi.lib().rrt_layernorm_backward(xd.data_ptr(), gd.data_ptr(), dyd.data_ptr(), dx.data_ptr(),
                                           dg.data_ptr(), db.data_ptr(), L, D, G.stream_ptr())
    cabi.check(rc, "rrt_layernorm_backward")
    torch.cuda.synchronize()
    assert rel(dx, xr.grad) < TOL_FP32
    assert rel(dg, gr.grad) < TOL_FP32
    assert rel(db, br.grad) < TOL_FP32


@pytest.mark.parametrize("rows,c_out,c_in", [(9216, 1536, 512), (9216, 512, 512), (192, 1536, 512), (700, 256, 256),
                                             (1000, 384, 128), (64, 128, 128), (1, 512, 512), (50000, 512, 512)])
def test_wgrad_gemm_mn_major_matches_fp64(rows, c_out, c_in):
    """dW = dY^T act on tcgen05 with both operands MN-major (read straight from the row-major
    activations): row tails (TMA zero fill), narrow outputs, one row, many split-K slices."""
    g = torch.Generator().manual_seed(rows + c_out)
    dy = torch.randn(rows, c_out, generator=g).half()
    act = torch.randn(rows, c_in, generator=g).half()
    dw = torch.empty(c_out, c_in, device="cuda")
    dyd, actd = dy.cuda(), act.cuda()
    rc = cabi.lib().rrt_linear_wgrad_f16(dyd.data_ptr(), actd.data_ptr(), dw.data_ptr(), rows,
                                         c_out, c_in, G.stream_ptr())
    cabi.check(rc, "rrt_linear_wgrad_f16")
    torch.cuda.synchronize()
    ref = dy.double().t() @ act.double()
    assert rel(dw, ref) < 5e-6      # same fp16 operands: fp32 accumulation error only (3.3e-6 over 50000 rows)


def attention_ref(qkv, taps, heads, P):
    """fp64 attention core in the EPEG-on-Q form (oracle order 'spec'); qkv [R*P, 3D] rows (3, heads, d)."""
    M, D3 = qkv.shape
    D, R = D3 // 3, M // P
    d = D // heads
    t = qkv.view(R, P, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    if taps is not None:
        kk = taps.shape[1]
        w = taps.repeat_interleave(d, 0).unsqueeze(1)
        qc = F.conv1d(q.permute(0, 1, 3, 2).reshape(R, heads * d, P), w, padding=kk // 2, groups=heads * d)
        q = q + qc.view(R, heads, d, P).permute(0, 1, 3, 2)
    a = torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2), -1)
    return (a @ v).permute(0, 2, 1, 3).reshape(M, D)


@pytest.mark.parametrize("R,P,D,heads,kk", [
    (3, 144, 512, 8, 15), (2, 9, 512, 8, 15), (3, 64, 512, 8, None), (2, 196, 512, 8, 21),
    (2, 256, 256, 8, 7), (1, 1, 128, 4, 3), (2, 100, 512, 8, None), (5, 144, 512, 8, 1)])
def test_attention_backward(R, P, D, heads, kk):
    g = torch.Generator().manual_seed(R * 1000 + P)
    M = R * P
    qkv16 = (torch.randn(M, 3 * D, generator=g) * 1.2).half()
    do16 = torch.randn(M, D, generator=g).half()
    taps = (0.2 * torch.randn(heads, kk, generator=g)).float() if kk else None
    qkv = qkv16.double().requires_grad_()
    tp = taps.double().requires_grad_() if kk else None
    o = attention_ref(qkv, tp, heads, P)
    (o * do16.double()).sum().backward()
    o16 = o.detach().half().cuda()
    qd, dod = qkv16.cuda(), do16.cuda()
    dqkv = torch.zeros(M, 3 * D, dtype=torch.float16, device="cuda")
    dt = torch.zeros(heads, kk, device="cuda") if kk else None
    td = taps.cuda() if kk else None
    rc = cabi.lib().rrt_attention_backward(qd.data_ptr(), o16.data_ptr(), dod.data_ptr(),
                                           td.data_ptr() if kk else None, dqkv.data_ptr(),
                                           dt.data_ptr() if kk else None, R, P, D, heads, kk or 1,
                                           G.stream_ptr())
    cabi.check(rc, "rrt_attention_backward")
    torch.cuda.synchronize()
    gq = qkv.grad
    if P == 1:   # softmax over one key: dq = dk = dtaps = 0 exactly, dv = dO
        assert float(dqkv[:, :2 * D].float().abs().max()) < 1e-2 and float(dt.abs().max()) < 1e-2
        assert rel(dqkv[:, 2 * D:], gq[:, 2 * D:]) < TOL_GRAD
        return
    errs = {n: rel(dqkv[:, i * D:(i + 1) * D], gq[:, i * D:(i + 1) * D]) for i, n in enumerate("qkv")}
    if kk:
        errs["taps"] = rel(dt, tp.grad)
    print(R, P, D, heads, kk, errs)
    assert torch.isfinite(dqkv.float()).all()
    for n, e in errs.items():
        assert e < TOL_GRAD, (n, e, errs)


def oracle_grads(cfg, w, x, gout):
    w64 = {k: v.double().clone().requires_grad_() for k, v in w.items()}
    x64 = x.double().clone().requires_grad_()
    y = O.encoder_forward(x64, w64, cfg, "spec")
    (y * gout.double()).sum().backward()
    return y.detach(), x64.grad, {k: v.grad for k, v in w64.items()}


ENC_CASES = [
    ("rmsa_only", 300, dict(cr_msa=False)),
    ("rmsa_only_noepeg", 200, dict(cr_msa=False, epeg=False, qkv_bias=False, mlp_dim=256)),
    ("crmsa_only", 300, dict(n_layers=1)),
    ("default_512", 512, dict()),
    ("default_1000", 1000, dict()),
    ("shortcut_d256", 300, dict(mlp_dim=256, region_num=4, epeg_k=5, crmsa_k=4, crmsa_heads=4, all_shortcut=True)),
    ("three_layers", 2500, dict(region_num=16, n_layers=3, epeg_k=21, crmsa_k=5)),
    ("tiny", 50, dict()),
    ("crmsa_mlp_k5", 600, dict(crmsa_mlp=True, crmsa_k=5)),           # README NSCLC-PLIP recipe's phi (README.md:119)
    ("crmsa_mlp_shortcut", 400, dict(crmsa_mlp=True, all_shortcut=True, region_num=4, epeg_k=7)),
    ("crmsa_heads1", 700, dict(crmsa_heads=1)),                        # BRCA-R50 / LUAD-PLIP recipes: head_dim 512
    ("nsclc_plip_mlp_h1", 500, dict(crmsa_mlp=True, crmsa_heads=1, crmsa_k=5)),   # README.md:119
    ("crmsa_heads1_d256", 300, dict(mlp_dim=256, n_heads=4, crmsa_heads=1, crmsa_k=4)),
    ("ppeg_front", 700, dict(pos="ppeg", pos_pos=-1)),                 # PEG / PPEG backward (f3)
    ("peg_between_3layers_d256", 900, dict(pos="peg", pos_pos=0, n_layers=3, peg_k=5, mlp_dim=256, n_heads=4,
                                           crmsa_heads=4)),
    ("ppeg_1d_shortcut", 400, dict(pos="ppeg", pos_pos=-1, peg_1d=True, all_shortcut=True)),
    ("n9000", 9000, dict()),
    ("n50000_g16", 50000, dict(region_num=16)),     # BASELINE configs[3] shape: P = 196 (R-MSA), 784 (CR-MSA)
]


@pytest.mark.parametrize("name,L,over", ENC_CASES, ids=[c[0] for c in ENC_CASES])
def test_encoder_backward_matches_oracle_autograd(name, L, over):
    cfg = O.EncoderConfig(**over)
    w = O.make_weights(cfg, 41)
    x = O.make_bag(L, cfg.mlp_dim, 42, kind="relu")
    gout = torch.randn(L, cfg.mlp_dim, generator=torch.Generator().manual_seed(43), dtype=torch.float64)
    if name == "n9000":
        gout = gout * 1e-7   # the automatic fp16 scaling must cope with tiny upstream gradients
    y_ref, dx_ref, dw_ref = oracle_grads(cfg, w, x, gout)

    m = G.make_encoder(cfg, w)
    xd = x.float().cuda().requires_grad_()
    y = m(xd)
    with torch.no_grad():
        y_inf = m(xd.detach())
    assert torch.equal(y.detach(), y_inf), "training forward must equal the inference forward bit for bit"
    assert O.rel_err(y.detach().cpu(), y_ref) < 1e-3
    (y * gout.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    errs = {"x": rel(xd.grad, dx_ref)}
    for n, p in m.named_parameters():
        ref = dw_ref[n]
        if ref is None or float(ref.norm()) < 1e-12 * max(1.0, float(gout.norm())):
            # exactly-zero reference gradient (pe.bias; taps when regions hold one token): rounding noise only
            assert p.grad is None or float(p.grad.abs().max()) <= 1e-4 * float(gout.abs().max()), n
            continue
        assert p.grad is not None, n
        errs[n] = rel(p.grad, ref)
    worst = max(errs, key=errs.get)
    print(name, "worst", worst, errs[worst], {k: f"{v:.1e}" for k, v in errs.items()})
    # The reference's min-max dispatch normaliser sends d/dmin, d/dmax to the argmin / argmax token of each
    # region, so its gradient wrt the CR-MSA logits JUMPS when two logits tie.  With 784 tokens per region
    # (N = 50000) the two smallest / largest logits of some region always sit within 1e-4 of the range
    # (crmsa_tie_gap; 2.3e-5 here), closer than the forward's fp16-operand rounding, so the parameters that
    # feed the logits (phi, cr_msa.norm) and everything upstream of them cannot be held to the smooth-function
    # bar; measured 1.1e-2 on phi, <= 2.7e-3 elsewhere.
    from oracle.make_golden import crmsa_tie_gap
    on_a_tie = name == "n50000_g16" and crmsa_tie_gap(x, w, cfg, (0.0, 0)) < 5e-5   # smaller cases: gap >= 5e-5, tight bar
    for n, e in errs.items():
        tol = 3e-2 if on_a_tie and (n.startswith("cr_msa.norm") or n == "cr_msa.attn.phi") else \
            (2 * TOL_GRAD if on_a_tie else TOL_GRAD)
        assert e < tol, (n, e)


@pytest.mark.parametrize("rows,D,p,seed,stream", [(700, 512, 0.1, 20240229, 0), (192, 512, 0.1, 5, 64),
                                                  (333, 256, 0.25, 2 ** 61 + 12345, 3), (8, 128, 0.5, 0, 1)])
def test_dropout_mask_matches_oracle(rows, D, p, seed, stream):
    """The library's counter-based mask, bit for bit against the numpy restatement."""
    out = torch.empty(rows, D, device="cuda")
    rc = cabi.lib().rrt_dropout_mask(out.data_ptr(), rows * D, p, seed, stream, G.stream_ptr())
    cabi.check(rc, "rrt_dropout_mask")
    ref = O.dropout_mask(rows, D, p, seed, stream, dtype=torch.float32)
    assert torch.equal(out.cpu(), ref)


def test_device_step_state_shifts_the_dropout_seed():
    """rrt_set_step_state: the mask a kernel evaluates is the one of (seed argument + the u64 in the device buffer),
    so a CUDA-graph replay draws a new mask by updating 8 bytes; switching the state off restores the plain seed."""
    lib = cabi.lib()
    state = torch.zeros(2, dtype=torch.int64, device="cuda")
    out = torch.empty(300, 512, device="cuda")
    try:
        for dev_seed in (0, 7, 2 ** 40 + 3):
            state[0] = dev_seed
            cabi.check(lib.rrt_set_step_state(state.data_ptr()), "rrt_set_step_state")
            cabi.check(lib.rrt_dropout_mask(out.data_ptr(), out.numel(), 0.25, 1000, 3, G.stream_ptr()), "mask")
            assert torch.equal(out.cpu(), O.dropout_mask(300, 512, 0.25, 1000 + dev_seed, 3, dtype=torch.float32))
    finally:
        cabi.check(lib.rrt_set_step_state(None), "rrt_set_step_state")
    cabi.check(lib.rrt_dropout_mask(out.data_ptr(), out.numel(), 0.25, 1000, 3, G.stream_ptr()), "mask")
    assert torch.equal(out.cpu(), O.dropout_mask(300, 512, 0.25, 1000, 3, dtype=torch.float32))
    assert lib.rrt_set_step_state(state.data_ptr() + 8) == cabi.RRT_E_INVALID      # 16-byte alignment


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_training_mode_matches_reference_fixture(name):
    """.train() forward (proj_drop active) + backward through the CUDA kernels against the fixture the
    REFERENCE produced in .train() with the same masks installed (oracle/make_golden.py)."""
    cfg, w, x, gout, (p, seed, dp_rate, dp_keep), gold = load_train_case(name)
    m = RRTEncoder(**cfg.to_dict(), drop_out=p, drop_path=dp_rate or 0.0).cuda().train()
    m.load_state_dict({k: v.float() for k, v in w.items()}, strict=True)
    m._dropout_seed = seed
    if dp_keep is not None:     # stochastic depth: pin the outcome of the Bernoulli draws the fixture was made with
        m._drop_path_keep = [bool(k) for k in dp_keep]
    xd = x.float().cuda().requires_grad_()
    y = m(xd)
    (y * gout.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    e = train_errors(y, xd.grad, {n: q.grad for n, q in m.named_parameters()}, gold)
    print(name, {k: f"{v:.1e}" for k, v in e.items()})
    for k, v in e.items():
        tol = 1e-3 if k.startswith("out") else TOL_GRAD
        assert v < tol, (k, v, e)
    if dp_keep is not None:
        assert m.last_drop_path_keep == [bool(k) for k in dp_keep]
    if p > 0:   # a different seed gives a different (but equally valid) result; eval ignores drop_out
        m._dropout_seed = seed + 1
        with torch.no_grad():
            y2 = m(xd.detach())
            y_eval = m.eval()(xd.detach())
        assert rel(y2, y.detach()) > 1e-2
        assert rel(y_eval, y.detach()) > 1e-2
        ref_eval = O.encoder_forward(x, w, cfg, "spec")
        assert O.rel_err(y_eval.cpu(), ref_eval) < 1e-3


def test_training_dropout_is_reproducible_under_manual_seed():
    cfg = O.EncoderConfig()
    m = G.make_encoder(cfg, O.make_weights(cfg, 3)).train()
    x = O.make_bag(300, 512, 4).float().cuda()
    torch.manual_seed(7)
    a = m(x)
    s1 = m.last_dropout_seed
    b = m(x)
    torch.manual_seed(7)
    c = m(x)
    assert m.last_dropout_seed == s1
    assert torch.equal(a, c) and not torch.equal(a, b)


def test_backward_rejects_unsupported():
    cfg = O.EncoderConfig(ffn=True)                     # the FFN ablation has no backward
    m = G.make_encoder(cfg, O.make_weights(cfg, 3))
    x = O.make_bag(200, 512, 4).float().cuda().requires_grad_()
    with pytest.raises(NotImplementedError):
        m(x)
    # drop_path (stochastic depth): the Bernoulli draws come from torch's CPU generator (reproducible), every block
    # is dropped now and then, and eval ignores it
    md = RRTEncoder(drop_path=0.5, n_layers=3, need_init=True).cuda().train()
    torch.manual_seed(11)
    keeps = []
    for _ in range(12):
        md(x.detach())
        keeps.append(tuple(md.last_drop_path_keep))
    assert len(set(keeps)) > 1 and all(len(k) == 3 for k in keeps)
    torch.manual_seed(11)
    md(x.detach())
    assert tuple(md.last_drop_path_keep) == keeps[0]
    with torch.no_grad():
        assert torch.equal(md.eval()(x.detach()), md(x.detach()))
    m2 = G.make_encoder(O.EncoderConfig(), O.make_weights(O.EncoderConfig(), 3)).train()
    with pytest.raises(NotImplementedError):     # the batch entry point is inference-only
        m2.forward_bags([x.detach()])


def test_backward_limits_are_reported_before_the_forward():
    """Limits that depend on the bag or on derived sizes raise at the call, not inside loss.backward()."""
    cases = [(dict(mlp_dim=256, n_heads=4, crmsa_heads=4, crmsa_mlp=True), 300),   # crmsa_mlp hidden width 64
             (dict(crmsa_heads=4), 300),                                           # CR-MSA head_dim 128
             (dict(), 20000)]                                                      # R-MSA regions of 324 tokens
    for over, L in cases:
        m = RRTEncoder(**over).cuda().train()
        x = torch.randn(1, L, m.final_dim, device="cuda")
        with pytest.raises(NotImplementedError, match="not covered"):
            m(x)
        with torch.no_grad():
            m.eval()(x)     # inference is unaffected
