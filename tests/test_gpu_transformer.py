"""DCP's transformer on the tcgen05 pipelines (utils/transformer_fused.py): channel-major linear layers, attention in
two score passes + the p.v GEMM, LayerNorm.  Floating-point kernels: checked against fp64 evaluations of the
reference's expressions (utils/transformer.py:17-23,128-137,175-194) with the fp32-GEMM error model, and the whole
forward against the module's own torch path (TF32 disabled)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_linear_cm_bias_relu_residual():
    from learning3d_b200.utils.transformer_fused import linear_cm
    torch.manual_seed(0)
    for (B, K, M, N) in [(3, 512, 512, 1024), (2, 512, 1024, 260), (1, 1024, 512, 64), (2, 72, 200, 132)]:
        lin = torch.nn.Linear(K, M).to(DEV)
        x = torch.randn(B, K, N, device=DEV)
        res = torch.randn(B, M, N, device=DEV)
        for relu, r in ((False, None), (True, None), (False, res)):
            got = linear_cm(x, lin, relu=relu, residual=r).double()
            y = torch.einsum("mk,bkn->bmn", lin.weight.double(), x.double()) + lin.bias.double()[None, :, None]
            mag = torch.einsum("mk,bkn->bmn", lin.weight.double().abs(), x.double().abs()) + 1.0
            if relu:
                y = y.clamp_min(0)
            if r is not None:
                y = y + r.double()
            err = ((got - y).abs() / mag).max().item()
            assert err < 4e-6, (B, K, M, N, relu, err)


def test_linear_cm_transposed_output_is_the_same_numbers():
    """l3d_linear_cm_t writes [B, N, M]: bit-identical to the transpose of l3d_linear_cm's result."""
    from learning3d_b200.utils.transformer_fused import linear_cm, linear_cm_t
    torch.manual_seed(7)
    for (B, K, M, N) in [(3, 512, 512, 1024), (2, 72, 200, 132), (1, 512, 512, 260)]:
        lin = torch.nn.Linear(K, M).to(DEV)
        x = torch.randn(B, K, N, device=DEV)
        assert torch.equal(linear_cm_t(x, lin), linear_cm(x, lin).transpose(1, 2).contiguous()), (B, K, M, N)


def test_layernorm_cm():
    from learning3d_b200.utils.transformer import _Norm
    from learning3d_b200.utils.transformer_fused import layernorm_cm
    torch.manual_seed(1)
    for (B, D, N) in [(2, 512, 1024), (3, 96, 77), (1, 8, 5)]:
        norm = _Norm(D).to(DEV)
        with torch.no_grad():
            norm.a_2.uniform_(0.5, 1.5); norm.b_2.normal_(0, 0.2)
        x = torch.randn(B, D, N, device=DEV) * 3 + 1
        got = layernorm_cm(x, norm)
        xd = x.double().transpose(1, 2)
        want = (norm.a_2.double() * (xd - xd.mean(-1, keepdim=True)) / (xd.std(-1, keepdim=True) + norm.eps)
                + norm.b_2.double()).transpose(1, 2)
        assert (got.double() - want).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,h,Nq,Nk", [(2, 4, 256, 256), (1, 4, 132, 520), (3, 2, 1024, 1024)])
def test_attention_cm_vs_fp64(B, h, Nq, Nk):
    from learning3d_b200.utils.transformer import MultiHeadedAttention
    from learning3d_b200.utils.transformer_fused import attention_cm
    torch.manual_seed(B + Nq)
    d = h * 128
    attn = MultiHeadedAttention(h, d).to(DEV).eval()
    xq = torch.randn(B, d, Nq, device=DEV)
    xkv = torch.randn(B, d, Nk, device=DEV) * 1.5
    res = torch.randn(B, d, Nq, device=DEV)
    with torch.no_grad():
        got = attention_cm(attn, xq, xkv, res).double()
        L = [(l.weight.double(), l.bias.double()) for l in attn.linears]
        q = (xq.double().transpose(1, 2) @ L[0][0].t() + L[0][1]).view(B, Nq, h, 128).transpose(1, 2)
        k = (xkv.double().transpose(1, 2) @ L[1][0].t() + L[1][1]).view(B, Nk, h, 128).transpose(1, 2)
        v = (xkv.double().transpose(1, 2) @ L[2][0].t() + L[2][1]).view(B, Nk, h, 128).transpose(1, 2)
        p = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(128), dim=-1)
        ctx = (p @ v).transpose(1, 2).reshape(B, Nq, d)
        want = (ctx @ L[3][0].t() + L[3][1]).transpose(1, 2) + res.double()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    print("attention B=%d h=%d Nq=%d Nk=%d: max |err| = %.3g (|out| max %.3g)" % (B, h, Nq, Nk, err, scale))
    assert err < 2e-5 * max(1.0, scale)


def test_transformer_forward_fused_vs_torch_path():
    from learning3d_b200.utils.transformer import Transformer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    net = Transformer(512, 1, 0.0, 1024, 4).to(DEV).eval()
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "a_2"):
                m.a_2.uniform_(0.8, 1.2); m.b_2.normal_(0, 0.1)
    src = torch.randn(4, 512, 1024, device=DEV)
    tgt = torch.randn(4, 512, 1024, device=DEV)
    with torch.no_grad():
        a, b = net(src, tgt)
        wa, wb = net._l3d_torch_forward(src, tgt)
    for got, want, name in ((a, wa, "src_embedding"), (b, wb, "tgt_embedding")):
        err = (got - want).abs().max().item()
        print("transformer %s: max |fused - torch| = %.3g (|x| max %.3g)" % (name, err, want.abs().max().item()))
        assert err < 3e-5 * max(1.0, want.abs().max().item())
    # autograd keeps working (torch path)
    src.requires_grad_(True)
    a, b = net(src, tgt)
    (a.mean() + b.mean()).backward()
    assert src.grad is not None


def test_attention_protocols_agree():
    """precise stats + normalised probabilities vs the fast protocol (one-pass TF32 max, unnormalised probabilities,
    row sums divided out by the p.v GEMM): same context to fp32-GEMM accuracy."""
    from learning3d_b200 import _C
    lib = _C.lib()
    torch.manual_seed(12)
    BH, D, Nq, Nk = 8, 128, 260, 512
    q = torch.randn(BH, D, Nq, device=DEV); k = torch.randn(BH, D, Nk, device=DEV) * 2
    vt = torch.randn(BH // 4, Nk, 4 * 128, device=DEV)
    st = _C.stream()
    outs = []
    for precise, normalized in ((1, 1), (0, 0)):
        stats = torch.empty(BH, Nq, 2, device=DEV); pt = torch.empty(BH, Nk, Nq, device=DEV)
        _C.check(lib.l3d_attention_stats(_C.ptr(q), _C.ptr(k), BH, D, Nq, Nk, precise, _C.ptr(stats), st))
        _C.check(lib.l3d_attention_probs_t(_C.ptr(q), _C.ptr(k), _C.ptr(stats), BH, D, Nq, Nk, normalized, _C.ptr(pt), st))
        div = None if normalized else stats[:, :, 1].contiguous()
        ctx = torch.empty(BH, 128, Nq, device=DEV)
        _C.check(lib.l3d_linear_cm(_C.ptr(vt), _C.ptr(pt), _C.ptr(None), _C.ptr(None), _C.ptr(div), BH, 128, Nk, Nq, 0, 4,
                                   _C.ptr(ctx), st))
        outs.append(ctx)
    p = torch.softmax(torch.einsum("bdq,bdk->bqk", q.double(), k.double()) / math.sqrt(D), dim=-1)
    v = vt.double().view(BH // 4, Nk, 4, 128).permute(0, 2, 3, 1).reshape(BH, 128, Nk)           # [BH, d_v, Nk]
    want = torch.einsum("bdk,bqk->bdq", v, p)
    # peaky softmax here (scores ~ N(0, 4)): the fp32-GEMM error of a score (<= 4e-6 * sum|q||k| ~ 6e-4) moves a
    # probability by up to ~8e-5 relative; both protocols sit at 1-3e-5 absolute on outputs of magnitude ~3
    errs = [(o.double() - want).abs().max().item() for o in outs]
    print("attention protocols: precise %.3g, fast %.3g, between %.3g" % (errs[0], errs[1], (outs[0] - outs[1]).abs().max().item()))
    assert max(errs) < 6e-5
    assert (outs[0] - outs[1]).abs().max().item() < 6e-5


def test_bound_referenced_softmax_and_its_fallback():
    """l3d_attention_bounds: the Cauchy-Schwarz row bound as the exponent reference (no statistics pass); with large
    scores the device flag is raised and l3d_attention_stats_if replaces the bounds by the true maxima."""
    from learning3d_b200 import _C
    lib = _C.lib()
    torch.manual_seed(13)
    BH, D, Nq, Nk = 8, 128, 260, 512
    st = _C.stream()
    for scale, want_flag in ((1.0, 0), (25.0, 1)):
        q = torch.randn(BH, D, Nq, device=DEV) * scale; k = torch.randn(BH, D, Nk, device=DEV)
        stats = torch.full((BH, Nq, 2), float("nan"), device=DEV)
        ws = torch.empty(BH + 1, dtype=torch.int32, device=DEV)
        assert lib.l3d_attention_bounds_ws_bytes(BH) == 4 * (BH + 1)
        _C.check(lib.l3d_attention_bounds(_C.ptr(q), _C.ptr(k), BH, D, Nq, Nk, _C.ptr(stats), _C.ptr(ws), st))
        bound = stats[:, :, 0].clone()
        s = torch.einsum("bdq,bdk->bqk", q.double(), k.double()) / math.sqrt(D) * math.log2(math.e)
        assert (bound.double() >= s.max(-1).values).all()                       # a true upper bound of every row
        assert int(ws[BH].item()) == want_flag, (scale, int(ws[BH].item()), bound.max().item())
        _C.check(lib.l3d_attention_stats_if(_C.ptr(q), _C.ptr(k), BH, D, Nq, Nk, _C._P(ws.data_ptr() + 4 * BH), _C.ptr(stats), st))
        if want_flag:
            got_max = stats[:, :, 0].double()
            assert ((got_max - s.max(-1).values).abs() <= 2e-3 * s.abs().max(-1).values + 1e-3).all()   # one TF32 pass
        else:
            assert torch.equal(stats[:, :, 0], bound)                           # the statistics launch returned at once
        pt = torch.empty(BH, Nk, Nq, device=DEV)
        _C.check(lib.l3d_attention_probs_t(_C.ptr(q), _C.ptr(k), _C.ptr(stats), BH, D, Nq, Nk, 0, _C.ptr(pt), st))
        p = pt.double().transpose(1, 2) / stats[:, :, 1].double().unsqueeze(-1)
        want = torch.softmax(s * math.log(2.0), dim=-1)
        assert torch.isfinite(pt).all()
        # a probability moves by (score error) x p: 3xTF32 scores carry <= ~4e-6 * sum|q||k|
        tol = 2e-5 if scale == 1.0 else 2e-3
        assert (p - want).abs().max().item() < tol, (scale, (p - want).abs().max().item())
    assert lib.l3d_soft_correspondence_status() == 0
