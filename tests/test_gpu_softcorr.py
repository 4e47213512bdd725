"""GPU: fused soft-correspondence kernel (tcgen05 3xTF32 + online softmax) against fp64 evaluation of
utils/svd.py:23-28, on all three operand pipelines (TMA MN-major with CTA pairs / TMA single CTA / generic LDG K-major).

Tolerances (stated, floating point): raw scores within 4e-6 * sum_d |a_d||b_d| of fp64 (3xTF32 keeps ~21
mantissa bits per product; the fp32 SGEMM the reference calls is in the same class), src_corr within 2e-5
absolute of fp64 for O(1) coordinates."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(src_emb, tgt_emb, tgt):
    s = torch.matmul(src_emb.double().transpose(2, 1), tgt_emb.double())
    p = torch.softmax(s / math.sqrt(src_emb.shape[1]), dim=2)
    bound = torch.matmul(src_emb.double().abs().transpose(2, 1), tgt_emb.double().abs())
    return s, torch.matmul(tgt.double(), p.transpose(2, 1)), bound


def _run_debug(src_emb, tgt_emb, tgt):
    from learning3d_b200 import _C
    lib = _C.lib()
    B, D, Ns = src_emb.shape
    Nt = tgt_emb.shape[2]
    out = torch.full((B, 3, Ns), float("nan"), device=DEV)
    sc = torch.full((B, Ns, Nt), float("nan"), device=DEV)
    _C.check(lib.l3d_debug_soft_correspondence_scores(_C.ptr(src_emb), _C.ptr(tgt_emb), _C.ptr(tgt), B, D, Ns, Nt,
                                                      _C.ptr(out), _C.ptr(sc), _C.stream()))
    assert lib.l3d_soft_correspondence_status() == 0
    return out, sc


@pytest.fixture(params=["auto", "generic", "tma_single_cta"])
def pipeline(request):
    """auto = TMA operands + CTA pairs (tcgen05 cta_group::2) when eligible; the other two force a fallback."""
    from learning3d_b200 import _C
    _C.lib().l3d_debug_soft_correspondence_force_generic({"auto": 0, "generic": 1, "tma_single_cta": 2}[request.param])
    yield request.param
    _C.lib().l3d_debug_soft_correspondence_force_generic(0)


@pytest.mark.parametrize("B,D,Ns,Nt", [
    (1, 32, 128, 128),      # one tile, one K block
    (2, 64, 128, 384),      # several target tiles (accumulator double buffering)
    (2, 96, 200, 332),      # ragged rows / columns, 3 K blocks
    (3, 80, 1000, 516),     # K tail (80 = 2.5 blocks), row and column tails
    (1, 8, 4, 4),           # tiny
    (2, 512, 1024, 1024),   # DCP shape (C3 per item)
    (1, 100, 131, 77),      # odd sizes: always the generic pipeline
    (2, 40, 512, 256),      # CTA pairs with a K tail (40 = 2.5 x 16 channels, zero-filled by TMA)
    (1, 72, 260, 516),      # CTA pair whose second member is mostly out of range, column tail
])
def test_scores_and_correspondences_vs_fp64(pipeline, B, D, Ns, Nt):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + D + Ns + Nt)
    a = torch.randn(B, D, Ns, device=DEV, generator=g)
    b = torch.randn(B, D, Nt, device=DEV, generator=g)
    t = torch.rand(B, 3, Nt, device=DEV, generator=g) * 2 - 1
    out, sc = _run_debug(a, b, t)
    s_ref, o_ref, bound = _ref(a, b, t)
    assert not torch.isnan(sc).any() and not torch.isnan(out).any()
    assert ((sc.double() - s_ref).abs() <= 4e-6 * bound + 1e-30).all(), (pipeline, float((sc.double() - s_ref).abs().max()))
    assert (out.double() - o_ref).abs().max().item() <= 2e-5, pipeline


def test_against_real_reference_fixture(pipeline, golden_dir):
    """src_corr computed by the unmodified reference SVDHead on CPU (tests/golden/make_golden.py)."""
    from learning3d_b200.utils.svd import soft_correspondence
    g = np.load(f"{golden_dir}/svd_head.npz")
    tgt = torch.from_numpy(np.ascontiguousarray(g["tgt"].transpose(0, 2, 1))).to(DEV)
    out = soft_correspondence(torch.from_numpy(g["src_emb"]).to(DEV), torch.from_numpy(g["tgt_emb"]).to(DEV), tgt)
    assert np.abs(out.cpu().numpy() - g["src_corr"]).max() <= 2e-5


def test_peaky_softmax_and_large_scores(pipeline):
    """Trained-like embeddings: scores of a few thousand, softmax essentially one-hot."""
    g = torch.Generator(device=DEV).manual_seed(7)
    B, D, N = 2, 512, 1024
    a = 3 * torch.randn(B, D, N, device=DEV, generator=g)
    b = a + 0.1 * torch.randn(B, D, N, device=DEV, generator=g)
    t = torch.rand(B, 3, N, device=DEV, generator=g)
    out, _ = _run_debug(a, b, t)
    _, o_ref, _ = _ref(a, b, t)
    assert (out.double() - o_ref).abs().max().item() <= 2e-5
    # the matched point dominates: src_corr ~ tgt
    assert (out - t).abs().max().item() < 1e-3


def test_bitwise_deterministic_and_public_wrapper(pipeline):
    from learning3d_b200.utils.svd import soft_correspondence
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(4, 256, 768, device=DEV, generator=g); b = torch.randn(4, 256, 768, device=DEV, generator=g)
    t = torch.rand(4, 3, 768, device=DEV, generator=g)
    o1 = soft_correspondence(a, b, t); o2 = soft_correspondence(a, b, t)
    assert torch.equal(o1, o2)
    _, o_ref, _ = _ref(a, b, t)
    assert (o1.double() - o_ref).abs().max().item() <= 2e-5


def test_argument_checks():
    from learning3d_b200 import _C
    from learning3d_b200.utils.svd import soft_correspondence
    lib = _C.lib()
    a = torch.randn(1, 8, 4, device=DEV); t = torch.rand(1, 3, 4, device=DEV); o = torch.empty(1, 3, 4, device=DEV)
    assert lib.l3d_soft_correspondence(_C.ptr(a), _C.ptr(a), _C.ptr(t), 0, 8, 4, 4, _C.ptr(o), _C.stream()) == 0   # B == 0
    assert lib.l3d_soft_correspondence(_C.ptr(a), _C.ptr(a), _C.ptr(t), 1, 8, 4, 0, _C.ptr(o), _C.stream()) < 0    # empty softmax
    assert lib.l3d_soft_correspondence(None, _C.ptr(a), _C.ptr(t), 1, 8, 4, 4, _C.ptr(o), _C.stream()) < 0
    with pytest.raises(RuntimeError):
        soft_correspondence(a.cpu(), a.cpu(), t.cpu())          # CUDA only, no CPU fallback
    with pytest.raises(ValueError):
        soft_correspondence(a, a, torch.rand(1, 3, 5, device=DEV))


def test_svd_head_inference_path_matches_reference_ops():
    """SVDHead.forward without grad runs the fused kernel; with grad the reference's torch ops: same R, t."""
    from learning3d_b200.utils.svd import SVDHead
    g = torch.Generator(device=DEV).manual_seed(3)
    B, D, N = 4, 512, 1024
    src = torch.rand(B, N, 3, device=DEV, generator=g) - 0.5
    ang = 0.6
    Rz = torch.tensor([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1.0]], device=DEV)
    tgt = src @ Rz.T + torch.tensor([0.1, -0.2, 0.3], device=DEV)
    emb = 2 * torch.randn(B, D, N, device=DEV, generator=g)
    emb_t = emb + 0.05 * torch.randn(B, D, N, device=DEV, generator=g)
    head = SVDHead(D, input_shape="bnc").to(DEV)
    with torch.no_grad():
        R1, t1 = head(emb, emb_t, src, tgt)
    emb_g = emb.clone().requires_grad_(True)
    R2, t2 = head(emb_g, emb_t, src, tgt)
    assert (R1 - R2).abs().max().item() <= 2e-5 and (t1 - t2).abs().max().item() <= 2e-5
    assert (R1 - Rz).abs().max().item() < 1e-2


@pytest.mark.parametrize("split", [0, 2, 3, 4, -1])
@pytest.mark.parametrize("B,D,Ns,Nt", [(1, 64, 128, 1024), (2, 512, 1024, 1024), (1, 96, 260, 1300)])
def test_target_range_split_and_merge(pipeline, split, B, D, Ns, Nt):
    """Small batches spread the target tiles of a row block over several CTAs (gridDim.z) and merge the
    partial softmax states: same result as the unsplit kernel, for automatic and forced split counts."""
    from learning3d_b200 import _C
    lib = _C.lib()
    g = torch.Generator(device=DEV).manual_seed(100 + split + B + Ns + Nt)
    a = torch.randn(B, D, Ns, device=DEV, generator=g)
    b = torch.randn(B, D, Nt, device=DEV, generator=g)
    t = torch.rand(B, 3, Nt, device=DEV, generator=g) * 2 - 1
    lib.l3d_debug_soft_correspondence_split(split)
    try:
        out, sc = _run_debug(a, b, t)
    finally:
        lib.l3d_debug_soft_correspondence_split(0)
    s_ref, o_ref, bound = _ref(a, b, t)
    assert not torch.isnan(out).any() and not torch.isnan(sc).any()
    assert ((sc.double() - s_ref).abs() <= 4e-6 * bound + 1e-30).all()
    assert (out.double() - o_ref).abs().max().item() <= 2e-5, (pipeline, split)


@pytest.mark.parametrize("B,D,Ns,Nt", [(2, 128, 256, 256), (3, 512, 1024, 1024), (1, 256, 132, 520)])
def test_soft_correspondence_backward_vs_fp64_autograd(B, D, Ns, Nt):
    """_SoftCorr (fused forward + tcgen05 recompute-backward) against fp64 autograd of the reference's three lines
    (utils/svd.py:23-28)."""
    import math
    from learning3d_b200.utils.svd import _SoftCorr
    torch.manual_seed(B * D + Ns)
    es = (0.3 * torch.randn(B, D, Ns, device="cuda")).requires_grad_(True)
    et = (0.3 * torch.randn(B, D, Nt, device="cuda")).requires_grad_(True)
    tgt = torch.rand(B, 3, Nt, device="cuda") - 0.5
    w = torch.randn(B, 3, Ns, device="cuda")
    corr = _SoftCorr.apply(es, et, tgt)
    (corr * w).sum().backward()
    es64 = es.detach().double().requires_grad_(True); et64 = et.detach().double().requires_grad_(True)
    scores = torch.softmax(torch.matmul(es64.transpose(2, 1), et64) / math.sqrt(D), dim=2)
    want = torch.matmul(tgt.double(), scores.transpose(2, 1))
    (want * w.double()).sum().backward()
    np.testing.assert_allclose(corr.detach().cpu().numpy(), want.detach().cpu().numpy(), atol=2e-5)
    for got, ref, name in ((es.grad, es64.grad, "d src_emb"), (et.grad, et64.grad, "d tgt_emb")):
        err = (got.double() - ref).abs().max().item()
        scale = ref.abs().max().item()
        print("softcorr backward %s B=%d D=%d: max |err| = %.3g (|grad| max %.3g)" % (name, B, D, err, scale))
        assert err <= 2e-5 * max(scale, 1e-3) + 1e-8


def test_svd_head_training_path_uses_tensor_core_backward():
    from learning3d_b200.utils import SVDHead
    torch.manual_seed(9)
    head = SVDHead(128).cuda()
    es = torch.randn(2, 128, 256, device="cuda", requires_grad=True)
    et = torch.randn(2, 128, 256, device="cuda", requires_grad=True)
    src = torch.rand(2, 256, 3, device="cuda"); tgt = torch.rand(2, 256, 3, device="cuda")
    from learning3d_b200 import _C
    n0 = _C.launch_count()
    R, t = head(es, et, src, tgt)
    (R.sum() + t.sum()).backward()
    assert _C.launch_count() - n0 >= 6          # fused forward, tail, tail backward, stats, dscores, 2 GEMMs
    assert torch.isfinite(es.grad).all() and torch.isfinite(et.grad).all() and es.grad.abs().sum() > 0
