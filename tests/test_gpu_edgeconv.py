"""EdgeConv stack of DGCNN (models/dgcnn.py:32-48) on tcgen05: l3d_edgeconv_layer1 + l3d_conv1x1_bn_relu_maxk.

Floating-point kernels: the checker is torch evaluated in fp64 on the same GPU (the oracle for a GEMM is the exact
product), with the fp32-GEMM error model  |err| <= tol * sum_k |w||x| * |scale|  (3xTF32: ~2^-21 per product).
The whole-module tests compare against the reference's own layer sequence (torch fp32, TF32 disabled) at the
1e-5 relative bar of north_star.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _layer(wt, x, scale, shift, G, relu=True, want_h=True, want_pool=True, coff=0, ctot=None):
    from learning3d_b200 import _C
    lib = _C.lib()
    K, M = wt.shape
    B, _, P = x.shape
    h = torch.empty(B, M, P, device=DEV) if want_h else None
    ctot = ctot or M
    pool = torch.full((B, ctot, P // G), float("nan"), device=DEV) if want_pool else None
    _C.check(lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(scale), _C.ptr(shift), B, M, K, P, G,
                                          1 if relu else 0, _C.ptr(h), _C.ptr(pool), ctot * (P // G), coff,
                                          _C.stream()), "conv1x1")
    torch.cuda.synchronize()
    assert lib.l3d_edgeconv_status() == 0
    return h, pool


def _want(wt, x, scale, shift, G, relu):
    y = torch.einsum("km,bkp->bmp", wt.double(), x.double())
    mag = torch.einsum("km,bkp->bmp", wt.double().abs(), x.double().abs()) * scale.double().abs()[None, :, None]
    y = y * scale.double()[None, :, None] + shift.double()[None, :, None]
    if relu:
        y = y.clamp_min(0)
    B, M, P = y.shape
    pool = y.view(B, M, P // G, G).max(-1)[0] if P % G == 0 else None
    pmag = mag.view(B, M, P // G, G).max(-1)[0] if P % G == 0 else None
    return y, mag, pool, pmag


@pytest.mark.parametrize("B,M,K,P,G", [
    (2, 64, 64, 20 * 256, 20),       # EdgeConv layer 2 shape (per item: N=256, k=20)
    (2, 128, 64, 20 * 256, 20),      # layer 3
    (2, 256, 128, 20 * 256, 20),     # layer 4: CTA pairs
    (3, 512, 512, 1024, 1),          # conv5 (emb 512): pairs, two channel blocks, K = 512
    (1, 96, 40, 8 * 100, 8),         # ragged: M, K not multiples of the tile, short last tile, k = 8
    (2, 200, 72, 4 * 333, 4),        # pairs with a partly empty second CTA, odd sizes
    (1, 64, 64, 20 * 12, 20),        # a single tile
    (70, 64, 16, 40, 20),            # many tiny items (more units than SMs would need for one wave of pairs)
])
def test_conv1x1_bn_relu_maxk_vs_fp64(B, M, K, P, G):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + M + K + P)
    wt = (torch.randn(K, M, generator=g) / np.sqrt(K)).to(DEV)
    x = torch.randn(B, K, P, generator=g).abs_().to(DEV)          # post-ReLU activations are non-negative
    scale = (0.5 + torch.rand(M, generator=g)).to(DEV)
    scale[::7] *= -1                                               # negative BN weights happen
    shift = (0.2 * torch.randn(M, generator=g)).to(DEV)
    for relu in (True, False):
        h, pool = _layer(wt, x, scale, shift, G, relu=relu)
        y, mag, wp, pmag = _want(wt, x, scale, shift, G, relu)
        err = ((h.double() - y).abs() / (mag + 1e-6)).max().item()
        perr = ((pool.double() - wp).abs() / (pmag + 1e-6)).max().item()
        print("B=%d M=%d K=%d P=%d G=%d relu=%d: max err / sum|w||x| = %.2e (pooled %.2e)" % (B, M, K, P, G, relu, err, perr))
        assert err < 4e-6 and perr < 4e-6
        # pooled output must be exactly the max of what was written
        assert torch.equal(pool, h.view(B, M, P // G, G).max(-1)[0])


def test_conv1x1_pool_only_and_channel_offset():
    """Layer 4 writes only the pooled rows, at a channel offset inside the concatenated [B, 512, N] buffer."""
    g = torch.Generator().manual_seed(7)
    B, M, K, N, k = 2, 256, 128, 64, 20
    wt = (torch.randn(K, M, generator=g) / np.sqrt(K)).to(DEV)
    x = torch.rand(B, K, N * k, generator=g).to(DEV)
    scale = torch.ones(M, device=DEV)
    shift = torch.zeros(M, device=DEV)
    _, pool = _layer(wt, x, scale, shift, k, want_h=False, coff=256, ctot=512)
    y, mag, wp, pmag = _want(wt, x, scale, shift, k, True)
    assert torch.isnan(pool[:, :256]).all()                       # rows outside the offset window untouched
    assert ((pool[:, 256:].double() - wp).abs() / (pmag + 1e-6)).max().item() < 4e-6


def test_conv1x1_argument_checks():
    from learning3d_b200 import _C
    lib = _C.lib()
    wt = torch.zeros(64, 64, device=DEV); x = torch.zeros(1, 64, 42, device=DEV); s = torch.zeros(64, device=DEV)
    h = torch.zeros(1, 64, 42, device=DEV)
    rc = lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(s), _C.ptr(s), 1, 64, 64, 42, 1, 1, _C.ptr(h),
                                      _C.ptr(None), 0, 0, _C.stream())
    assert rc == -2                                                # P % 4 != 0: unsupported, not garbage
    rc = lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(s), _C.ptr(s), 1, 64, 64, 40, 3, 1, _C.ptr(h),
                                      _C.ptr(h), 0, 0, _C.stream())
    assert rc == -1                                                # P % G != 0
    assert lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(s), _C.ptr(s), 0, 64, 64, 40, 1, 1, _C.ptr(h),
                                        _C.ptr(None), 0, 0, _C.stream()) == 0


def test_edgeconv_layer1_vs_torch():
    from learning3d_b200 import _C
    from learning3d_b200.utils import knn
    from oracle import ref_torch
    lib = _C.lib()
    torch.manual_seed(3)
    B, N, k = 3, 512, 20
    x = torch.rand(B, 3, N, device=DEV)
    idx = knn(x, k)
    w = torch.randn(64, 6) * 0.4
    scale = 0.5 + torch.rand(64); shift = 0.1 * torch.randn(64)
    h1 = torch.empty(B, 64, N * k, device=DEV)
    pool = torch.empty(B, 64, N, device=DEV)
    _C.check(lib.l3d_edgeconv_layer1(_C.ptr(x), _C.ptr(idx), _C._P(w.data_ptr()), _C._P(scale.data_ptr()),
                                     _C._P(shift.data_ptr()), B, N, k, 64, _C.ptr(h1), _C.ptr(pool), 64 * N, 0,
                                     _C.stream()), "layer1")
    feat = ref_torch.get_graph_feature(x, k=k).double()              # the reference's gather [B,6,N,k]
    want = torch.einsum("ci,bink->bcnk", w.double().to(DEV), feat)
    want = (want * scale.double().to(DEV)[None, :, None, None] + shift.double().to(DEV)[None, :, None, None]).clamp_min(0)
    # rows where the reference's own topk picked a different (tied) neighbour set are excluded
    ridx = ref_torch.knn(x, k)
    same = (ridx.sort(-1)[0] == idx.sort(-1)[0]).all(-1)             # [B, N]
    assert same.float().mean() > 0.999
    got = h1.view(B, 64, N, k).double()
    # neighbour ORDER may differ between tied keys; the max over k and the multiset of values do not
    diff = (got.sort(-1)[0] - want.sort(-1)[0]).abs().amax(dim=(1, 3))
    assert diff[same].max().item() < 1e-5
    assert torch.equal(pool, h1.view(B, 64, N, k).max(-1)[0])


@pytest.mark.parametrize("emb,B,N", [(512, 4, 1024), (1024, 2, 256)])
def test_dgcnn_forward_fused_vs_reference_layers(emb, B, N):
    """DGCNN.forward in eval mode (fused tcgen05 stack) against the reference's layer sequence
    (models/dgcnn.py:32-48) in torch fp32 with TF32 disabled, on the reference's own graph construction."""
    from learning3d_b200.models import DGCNN
    from oracle import ref_torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(11)
    net = DGCNN(emb_dims=emb).to(DEV)
    with torch.no_grad():                                            # non-trivial BatchNorm statistics
        for i in range(1, 6):
            bn = getattr(net, "bn%d" % i)
            bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
    net.eval()
    x = torch.rand(B, N, 3, device=DEV)
    with torch.no_grad():
        got = net(x)
        feat = ref_torch.get_graph_feature(x.permute(0, 2, 1).contiguous(), k=20)
        h, pooled = feat, []
        for i in range(1, 5):
            h = torch.relu(getattr(net, "bn%d" % i)(getattr(net, "conv%d" % i)(h)))
            pooled.append(h.max(dim=-1, keepdim=True)[0])
        want = torch.relu(net.bn5(net.conv5(torch.cat(pooled, 1)))).view(B, -1, N)
        from learning3d_b200.utils import knn
        xt = x.permute(0, 2, 1).contiguous()
        same = (ref_torch.knn(xt, 20).sort(-1)[0] == knn(xt, 20).sort(-1)[0]).all(-1)       # [B, N]
    assert got.shape == want.shape == (B, emb, N)
    scale = want.abs().max().item()
    diff = (got - want).abs().amax(dim=1)                            # [B, N]
    print("DGCNN-%d fused vs torch layers: max |diff| = %.3g on %d/%d untied points (|y| max %.3g)" % (
        emb, diff[same].max().item(), int(same.sum()), same.numel(), scale))
    assert same.float().mean() > 0.999
    assert diff[same].max().item() <= 1e-5 * max(1.0, scale)
    # training mode keeps the torch layers (autograd) and still runs
    net.train()
    y = net(x)
    y.mean().backward()
    assert net.conv1.weight.grad is not None


def test_chunked_tensor_maps_move_the_same_bytes():
    """The 4-D tensor maps (one bulk-tensor instruction per operand tile, tc05.cuh:make_dn_tmap4) must fill shared memory
    exactly like the 32-point boxes they replace: results bit-identical, including the k = 20 tiles that start 16
    positions past a chunk boundary (shifted map) and the ragged last tile of a row (falls back to the boxes)."""
    from learning3d_b200 import _C
    lib = _C.lib()
    torch.manual_seed(21)
    cases = [(2, 64, 64, 5120, 20), (2, 256, 128, 2560, 20), (1, 128, 64, 1024 * 16, 16), (2, 512, 512, 1024, 1), (1, 64, 64, 20 * 112, 20), (1, 64, 64, 20 * 77, 20)]
    try:
        for (B, M, K, P, G) in cases:
            wt = torch.randn(K, M, device=DEV) * 0.2
            x = torch.randn(B, K, P, device=DEV)
            sc = torch.rand(M, device=DEV) + 0.5
            sh = torch.randn(M, device=DEV) * 0.1
            res = []
            for mode in (3, 0):                                   # 3: force the 32-point boxes, 0: automatic
                lib.l3d_debug_soft_correspondence_force_generic(mode)
                res.append(_layer(wt, x, sc, sh, G, want_h=(M <= 128 or G == 1), want_pool=(G > 1)))
            for a, b in zip(res[0], res[1]):
                assert (a is None and b is None) or torch.equal(a, b), (B, M, K, P, G)
        src = torch.randn(3, 512, 1024, device=DEV); tgt = torch.randn(3, 512, 1024, device=DEV); xyz = torch.randn(3, 3, 1024, device=DEV)
        outs = []
        for mode in (3, 0):
            lib.l3d_debug_soft_correspondence_force_generic(mode)
            o = torch.empty(3, 3, 1024, device=DEV)
            _C.check(lib.l3d_soft_correspondence(_C.ptr(src), _C.ptr(tgt), _C.ptr(xyz), 3, 512, 1024, 1024, _C.ptr(o), _C.stream()))
            outs.append(o)
        assert torch.equal(outs[0], outs[1]) and lib.l3d_soft_correspondence_status() == 0
    finally:
        lib.l3d_debug_soft_correspondence_force_generic(0)
