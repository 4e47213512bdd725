"""GPU: knn() for C != 3 (PRNet's feature-space graphs) — tensor-core Gram matrix + warp selection.

A GEMM is not bit-reproducible across libraries (MKL, cuBLAS and our 3xTF32 all round differently), so the
bar is stated as a tolerance: with t = fp64 evaluation of the reference formula
(pd = -xx - inner - xx^T, model_common_utils.py:5-7), every returned row must (a) be sorted by t up to eps,
(b) contain only neighbours whose t is within eps of the true k-th best, eps = 4e-6 * (|x_i|^2 + |x_j|^2 scale),
and (c) agree with the fp64 top-k set wherever the k-th / (k+1)-th gap exceeds eps.  Exact ties (duplicated
feature vectors) must come out lower-index-first."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _true_keys(x):
    xd = x.double()
    inner = -2 * torch.matmul(xd.transpose(2, 1), xd)
    xx = (xd ** 2).sum(dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def _check(x, k, idx):
    t = _true_keys(x)                                   # [B,N,N]
    B, C, N = x.shape
    assert idx.shape == (B, N, k) and idx.dtype == torch.int64
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    got = torch.gather(t, 2, idx)                       # true keys of the returned neighbours, in returned order
    scale = (x.double() ** 2).sum(dim=1).max().item()
    eps = 8e-6 * max(scale, 1e-30)
    assert (got[:, :, 1:] <= got[:, :, :-1] + eps).all(), "rows not sorted by true key"
    top = torch.topk(t, k + 1 if k < N else k, dim=2).values
    kth = top[:, :, k - 1:k]
    assert (got >= kth - eps).all(), "a returned neighbour is worse than the true k-th by more than eps"
    # distinct indices per row
    s = torch.sort(idx, dim=2).values
    assert (s[:, :, 1:] != s[:, :, :-1]).all()
    if k < N:
        clear = (top[:, :, k - 1] - top[:, :, k]) > 2 * eps            # unambiguous rows
        want = torch.sort(torch.topk(t, k, dim=2).indices, dim=2).values
        assert torch.equal(s[clear], want[clear])
        assert clear.float().mean().item() > 0.9


@pytest.mark.parametrize("B,C,N,k", [
    (2, 64, 1024, 20),      # PRNet layer 2/3 shape
    (2, 128, 768, 20),      # PRNet layer 4
    (1, 64, 200, 20),       # ragged (generic operand pipeline: 200 % 4 == 0 -> TMA; 3-D tails)
    (2, 7, 131, 5),         # odd everything -> generic pipeline
    (1, 512, 2048, 16),     # two selection tiles, deep K
    (1, 32, 64, 40),        # k > 24: exact k-round path
    (3, 64, 40, 40),        # k == N
])
def test_feature_knn_against_fp64(B, C, N, k):
    from learning3d_b200.utils import knn
    g = torch.Generator(device=DEV).manual_seed(B + C + N + k)
    x = torch.randn(B, C, N, device=DEV, generator=g)
    _check(x, k, knn(x, k))


def test_relu_features_with_exact_duplicates():
    """Post-ReLU features: many zeros, and duplicated columns give exactly equal keys -> lower index first."""
    from learning3d_b200.utils import knn
    g = torch.Generator(device=DEV).manual_seed(9)
    B, C, N, k = 2, 64, 512, 20
    x = torch.relu(torch.randn(B, C, N, device=DEV, generator=g))
    x[:, :, 256:] = x[:, :, :256]                       # every point has an exact twin
    idx = knn(x, k)
    _check(x, k, idx)
    # the twin of column j (j + 256 or j - 256) has exactly the same key column -> they are adjacent and ordered
    t = _true_keys(x)
    pos_self = (idx == torch.arange(N, device=DEV).view(1, N, 1)).float().argmax(dim=2)
    twin = (torch.arange(N, device=DEV) + 256) % N
    pos_twin = (idx == twin.view(1, N, 1)).float().argmax(dim=2)
    has_both = ((idx == torch.arange(N, device=DEV).view(1, N, 1)).any(dim=2) & (idx == twin.view(1, N, 1)).any(dim=2))
    assert has_both.float().mean().item() > 0.99
    lower_first = torch.where(torch.arange(N, device=DEV).view(1, N) < twin.view(1, N), pos_self < pos_twin, pos_twin < pos_self)
    assert lower_first[has_both].all()


def test_forced_exact_path_matches_fast_path():
    from learning3d_b200 import _C
    from learning3d_b200.utils import knn
    g = torch.Generator(device=DEV).manual_seed(21)
    x = torch.randn(2, 64, 1500, device=DEV, generator=g)
    fast = knn(x, 20)
    _C.lib().l3d_debug_force_slow_path(1)
    try:
        slow = knn(x, 20)
    finally:
        _C.lib().l3d_debug_force_slow_path(0)
    assert torch.equal(fast, slow)


def test_graph_feature_on_features_and_add_one_to_k():
    from learning3d_b200.utils import get_graph_feature, knn
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(2, 64, 256, device=DEV, generator=g)
    idx = knn(x, 20)
    assert knn(x, 19, add_one_to_k=True).shape == (2, 256, 20)
    feat = get_graph_feature(x, k=20)
    assert feat.shape == (2, 128, 256, 20)
    nbr = torch.gather(x.unsqueeze(2).expand(-1, -1, 256, -1), 3, idx.unsqueeze(1).expand(-1, 64, -1, -1))
    assert torch.equal(feat[:, :64], nbr) and torch.equal(feat[:, 64:], x.unsqueeze(3).expand(-1, -1, -1, 20))


@pytest.mark.parametrize("split", [2, 4])
def test_feature_knn_with_split_target_range(split):
    from learning3d_b200 import _C
    from learning3d_b200.utils import knn
    g = torch.Generator(device=DEV).manual_seed(50 + split)
    x = torch.randn(1, 64, 1024, device=DEV, generator=g)
    _C.lib().l3d_debug_soft_correspondence_split(split)
    try:
        idx = knn(x, 20)
    finally:
        _C.lib().l3d_debug_soft_correspondence_split(0)
    _check(x, 20, idx)
    assert torch.equal(idx, knn(x, 20))      # automatic split (B = 1) gives the same keys, hence the same graph
