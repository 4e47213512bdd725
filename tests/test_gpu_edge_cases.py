"""GPU: edge cases — empty batches, single points, maximum supported sizes, unsupported shapes, ragged
clouds, non-default streams."""
import numpy as np
import pytest
import torch

from oracle import group as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_empty_batches_return_empty_without_error():
    from learning3d_b200.utils import knn, get_graph_feature, query_ball_point, farthest_point_sample, index_points
    from learning3d_b200.utils.lib import pointnet2_utils as pu
    from learning3d_b200.losses.cuda.chamfer_distance import ChamferDistanceFunction
    assert knn(torch.empty(0, 3, 64, device=DEV), 4).shape == (0, 64, 4)
    assert get_graph_feature(torch.empty(0, 3, 64, device=DEV), k=4).shape == (0, 6, 64, 4)
    e = torch.empty(0, 16, 3, device=DEV)
    assert query_ball_point(0.1, 4, e, e).shape == (0, 16, 4)
    assert farthest_point_sample(e, 5, start_with_first_point=True).shape == (0, 5)
    assert pu.ball_query(0.1, 4, e, e).shape == (0, 16, 4)
    assert pu.furthest_point_sample(e, 5).shape == (0, 5)
    d1, d2 = ChamferDistanceFunction.apply(e, e)
    assert d1.shape == (0, 16) and d2.shape == (0, 16)
    assert index_points(e, torch.empty(0, 7, dtype=torch.int64, device=DEV)).shape == (0, 7, 3)
    torch.cuda.synchronize()


def test_single_point_and_k_equals_one(oracle_mod):
    from learning3d_b200.utils import knn
    from learning3d_b200.losses import chamfer_distance
    x = torch.rand(2, 3, 1, device=DEV)
    assert torch.equal(knn(x, 1), torch.zeros(2, 1, 1, dtype=torch.int64, device=DEV))
    a = torch.rand(3, 1, 3, device=DEV); b = torch.rand(3, 1, 3, device=DEV)
    want = oracle_mod.chamfer_loss(a.cpu().numpy(), b.cpu().numpy())
    assert abs(chamfer_distance(a, b).item() - want) < 1e-6
    rng = np.random.default_rng(0)
    y = rng.random((2, 3, 37), dtype=np.float32)
    assert np.array_equal(knn(T(y), 1).cpu().numpy(), oracle_mod.knn_expansion(y, 1))


def test_maximum_cloud_size_and_beyond(oracle_mod):
    from learning3d_b200.utils import knn
    rng = np.random.default_rng(1)
    x = rng.random((1, 3, 8192), dtype=np.float32)                   # L3D_KNN_MAX_N: 8 candidate tiles
    assert np.array_equal(knn(T(x), 20).cpu().numpy(), oracle_mod.knn_expansion(x, 20, mt=True))
    y = rng.random((1, 3, 700), dtype=np.float32)                    # k > 128: exact k-round scan path
    assert np.array_equal(knn(T(y), 200).cpu().numpy(), oracle_mod.knn_expansion(y, 200, mt=True))


def test_clouds_beyond_the_resident_kernel_stream(oracle_mod):
    """N > L3D_KNN_MAX_N (the reference's matmul + topk has no size limit): the streamed selection, every key mode,
    bit-exact against the oracle including ties, k up to 200, the fused graph feature and the host entry point."""
    from learning3d_b200 import _C
    from learning3d_b200.utils import knn, get_graph_feature, knn_point
    rng = np.random.default_rng(11)
    for B, N, k in ((1, 8196, 4), (2, 10000, 20), (1, 8300, 200), (1, 20000, 33)):
        x = rng.random((B, 3, N), dtype=np.float32)
        assert np.array_equal(knn(T(x), k).cpu().numpy(), oracle_mod.knn_expansion(x, k, mt=True)), (B, N, k)
    lat = rng.integers(0, 12, size=(1, 3, 9000)).astype(np.float32)          # a 12^3 lattice: every row is full of ties
    assert np.array_equal(knn(T(lat), 20).cpu().numpy(), oracle_mod.knn_expansion(lat, 20, mt=True))
    x = rng.random((1, 3, 8200), dtype=np.float32)
    idx = knn(T(x), 6)
    feat = get_graph_feature(T(x), k=6)                                        # knn + gather on this path
    xt = torch.from_numpy(x).to(DEV)
    nb = torch.gather(xt.unsqueeze(2).expand(1, 3, 8200, 8200), 3, idx.unsqueeze(1).expand(1, 3, 8200, 6))
    want = torch.cat([nb, xt.unsqueeze(3).expand(1, 3, 8200, 6)], 1)
    assert torch.equal(feat, want)
    data = rng.random((2, 9001, 3), dtype=np.float32); q = rng.random((2, 77, 3), dtype=np.float32)
    val, idx = knn_point(5, T(data), T(q))
    ov, oi = oracle_mod.knn_point(5, data, q)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(val.cpu().numpy(), ov)
    dd, qd = T(data), T(q)
    for kk in (3, 64):
        d2 = torch.empty((2, 77, kk), dtype=torch.float32, device=DEV)
        i32 = torch.empty((2, 77, kk), dtype=torch.int32, device=DEV)
        _C.check(_C.lib().l3d_pn2_knn(2, 77, 9001, kk, _C.ptr(qd), _C.ptr(dd), _C.ptr(d2), _C.ptr(i32), _C.stream()))
        od2, oi = oracle_mod.pn2_knn(kk, q, data)
        assert np.array_equal(i32.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od2)
        i64 = torch.empty((2, 77, kk), dtype=torch.int64, device=DEV)
        _C.check(_C.lib().l3d_knn_sqdist(_C.ptr(dd), _C.ptr(qd), 2, 9001, 77, kk, _C.ptr(i64), _C.stream()))
        assert np.array_equal(i64.cpu().numpy(), oracle_mod.knn_sqdist(data, q, kk))
    xh = torch.rand(2, 3, 9000)
    ih = torch.empty(2, 9000, 8, dtype=torch.int64)
    _C.check(_C.lib().l3d_knn_expansion_host(_C._P(xh.data_ptr()), 2, 9000, 8, _C._P(ih.data_ptr())))
    assert np.array_equal(ih.numpy(), oracle_mod.knn_expansion(xh.numpy(), 8, mt=True))


def test_fps_beyond_the_register_resident_kernel():
    """N > 8192: running minima in `temp` (global), same rounds / tie rules as the resident kernel."""
    from learning3d_b200.utils.lib import pointnet2_utils as pu
    from learning3d_b200.utils.pointconv_util import farthest_point_sample
    rng = np.random.default_rng(12)
    x = rng.random((2, 9000, 3), dtype=np.float32)
    assert np.array_equal(pu.furthest_point_sample(T(x), 96).cpu().numpy(), og.pn2_fps(x, 96)[0])
    assert np.array_equal(farthest_point_sample(T(x), 50).cpu().numpy(), og.farthest_point_sample(x, 50))
    lat = rng.integers(0, 6, size=(1, 8500, 3)).astype(np.float32)             # heavy ties
    assert np.array_equal(pu.furthest_point_sample(T(lat), 40).cpu().numpy(), og.pn2_fps(lat, 40)[0])
    assert np.array_equal(farthest_point_sample(T(lat), 40).cpu().numpy(), og.farthest_point_sample(lat, 40))


def test_k_ranges_take_every_kernel_variant(oracle_mod):
    """k <= 24 (two rows per warp, 64-bit composite network), 25..48 (KS=2), 49..128 (KS=4), whole-row sort."""
    from learning3d_b200.utils import knn
    rng = np.random.default_rng(2)
    x = rng.random((2, 3, 1000), dtype=np.float32)
    for k in (1, 7, 24, 25, 48, 49, 100, 128):
        assert np.array_equal(knn(T(x), k).cpu().numpy(), oracle_mod.knn_expansion(x, k)), k
    xs = rng.random((3, 3, 200), dtype=np.float32)
    for k in (25, 64, 128, 200):                                        # k*8 >= N: whole-row sort / slow path
        if k <= 128:
            assert np.array_equal(knn(T(xs), k).cpu().numpy(), oracle_mod.knn_expansion(xs, k)), k


def test_odd_row_counts_and_batch_boundaries(oracle_mod):
    """Row pairs straddling CTA / batch-item boundaries (M odd, B*M not a multiple of the grid)."""
    from learning3d_b200.utils import knn, knn_point
    rng = np.random.default_rng(3)
    for B, N in ((5, 333), (7, 1025), (37, 129), (1, 3)):
        x = rng.random((B, 3, N), dtype=np.float32)
        k = min(9, N)
        assert np.array_equal(knn(T(x), k).cpu().numpy(), oracle_mod.knn_expansion(x, k)), (B, N)
    data = rng.random((3, 257, 3), dtype=np.float32); q = rng.random((3, 11, 3), dtype=np.float32)
    val, idx = knn_point(5, T(data), T(q))
    ov, oi = oracle_mod.knn_point(5, data, q)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(val.cpu().numpy(), ov)


def test_fps_extremes():
    from learning3d_b200.utils.lib import pointnet2_utils as pu
    rng = np.random.default_rng(4)
    x = rng.random((2, 100, 3), dtype=np.float32)
    assert torch.equal(pu.furthest_point_sample(T(x), 1), torch.zeros(2, 1, dtype=torch.int32, device=DEV))
    full = pu.furthest_point_sample(T(x), 100).cpu().numpy()            # npoint == N: a permutation
    assert np.array_equal(full, og.pn2_fps(x, 100)[0])
    assert all(sorted(r) == list(range(100)) for r in full)
    one = rng.random((1, 1, 3), dtype=np.float32)
    assert pu.furthest_point_sample(T(one), 1).item() == 0


def test_runs_on_a_side_stream(oracle_mod):
    from learning3d_b200.utils import knn
    from learning3d_b200.losses import ChamferDistanceLoss
    rng = np.random.default_rng(5)
    x = rng.random((4, 3, 512), dtype=np.float32)
    s = torch.cuda.Stream()
    xd = T(x)
    a = torch.rand(4, 256, 3, device=DEV, requires_grad=True); b = torch.rand(4, 256, 3, device=DEV)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        idx = knn(xd, 8)
        loss = ChamferDistanceLoss()(a, b)
        loss.backward()
    s.synchronize()
    assert np.array_equal(idx.cpu().numpy(), oracle_mod.knn_expansion(x, 8))
    assert abs(loss.item() - oracle_mod.chamfer_loss(a.detach().cpu().numpy(), b.cpu().numpy())) < 1e-6


def test_non_contiguous_inputs_are_accepted_like_the_reference():
    from learning3d_b200.utils import knn
    x = torch.rand(2, 300, 3, device=DEV)
    xt = x.permute(0, 2, 1)                      # DGCNN passes exactly this view (models/dgcnn.py:27)
    assert torch.equal(knn(xt, 6), knn(xt.contiguous(), 6))
