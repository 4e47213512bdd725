"""GPU: seeded randomized sweep over shapes / k / point distributions for every selection kernel, bit-exact
against the oracle.  Distributions include clustered clouds, lattices (many exact ties) and huge offsets
(catastrophic cancellation in the expansion form) — the cases where a kernel that is merely 'close' fails."""
import numpy as np
import pytest
import torch

from oracle import group as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def cloud(rng, B, N, kind):
    if kind == "uniform":
        x = rng.random((B, N, 3), dtype=np.float32)
    elif kind == "normal":
        x = rng.standard_normal((B, N, 3)).astype(np.float32)
    elif kind == "clusters":
        centres = rng.standard_normal((B, 8, 3)).astype(np.float32) * 3
        x = centres[:, rng.integers(0, 8, N)] + 0.01 * rng.standard_normal((B, N, 3)).astype(np.float32)
    elif kind == "lattice":
        x = rng.integers(0, 6, (B, N, 3)).astype(np.float32) * 0.25          # duplicates + exact ties
    elif kind == "offset":
        x = rng.random((B, N, 3), dtype=np.float32) + np.float32(100.0)       # |x|^2 ~ 3e4, d2 ~ 1e-2
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x.astype(np.float32))


KINDS = ["uniform", "normal", "clusters", "lattice", "offset"]


@pytest.mark.parametrize("seed", range(12))
def test_knn_family_random_configs(oracle_mod, seed):
    from learning3d_b200 import _C
    from learning3d_b200.utils import knn, knn_point
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 5)); N = int(rng.integers(1, 2300)); kind = KINDS[seed % len(KINDS)]
    k = int(rng.integers(1, min(N, 70) + 1))
    x = cloud(rng, B, N, kind)
    xb = np.ascontiguousarray(x.transpose(0, 2, 1))
    assert np.array_equal(knn(T(xb), k).cpu().numpy(), oracle_mod.knn_expansion(xb, k)), (B, N, k, kind)
    M = int(rng.integers(1, 400))
    q = cloud(rng, B, M, kind)
    val, idx = knn_point(k, T(x), T(q))
    ov, oi = oracle_mod.knn_point(k, x, q)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(val.cpu().numpy(), ov), (B, N, M, k, kind)
    lib = _C.lib()
    i64 = torch.empty((B, M, k), dtype=torch.int64, device=DEV)
    xd, qd = T(x), T(q)
    _C.check(lib.l3d_knn_sqdist(_C.ptr(xd), _C.ptr(qd), B, N, M, k, _C.ptr(i64), _C.stream()))
    assert np.array_equal(i64.cpu().numpy(), oracle_mod.knn_sqdist(x, q, k)), (B, N, M, k, kind)
    d2 = torch.empty((B, M, k), device=DEV); i32 = torch.empty((B, M, k), dtype=torch.int32, device=DEV)
    _C.check(lib.l3d_pn2_knn(B, M, N, k, _C.ptr(qd), _C.ptr(xd), _C.ptr(d2), _C.ptr(i32), _C.stream()))
    od2, oi32 = oracle_mod.pn2_knn(k, q, x)
    assert np.array_equal(i32.cpu().numpy(), oi32) and np.array_equal(d2.cpu().numpy(), od2), (B, N, M, k, kind)


@pytest.mark.parametrize("seed", range(10))
def test_chamfer_random_configs(oracle_mod, seed):
    from learning3d_b200.losses.cuda.chamfer_distance import ChamferDistanceFunction
    rng = np.random.default_rng(2000 + seed)
    B = int(rng.integers(1, 6)); n = int(rng.integers(1, 3000)); m = int(rng.integers(1, 3000))
    kind = KINDS[seed % len(KINDS)]
    a_np, b_np = cloud(rng, B, n, kind), cloud(rng, B, m, kind)
    a = T(a_np).requires_grad_(True); b = T(b_np).requires_grad_(True)
    d1, d2 = ChamferDistanceFunction.apply(a, b)
    od1, od2, oi1, oi2 = oracle_mod.chamfer_forward(a_np, b_np)
    assert np.array_equal(d1.detach().cpu().numpy(), od1) and np.array_equal(d2.detach().cpu().numpy(), od2), (B, n, m, kind)
    g1 = rng.standard_normal((B, n)).astype(np.float32); g2 = rng.standard_normal((B, m)).astype(np.float32)
    ga, gb = torch.autograd.grad([d1, d2], [a, b], [T(g1), T(g2)])
    oa, ob = oracle_mod.chamfer_backward(a_np, b_np, g1, g2, oi1, oi2)
    assert np.array_equal(ga.cpu().numpy(), oa) and np.array_equal(gb.cpu().numpy(), ob), (B, n, m, kind)


@pytest.mark.parametrize("seed", range(10))
def test_grouping_random_configs(seed):
    from learning3d_b200.utils import query_ball_point
    from learning3d_b200.utils.lib import pointnet2_utils as pu
    from learning3d_b200.utils.pointconv_util import farthest_point_sample
    rng = np.random.default_rng(3000 + seed)
    B = int(rng.integers(1, 5)); N = int(rng.integers(2, 3000)); S = int(rng.integers(1, min(N, 600) + 1))
    kind = KINDS[seed % len(KINDS)]
    xyz = cloud(rng, B, N, kind)
    new_xyz = np.ascontiguousarray(xyz[:, rng.permutation(N)[:S]])
    r = float(rng.choice([0.05, 0.2, 0.6])); ns = int(rng.integers(1, 70))
    assert np.array_equal(pu.ball_query(r, ns, T(xyz), T(new_xyz)).cpu().numpy(), og.pn2_ball_query(r, ns, xyz, new_xyz)), (B, N, S, r, ns, kind)
    idx, cnt = query_ball_point(r, ns, T(xyz), T(new_xyz), get_cnt=True)
    oi, oc = og.query_ball_point(r, ns, xyz, new_xyz, want_cnt=True)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(cnt.cpu().numpy(), oc), (B, N, S, r, ns, kind)
    M = int(rng.integers(1, min(N, 300) + 1))
    assert np.array_equal(pu.furthest_point_sample(T(xyz), M).cpu().numpy(), og.pn2_fps(xyz, M)[0]), (B, N, M, kind)
    assert np.array_equal(farthest_point_sample(T(xyz), M).cpu().numpy(), og.farthest_point_sample(xyz, M)), (B, N, M, kind)
