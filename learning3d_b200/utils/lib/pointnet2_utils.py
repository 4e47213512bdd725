"""pointnet2 ops on libl3d_b200.so — the public surface of learning3d/utils/lib/pointnet2_utils.py
(:10-318): furthest_point_sample, gather_operation, knn, three_nn, three_interpolate,
grouping_operation, ball_query, QueryAndGroup, GroupAll; int32 indices, [B,C,N] features, gradients only
through gather / group / interpolate.  The reference binds the `pointnet2_cuda` extension
(utils/lib/src/pointnet2_api.cpp:10-25), which needs THC and no longer builds; here each op is one call
into the C ABI (include/l3d_b200.h, "pointnet2_cuda replacements") on the caller's current stream.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from ... import _C

_F32, _I32 = torch.float32, torch.int32


def _run(symbol, anchor, *args):
    """Launch `symbol` on anchor's device / current stream; raise on a non-zero return code."""
    with _C.on_device(anchor.device):
        _C.check(getattr(_C.lib(), symbol)(*args, _C.stream()), symbol)


def _feat(t, what):
    if not t.is_contiguous():
        raise AssertionError("%s must be contiguous" % what)     # the reference asserts contiguity
    return _C.require_cuda(t, what)


def _index(t, what):
    if not t.is_cuda:
        raise RuntimeError("learning3d_b200: %s must be a CUDA tensor (no CPU fallback)" % what)
    if not t.is_contiguous():
        raise AssertionError("%s must be contiguous" % what)
    return t if t.dtype == _I32 else t.to(_I32)


class FurthestPointSampling(Function):
    """(B, N, 3) cloud -> (B, npoint) int32 sample indices, index 0 first (pointnet2_utils.py:10-36)."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        xyz = _feat(xyz, "xyz")
        B, N = xyz.shape[0], xyz.shape[1]
        picked = torch.empty((B, npoint), dtype=_I32, device=xyz.device)
        running_min = torch.full((B, N), 1e10, dtype=_F32, device=xyz.device)
        _run("l3d_pn2_furthest_point_sampling", xyz, B, N, npoint, _C.ptr(xyz), _C.ptr(running_min), _C.ptr(picked))
        ctx.mark_non_differentiable(picked)
        return picked

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


class GatherOperation(Function):
    """features (B, C, N), idx (B, npoint) -> (B, C, npoint) (pointnet2_utils.py:38-70)."""

    @staticmethod
    def forward(ctx, features, idx):
        features, idx = _feat(features, "features"), _index(idx, "idx")
        (B, C, N), npoint = features.shape, idx.shape[1]
        out = torch.empty((B, C, npoint), dtype=_F32, device=features.device)
        _run("l3d_pn2_gather_points", features, B, C, N, npoint, _C.ptr(features), _C.ptr(idx), _C.ptr(out))
        ctx.for_backwards = (idx, C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        g = grad_out.contiguous()
        B, npoint = idx.shape
        acc = torch.zeros((B, C, N), dtype=_F32, device=g.device)
        _run("l3d_pn2_gather_points_grad", g, B, C, N, npoint, _C.ptr(g), _C.ptr(idx), _C.ptr(acc))
        return acc, None


def _nearest(symbol, k, unknown, known, with_k):
    unknown, known = _feat(unknown, "unknown"), _feat(known, "known")
    B, n, m = unknown.shape[0], unknown.shape[1], known.shape[1]
    d2 = torch.empty((B, n, k), dtype=_F32, device=unknown.device)
    idx = torch.empty((B, n, k), dtype=_I32, device=unknown.device)
    dims = (B, n, m, k) if with_k else (B, n, m)
    _run(symbol, unknown, *dims, _C.ptr(unknown), _C.ptr(known), _C.ptr(d2), _C.ptr(idx))
    return torch.sqrt(d2), idx


class KNN(Function):
    """k nearest `known` (B, M, 3) points of every `unknown` (B, N, 3) query: (L2 distance, int32 index),
    nearest first (pointnet2_utils.py:72-101)."""

    @staticmethod
    def forward(ctx, k, unknown, known):
        dist, idx = _nearest("l3d_pn2_knn", k, unknown, known, True)
        ctx.mark_non_differentiable(idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


class ThreeNN(Function):
    """The k = 3 case used by feature propagation (pointnet2_utils.py:104-130)."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist, idx = _nearest("l3d_pn2_three_nn", 3, unknown, known, False)
        ctx.mark_non_differentiable(idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


class ThreeInterpolate(Function):
    """features (B, C, M), idx / weight (B, n, 3) -> weighted sum of three gathered columns (B, C, n)
    (pointnet2_utils.py:135-185)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        features, weight, idx = _feat(features, "features"), _feat(weight, "weight"), _index(idx, "idx")
        (B, C, m), n = features.shape, idx.shape[1]
        out = torch.empty((B, C, n), dtype=_F32, device=features.device)
        _run("l3d_pn2_three_interpolate", features, B, C, m, n, _C.ptr(features), _C.ptr(idx), _C.ptr(weight), _C.ptr(out))
        ctx.three_interpolate_for_backward = (idx, weight, m)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        g = grad_out.contiguous()
        B, C, n = g.shape
        acc = torch.zeros((B, C, m), dtype=_F32, device=g.device)
        _run("l3d_pn2_three_interpolate_grad", g, B, C, n, m, _C.ptr(g), _C.ptr(idx), _C.ptr(weight), _C.ptr(acc))
        return acc, None, None


class GroupingOperation(Function):
    """features (B, C, N), idx (B, npoint, nsample) -> (B, C, npoint, nsample) (pointnet2_utils.py:188-225)."""

    @staticmethod
    def forward(ctx, features, idx):
        features, idx = _feat(features, "features"), _index(idx, "idx")
        (B, C, N), (_, npoint, nsample) = features.shape, idx.shape
        out = torch.empty((B, C, npoint, nsample), dtype=_F32, device=features.device)
        _run("l3d_pn2_group_points", features, B, C, N, npoint, nsample, _C.ptr(features), _C.ptr(idx), _C.ptr(out))
        ctx.for_backwards = (idx, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        g = grad_out.contiguous()
        B, C, npoint, nsample = g.shape
        acc = torch.zeros((B, C, N), dtype=_F32, device=g.device)
        _run("l3d_pn2_group_points_grad", g, B, C, N, npoint, nsample, _C.ptr(g), _C.ptr(idx), _C.ptr(acc))
        return acc, None


class BallQuery(Function):
    """First `nsample` indices (ascending) of xyz (B, N, 3) within `radius` of each centre new_xyz
    (B, npoint, 3), padded with the first hit (pointnet2_utils.py:228-254)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        new_xyz, xyz = _feat(new_xyz, "new_xyz"), _feat(xyz, "xyz")
        B, N, npoint = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
        idx = torch.empty((B, npoint, nsample), dtype=_I32, device=xyz.device)
        _run("l3d_pn2_ball_query", xyz, B, N, npoint, float(radius), nsample, _C.ptr(new_xyz), _C.ptr(xyz), _C.ptr(idx))
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


furthest_point_sample = FurthestPointSampling.apply
gather_operation = GatherOperation.apply
knn = KNN.apply
three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply
grouping_operation = GroupingOperation.apply
ball_query = BallQuery.apply


def _stack_xyz_and_features(grouped_xyz, grouped_features, use_xyz):
    if grouped_features is None:
        if not use_xyz:
            raise AssertionError("Cannot have not features and not use xyz as a feature!")
        return grouped_xyz
    return torch.cat([grouped_xyz, grouped_features], dim=1) if use_xyz else grouped_features


class QueryAndGroup(nn.Module):
    """Ball query + grouping with centre-relative coordinates: (B, 3 + C, npoint, nsample)
    (pointnet2_utils.py:257-295)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        rel = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        grouped = None if features is None else grouping_operation(features, idx)
        return _stack_xyz_and_features(rel, grouped, self.use_xyz)


class GroupAll(nn.Module):
    """One group holding the whole cloud: (B, 3 + C, 1, N) (pointnet2_utils.py:298-318)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        whole = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return whole
        return _stack_xyz_and_features(whole, features.unsqueeze(2), self.use_xyz)
