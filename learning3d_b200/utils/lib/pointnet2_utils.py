"""Drop-in for learning3d/utils/lib/pointnet2_utils.py:10-318.

Same public names (furthest_point_sample, gather_operation, knn, three_nn, three_interpolate,
grouping_operation, ball_query, QueryAndGroup, GroupAll) and tensor contracts (int32 indices,
[B,C,N] feature layout, gradients only through gather / group / interpolate).  The reference binds
`pointnet2_cuda` (utils/lib/src/pointnet2_api.cpp:10-25), which no longer builds (THC removed);
here every call goes to libl3d_b200.so (include/l3d_b200.h, "pointnet2_cuda replacements").
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from ... import _C


def _f32(t, name):
    return _C.require_cuda(t, name)


def _i32(t, name):
    if not t.is_cuda:
        raise RuntimeError("learning3d_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    return t.to(torch.int32).contiguous()


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B, N, 3) -> (B, npoint) int32, first index 0   (pointnet2_utils.py:12-29)."""
        assert xyz.is_contiguous()
        xyz = _f32(xyz, "xyz")
        B, N, _ = xyz.size()
        output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        with _C.on_device(xyz.device):
            _C.check(_C.lib().l3d_pn2_furthest_point_sampling(B, N, npoint, _C.ptr(xyz), _C.ptr(temp),
                                                              _C.ptr(output), _C.stream()),
                     "furthest_point_sample")
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint) -> (B, C, npoint)   (pointnet2_utils.py:40-60)."""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        features = _f32(features, "features")
        idx = _i32(idx, "idx")
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        with _C.on_device(features.device):
            _C.check(_C.lib().l3d_pn2_gather_points(B, C, N, npoint, _C.ptr(features), _C.ptr(idx),
                                                    _C.ptr(output), _C.stream()), "gather_operation")
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.contiguous()
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_pn2_gather_points_grad(B, C, N, npoint, _C.ptr(grad_out_data),
                                                         _C.ptr(idx), _C.ptr(grad_features),
                                                         _C.stream()), "gather_operation backward")
        return grad_features, None


gather_operation = GatherOperation.apply


class KNN(Function):
    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B, N, 3) queries, known (B, M, 3) -> (sqrt(d2) (B, N, k), idx int32)
        (pointnet2_utils.py:74-97)."""
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        unknown, known = _f32(unknown, "unknown"), _f32(known, "known")
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty((B, N, k), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, k), dtype=torch.int32, device=unknown.device)
        with _C.on_device(unknown.device):
            _C.check(_C.lib().l3d_pn2_knn(B, N, m, k, _C.ptr(unknown), _C.ptr(known), _C.ptr(dist2),
                                          _C.ptr(idx), _C.stream()), "knn")
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(pointnet2_utils.py:104-126)."""
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        unknown, known = _f32(unknown, "unknown"), _f32(known, "known")
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        with _C.on_device(unknown.device):
            _C.check(_C.lib().l3d_pn2_three_nn(B, N, m, _C.ptr(unknown), _C.ptr(known), _C.ptr(dist2),
                                               _C.ptr(idx), _C.stream()), "three_nn")
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B, C, M), idx (B, n, 3), weight (B, n, 3) -> (B, C, n)   (pointnet2_utils.py:137-161)."""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        features, weight = _f32(features, "features"), _f32(weight, "weight")
        idx = _i32(idx, "idx")
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        with _C.on_device(features.device):
            _C.check(_C.lib().l3d_pn2_three_interpolate(B, c, m, n, _C.ptr(features), _C.ptr(idx),
                                                        _C.ptr(weight), _C.ptr(output), _C.stream()),
                     "three_interpolate")
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.contiguous()
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_pn2_three_interpolate_grad(B, c, n, m, _C.ptr(grad_out_data),
                                                             _C.ptr(idx), _C.ptr(weight),
                                                             _C.ptr(grad_features), _C.stream()),
                     "three_interpolate backward")
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint, nsample) -> (B, C, npoint, nsample)
        (pointnet2_utils.py:186-205)."""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        features = _f32(features, "features")
        idx = _i32(idx, "idx")
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
        with _C.on_device(features.device):
            _C.check(_C.lib().l3d_pn2_group_points(B, C, N, nfeatures, nsample, _C.ptr(features),
                                                   _C.ptr(idx), _C.ptr(output), _C.stream()),
                     "grouping_operation")
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.contiguous()
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_pn2_group_points_grad(B, C, N, npoint, nsample, _C.ptr(grad_out_data),
                                                        _C.ptr(idx), _C.ptr(grad_features),
                                                        _C.stream()), "grouping_operation backward")
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B, N, 3), new_xyz (B, npoint, 3) -> idx (B, npoint, nsample) int32
        (pointnet2_utils.py:231-248)."""
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        xyz, new_xyz = _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        with _C.on_device(xyz.device):
            _C.check(_C.lib().l3d_pn2_ball_query(B, N, npoint, float(radius), nsample, _C.ptr(new_xyz),
                                                 _C.ptr(xyz), _C.ptr(idx), _C.stream()), "ball_query")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:259-295."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return new_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:298-318."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            new_features = grouped_xyz
        return new_features
