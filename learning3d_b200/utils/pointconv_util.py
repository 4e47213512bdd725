"""Drop-in for the grouping functions of learning3d/utils/pointconv_util.py:18-209 (the module
classes built on them — DensityNet, PointConvDensitySetAbstraction ... — are callers and stay with
the reference)."""
import torch

from . import _ops


def square_distance(src, dst):
    """utils/pointconv_util.py:18-39."""
    return _ops.square_distance(src, dst)


def index_points(points, idx):
    """utils/pointconv_util.py:41-58."""
    return _ops.index_points(points, idx)


def farthest_point_sample(xyz, npoint):
    """utils/pointconv_util.py:60-83 (always starts at index 0)."""
    return _ops.farthest_point_sample(xyz, npoint, None)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """utils/pointconv_util.py:85-105."""
    return _ops.query_ball_point(radius, nsample, xyz, new_xyz)


def knn_point(nsample, xyz, new_xyz):
    """utils/pointconv_util.py:107-118 (topk sorted=False in the reference: same set, here in
    ascending-distance order)."""
    return _ops.knn_sqdist(nsample, xyz, new_xyz)


def sample_and_group(npoint, nsample, xyz, points, density_scale=None):
    """utils/pointconv_util.py:120-147."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = knn_point(nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        grouped_points = index_points(points, idx)
        new_points = torch.cat([grouped_xyz_norm, grouped_points], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if density_scale is None:
        return new_xyz, new_points, grouped_xyz_norm, idx
    grouped_density = index_points(density_scale, idx)
    return new_xyz, new_points, grouped_xyz_norm, idx, grouped_density


def sample_and_group_all(xyz, points, density_scale=None):
    """utils/pointconv_util.py:149-172."""
    B, N, C = xyz.shape
    new_xyz = xyz.mean(dim=1, keepdim=True)
    grouped_xyz = xyz.view(B, 1, N, C) - new_xyz.view(B, 1, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    if density_scale is None:
        return new_xyz, new_points, grouped_xyz
    grouped_density = density_scale.view(B, 1, N, 1)
    return new_xyz, new_points, grouped_xyz, grouped_density


def group(nsample, xyz, points):
    """utils/pointconv_util.py:174-197."""
    B, N, C = xyz.shape
    S = N
    new_xyz = xyz
    idx = knn_point(nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        grouped_points = index_points(points, idx)
        new_points = torch.cat([grouped_xyz_norm, grouped_points], dim=-1)
    else:
        new_points = grouped_xyz_norm
    return new_points, grouped_xyz_norm


def compute_density(xyz, bandwidth):
    """utils/pointconv_util.py:199-209."""
    return _ops.compute_density(xyz, bandwidth)
