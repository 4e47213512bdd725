"""Grouping functions of learning3d/utils/pointconv_util.py (:18-209) on the CUDA hot path.  The nn.Module
classes of that file (DensityNet, WeightNet, PointConvDensitySetAbstraction ...) are callers of these
functions and stay with the reference."""
import torch

from . import _ops

# one-to-one replacements (same signatures)
square_distance = _ops.square_distance            # pointconv_util.py:18-39
index_points = _ops.index_points                  # :41-58
compute_density = _ops.compute_density            # :199-209, fused: no N x N matrix


def farthest_point_sample(xyz, npoint):
    """:60-83 — this variant always starts from point 0."""
    return _ops.farthest_point_sample(xyz, npoint, None)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """:85-105."""
    return _ops.query_ball_point(radius, nsample, xyz, new_xyz)


def knn_point(nsample, xyz, new_xyz):
    """:107-118 — the reference asks topk(sorted=False); the same neighbour SET comes back here in
    ascending-distance order."""
    return _ops.knn_sqdist(nsample, xyz, new_xyz)


def sample_and_group(npoint, nsample, xyz, points, density_scale=None):
    """:120-147 — FPS centres, kNN neighbourhoods, centre-relative coordinates (+ features, + density)."""
    centres = index_points(xyz, farthest_point_sample(xyz, npoint))
    idx = knn_point(nsample, xyz, centres)
    _, rel, merged = _ops.group_around(xyz, centres, idx, points)
    if density_scale is None:
        return centres, merged, rel, idx
    return centres, merged, rel, idx, index_points(density_scale, idx)


def sample_and_group_all(xyz, points, density_scale=None):
    """:149-172 — a single group around the centroid."""
    B, N, C = xyz.shape
    centre = xyz.mean(dim=1, keepdim=True)
    rel = (xyz - centre).view(B, 1, N, C)
    merged = rel if points is None else torch.cat([rel, points.view(B, 1, N, -1)], dim=-1)
    if density_scale is None:
        return centre, merged, rel
    return centre, merged, rel, density_scale.view(B, 1, N, 1)


def group(nsample, xyz, points):
    """:174-197 — every point is a centre."""
    _, rel, merged = _ops.group_around(xyz, xyz, knn_point(nsample, xyz, xyz), points)
    return merged, rel
