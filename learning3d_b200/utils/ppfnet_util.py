"""Drop-in for learning3d/utils/ppfnet_util.py:11-244 (PPFNet / RPMNet grouping)."""
import torch

from . import _ops


def angle_difference(src, dst):
    """utils/ppfnet_util.py:11-26 (dense [B,N,M] acos of a GEMM: not a neighbour-search op, plain torch)."""
    dist = torch.matmul(src, dst.permute(0, 2, 1))
    return torch.acos(dist)


def square_distance(src, dst):
    """utils/ppfnet_util.py:29-48."""
    return _ops.square_distance(src, dst)


def index_points(points, idx):
    """utils/ppfnet_util.py:51-68."""
    return _ops.index_points(points, idx)


def farthest_point_sample(xyz, npoint):
    """utils/ppfnet_util.py:71-93 (random start, drawn as the reference draws it)."""
    B, N, _ = xyz.shape
    farthest = torch.randint(0, N, (B,), dtype=torch.long)
    return _ops.farthest_point_sample(xyz, npoint, farthest)


def query_ball_point(radius, nsample, xyz, new_xyz, itself_indices=None):
    """utils/ppfnet_util.py:96-131."""
    return _ops.query_ball_point(radius, nsample, xyz, new_xyz, itself_indices)


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """utils/ppfnet_util.py:134-170."""
    B, N, C = xyz.shape
    if npoint > 0:
        S = npoint
        fps_idx = farthest_point_sample(xyz, npoint)
        new_xyz = index_points(xyz, fps_idx)
    else:
        S = xyz.shape[1]
        fps_idx = torch.arange(0, xyz.shape[1])[None, ...].repeat(xyz.shape[0], 1)
        new_xyz = xyz
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        grouped_points = index_points(points, idx)
        new_points = torch.cat([grouped_xyz_norm, grouped_points], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def angle(v1, v2):
    """utils/ppfnet_util.py:173-194."""
    cross_prod = torch.stack([v1[..., 1] * v2[..., 2] - v1[..., 2] * v2[..., 1],
                              v1[..., 2] * v2[..., 0] - v1[..., 0] * v2[..., 2],
                              v1[..., 0] * v2[..., 1] - v1[..., 1] * v2[..., 0]], dim=-1)
    cross_prod_norm = torch.norm(cross_prod, dim=-1)
    dot_prod = torch.sum(v1 * v2, dim=-1)
    return torch.atan2(cross_prod_norm, dot_prod)


def sample_and_group_multi(npoint, radius, nsample, xyz, normals, returnfps=False):
    """utils/ppfnet_util.py:197-244."""
    B, N, C = xyz.shape
    if npoint > 0:
        S = npoint
        fps_idx = farthest_point_sample(xyz, npoint)
        new_xyz = index_points(xyz, fps_idx)
        nr = index_points(normals, fps_idx)[:, :, None, :]
    else:
        S = xyz.shape[1]
        fps_idx = torch.arange(0, xyz.shape[1])[None, ...].repeat(xyz.shape[0], 1).to(xyz.device)
        new_xyz = xyz
        nr = normals[:, :, None, :]
    idx = query_ball_point(radius, nsample, xyz, new_xyz, fps_idx)
    grouped_xyz = index_points(xyz, idx)
    d = grouped_xyz - new_xyz.view(B, S, 1, C)
    ni = index_points(normals, idx)
    nr_d = angle(nr, d)
    ni_d = angle(ni, d)
    nr_ni = angle(nr, ni)
    d_norm = torch.norm(d, dim=-1)
    xyz_feat = d
    ppf_feat = torch.stack([nr_d, ni_d, nr_ni, d_norm], dim=-1)
    if returnfps:
        return {'xyz': new_xyz, 'dxyz': xyz_feat, 'ppf': ppf_feat}, grouped_xyz, fps_idx
    return {'xyz': new_xyz, 'dxyz': xyz_feat, 'ppf': ppf_feat}
