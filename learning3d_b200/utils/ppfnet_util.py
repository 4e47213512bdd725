"""PPFNet / RPMNet grouping of learning3d/utils/ppfnet_util.py (:11-244) on the CUDA hot path."""
import torch

from . import _ops

square_distance = _ops.square_distance            # ppfnet_util.py:29-48
index_points = _ops.index_points                  # :51-68


def angle_difference(src, dst):
    """:11-26 — acos of a dense [B,N,M] GEMM; not a neighbour-search op, plain torch as in the reference."""
    return torch.acos(torch.matmul(src, dst.permute(0, 2, 1)))


def farthest_point_sample(xyz, npoint):
    """:71-93 — random start drawn like the reference does (host generator), so seeded runs agree."""
    start = torch.randint(0, xyz.shape[1], (xyz.shape[0],), dtype=torch.long)
    return _ops.farthest_point_sample(xyz, npoint, start)


def query_ball_point(radius, nsample, xyz, new_xyz, itself_indices=None):
    """:96-131 — with itself_indices the centre is excluded from its own ball and used as padding."""
    return _ops.query_ball_point(radius, nsample, xyz, new_xyz, itself_indices)


def _centres(npoint, xyz):
    """FPS centres (npoint > 0) or every point (npoint <= 0): (centres, their indices)."""
    if npoint > 0:
        fps_idx = farthest_point_sample(xyz, npoint)
        return index_points(xyz, fps_idx), fps_idx
    every = torch.arange(0, xyz.shape[1], device=xyz.device)[None, ...].repeat(xyz.shape[0], 1)
    return xyz, every


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """:134-170."""
    centres, fps_idx = _centres(npoint, xyz)
    absolute, _, merged = _ops.group_around(xyz, centres, query_ball_point(radius, nsample, xyz, centres), points)
    return (centres, merged, absolute, fps_idx) if returnfps else (centres, merged)


def angle(v1, v2):
    """:173-194 — atan2(|v1 x v2|, v1 . v2), robust at zero vectors."""
    a, b = v1.unbind(-1), v2.unbind(-1)
    cross = torch.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dim=-1)
    return torch.atan2(torch.norm(cross, dim=-1), torch.sum(v1 * v2, dim=-1))


def sample_and_group_multi(npoint, radius, nsample, xyz, normals, returnfps=False):
    """:197-244 — xyz, centre-relative xyz and the 4-d point-pair features of every neighbourhood."""
    centres, fps_idx = _centres(npoint, xyz)
    centre_normals = (index_points(normals, fps_idx) if npoint > 0 else normals)[:, :, None, :]
    idx = query_ball_point(radius, nsample, xyz, centres, fps_idx)
    absolute, d, _ = _ops.group_around(xyz, centres, idx)
    neighbour_normals = index_points(normals, idx)
    ppf = torch.stack([angle(centre_normals, d), angle(neighbour_normals, d),
                       angle(centre_normals, neighbour_normals), torch.norm(d, dim=-1)], dim=-1)
    out = {'xyz': centres, 'dxyz': d, 'ppf': ppf}
    return (out, absolute, fps_idx) if returnfps else out
