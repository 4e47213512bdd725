"""Shared torch-facing wrappers over the C ABI for the pure-torch grouping helpers of the reference
(utils/model_common_utils.py, utils/pointconv_util.py, utils/ppfnet_util.py).  The three reference
modules carry near-identical copies of these functions; their drop-ins all route here."""
import numpy as np
import torch

from .. import _C


def _no_grad(name, *tensors):
    """These reference functions are differentiable torch expressions; the kernels behind the drop-ins are
    forward-only, so a gradient request is an error instead of a silently dropped gradient."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError("learning3d_b200.%s is forward-only: an input requires grad. Detach it (the models in "
                           "learning3d only use this op to build indices / weights) or call it under "
                           "torch.no_grad()." % name)


def _xyz(t, name):
    t = _C.require_cuda(t, name)
    if t.dim() != 3 or t.size(2) != 3:
        raise NotImplementedError("learning3d_b200: %s must be [B, N, 3] (got %s); only xyz clouds are "
                                  "on the built hot path" % (name, tuple(t.shape)))
    return t


def feature_square_distance(src, dst, beta=None, alpha=None):
    """square_distance(src [B,N,C], dst [B,M,C]) -> [B,N,M] on the tensor cores (any C); with beta/alpha [B] the
    epilogue returns RPMNet's affinity -beta * (dist - alpha) instead."""
    src, dst = _C.require_cuda(src, "src"), _C.require_cuda(dst, "dst")
    B, N, C = src.shape
    M = dst.shape[1]
    if dst.shape[0] != B or dst.shape[2] != C:
        raise ValueError("square_distance: inconsistent shapes %s %s" % (tuple(src.shape), tuple(dst.shape)))
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    lib = _C.lib()
    if beta is not None:
        beta = _C.require_cuda(beta.reshape(B), "beta")
        alpha = _C.require_cuda(alpha.reshape(B) if isinstance(alpha, torch.Tensor)
                                else torch.full((B,), float(alpha), device=src.device), "alpha")
    with _C.on_device(src.device):
        ws = torch.empty(max(int(lib.l3d_feature_square_distance_ws_bytes(B, N, M)), 16), dtype=torch.uint8,
                         device=src.device)
        _C.check(lib.l3d_feature_square_distance(_C.ptr(src), _C.ptr(dst), B, N, M, C, _C.ptr(beta), _C.ptr(alpha),
                                                 _C.ptr(out), _C.ptr(ws), _C.stream()), "square_distance")
    return out



def square_distance(src, dst):
    """model_common_utils.py:19-38 — [B,N,C], [B,M,C] -> [B,N,M] expansion-form squared distance.  C = 3 (clouds):
    bit-exact SIMT kernel; any other C (RPMNet's 96-d features): Gram matrix on the tensor cores (toleranced)."""
    _no_grad("square_distance", src, dst)
    if src.dim() == 3 and src.size(2) != 3:
        return feature_square_distance(src, dst)
    src, dst = _xyz(src, "src"), _xyz(dst, "dst")
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    with _C.on_device(src.device):
        _C.check(_C.lib().l3d_square_distance(_C.ptr(src), _C.ptr(dst), B, N, M, _C.ptr(out),
                                              _C.stream()), "square_distance")
    return out


class _IndexPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        B, N, C = points.shape
        R = idx.numel() // B if B > 0 else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=points.device)
        with _C.on_device(points.device):
            _C.check(_C.lib().l3d_index_points(_C.ptr(points), _C.ptr(idx), B, N, R, C, _C.ptr(out),
                                               _C.stream()), "index_points")
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, C, R)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, N, C, R = ctx.dims
        grad_out = grad_out.contiguous()
        gp = torch.zeros((B, N, C), dtype=torch.float32, device=grad_out.device)
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_index_points_grad(_C.ptr(grad_out), _C.ptr(idx), B, N, R, C,
                                                    _C.ptr(gp), _C.stream()), "index_points backward")
        return gp, None


def index_points(points, idx):
    """model_common_utils.py:40-56 — points [B,N,C], idx [B,S] or [B,S,K] int64 -> [B,S(,K),C]."""
    points = _C.require_cuda(points, "points")
    if not idx.is_cuda:
        raise RuntimeError("learning3d_b200: idx must be a CUDA tensor (no CPU fallback)")
    idx = idx.to(torch.int64).contiguous()
    return _IndexPoints.apply(points, idx)


def farthest_point_sample(xyz, npoint, start=None):
    """FPS with the torch reference semantics; `start` None -> index 0, else an int64 [B] tensor."""
    xyz = _xyz(xyz, "xyz")
    B, N, _ = xyz.shape
    cent = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    if start is not None:
        start = start.to(device=xyz.device, dtype=torch.int64).contiguous()
    with _C.on_device(xyz.device):
        _C.check(_C.lib().l3d_farthest_point_sample(_C.ptr(xyz), B, N, npoint, _C.ptr(start),
                                                    _C.ptr(cent), _C.stream()), "farthest_point_sample")
    return cent


def query_ball_point(radius, nsample, xyz, new_xyz, itself_indices=None, get_cnt=False):
    xyz, new_xyz = _xyz(xyz, "xyz"), _xyz(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty((B, S, nsample), dtype=torch.int64, device=xyz.device)
    cnt = torch.empty((B, S), dtype=torch.int64, device=xyz.device) if get_cnt else None
    if itself_indices is not None:
        itself_indices = itself_indices.to(device=xyz.device, dtype=torch.int64).contiguous()
    # `sqrdists > radius ** 2`: python computes radius**2 in double, the comparison casts it to fp32
    r2 = float(np.float32(radius ** 2))
    with _C.on_device(xyz.device):
        _C.check(_C.lib().l3d_query_ball_point(_C.ptr(xyz), _C.ptr(new_xyz), B, N, S, r2, nsample,
                                               _C.ptr(itself_indices), _C.ptr(idx), _C.ptr(cnt),
                                               _C.stream()), "query_ball_point")
    return (idx, cnt) if get_cnt else idx


def knn_sqdist(nsample, xyz, new_xyz):
    """pointconv_util.knn_point (pointconv_util.py:107-118)."""
    xyz, new_xyz = _xyz(xyz, "xyz"), _xyz(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if nsample > N:
        raise RuntimeError("selected index k out of range")
    idx = torch.empty((B, S, nsample), dtype=torch.int64, device=xyz.device)
    with _C.on_device(xyz.device):
        _C.check(_C.lib().l3d_knn_sqdist(_C.ptr(xyz), _C.ptr(new_xyz), B, N, S, nsample, _C.ptr(idx),
                                         _C.stream()), "knn_point")
    return idx


def group_around(xyz, centres, idx, feats=None):
    """Gather the neighbourhoods `idx` [B,S,K] of `xyz` [B,N,C], express them relative to their centres
    [B,S,C] and optionally append gathered per-point features.  Returns (absolute, relative, merged) —
    the tail shared by every sample_and_group / group composition of the reference."""
    absolute = index_points(xyz, idx)
    relative = absolute - centres.unsqueeze(2)
    merged = relative if feats is None else torch.cat([relative, index_points(feats, idx)], dim=-1)
    return absolute, relative, merged


def compute_density(xyz, bandwidth):
    """pointconv_util.py:199-209 — fused row reduction, the N x N matrix is never materialised."""
    _no_grad("compute_density", xyz)
    xyz = _xyz(xyz, "xyz")
    B, N, _ = xyz.shape
    out = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
    two_bw2 = float(np.float32(2.0 * bandwidth * bandwidth))
    norm = float(np.float32(2.5 * bandwidth))
    with _C.on_device(xyz.device):
        _C.check(_C.lib().l3d_compute_density(_C.ptr(xyz), B, N, two_bw2, norm, _C.ptr(out),
                                              _C.stream()), "compute_density")
    return out
