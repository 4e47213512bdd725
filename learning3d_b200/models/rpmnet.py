"""Drop-ins for the matching tail of learning3d/models/rpmnet.py:

    match_features(feat_src, feat_ref, metric)      rpmnet.py:130-154   tcgen05 Gram matrix (l3d_feature_square_distance)
    sinkhorn(log_alpha, n_iters, slack, eps)        rpmnet.py:157-218   potentials + sweeps (l3d_sinkhorn)
    compute_rigid_transform(a, b, weights)          rpmnet.py:221-254   weighted Kabsch (l3d_weighted_rigid_transform)
    match_tail(...)                                 rpmnet.py:280-287   affinity -> Sinkhorn -> exp -> weighted template,
                                                                        fused (l3d_rpm_match_tail)

Same names, arguments and return values as the reference functions; learning3d_b200.bind swaps them into an
imported reference package.  The kernels are forward-only: when autograd needs a gradient through one of these
calls (RPMNet training) the same torch expressions as the reference run on the GPU instead.
"""
import torch
import torch.nn as nn

from .. import _C
from ..utils._ops import feature_square_distance

_EPS = 1e-5        # rpmnet.py:11 ("to prevent division by zero")


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _ws(dev, nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


def _square_distance_torch(src, dst):
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist = dist + torch.sum(src ** 2, dim=-1)[:, :, None]
    return dist + torch.sum(dst ** 2, dim=-1)[:, None, :]


def match_features(feat_src, feat_ref, metric='l2'):
    """rpmnet.py:130-154 — (B,J,C), (B,K,C) -> (B,J,K)."""
    assert feat_src.shape[-1] == feat_ref.shape[-1]
    if metric == 'l2':
        if _needs_grad(feat_src, feat_ref):
            return _square_distance_torch(feat_src, feat_ref)
        return feature_square_distance(feat_src, feat_ref)
    if metric == 'angle':
        from ..utils import angle_difference
        a = feat_src / (torch.norm(feat_src, dim=-1, keepdim=True) + _EPS)
        b = feat_ref / (torch.norm(feat_ref, dim=-1, keepdim=True) + _EPS)
        return angle_difference(a, b)
    raise NotImplementedError


def _sinkhorn_torch(log_alpha, n_iters, slack, eps):
    """The reference's expressions (rpmnet.py:177-218), used when a gradient or the eps early exit is requested."""
    prev_alpha = None
    if slack:
        lap = torch.squeeze(nn.ZeroPad2d((0, 1, 0, 1))(log_alpha[:, None, :, :]), dim=1)
        for _ in range(n_iters):
            lap = torch.cat((lap[:, :-1, :] - torch.logsumexp(lap[:, :-1, :], dim=2, keepdim=True), lap[:, -1, None, :]), dim=1)
            lap = torch.cat((lap[:, :, :-1] - torch.logsumexp(lap[:, :, :-1], dim=1, keepdim=True), lap[:, :, -1, None]), dim=2)
            if eps > 0:
                if prev_alpha is not None:
                    if torch.max(torch.sum(torch.abs(torch.exp(lap[:, :-1, :-1]) - prev_alpha), dim=[1, 2])) < eps:
                        break
                prev_alpha = torch.exp(lap[:, :-1, :-1]).clone()
        return lap[:, :-1, :-1]
    for _ in range(n_iters):
        log_alpha = log_alpha - torch.logsumexp(log_alpha, dim=2, keepdim=True)
        log_alpha = log_alpha - torch.logsumexp(log_alpha, dim=1, keepdim=True)
        if eps > 0:
            if prev_alpha is not None:
                if torch.max(torch.sum(torch.abs(torch.exp(log_alpha) - prev_alpha), dim=[1, 2])) < eps:
                    break
            prev_alpha = torch.exp(log_alpha).clone()
    return log_alpha


def sinkhorn(log_alpha, n_iters: int = 5, slack: bool = True, eps: float = -1):
    """rpmnet.py:157-218 — log of the (near) doubly stochastic matrix, (B,J,K)."""
    if eps > 0 or _needs_grad(log_alpha):
        return _sinkhorn_torch(log_alpha, n_iters, slack, eps)
    la = _C.require_cuda(log_alpha, "log_alpha")
    B, J, K = la.shape
    out = torch.empty_like(la)
    lib = _C.lib()
    with _C.on_device(la.device):
        ws = _ws(la.device, lib.l3d_sinkhorn_ws_bytes(B, J, K))
        _C.check(lib.l3d_sinkhorn(_C.ptr(la), B, J, K, int(n_iters), 1 if slack else 0, _C.ptr(out), _C.ptr(ws),
                                  _C.stream()), "sinkhorn")
    return out


def match_tail(affinity, xyz_ref, n_iters: int = 5, slack: bool = True):
    """RPMNet.spam's tail (rpmnet.py:283-287) in one call: perm = exp(sinkhorn(affinity)), weights = sum_k perm,
    weighted_ref = perm @ xyz_ref / (weights + eps).  Returns (perm [B,J,K], weighted_ref [B,J,3], weights [B,J])."""
    aff, xyz = _C.require_cuda(affinity, "affinity"), _C.require_cuda(xyz_ref, "xyz_ref")
    B, J, K = aff.shape
    perm = torch.empty_like(aff)
    weighted = torch.empty((B, J, 3), dtype=torch.float32, device=aff.device)
    rowsum = torch.empty((B, J), dtype=torch.float32, device=aff.device)
    lib = _C.lib()
    with _C.on_device(aff.device):
        ws = _ws(aff.device, lib.l3d_sinkhorn_ws_bytes(B, J, K))
        _C.check(lib.l3d_rpm_match_tail(_C.ptr(aff), _C.ptr(xyz), B, J, K, int(n_iters), 1 if slack else 0, _EPS,
                                        _C.ptr(perm), _C.ptr(weighted), _C.ptr(rowsum), _C.ptr(ws), _C.stream()),
                 "rpm_match_tail")
    return perm, weighted, rowsum


def _rigid_torch(a, b, weights):
    wn = weights[..., None] / (torch.sum(weights[..., None], dim=1, keepdim=True) + _EPS)
    ca, cb = torch.sum(a * wn, dim=1), torch.sum(b * wn, dim=1)
    cov = (a - ca[:, None, :]).transpose(-2, -1) @ ((b - cb[:, None, :]) * wn)
    u, s, v = torch.svd(cov, some=False, compute_uv=True)
    pos = v @ u.transpose(-1, -2)
    vn = v.clone()
    vn[:, :, 2] *= -1
    neg = vn @ u.transpose(-1, -2)
    rot = torch.where(torch.det(pos)[:, None, None] > 0, pos, neg)
    t = -rot @ ca[:, :, None] + cb[:, :, None]
    return torch.cat((rot, t), dim=2)


def compute_rigid_transform(a, b, weights):
    """rpmnet.py:221-254 — a (B,M,3), b (B,M,3), weights (B,M) -> T (B,3,4) with T*a = b."""
    if _needs_grad(a, b, weights):
        return _rigid_torch(a, b, weights)
    a, b, w = _C.require_cuda(a, "a"), _C.require_cuda(b, "b"), _C.require_cuda(weights, "weights")
    B, M, _ = a.shape
    T = torch.empty((B, 3, 4), dtype=torch.float32, device=a.device)
    with _C.on_device(a.device):
        _C.check(_C.lib().l3d_weighted_rigid_transform(_C.ptr(a), _C.ptr(b), _C.ptr(w), B, M, _EPS, _C.ptr(T),
                                                       _C.stream()), "compute_rigid_transform")
    return T
