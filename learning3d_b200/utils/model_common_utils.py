"""Drop-in for learning3d/utils/model_common_utils.py — same names, arguments and error
behaviour; every function launches hand-written sm_100a kernels through the C ABI.

Reference: utils/model_common_utils.py:3-155.
"""
import torch

from .. import _C
from . import _ops


def knn(x, k, add_one_to_k=False):
    """utils/model_common_utils.py:3-9.  x [B,C,N] fp32 -> idx [B,N,k] int64, nearest first.

    The B x N x N matrix of the reference (three 134 MB passes at B=32, N=1024) is never
    materialised: distance evaluation and top-k selection are one fused kernel.
    """
    if add_one_to_k:
        k = k + 1
    x = _C.require_cuda(x, "x")
    if x.dim() != 3:
        raise ValueError("knn expects x of shape [B, C, N], got %s" % (tuple(x.shape),))
    B, C, N = x.shape
    if k > N:
        # same failure class as torch.topk in the reference
        raise RuntimeError("selected index k out of range")
    idx = torch.empty((B, N, k), dtype=torch.int64, device=x.device)
    if C != 3:
        # feature-space graph (PRNet's dynamic DGCNN, models/prnet.py:78-90): Gram matrix on the tensor
        # cores (tcgen05 3xTF32) fused with the expansion keys, then the same warp selection as C == 3
        lib = _C.lib()
        ws = torch.empty(lib.l3d_knn_features_ws_bytes(B, C, N), dtype=torch.uint8, device=x.device)
        with _C.on_device(x.device):
            _C.check(lib.l3d_knn_features(_C.ptr(x), B, C, N, k, _C.ptr(idx), _C.ptr(ws), _C.stream()), "knn")
        return idx
    with _C.on_device(x.device):
        _C.check(_C.lib().l3d_knn_expansion(_C.ptr(x), B, N, k, _C.ptr(idx), _C.ptr(None),
                                            _C.stream()), "knn")
    return idx


class _GraphFeature(torch.autograd.Function):
    """cat(x[nbr], x[centre]) gather of get_graph_feature (:149-154) and its scatter backward."""

    @staticmethod
    def forward(ctx, x, idx):
        B, C, N = x.shape
        k = idx.shape[-1]
        out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=x.device)
        with _C.on_device(x.device):
            _C.check(_C.lib().l3d_graph_feature(_C.ptr(x), _C.ptr(idx), B, C, N, k, _C.ptr(out),
                                                _C.stream()), "get_graph_feature")
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, N, k)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, N, k = ctx.dims
        grad_out = grad_out.contiguous()
        gx = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_graph_feature_grad(_C.ptr(grad_out), _C.ptr(idx), B, C, N, k,
                                                     _C.ptr(gx), _C.stream()),
                     "get_graph_feature backward")
        return gx, None


class _KnnGraphFeature(torch.autograd.Function):
    """knn() + the gather of get_graph_feature in one kernel launch (xyz graph, C == 3): the neighbour
    coordinates are read from the shared-memory copy of the cloud the selection just used.  Backward is the
    same scatter as _GraphFeature (indices are saved; they carry no gradient, as in the reference)."""

    @staticmethod
    def forward(ctx, x, k):
        B, C, N = x.shape
        idx = torch.empty((B, N, k), dtype=torch.int64, device=x.device)
        out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=x.device)
        with _C.on_device(x.device):
            _C.check(_C.lib().l3d_knn_graph_feature(_C.ptr(x), B, N, k, _C.ptr(idx), _C.ptr(out), _C.stream()),
                     "get_graph_feature")
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, N, k)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, N, k = ctx.dims
        grad_out = grad_out.contiguous()
        gx = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        with _C.on_device(grad_out.device):
            _C.check(_C.lib().l3d_graph_feature_grad(_C.ptr(grad_out), _C.ptr(idx), B, C, N, k,
                                                     _C.ptr(gx), _C.stream()),
                     "get_graph_feature backward")
        return gx, None


def get_graph_feature(x, k=20, device=None):
    """utils/model_common_utils.py:132-155.  x [B,C,N(,1)] -> [B,2C,N,k] = cat(neighbour, centre).

    `device` is accepted for signature compatibility; the result lives on x's device (the
    reference's default picks 'cuda' whenever CUDA is available, :138-139).
    """
    x = x.view(*x.size()[:3])
    x = _C.require_cuda(x, "x")
    if x.shape[1] == 3 and k <= x.shape[2]:
        return _KnnGraphFeature.apply(x, k)          # one launch: selection + gather
    idx = knn(x, k=k)
    return _GraphFeature.apply(x, idx)


def knn_point(k, pos1, pos2):
    """utils/model_common_utils.py:84-100.  pos1 [B,N,C] data, pos2 [B,M,C] queries ->
    (sqrt(d2) [B,M,k], idx [B,M,k] int64), nearest first."""
    _ops._no_grad("knn_point", pos1, pos2)      # the reference's sqrt(d2) values are differentiable; ours are not
    pos1 = _C.require_cuda(pos1, "pos1")
    pos2 = _C.require_cuda(pos2, "pos2")
    B, N, C = pos1.shape
    M = pos2.shape[1]
    if k > N:
        raise RuntimeError("selected index k out of range")
    if C != 3 or pos2.shape[2] != 3:
        # C-dimensional features: Gram matrix on the tensor cores, then the row-wise selection kernel (toleranced
        # like every GEMM-based path; the reference evaluates the direct differences in fp32)
        d2 = _ops.feature_square_distance(pos2, pos1)                 # [B, M, N]
        keys = torch.neg(d2)
        idx = torch.empty((B, M, k), dtype=torch.int64, device=pos1.device)
        with _C.on_device(pos1.device):
            _C.check(_C.lib().l3d_topk_rows(_C.ptr(keys), B * M, N, k, _C.ptr(idx), _C.stream()), "knn_point")
        return torch.sqrt(torch.gather(d2, 2, idx).clamp_min_(0)), idx
    val = torch.empty((B, M, k), dtype=torch.float32, device=pos1.device)
    idx = torch.empty((B, M, k), dtype=torch.int64, device=pos1.device)
    with _C.on_device(pos1.device):
        _C.check(_C.lib().l3d_knn_point(_C.ptr(pos1), _C.ptr(pos2), B, N, M, k, _C.ptr(val),
                                        _C.ptr(idx), _C.stream()), "knn_point")
    return val, idx


# ---- the remaining helpers of utils/model_common_utils.py --------------------------------------
import numpy as np  # noqa: E402


def pc_normalize(pc):
    """utils/model_common_utils.py:11-17 (numpy, host side; unchanged semantics)."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    return pc / m


def square_distance(src, dst):
    """utils/model_common_utils.py:19-38."""
    return _ops.square_distance(src, dst)


def index_points(points, idx):
    """utils/model_common_utils.py:40-56."""
    return _ops.index_points(points, idx)


def farthest_point_sample(xyz, npoint, start_with_first_point=False):
    """utils/model_common_utils.py:58-82.  The random start is drawn exactly as the reference draws
    it (CPU generator, torch.randint(0, N, (B,))) so seeded runs pick the same points."""
    B, N, _ = xyz.shape
    farthest = torch.randint(0, N, (B,), dtype=torch.long)
    if start_with_first_point:
        farthest = farthest * 0
    return _ops.farthest_point_sample(xyz, npoint, farthest)


def query_ball_point(radius, nsample, xyz, new_xyz, get_cnt=False):
    """utils/model_common_utils.py:102-130."""
    return _ops.query_ball_point(radius, nsample, xyz, new_xyz, None, get_cnt)
