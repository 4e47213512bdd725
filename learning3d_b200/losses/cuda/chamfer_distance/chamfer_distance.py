"""Drop-in for learning3d/losses/cuda/chamfer_distance/chamfer_distance.py:14-66.

Same class names and tensor contract (squared distances out, int32 arg-mins saved for backward,
gradients for both clouds).  Instead of JIT-compiling `cd` (chamfer_distance.py:11) the calls go
to libl3d_b200.so: l3d_chamfer_forward / l3d_chamfer_backward (include/l3d_b200.h).
"""
import torch

from .... import _C


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = _C.require_cuda(xyz1, "xyz1")
        xyz2 = _C.require_cuda(xyz2, "xyz2")
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        dev = xyz1.device
        dist1 = torch.empty(batchsize, n, device=dev)
        dist2 = torch.empty(batchsize, m, device=dev)
        idx1 = torch.empty(batchsize, n, dtype=torch.int, device=dev)
        idx2 = torch.empty(batchsize, m, dtype=torch.int, device=dev)
        with _C.on_device(dev):
            _C.check(_C.lib().l3d_chamfer_forward(_C.ptr(xyz1), _C.ptr(xyz2), batchsize, n, m,
                                                  _C.ptr(dist1), _C.ptr(dist2), _C.ptr(idx1),
                                                  _C.ptr(idx2), _C.stream()), "chamfer forward")
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        batchsize, n, _ = xyz1.size()
        m = xyz2.size(1)
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        with _C.on_device(xyz1.device):
            _C.check(_C.lib().l3d_chamfer_backward(
                _C.ptr(xyz1), _C.ptr(xyz2), batchsize, n, m, _C.ptr(graddist1), _C.ptr(graddist2),
                _C.ptr(idx1), _C.ptr(idx2), _C.ptr(gradxyz1), _C.ptr(gradxyz2), _C.stream()),
                "chamfer backward")
        return gradxyz1, gradxyz2


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)
