"""Drop-in for learning3d/losses/cuda/emd_torch/pkg/layer/emd_loss_layer.py:7-40.

Same names and contract: EMDFunction.forward returns cost [B] and saves (xyz1, xyz2, match);
backward returns the gradients of cost with the matching held constant and — exactly like the
reference (emd_loss_layer.py:16-19) — IGNORES grad_output unless `scale_by_grad_output` is set.
The `_emd_ext._emd` extension (which no longer compiles) is replaced by l3d_emd_forward /
l3d_emd_backward of libl3d_b200.so.
"""
import torch
import torch.nn as nn

from ...... import _C


def _ws(dev, nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


class EMDFunction(torch.autograd.Function):
    @staticmethod
    def forward(self, xyz1, xyz2, scale_by_grad_output=False):
        # CHECK_INPUT of the reference: CUDA + contiguous (cuda_helper.h:12-16)
        if not (xyz1.is_cuda and xyz2.is_cuda):
            raise RuntimeError("xyz1 / xyz2 must be a CUDA tensor")
        if not (xyz1.is_contiguous() and xyz2.is_contiguous()):
            raise RuntimeError("xyz1 / xyz2 must be contiguous")
        xyz1 = _C.require_cuda(xyz1, "xyz1")
        xyz2 = _C.require_cuda(xyz2, "xyz2")
        B, n, d = xyz1.shape
        m = xyz2.shape[1]
        if d != 3 or xyz2.shape[2] != 3:
            raise NotImplementedError("learning3d_b200 EMD: only 3-D points are built")
        dev = xyz1.device
        lib = _C.lib()
        cost = torch.empty((B,), dtype=torch.float32, device=dev)
        match = torch.empty((B, n, m), dtype=torch.float32, device=dev)     # emd.cu:18
        with _C.on_device(dev):
            ws = _ws(dev, lib.l3d_emd_forward_ws_bytes(B, n, m))
            _C.check(lib.l3d_emd_forward(_C.ptr(xyz1), _C.ptr(xyz2), B, n, m, _C.ptr(cost),
                                         _C.ptr(match), _C.ptr(ws), _C.stream()), "emd_forward")
        self.save_for_backward(xyz1, xyz2, match)
        self.scale = bool(scale_by_grad_output)
        return cost

    @staticmethod
    def backward(self, grad_output):
        xyz1, xyz2, match = self.saved_tensors
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        lib = _C.lib()
        g1 = torch.empty_like(xyz1)
        g2 = torch.empty_like(xyz2)
        with _C.on_device(xyz1.device):
            ws = _ws(xyz1.device, lib.l3d_emd_backward_ws_bytes(B, n, m))
            _C.check(lib.l3d_emd_backward(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(match), B, n, m,
                                          _C.ptr(g1), _C.ptr(g2), _C.ptr(ws), _C.stream()), "emd_backward")
        if self.scale:
            g = grad_output.to(torch.float32).view(B, 1, 1)
            g1, g2 = g1 * g, g2 * g
        return g1, g2, None


class EMDLoss(nn.Module):
    """Approximate EMD between two point sets (emd_loss_layer.py:24-40): returns cost [B]."""

    def __init__(self):
        super(EMDLoss, self).__init__()

    def forward(self, xyz1, xyz2):
        assert xyz1.shape[-1] == xyz2.shape[-1], 'Both point sets must have the same dimensionality'
        return EMDFunction.apply(xyz1, xyz2)
