"""Drop-in for learning3d/losses/chamfer_distance.py:5-51.

`chamfer_distance()` keeps the reference's value — (mean sqrt d1 + mean sqrt d2) / 2 — but runs as
ONE fused forward launch and ONE backward launch (l3d_chamfer_loss_forward/backward) instead of the
reference's native call plus ~10 elementwise launches, and never falls back: the reference's bare
`except:` (chamfer_distance.py:41) that silently switches to the O(B*N*M*3)-memory torch path is
replaced by a hard error.
"""
import torch
import torch.nn as nn

from .. import _C

_ws_cache = {}


def _workspace(dev, nbytes):
    """Device scratch for the fused loss, one per (device, stream); zero-filled once."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def pairwise_distances(a: torch.Tensor, b: torch.Tensor, p=2):
    """losses/chamfer_distance.py:5-19 — kept for API compatibility (materialises [m,n,n])."""
    if len(a.shape) != 3:
        raise ValueError("Invalid shape for a. Must be [m, n, d] but got", a.shape)
    if len(b.shape) != 3:
        raise ValueError("Invalid shape for a. Must be [m, n, d] but got", b.shape)
    return (a.unsqueeze(2) - b.unsqueeze(1)).abs().pow(p).sum(3)


class _ChamferLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, template, source):
        xyz1 = _C.require_cuda(template, "template")
        xyz2 = _C.require_cuda(source, "source")
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3:
            raise ValueError("chamfer_distance expects [B, N, 3] clouds, got %s and %s"
                             % (tuple(xyz1.shape), tuple(xyz2.shape)))
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        tot = B * (n + m)
        buf = torch.empty(2 * tot + 1, device=dev)          # [dist | idx (int32 bits) | loss]
        dist, idx, loss = buf, buf, buf[2 * tot]
        dptr = buf.data_ptr()
        lib = _C.lib()
        with _C.on_device(dev):
            ws = _workspace(dev, int(lib.l3d_chamfer_ws_bytes(B, n, m)))
            _C.check(lib.l3d_chamfer_loss_forward(
                _C.ptr(xyz1), _C.ptr(xyz2), B, n, m, _C._P(dptr), _C._P(dptr + 4 * B * n),
                _C._P(dptr + 4 * tot), _C._P(dptr + 4 * (tot + B * n)), _C._P(dptr + 8 * tot), _C.ptr(ws),
                _C.stream()), "chamfer loss forward")
        ctx.save_for_backward(xyz1, xyz2, buf)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        xyz1, xyz2, buf = ctx.saved_tensors
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        tot = B * (n + m)
        dptr = buf.data_ptr()
        if grad_loss.dtype != torch.float32 or not grad_loss.is_contiguous():
            grad_loss = grad_loss.contiguous().to(torch.float32)
        g1 = torch.empty_like(xyz1)
        g2 = torch.empty_like(xyz2)
        with _C.on_device(xyz1.device):
            _C.check(_C.lib().l3d_chamfer_loss_backward(
                _C.ptr(xyz1), _C.ptr(xyz2), B, n, m, _C._P(dptr), _C._P(dptr + 4 * B * n),
                _C._P(dptr + 4 * tot), _C._P(dptr + 4 * (tot + B * n)), _C.ptr(grad_loss), _C.ptr(g1),
                _C.ptr(g2), _C.stream()), "chamfer loss backward")
        return g1, g2


def chamfer(a, b):
    """losses/chamfer_distance.py:21-31: same value as chamfer_distance() (the reference's
    pure-torch formulation); here it is the same fused CUDA path."""
    return _ChamferLoss.apply(a, b)


def chamfer_distance(template: torch.Tensor, source: torch.Tensor):
    """losses/chamfer_distance.py:34-43."""
    return _ChamferLoss.apply(template, source)


class ChamferDistanceLoss(nn.Module):
    def __init__(self):
        super(ChamferDistanceLoss, self).__init__()

    def forward(self, template, source):
        return chamfer_distance(template, source)
