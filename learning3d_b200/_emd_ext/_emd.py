"""`_emd_ext._emd` on libl3d_b200.so — the two functions of losses/cuda/emd_torch/pkg/include/emd.h:25-46
with the reference's contract: callee-allocated outputs, CUDA + contiguous inputs checked
(cuda_helper.h:12-16 CHECK_INPUT), `emd_forward(xyz1, xyz2) -> [cost[B], match[B,N1,N2]]`,
`emd_backward(xyz1, xyz2, match) -> [grad1, grad2]`.  The reference's emd_loss_layer.py runs unmodified with

    import sys, learning3d_b200._emd_ext
    sys.modules["_emd_ext"] = learning3d_b200._emd_ext
    sys.modules["_emd_ext._emd"] = learning3d_b200._emd_ext._emd

Only float32 is built (the reference dispatches float64 too, emd.cuh:188); no device synchronisation inside
(the reference calls cudaDeviceSynchronize in the forward, emd.cuh:197).
"""
import torch

from .. import _C


def _check_input(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if t.dtype != torch.float32:
        raise NotImplementedError("learning3d_b200 _emd: only float32 is built, got %s" % t.dtype)
    if t.dim() != 3 or t.shape[2] != 3:
        raise NotImplementedError("learning3d_b200 _emd: point sets must be [B, N, 3]")


def emd_forward(xyz1, xyz2):
    _check_input(xyz1, "xyz1")
    _check_input(xyz2, "xyz2")
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    lib = _C.lib()
    cost = torch.empty((B,), dtype=torch.float32, device=xyz1.device)
    match = torch.empty((B, n, m), dtype=torch.float32, device=xyz1.device)
    with _C.on_device(xyz1.device):
        ws = torch.empty(max(int(lib.l3d_emd_forward_ws_bytes(B, n, m)), 16), dtype=torch.uint8, device=xyz1.device)
        _C.check(lib.l3d_emd_forward(_C.ptr(xyz1), _C.ptr(xyz2), B, n, m, _C.ptr(cost), _C.ptr(match), _C.ptr(ws),
                                     _C.stream()), "emd_forward")
    return [cost, match]


def emd_backward(xyz1, xyz2, match):
    _check_input(xyz1, "xyz1")
    _check_input(xyz2, "xyz2")
    if not (match.is_cuda and match.is_contiguous()):
        raise RuntimeError("match must be a contiguous CUDA tensor")
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    lib = _C.lib()
    g1 = torch.empty_like(xyz1)
    g2 = torch.empty_like(xyz2)
    with _C.on_device(xyz1.device):
        ws = torch.empty(max(int(lib.l3d_emd_backward_ws_bytes(B, n, m)), 16), dtype=torch.uint8, device=xyz1.device)
        _C.check(lib.l3d_emd_backward(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(match), B, n, m, _C.ptr(g1), _C.ptr(g2),
                                      _C.ptr(ws), _C.stream()), "emd_backward")
    return [g1, g2]
