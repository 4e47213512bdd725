"""Shared-MLP stacks (1x1 conv + BatchNorm + ReLU, optionally followed by a max over the last axis) on the tcgen05
GEMM pipeline of csrc/edgeconv.cu — the pattern of DGCNN's EdgeConv (models/dgcnn.py:34-46) and of every
set-abstraction / flow-embedding / up-conv level of FlowNet3D (models/flownet3d.py:115-121,168-174,231-240,282-286).

Eval mode / no-grad only (BatchNorm is folded into a per-channel scale + shift); callers keep their torch layers
for training.  Activations are channel-major [B, C, positions] exactly as the grouping ops produce them.
"""
import torch

from .. import _C


ENABLED = True        # tests flip this to compare against the torch layers


def fold_bn(conv, bn):
    """(scale, shift) such that bn(conv(x)) == scale * conv_nobias(x) + shift in eval mode."""
    with torch.no_grad():
        if bn is None:
            scale = torch.ones(conv.out_channels, dtype=torch.float32, device=conv.weight.device)
            shift = torch.zeros_like(scale)
        else:
            scale = torch.rsqrt(bn.running_var.float() + bn.eps)
            if bn.weight is not None:
                scale = scale * bn.weight.float()
            shift = -bn.running_mean.float() * scale
            if bn.bias is not None:
                shift = shift + bn.bias.float()
        if conv.bias is not None:
            shift = shift + conv.bias.float() * scale
    return scale.contiguous(), shift.contiguous()


def _layer_cache(conv, bn):
    tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
    key = tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)
    c = conv.__dict__.get("_l3d_fused")
    if c is None or c[0] != key:
        with torch.no_grad():
            w = conv.weight.detach().float().reshape(conv.out_channels, -1)
            c = (key, w.t().contiguous()) + fold_bn(conv, bn)
        conv.__dict__["_l3d_fused"] = c
    return c[1], c[2], c[3]


def _unpack(layer):
    """(conv, bn, relu) of one MLP level given as (conv, bn) or as nn.Sequential(conv, bn, relu)."""
    if isinstance(layer, (tuple, list)):
        return layer[0], layer[1], True
    mods = list(layer.children())
    conv = mods[0]
    bn = next((m for m in mods[1:] if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d))), None)
    return conv, bn, any(isinstance(m, torch.nn.ReLU) for m in mods[1:])


def usable(x, layers, training):
    """True when the fused stack applies: eval, no autograd, CUDA fp32, TMA-friendly shapes."""
    if not ENABLED or training or not x.is_cuda or x.dtype != torch.float32:
        return False
    pos = 1
    for d in x.shape[2:]:
        pos *= d
    if pos & 3 or x.shape[0] > 65535 or x.shape[0] < 1:
        return False
    convs = [_unpack(l)[0] for l in layers]
    if torch.is_grad_enabled() and (x.requires_grad or any(c.weight.requires_grad for c in convs)):
        return False
    return all((c.out_channels & 3) == 0 and all(k == 1 for k in c.kernel_size) for c in convs) and len(convs) > 0


def mlp_forward(x, layers, pool=False):
    """x [B, C, S] or [B, C, S, K] -> the stack's output [B, C_out, S(, K)]; with pool=True the max over the last
    axis is taken in the last layer's epilogue and [B, C_out, S] is returned (its full activation is never written)."""
    lib = _C.lib()
    shape = x.shape
    B = shape[0]
    G = shape[-1] if (pool and x.dim() == 4) else 1
    P = 1
    for d in shape[2:]:
        P *= d
    h = x.contiguous().view(B, shape[1], P)
    with _C.on_device(x.device):
        st = _C.stream()
        for i, layer in enumerate(layers):
            conv, bn, relu = _unpack(layer)
            wt, scale, shift = _layer_cache(conv, bn)
            last = i == len(layers) - 1
            K, M = wt.shape
            if last and pool and G > 1:
                out = torch.empty((B, M, P // G), dtype=torch.float32, device=x.device)
                _C.check(lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(h), _C.ptr(scale), _C.ptr(shift), B, M, K, P, G,
                                                      1 if relu else 0, _C.ptr(None), _C.ptr(out), M * (P // G), 0, st),
                         "fused mlp")
                return out
            out = torch.empty((B, M, P), dtype=torch.float32, device=x.device)
            _C.check(lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(h), _C.ptr(scale), _C.ptr(shift), B, M, K, P, 1,
                                                  1 if relu else 0, _C.ptr(out), _C.ptr(None), 0, 0, st), "fused mlp")
            h = out
    return h.view((B, h.shape[1]) + tuple(shape[2:]))
