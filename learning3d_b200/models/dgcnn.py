"""DGCNN embedding network with the interface and state_dict keys of learning3d/models/dgcnn.py:6-49
(conv1..conv5, bn1..bn5): ONE static kNN graph on xyz (get_graph_feature), four EdgeConv 1x1 convs each
followed by a max over the k neighbours, then conv5 on the concatenation."""
import torch
import torch.nn.functional as F

from .. import _C
from ..utils import get_graph_feature, knn
from ..utils.fused_mlp import fold_bn as _fold_bn


def _edge_cache(net, dev):
    """W^T [C_in, C_out] (the MN-major A operand of l3d_conv1x1_bn_relu_maxk), folded BN and layer 1's host copy,
    rebuilt whenever a parameter / buffer of the five conv+bn pairs changes (tensor version counters)."""
    tensors = []
    for i in range(1, 6):
        conv, bn = getattr(net, "conv%d" % i), getattr(net, "bn%d" % i)
        tensors += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    key = (str(dev),) + tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors)
    cache = net.__dict__.get("_l3d_edge_cache")
    if cache is not None and cache["key"] == key:
        return cache
    cache = {"key": key, "wt": [], "scale": [], "shift": []}
    with torch.no_grad():
        for i in range(1, 6):
            conv, bn = getattr(net, "conv%d" % i), getattr(net, "bn%d" % i)
            w = conv.weight.detach().float().reshape(conv.out_channels, conv.in_channels)
            sc, sh = _fold_bn(conv, bn)
            cache["wt"].append(w.t().contiguous())
            cache["scale"].append(sc)
            cache["shift"].append(sh)
            if i == 1:
                cache["l1_host"] = (w.contiguous().cpu(), sc.cpu(), sh.cpu())
    net.__dict__["_l3d_edge_cache"] = cache
    return cache


def _fused_ok(net, x):
    if net.training or not x.is_cuda or x.dtype != torch.float32:
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in net.parameters())):
        return False                      # autograd needs the torch layers (the fused stack is forward-only)
    B, _, N = x.shape
    k = 20
    if N < k or (N & 3) or B < 1 or B > 65535:
        return False
    convs = [getattr(net, "conv%d" % i) for i in range(1, 6)]
    if tuple(convs[0].weight.shape[:2]) != (64, 6):
        return False
    for c in convs:
        if c.kernel_size != (1, 1) or (c.out_channels & 3):
            return False
    return not any(isinstance(getattr(net, "bn%d" % i).running_var, type(None)) for i in range(1, 6))


def edgeconv_stack(net, x, k=20):
    """models/dgcnn.py:32-48 for an eval-mode module `net` (conv1..5, bn1..5) and a cloud x [B,3,N] on CUDA:
    kNN indices -> layer 1 from the indices (gather-on-load) -> layers 2-4 + conv5 on tcgen05 (3xTF32), BatchNorm
    folded, ReLU and the max over the k neighbours in the TMEM epilogue.  Neither the [B,6,N,k] graph feature
    nor the layer-4 activations (671 MB at B=32) are ever written."""
    lib = _C.lib()
    B, _, N = x.shape
    P = N * k
    dev = x.device
    cache = _edge_cache(net, dev)
    widths = [w.shape[1] for w in cache["wt"]]           # 64, 64, 128, 256, emb
    ctot = sum(widths[:4])
    idx = knn(x, k)
    cat = torch.empty((B, ctot, N), dtype=torch.float32, device=dev)
    h = torch.empty((B, widths[0], P), dtype=torch.float32, device=dev)
    w1, s1, t1 = cache["l1_host"]
    with _C.on_device(dev):
        st = _C.stream()
        _C.check(lib.l3d_edgeconv_layer1(_C.ptr(x), _C.ptr(idx), _C._P(w1.data_ptr()), _C._P(s1.data_ptr()),
                                         _C._P(t1.data_ptr()), B, N, k, widths[0], _C.ptr(h), _C.ptr(cat),
                                         ctot * N, 0, st), "edgeconv layer 1")
        coff = widths[0]
        for i in (1, 2, 3):
            last = i == 3
            nxt = None if last else torch.empty((B, widths[i], P), dtype=torch.float32, device=dev)
            _C.check(lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(cache["wt"][i]), _C.ptr(h), _C.ptr(cache["scale"][i]),
                                                  _C.ptr(cache["shift"][i]), B, widths[i], widths[i - 1], P, k, 1,
                                                  _C.ptr(nxt), _C.ptr(cat), ctot * N, coff, st),
                     "edgeconv layer %d" % (i + 1))
            coff += widths[i]
            h = nxt
        out = torch.empty((B, widths[4], N), dtype=torch.float32, device=dev)
        _C.check(lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(cache["wt"][4]), _C.ptr(cat), _C.ptr(cache["scale"][4]),
                                              _C.ptr(cache["shift"][4]), B, widths[4], ctot, N, 1, 1,
                                              _C.ptr(out), _C.ptr(None), 0, 0, st), "edgeconv conv5")
    return out


def dgcnn_forward(self, input_data):
    """DGCNN.forward (models/dgcnn.py:25-49) for this package's DGCNN AND for the reference's own class
    (learning3d_b200.bind swaps it in): the fused tensor-core stack in eval mode / no-grad, the torch layers on the
    fused kNN graph feature otherwise."""
    if self.input_shape == "bnc":
        input_data = input_data.permute(0, 2, 1)
    if input_data.shape[1] != 3:
        raise RuntimeError("shape of x must be of [Batch x 3 x NumInPoints]")
    batch_size, _, num_points = input_data.size()
    x = input_data.contiguous()
    if _fused_ok(self, x):
        return edgeconv_stack(self, x)
    x = get_graph_feature(x)          # fused kNN + gather: [B, 6, N, k]
    pooled = []
    for i in range(1, 5):
        x = F.relu(getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(x)))
        pooled.append(x.max(dim=-1, keepdim=True)[0])
    x = torch.cat(pooled, dim=1)
    return F.relu(self.bn5(self.conv5(x))).view(batch_size, -1, num_points)


class DGCNN(torch.nn.Module):
    WIDTHS = (64, 64, 128, 256)

    def __init__(self, emb_dims=1024, input_shape="bnc"):
        super().__init__()
        if input_shape not in ["bcn", "bnc"]:
            raise ValueError("Allowed shapes are 'bcn' (batch * channels * num_in_points), 'bnc' ")
        self.input_shape = input_shape
        self.emb_dims = emb_dims
        c_in = 6
        for i, c_out in enumerate(self.WIDTHS, start=1):
            setattr(self, "conv%d" % i, torch.nn.Conv2d(c_in, c_out, kernel_size=1, bias=False))
            setattr(self, "bn%d" % i, torch.nn.BatchNorm2d(c_out))
            c_in = c_out
        self.conv5 = torch.nn.Conv2d(sum(self.WIDTHS), emb_dims, kernel_size=1, bias=False)
        self.bn5 = torch.nn.BatchNorm2d(emb_dims)

    forward = dgcnn_forward
