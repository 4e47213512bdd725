"""Torch-CPU restatements of the reference's pure-PyTorch hot-path functions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  These issue exactly the library calls the
reference issues (same torch ops in the same order), so on the host CPU they ARE the
reference's CPU path; `/root/reference` itself cannot travel to the GPU box.  Used by
`bench.py --impl reference` (timed with every host thread) and by tests as a second checker.
Each function cites the lines it restates.
"""
import torch


def knn(x, k):
    """utils/model_common_utils.py:3-9."""
    inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pairwise_distance = -xx - inner - xx.transpose(2, 1).contiguous()
    return pairwise_distance.topk(k=k, dim=-1)[1]


def get_graph_feature(x, k=20):
    """utils/model_common_utils.py:132-155 (device taken from x)."""
    x = x.view(*x.size()[:3])
    idx = knn(x, k=k)
    batch_size, num_points, _ = idx.size()
    idx_base = torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points
    idx = (idx + idx_base).view(-1)
    _, num_dims, _ = x.size()
    x = x.transpose(2, 1).contiguous()
    feature = x.view(batch_size * num_points, -1)[idx, :]
    feature = feature.view(batch_size, num_points, k, num_dims)
    x = x.view(batch_size, num_points, 1, num_dims).repeat(1, 1, k, 1)
    return torch.cat((feature, x), dim=3).permute(0, 3, 1, 2)


def chamfer(a, b):
    """losses/chamfer_distance.py:5-31 (the pure-torch fallback the reference takes on CPU when
    its JIT extension is unavailable)."""
    M = (a.unsqueeze(2) - b.unsqueeze(1)).abs().pow(2).sum(3)
    dist1 = torch.mean(torch.sqrt(M.min(1)[0]))
    dist2 = torch.mean(torch.sqrt(M.min(2)[0]))
    return (dist1 + dist2) / 2.0
