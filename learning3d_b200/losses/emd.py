"""Drop-in for learning3d/losses/emd.py:5-15.

The reference file is broken as shipped (imports a top-level package `emd` and calls `self.emd`
inside a free function -> NameError, losses/emd.py:5-8).  This module implements its evident
intent, mean_b(EMD(template, source)) / N, on the CUDA approxmatch path, with a mathematically
correct gradient (the upstream gradient is applied; the raw reference Function ignores it, see
losses/cuda/emd_torch/pkg/layer/emd_loss_layer.py).  Deviation documented in DESIGN.md §6.
"""
import torch
import torch.nn as nn

from .cuda.emd_torch.pkg.layer.emd_loss_layer import EMDFunction


def emd(template: torch.Tensor, source: torch.Tensor):
    cost = EMDFunction.apply(template.contiguous(), source.contiguous(), True)
    return torch.mean(cost) / (template.size()[1])


class EMDLoss(nn.Module):
    def __init__(self):
        super(EMDLoss, self).__init__()

    def forward(self, template, source):
        return emd(template, source)
