"""DGCNN embedding network with the interface and state_dict keys of learning3d/models/dgcnn.py:6-49
(conv1..conv5, bn1..bn5): ONE static kNN graph on xyz (get_graph_feature), four EdgeConv 1x1 convs each
followed by a max over the k neighbours, then conv5 on the concatenation."""
import torch
import torch.nn.functional as F

from ..utils import get_graph_feature


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

    def forward(self, input_data):
        if self.input_shape == "bnc":
            input_data = input_data.permute(0, 2, 1)
        if input_data.shape[1] != 3:
            raise RuntimeError("shape of x must be of [Batch x 3 x NumInPoints]")
        batch_size, _, num_points = input_data.size()
        x = get_graph_feature(input_data.contiguous())          # fused kNN + gather: [B, 6, N, k]
        pooled = []
        for i in range(1, 5):
            x = F.relu(getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(x)))
            pooled.append(x.max(dim=-1, keepdim=True)[0])
        x = torch.cat(pooled, dim=1)
        return F.relu(self.bn5(self.conv5(x))).view(batch_size, -1, num_points)
