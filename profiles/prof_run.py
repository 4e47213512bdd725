"""Tiny drivers for `ncu` captures (one short pass of one kernel family; run under gpurun):
    python profiles/prof_run.py edge|emd|chamfer|fps|group|kabsch
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = torch.device("cuda:0")


def main(which):
    torch.manual_seed(0)
    if which == "edge":
        from learning3d_b200.models import DGCNN
        net = DGCNN(emb_dims=512).to(DEV).eval()
        x = torch.rand(32, 1024, 3, device=DEV)
        with torch.no_grad():
            for _ in range(2):
                net(x)
    elif which == "emd":
        from learning3d_b200.losses import EMDLoss
        a = torch.rand(8, 1024, 3, device=DEV, requires_grad=True)
        b = torch.rand(8, 1024, 3, device=DEV)
        for _ in range(2):
            a.grad = None
            EMDLoss()(a, b).backward()
    elif which == "chamfer":
        from learning3d_b200.losses import ChamferDistanceLoss
        for B in (4, 32):
            a = torch.rand(B, 1024, 3, device=DEV, requires_grad=True)
            b = torch.rand(B, 1024, 3, device=DEV, requires_grad=True)
            for _ in range(2):
                a.grad = b.grad = None
                ChamferDistanceLoss()(a, b).backward()
    elif which == "fps":
        from learning3d_b200.utils.lib import pointnet2_utils as pu
        pc = (torch.rand(16, 2048, 3, device=DEV) * 4 - 2).contiguous()
        for _ in range(2):
            fps = pu.furthest_point_sample(pc, 1024)
        new = pu.gather_operation(pc.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        idx = pu.ball_query(0.5, 16, pc, new)
        pu.grouping_operation(pc.transpose(1, 2).contiguous(), idx)
        p1 = torch.rand(16, 256, 3, device=DEV); p2 = torch.rand(16, 256, 3, device=DEV)
        _, kidx = pu.knn(64, p1, p2)
        pu.grouping_operation(torch.rand(16, 128, 256, device=DEV), kidx)
    elif which == "kabsch":
        from learning3d_b200.utils import SVDHead, knn
        es = torch.randn(32, 512, 1024, device=DEV); et = torch.randn(32, 512, 1024, device=DEV)
        src = torch.rand(32, 1024, 3, device=DEV); tgt = torch.rand(32, 1024, 3, device=DEV)
        head = SVDHead(512).to(DEV)
        with torch.no_grad():
            head(es, et, src, tgt)
        knn(torch.randn(32, 64, 1024, device=DEV), 20)          # feature-space graph: Gram + knn_matrix
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1])
