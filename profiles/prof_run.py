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
    elif which == "knn":
        from learning3d_b200.utils import knn
        x = torch.rand(32, 3, 1024, device=DEV)
        for _ in range(5):
            knn(x, 20)
    elif which == "knnstream":
        from learning3d_b200.utils import knn
        x = torch.rand(2, 3, 16384, device=DEV)           # beyond the resident kernel: streamed selection
        for _ in range(3):
            knn(x, 20)
    elif which == "dcp":
        from learning3d_b200.models import DCP, DGCNN
        net = DCP(feature_model=DGCNN(emb_dims=512), cycle=True).to(DEV).eval()
        a = torch.rand(32, 1024, 3, device=DEV); b = torch.rand(32, 1024, 3, device=DEV)
        with torch.no_grad():
            for _ in range(2):
                net(a, b)
                torch.cuda.synchronize()
    elif which == "flownet":
        from learning3d_b200.models import FlowNet3D
        fn = FlowNet3D().to(DEV).eval()
        pc1 = torch.rand(16, 3, 2048, device=DEV) * 4 - 2
        pc2 = pc1 + 0.05 * torch.randn_like(pc1)
        f1 = torch.rand(16, 3, 2048, device=DEV); f2 = torch.rand(16, 3, 2048, device=DEV)
        with torch.no_grad():
            for _ in range(2):
                fn(pc1, pc2, f1, f2)
                torch.cuda.synchronize()
    elif which == "attn":
        from learning3d_b200.utils.transformer import MultiHeadedAttention
        from learning3d_b200.utils.transformer_fused import attention_cm
        attn = MultiHeadedAttention(4, 512).to(DEV).eval()
        x = torch.randn(32, 512, 1024, device=DEV)
        with torch.no_grad():
            for _ in range(2):
                attention_cm(attn, x, x, x)
    elif which == "rpm":
        from learning3d_b200.models import rpmnet as R
        from learning3d_b200.utils._ops import feature_square_distance
        fs = 0.3 * torch.randn(8, 717, 96, device=DEV); fr = 0.3 * torch.randn(8, 717, 96, device=DEV)
        beta = torch.ones(8, device=DEV); alpha = torch.full((8,), 0.5, device=DEV)
        xyz = torch.rand(8, 717, 3, device=DEV)
        for _ in range(2):
            aff = feature_square_distance(fs, fr, beta, alpha)
            perm, wt, rs = R.match_tail(aff, xyz, 5, True)
            R.compute_rigid_transform(xyz, wt, rs)
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
