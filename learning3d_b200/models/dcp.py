"""Deep Closest Point assembled from the hot-path pieces: DGCNN (fused kNN graph) -> Transformer pointer ->
SVDHead (batched Kabsch tail).  Interface and state_dict layout of learning3d/models/dcp.py:10-55
(`emb_nn.*`, `pointer.model.*`, `head.reflect`); result dictionary with the same keys."""
import torch
import torch.nn as nn

from ..utils import SVDHead
from ..utils.transformer import Identity, Transformer
from .dgcnn import DGCNN


def transform_point_cloud(point_cloud, rotation, translation):
    """ops/transform_functions.py:24-29 — R p + t for [B,N,3] clouds."""
    return torch.matmul(rotation, point_cloud.permute(0, 2, 1)).permute(0, 2, 1) + translation.unsqueeze(1)


def convert2transformation(rotation_matrix, translation_vector):
    """ops/transform_functions.py:31-35 — [B,4,4] homogeneous transform."""
    B = rotation_matrix.shape[0]
    T = torch.eye(4, device=rotation_matrix.device, dtype=rotation_matrix.dtype).repeat(B, 1, 1)
    T[:, :3, :3] = rotation_matrix
    T[:, :3, 3] = translation_vector
    return T


class DCP(nn.Module):
    def __init__(self, feature_model=None, cycle=False, pointer_='transformer', head='svd'):
        super().__init__()
        self.cycle = cycle
        self.emb_nn = feature_model if feature_model is not None else DGCNN()
        if pointer_ == 'identity':
            self.pointer = Identity()
        elif pointer_ == 'transformer':
            self.pointer = Transformer(self.emb_nn.emb_dims, n_blocks=1, dropout=0.0, ff_dims=1024, n_heads=4)
        else:
            raise Exception("Not implemented")
        if head != 'svd':
            raise Exception('Not implemented')      # the MLP head of the reference is outside the hot path
        self.head = SVDHead(self.emb_nn.emb_dims)

    def forward(self, template, source):
        if not self.training and source.shape == template.shape:
            # eval: one 2B-cloud pass through the embedding network (BatchNorm uses running statistics, so this
            # equals the two separate calls of models/dcp.py:31-32) — half the launches, fuller waves
            both = self.emb_nn(torch.cat([source, template], 0))
            src_f, tpl_f = both[:source.shape[0]].contiguous(), both[source.shape[0]:].contiguous()
        else:
            src_f, tpl_f = self.emb_nn(source), self.emb_nn(template)
        src_p, tpl_p = self.pointer(src_f, tpl_f)
        src_f, tpl_f = src_f + src_p, tpl_f + tpl_p
        rot_ab, trans_ab = self.head(src_f, tpl_f, source, template)
        if self.cycle:
            rot_ba, trans_ba = self.head(tpl_f, src_f, template, source)
        else:
            rot_ba = rot_ab.transpose(2, 1).contiguous()
            trans_ba = -torch.matmul(rot_ba, trans_ab.unsqueeze(2)).squeeze(2)
        return {'est_R': rot_ab, 'est_t': trans_ab, 'est_R_': rot_ba, 'est_t_': trans_ba,
                'est_T': convert2transformation(rot_ab, trans_ab), 'r': tpl_f - src_f,
                'transformed_source': transform_point_cloud(source, rot_ab, trans_ab)}
