"""Callers of the hot path kept for API-compatibility checks (SURVEY.md §2.1 "caller of *", §8a a16):
the dense layers (1x1 convs, BatchNorm) are plain torch as in the reference — out of scope for the
kernels — while every neighbour search / grouping call lands in libl3d_b200.so."""
from .dgcnn import DGCNN
from .flownet3d import FlowNet3D
from .dcp import DCP
