"""learning3d_b200 — B200 (sm_100a) drop-in for learning3d's data-parallel hot path.

Mirrors the reference's module paths for the hot-path symbols only:
  learning3d_b200.utils   <-> learning3d.utils   (knn, get_graph_feature, square_distance, ...)
  learning3d_b200.losses  <-> learning3d.losses  (ChamferDistanceLoss, EMDLoss)
  learning3d_b200.models  <-> the callers kept for API-compat checks (DGCNN ...)
All compute goes through libl3d_b200.so (C ABI in include/l3d_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
