"""Rebind an imported `learning3d` package (the reference, unmodified) to libl3d_b200.so.

This is the reference-side binding of INTEGRATION.md as code: the reference's own models
(learning3d.models.DCP / DGCNN / FlowNet3D ...) keep running, but the hot-path functions they call are the
C-ABI kernels.

    import learning3d, learning3d_b200.bind
    learning3d_b200.bind.bind(learning3d)            # ... run examples/test_dcp.py as usual
    learning3d_b200.bind.unbind(learning3d)

What is replaced (reference file:line -> ours):
  utils/model_common_utils.py:3-155  knn, get_graph_feature, square_distance, index_points,
                                     farthest_point_sample, knn_point, query_ball_point (and the names
                                     re-exported by utils/__init__.py and imported into models/*.py)
  utils/svd.py:13-59                 SVDHead.forward  (fused soft correspondences + batched Kabsch)
  utils/transformer.py:255-263       Transformer.forward (eval: linear layers, attention, LayerNorm on tcgen05)
  models/dgcnn.py:25-49              DGCNN.forward    (eval mode: kNN graph + EdgeConv stack on tcgen05;
                                     training mode keeps the torch layers on the fused graph feature)
  models/flownet3d.py:73-328         forward of PointNetSetAbstraction / FlowEmbedding / PointNetSetUpConv /
                                     PointNetFeaturePropogation / FlowNet3D (eval: shared MLPs + max on tcgen05)
  models/rpmnet.py:130-254           match_features, sinkhorn, compute_rigid_transform (RPMNet's matching tail)
  utils/lib/pointnet2_utils.py:8     the `pointnet2_cuda` extension module
Nothing is copied from the reference; only attributes of the live package objects are swapped.
"""
import sys

_SAVED = {}     # id(pkg) -> list of (obj, attr, old)


def _swap(rec, obj, attr, new):
    if not hasattr(obj, attr):
        if attr.startswith("_"):                 # helper methods our forwards rely on (added, removed by unbind)
            setattr(obj, attr, new)
            rec.append((obj, attr, None))
        return
    rec.append((obj, attr, getattr(obj, attr)))
    setattr(obj, attr, new)


def bind(pkg, edgeconv=True):
    """Swap the hot-path callables of the imported reference package `pkg` for the libl3d_b200.so ones."""
    if id(pkg) in _SAVED:
        return pkg
    from . import utils as U
    from .utils import svd as our_svd
    rec = []
    names = ("knn", "get_graph_feature", "square_distance", "index_points", "farthest_point_sample",
             "knn_point", "query_ball_point")
    mcu = sys.modules.get(pkg.__name__ + ".utils.model_common_utils")
    targets = [mcu, sys.modules.get(pkg.__name__ + ".utils")]
    for mod_name in ("dgcnn", "flownet3d", "prnet", "curvenet"):
        targets.append(sys.modules.get(pkg.__name__ + ".models." + mod_name))
    for mod in targets:
        if mod is None:
            continue
        for n in names:
            _swap(rec, mod, n, getattr(U, n))
    svd_mod = sys.modules.get(pkg.__name__ + ".utils.svd")
    if svd_mod is not None:
        _swap(rec, svd_mod.SVDHead, "forward", our_svd.SVDHead.forward)
    dg = sys.modules.get(pkg.__name__ + ".models.dgcnn")
    if edgeconv and dg is not None:
        from .models import dgcnn as our_dgcnn
        if hasattr(our_dgcnn, "dgcnn_forward"):
            _swap(rec, dg.DGCNN, "forward", our_dgcnn.dgcnn_forward)
    fl = sys.modules.get(pkg.__name__ + ".models.flownet3d")
    if fl is not None and hasattr(fl, "FlowNet3D"):
        # same attribute names as ours (the checkpoints force them), so our forward methods run on the reference's
        # objects: grouping on the C ABI, shared MLPs + max on tcgen05 in eval mode, both frames batched
        from .models import flownet3d as our_fl

        def guarded(ours, theirs, ok):
            def forward(self, *a, **k):
                return ours(self, *a, **k) if ok(self) else theirs(self, *a, **k)
            return forward
        for name, ok in (("PointNetSetAbstraction", lambda m: True),
                         ("FlowEmbedding", lambda m: m.knn and m.corr_func == "concat" and m.pooling == "max"),
                         ("PointNetSetUpConv", lambda m: m.knn),
                         ("PointNetFeaturePropogation", lambda m: True)):
            cls = getattr(fl, name, None)
            if cls is not None:
                _swap(rec, cls, "forward", guarded(getattr(our_fl, name).forward, cls.forward, ok))
        _swap(rec, fl.FlowNet3D, "_encode", our_fl.FlowNet3D._encode)
        _swap(rec, fl.FlowNet3D, "forward", our_fl.FlowNet3D.forward)
    tr = sys.modules.get(pkg.__name__ + ".utils.transformer")
    if tr is not None and hasattr(tr, "Transformer"):
        from .utils.transformer_fused import transformer_forward
        if not hasattr(tr.Transformer, "_l3d_torch_forward"):
            tr.Transformer._l3d_torch_forward = tr.Transformer.forward      # kept for training / autograd
            rec.append((tr.Transformer, "_l3d_torch_forward", None))
        _swap(rec, tr.Transformer, "forward", transformer_forward)
    rpm = sys.modules.get(pkg.__name__ + ".models.rpmnet")
    if rpm is not None:
        from .models import rpmnet as our_rpm
        for n in ("match_features", "sinkhorn", "compute_rigid_transform"):
            _swap(rec, rpm, n, getattr(our_rpm, n))
    pn2 = sys.modules.get(pkg.__name__ + ".utils.lib.pointnet2_utils")
    if pn2 is not None:
        from . import pointnet2_cuda
        _swap(rec, pn2, "pointnet2", pointnet2_cuda)
    _SAVED[id(pkg)] = rec
    return pkg


def unbind(pkg):
    """Undo bind(pkg)."""
    for obj, attr, old in reversed(_SAVED.pop(id(pkg), [])):
        if old is None:
            delattr(obj, attr)
        else:
            setattr(obj, attr, old)
    return pkg
