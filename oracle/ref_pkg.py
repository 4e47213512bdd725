"""TEST INFRASTRUCTURE ONLY — imports the reference's OWN Python package as `learning3d`.

Where it comes from: /root/reference in the build container, else the copy oracle/build_ref.py staged
under oracle/_ref/learning3d/ (git-ignored; it travels to the GPU box with the snapshot like the
oracle/_ref/*.so files).  Nothing under learning3d_b200/ imports this module.

`pointnet2_cuda` — the extension utils/lib/pointnet2_utils.py:8 imports and which no longer builds
(THC) — is provided as a tiny module whose ten functions have the signatures of
utils/lib/src/pointnet2_api.cpp:10-25 and forward to either
  * backend "ref": the reference's own CUDA kernels compiled in place into oracle/_ref/libpn2_ref.so, or
  * backend "l3d": libl3d_b200.so (exactly the binding INTEGRATION.md §4 describes).
The reference's FlowNet3D then runs unmodified on either.
"""
import ctypes
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(HERE, "_ref", "learning3d")
_LIVE = "/root/reference"
_pkg = None


def reference_root():
    if os.path.isdir(os.path.join(_LIVE, "models")):
        return _LIVE
    if os.path.isdir(os.path.join(_STAGED, "models")):
        return _STAGED
    return None


def checkpoint(rel):
    """Path of pretrained/<rel> (e.g. 'exp_dcp/models/best_model.t7') or None."""
    root = reference_root()
    p = os.path.join(root, "pretrained", rel) if root else None
    return p if p and os.path.exists(p) else None


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def pointnet2_module(backend):
    """A stand-in for the `pointnet2_cuda` extension (pointnet2_api.cpp:10-25) on the chosen backend."""
    m = types.ModuleType("pointnet2_cuda")
    if backend == "ref":
        lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libpn2_ref.so"))

        def call(name, *a):
            fn = getattr(lib, "ref_" + name, None)
            if fn is None:
                raise NotImplementedError("libpn2_ref.so has no %s (forward-only shim)" % name)
            fn.restype = None
            fn(*a, _s())
        names = {"ball_query": "ball_query", "group_points": "group_points", "group_points_grad": "group_points_grad",
                 "gather_points": "gather_points", "gather_points_grad": "gather_points_grad", "fps": "fps",
                 "knn": "knn", "three_nn": "three_nn", "three_interpolate": "three_interpolate",
                 "three_interpolate_grad": "three_interpolate_grad"}
    elif backend == "l3d":
        # the product's own stand-in for the extension
        import learning3d_b200.pointnet2_cuda as product
        return product
    else:
        raise ValueError(backend)
    F = ctypes.c_float

    def ball_query_wrapper(b, n, m_, radius, nsample, new_xyz, xyz, idx):
        call(names["ball_query"], b, n, m_, F(radius), nsample, _p(new_xyz), _p(xyz), _p(idx))

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        call(names["group_points"], b, c, n, npoints, nsample, _p(points), _p(idx), _p(out))

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        call(names["group_points_grad"], b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(grad_points))

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        call(names["gather_points"], b, c, n, npoints, _p(points), _p(idx), _p(out))

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        call(names["gather_points_grad"], b, c, n, npoints, _p(grad_out), _p(idx), _p(grad_points))

    def furthest_point_sampling_wrapper(b, n, m_, points, temp, idx):
        call(names["fps"], b, n, m_, _p(points), _p(temp), _p(idx))

    def knn_wrapper(b, n, m_, k, unknown, known, dist2, idx):
        call(names["knn"], b, n, m_, k, _p(unknown), _p(known), _p(dist2), _p(idx))

    def three_nn_wrapper(b, n, m_, unknown, known, dist2, idx):
        call(names["three_nn"], b, n, m_, _p(unknown), _p(known), _p(dist2), _p(idx))

    def three_interpolate_wrapper(b, c, m_, n, points, idx, weight, out):
        call(names["three_interpolate"], b, c, m_, n, _p(points), _p(idx), _p(weight), _p(out))

    def three_interpolate_grad_wrapper(b, c, n, m_, grad_out, idx, weight, grad_points):
        call(names["three_interpolate_grad"], b, c, n, m_, _p(grad_out), _p(idx), _p(weight), _p(grad_points))

    for f in (ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper, gather_points_wrapper,
              gather_points_grad_wrapper, furthest_point_sampling_wrapper, knn_wrapper, three_nn_wrapper,
              three_interpolate_wrapper, three_interpolate_grad_wrapper):
        setattr(m, f.__name__, f)
    return m


def emd_module(backend):
    """A stand-in for `_emd_ext._emd` (losses/cuda/emd_torch/pkg/include/emd.h:25-46): backend "ref" = the
    reference's own kernels (oracle/_ref/libemd_ref.so), "l3d" = the product's module."""
    if backend == "l3d":
        import learning3d_b200._emd_ext._emd as product
        return product
    lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libemd_ref.so"))
    m = types.ModuleType("_emd_ext._emd")

    def emd_forward(xyz1, xyz2):
        B, n, _ = xyz1.shape
        mm = xyz2.shape[1]
        match = torch.zeros(B, n, mm, device=xyz1.device)            # emd.cu:18-22 (zeros, callee-allocated)
        temp = torch.zeros(B, 2 * (n + mm), device=xyz1.device)
        cost = torch.zeros(B, device=xyz1.device)
        torch.cuda.current_stream().synchronize()                   # the launchers use the legacy default stream
        lib.ref_emd_forward(B, n, mm, _p(xyz1), _p(xyz2), _p(match), _p(temp), _p(cost))
        torch.cuda.synchronize()
        return [cost, match]

    def emd_backward(xyz1, xyz2, match):
        B, n, _ = xyz1.shape
        mm = xyz2.shape[1]
        g1 = torch.zeros_like(xyz1)
        g2 = torch.zeros_like(xyz2)
        torch.cuda.current_stream().synchronize()
        lib.ref_emd_backward(B, n, mm, _p(xyz1), _p(xyz2), _p(match), _p(g1), _p(g2))
        torch.cuda.synchronize()
        return [g1, g2]
    m.emd_forward, m.emd_backward = emd_forward, emd_backward
    return m


def import_reference(pointnet2_backend="ref"):
    """Import the reference as the package `learning3d` (once per process) and return it.
    `h5py` is stubbed (utils/transformer.py:4 and data_utils import it; neither is used here)."""
    global _pkg
    if _pkg is not None:
        return _pkg
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package not found: run oracle/build_ref.py in the build container")
    tmp = tempfile.mkdtemp(prefix="l3dref_")
    os.symlink(root, os.path.join(tmp, "learning3d"))
    sys.path.insert(0, tmp)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if "pointnet2_cuda" not in sys.modules and os.path.exists(os.path.join(HERE, "_ref", "libpn2_ref.so")):
        sys.modules["pointnet2_cuda"] = pointnet2_module(pointnet2_backend)
    if "_emd_ext" not in sys.modules and os.path.exists(os.path.join(HERE, "_ref", "libemd_ref.so")):
        ext = types.ModuleType("_emd_ext")
        ext._emd = emd_module("ref")
        sys.modules["_emd_ext"] = ext
        sys.modules["_emd_ext._emd"] = ext._emd
    import learning3d  # noqa: F401
    import learning3d.models  # noqa: F401
    import learning3d.utils  # noqa: F401
    _pkg = learning3d
    return _pkg


def set_pointnet2_backend(backend):
    """Swap the extension module the reference's pointnet2_utils calls (it binds `pointnet2` at import)."""
    pkg = import_reference(backend)
    mod = pointnet2_module(backend)
    pkg.utils.lib.pointnet2_utils.pointnet2 = mod
    return mod
