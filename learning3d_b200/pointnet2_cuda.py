"""Drop-in for the reference's `pointnet2_cuda` extension module (utils/lib/src/pointnet2_api.cpp:10-25).

The reference's utils/lib/pointnet2_utils.py:8 does `import pointnet2_cuda as pointnet2` and calls ten
`*_wrapper` functions with ints + caller-allocated CUDA tensors.  That extension needs THC and no longer
builds; this module has the same ten names and argument orders and forwards each to the C ABI
(`l3d_pn2_*`, include/l3d_b200.h) on the caller's current stream, so the reference file runs unmodified:

    import sys, learning3d_b200.pointnet2_cuda
    sys.modules["pointnet2_cuda"] = learning3d_b200.pointnet2_cuda      # before importing learning3d

Errors raise RuntimeError (the reference extension prints and calls exit(-1)).
"""
import ctypes

import torch

from . import _C


def _run(symbol, anchor, *args):
    with _C.on_device(anchor.device):
        _C.check(getattr(_C.lib(), symbol)(*args, _C.stream()), symbol)


def _f(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise RuntimeError("pointnet2_cuda.%s: expected a contiguous CUDA float32 tensor" % name)
    return _C.ptr(t)


def _i(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == torch.int32):
        raise RuntimeError("pointnet2_cuda.%s: expected a contiguous CUDA int32 tensor" % name)
    return _C.ptr(t)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _run("l3d_pn2_ball_query", xyz, b, n, m, ctypes.c_float(radius), nsample, _f(new_xyz, "new_xyz"), _f(xyz, "xyz"),
         _i(idx, "idx"))


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _run("l3d_pn2_group_points", points, b, c, n, npoints, nsample, _f(points, "points"), _i(idx, "idx"), _f(out, "out"))


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _run("l3d_pn2_group_points_grad", grad_out, b, c, n, npoints, nsample, _f(grad_out, "grad_out"), _i(idx, "idx"),
         _f(grad_points, "grad_points"))


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _run("l3d_pn2_gather_points", points, b, c, n, npoints, _f(points, "points"), _i(idx, "idx"), _f(out, "out"))


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _run("l3d_pn2_gather_points_grad", grad_out, b, c, n, npoints, _f(grad_out, "grad_out"), _i(idx, "idx"),
         _f(grad_points, "grad_points"))


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _run("l3d_pn2_furthest_point_sampling", points, b, n, m, _f(points, "points"), _f(temp, "temp"), _i(idx, "idx"))


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    _run("l3d_pn2_knn", unknown, b, n, m, k, _f(unknown, "unknown"), _f(known, "known"), _f(dist2, "dist2"),
         _i(idx, "idx"))


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _run("l3d_pn2_three_nn", unknown, b, n, m, _f(unknown, "unknown"), _f(known, "known"), _f(dist2, "dist2"),
         _i(idx, "idx"))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _run("l3d_pn2_three_interpolate", points, b, c, m, n, _f(points, "points"), _i(idx, "idx"), _f(weight, "weight"),
         _f(out, "out"))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _run("l3d_pn2_three_interpolate_grad", grad_out, b, c, n, m, _f(grad_out, "grad_out"), _i(idx, "idx"),
         _f(weight, "weight"), _f(grad_points, "grad_points"))
