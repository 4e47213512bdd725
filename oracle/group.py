"""ctypes wrappers of oracle/l3d_oracle_group.c (grouping family).  TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import lib, _f32, _i64, _i32, _out


def _opt_i64(a):
    if a is None:
        return None, None
    return _i64(a)


def pn2_ball_query(radius, nsample, xyz, new_xyz):
    xyz, xp = _f32(xyz)
    new_xyz, qp = _f32(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx, ip = _out((b, m, nsample), np.int32)
    lib().l3d_oracle_pn2_ball_query(b, n, m, ctypes.c_float(radius), nsample, qp, xp, ip)
    return idx


def query_ball_point(radius, nsample, xyz, new_xyz, itself=None, want_cnt=False):
    xyz, xp = _f32(xyz)
    new_xyz, qp = _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    it, itp = _opt_i64(itself)
    out, op = _out((B, S, nsample), np.int64)
    cnt, cp = _out((B, S), np.int64) if want_cnt else (None, None)
    r2 = np.float32(radius ** 2)
    lib().l3d_oracle_query_ball_point(xp, qp, B, N, S, ctypes.c_float(float(r2)), nsample, itp, op, cp)
    return (out, cnt) if want_cnt else out


def pn2_group_points(points, idx):
    """points [b,c,n], idx [b,npoints,nsample] (or [b,npoints]) int32."""
    points, pp = _f32(points)
    idx, ip = _i32(idx)
    b, c, n = points.shape
    P = int(np.prod(idx.shape[1:]))
    out, op = _out((b, c) + idx.shape[1:], np.float32)
    lib().l3d_oracle_pn2_group_points(b, c, n, ctypes.c_long(P), pp, ip, op)
    return out


def pn2_group_points_grad(grad_out, idx, n):
    grad_out, gp = _f32(grad_out)
    idx, ip = _i32(idx)
    b, c = grad_out.shape[:2]
    P = int(np.prod(idx.shape[1:]))
    out, op = _out((b, c, n), np.float32)
    lib().l3d_oracle_pn2_group_points_grad(b, c, n, ctypes.c_long(P), gp, ip, op)
    return out


def pn2_three_interpolate(points, idx, weight):
    points, pp = _f32(points)
    idx, ip = _i32(idx)
    weight, wp = _f32(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out, op = _out((b, c, n), np.float32)
    lib().l3d_oracle_pn2_three_interpolate(b, c, m, n, pp, ip, wp, op)
    return out


def pn2_three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, gp = _f32(grad_out)
    idx, ip = _i32(idx)
    weight, wp = _f32(weight)
    b, c, n = grad_out.shape
    out, op = _out((b, c, m), np.float32)
    lib().l3d_oracle_pn2_three_interpolate_grad(b, c, n, m, gp, ip, wp, op)
    return out


def pn2_fps(xyz, m):
    xyz, xp = _f32(xyz)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, np.float32)
    idx, ip = _out((b, m), np.int32)
    lib().l3d_oracle_pn2_fps(b, n, m, xp, temp.ctypes.data_as(ctypes.c_void_p), ip)
    return idx, temp


def farthest_point_sample(xyz, npoint, start=None):
    xyz, xp = _f32(xyz)
    B, N, _ = xyz.shape
    st, sp = _opt_i64(start)
    out, op = _out((B, npoint), np.int64)
    lib().l3d_oracle_farthest_point_sample(xp, B, N, npoint, sp, op)
    return out


def index_points(points, idx):
    points, pp = _f32(points)
    idx, ip = _i64(idx)
    B, N, C = points.shape
    R = int(np.prod(idx.shape[1:]))
    out, op = _out(idx.shape + (C,), np.float32)
    lib().l3d_oracle_index_points(pp, ip, B, N, ctypes.c_long(R), C, op)
    return out


def compute_density(xyz, bandwidth):
    xyz, xp = _f32(xyz)
    B, N, _ = xyz.shape
    out, op = _out((B, N), np.float32)
    two_bw2 = float(np.float32(2.0 * bandwidth * bandwidth))
    norm = float(np.float32(2.5 * bandwidth))
    lib().l3d_oracle_compute_density(xp, B, N, ctypes.c_float(two_bw2), ctypes.c_float(norm), op)
    return out


def ref_pn2():
    """ctypes handle on oracle/_ref/libpn2_ref.so — the reference's own pointnet2 CUDA kernels
    (compiled from /root/reference by oracle/build_ref.py) behind the extern "C" shim in
    oracle/ref_shims/pn2_shim.cu.  None when not built.  GPU only."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpn2_ref.so")
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)
