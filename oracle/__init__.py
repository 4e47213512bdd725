"""CPU oracle for the learning3d hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``learning3d_b200``) never
does and fails loudly when its CUDA library is missing.

Two layers:
  * ``liboracle.so`` (``l3d_oracle.c``): plain-C restatement with the fp32 rounding sequence
    spelled out; the parity checker.  Build with ``make -C oracle``.
  * ``ref_torch``: line-by-line torch-CPU restatements of the reference's pure-PyTorch functions
    (same library calls, hence the same arithmetic as the reference's own CPU path); used to
    time the reference arm and to cross-check the C restatement.

Pinning status: see oracle/README.md.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
        _LIB.l3d_oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, dtype):
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().l3d_oracle_num_threads())


def knn_expansion(x, k, want_val=False, mt=False):
    """x [B,3,N] -> idx [B,N,k] int64 (and pd values)."""
    x, xp = _f32(x)
    B, C, N = x.shape
    assert C == 3
    idx, ip = _out((B, N, k), np.int64)
    if mt and not want_val:
        lib().l3d_oracle_knn_expansion_mt(xp, B, N, k, ip)
        return idx
    if want_val:
        val, vp = _out((B, N, k), np.float32)
    else:
        val, vp = None, None
    lib().l3d_oracle_knn_expansion(xp, B, N, k, ip, vp)
    return (idx, val) if want_val else idx


def graph_feature(x, idx):
    x, xp = _f32(x)
    idx, ip = _i64(idx)
    B, C, N = x.shape
    k = idx.shape[-1]
    out, op = _out((B, 2 * C, N, k), np.float32)
    lib().l3d_oracle_graph_feature(xp, ip, B, C, N, k, op)
    return out


def graph_feature_grad(go, idx, C):
    go, gp = _f32(go)
    idx, ip = _i64(idx)
    B, N, k = idx.shape
    gx, xp = _out((B, C, N), np.float32)
    lib().l3d_oracle_graph_feature_grad(gp, ip, B, C, N, k, xp)
    return gx


def square_distance(src, dst):
    src, sp = _f32(src)
    dst, dp = _f32(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out, op = _out((B, N, M), np.float32)
    lib().l3d_oracle_square_distance(sp, dp, B, N, M, op)
    return out


def knn_sqdist(xyz, new_xyz, nsample):
    xyz, xp = _f32(xyz)
    new_xyz, np_ = _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx, ip = _out((B, S, nsample), np.int64)
    lib().l3d_oracle_knn_sqdist(xp, np_, B, N, S, nsample, ip)
    return idx


def knn_point(k, data, query):
    data, dp = _f32(data)
    query, qp = _f32(query)
    B, N, _ = data.shape
    M = query.shape[1]
    val, vp = _out((B, M, k), np.float32)
    idx, ip = _out((B, M, k), np.int64)
    lib().l3d_oracle_knn_point(dp, qp, B, N, M, k, vp, ip)
    return val, idx


def pn2_knn(k, unknown, known):
    unknown, up = _f32(unknown)
    known, kp = _f32(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2, dp = _out((b, n, k), np.float32)
    idx, ip = _out((b, n, k), np.int32)
    lib().l3d_oracle_pn2_knn(b, n, m, k, up, kp, dp, ip)
    return d2, idx


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def chamfer_forward(xyz1, xyz2):
    """cd.forward semantics: (dist1 [B,n], dist2 [B,m], idx1, idx2 int32)."""
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1, dp1 = _out((b, n), np.float32)
    d2, dp2 = _out((b, m), np.float32)
    i1, ip1 = _out((b, n), np.int32)
    i2, ip2 = _out((b, m), np.int32)
    lib().l3d_oracle_chamfer_forward(p1, p2, b, n, m, dp1, dp2, ip1, ip2)
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, gd1, gd2, idx1, idx2):
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    gd1, g1 = _f32(gd1)
    gd2, g2 = _f32(gd2)
    idx1, i1 = _i32(idx1)
    idx2, i2 = _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1, o1 = _out((b, n, 3), np.float32)
    gx2, o2 = _out((b, m, 3), np.float32)
    lib().l3d_oracle_chamfer_backward(p1, p2, b, n, m, g1, g2, i1, i2, o1, o2)
    return gx1, gx2


def chamfer_loss(xyz1, xyz2):
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    f = lib().l3d_oracle_chamfer_loss
    f.restype = ctypes.c_double
    return float(f(p1, p2, b, n, m))


def chamfer_loss_grads(xyz1, xyz2, grad_loss=1.0):
    """Gradients of the loss above w.r.t. both clouds, float32 chain rule as torch applies it:
    d/d dist = ((g/2)/numel) / (2*sqrt(dist)), then the native backward."""
    d1, d2, i1, i2 = chamfer_forward(xyz1, xyz2)
    g = np.float32(grad_loss) * np.float32(0.5)
    with np.errstate(divide="ignore", invalid="ignore"):
        gd1 = (g / np.float32(d1.size)) / (np.float32(2.0) * np.sqrt(d1))
        gd2 = (g / np.float32(d2.size)) / (np.float32(2.0) * np.sqrt(d2))
    return chamfer_backward(xyz1, xyz2, gd1.astype(np.float32), gd2.astype(np.float32), i1, i2)


def ref_cd():
    """The reference's own Chamfer extension compiled from /root/reference by oracle/build_ref.py
    (module with forward/backward [CPU nnsearch] and forward_cuda/backward_cuda).  None when
    oracle/_ref/cd_ref.so has not been built."""
    path = os.path.join(_HERE, "_ref", "cd_ref.so")
    if not os.path.exists(path):
        return None
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    loader = importlib.machinery.ExtensionFileLoader("cd_ref", path)
    spec = importlib.util.spec_from_loader("cd_ref", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod
