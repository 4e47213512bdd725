"""ctypes wrappers of oracle/l3d_oracle_emd.c and a numpy restatement of the SVD head.
TEST INFRASTRUCTURE ONLY."""
import numpy as np

from . import lib, _f32, _out


def approxmatch(xyz1, xyz2):
    """-> match [B, n, m] as the reference allocates it (emd.cu:18), memory index l*n + k."""
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match, mp = _out((b, n, m), np.float32)
    lib().l3d_oracle_emd_approxmatch(b, n, m, p1, p2, mp)
    return match


def matchcost(xyz1, xyz2, match):
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    match, mp = _f32(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost, cp = _out((b,), np.float32)
    lib().l3d_oracle_emd_matchcost(b, n, m, p1, p2, mp, cp)
    return cost


def grads(xyz1, xyz2, match):
    xyz1, p1 = _f32(xyz1)
    xyz2, p2 = _f32(xyz2)
    match, mp = _f32(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1, gp1 = _out((b, n, 3), np.float32)
    g2, gp2 = _out((b, m, 3), np.float32)
    lib().l3d_oracle_emd_grads(b, n, m, p1, p2, mp, gp1, gp2)
    return g1, g2


def emd_forward(xyz1, xyz2):
    match = approxmatch(xyz1, xyz2)
    return matchcost(xyz1, xyz2, match), match


def soft_correspondence(src_emb, tgt_emb, tgt):
    """utils/svd.py:23-28 restated with numpy in fp64 (checker for the fused tcgen05 kernel):
    src_emb [B,D,Ns], tgt_emb [B,D,Nt], tgt [B,3,Nt] -> src_corr [B,3,Ns] (returned as fp64).
    Pinned against the real reference's fp32 output in tests/golden/svd_head.npz (tests/test_oracle_emd_svd.py)."""
    a = np.asarray(src_emb, np.float64)
    b = np.asarray(tgt_emb, np.float64)
    scores = np.matmul(a.transpose(0, 2, 1), b) / np.sqrt(a.shape[1])          # :24
    scores = scores - scores.max(axis=2, keepdims=True)
    p = np.exp(scores)
    p /= p.sum(axis=2, keepdims=True)                                             # :25 softmax(dim=2)
    return np.matmul(np.asarray(tgt, np.float64), p.transpose(0, 2, 1))           # :27


def svd_head_tail(src, src_corr):
    """utils/svd.py:29-58 restated with numpy (LAPACK gesdd, the routine torch.svd calls on CPU):
    src, src_corr [B,3,N] fp32 -> R [B,3,3], t [B,3]."""
    src = np.asarray(src, np.float32)
    src_corr = np.asarray(src_corr, np.float32)
    B = src.shape[0]
    reflect = np.eye(3, dtype=np.float32)
    reflect[2, 2] = -1
    src_centered = src - src.mean(axis=2, keepdims=True)
    corr_centered = src_corr - src_corr.mean(axis=2, keepdims=True)
    H = np.matmul(src_centered, corr_centered.transpose(0, 2, 1))
    R = np.empty((B, 3, 3), np.float32)
    for i in range(B):
        u, s, vt = np.linalg.svd(H[i])
        v = vt.T
        r = v @ u.T
        if np.linalg.det(r) < 0:
            v = v @ reflect
            r = v @ u.T
        R[i] = r
    t = np.matmul(-R, src.mean(axis=2, keepdims=True)) + src_corr.mean(axis=2, keepdims=True)
    return R, t.reshape(B, 3)


def ref_emd():
    """ctypes handle on oracle/_ref/libemd_ref.so — the reference's own EMD CUDA kernels (emd.cuh) behind
    oracle/ref_shims/emd_shim.cu.  None when not built.  GPU only; import torch first (libtorch symbols)."""
    import ctypes
    import os
    import torch  # noqa: F401
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libemd_ref.so")
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)
