"""CPU restatement (numpy, float64 inside) of RPMNet's matching tail — TEST INFRASTRUCTURE ONLY.

Follows models/rpmnet.py of the reference: sinkhorn :157-218 (eps <= 0 branch), RPMNet.spam's tail :283-287,
compute_rigid_transform :221-254, and utils/ppfnet_util.py:29-48 square_distance.  Pinned by
tests/golden/rpm_tail.npz, which tests/golden/make_golden.py generates by running the REAL reference functions.
"""
import numpy as np

_EPS = 1e-5          # rpmnet.py:11


def square_distance(src, dst):
    """ppfnet_util.py:45-47 — src [B,N,C], dst [B,M,C] -> [B,N,M]."""
    src = src.astype(np.float64); dst = dst.astype(np.float64)
    d = -2.0 * np.matmul(src, dst.transpose(0, 2, 1))
    d += (src ** 2).sum(-1)[:, :, None]
    d += (dst ** 2).sum(-1)[:, None, :]
    return d


def _lse(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))


def sinkhorn(log_alpha, n_iters=5, slack=True):
    """rpmnet.py:177-218 (no early exit): returns log(perm) [B,J,K]."""
    la = log_alpha.astype(np.float64)
    if slack:
        B, J, K = la.shape
        lap = np.zeros((B, J + 1, K + 1))                                 # ZeroPad2d((0,1,0,1)), :181-184
        lap[:, :J, :K] = la
        for _ in range(n_iters):
            lap[:, :-1, :] = lap[:, :-1, :] - _lse(lap[:, :-1, :], 2)     # rows, last row untouched (:187-191)
            lap[:, :, :-1] = lap[:, :, :-1] - _lse(lap[:, :, :-1], 1)     # columns, last column untouched (:193-197)
        return lap[:, :-1, :-1]
    for _ in range(n_iters):
        la = la - _lse(la, 2)
        la = la - _lse(la, 1)
    return la


def match_tail(affinity, xyz_ref, n_iters=5, slack=True):
    """rpmnet.py:283-287: perm, weighted template, row sums."""
    perm = np.exp(sinkhorn(affinity, n_iters, slack))
    rs = perm.sum(2, keepdims=True)
    return perm, perm @ xyz_ref.astype(np.float64) / (rs + _EPS), rs[..., 0]


def compute_rigid_transform(a, b, weights):
    """rpmnet.py:221-254 -> T [B,3,4]."""
    a = a.astype(np.float64); b = b.astype(np.float64); w = weights.astype(np.float64)
    wn = w[..., None] / (w[..., None].sum(1, keepdims=True) + _EPS)
    ca, cb = (a * wn).sum(1), (b * wn).sum(1)
    cov = (a - ca[:, None, :]).transpose(0, 2, 1) @ ((b - cb[:, None, :]) * wn)
    u, s, vt = np.linalg.svd(cov)
    v = vt.transpose(0, 2, 1)
    pos = v @ u.transpose(0, 2, 1)
    vn = v.copy(); vn[:, :, 2] *= -1
    neg = vn @ u.transpose(0, 2, 1)
    rot = np.where(np.linalg.det(pos)[:, None, None] > 0, pos, neg)
    t = -rot @ ca[:, :, None] + cb[:, :, None]
    return np.concatenate((rot, t), axis=2)
