"""CPU: EMD restatement properties (the reference EMD is CUDA-only and does not compile -> parity
unpinned; these are the algorithm's invariants) and the SVD-head restatement vs the reference fixture."""
import numpy as np

from oracle import emd as oe


def test_approxmatch_is_a_transport_plan(oracle_mod):
    rng = np.random.default_rng(0)
    a = rng.random((2, 200, 3), dtype=np.float32)
    b = rng.random((2, 200, 3), dtype=np.float32)
    cost, match = oe.emd_forward(a, b)
    m = match.reshape(2, 200, 200)                 # memory order [l (xyz2), k (xyz1)]
    assert (m >= 0).all()
    np.testing.assert_allclose(m.sum(1), 1.0, atol=2e-5)     # every xyz1 point ships its unit mass
    np.testing.assert_allclose(m.sum(2), 1.0, atol=2e-3)     # every xyz2 point receives ~ unit mass
    # identical clouds: the plan concentrates on the diagonal and the cost is ~0 per point
    c0, m0 = oe.emd_forward(a, a.copy())
    assert (c0 / 200 < 2e-3).all() and (np.diagonal(m0.reshape(2, 200, 200), axis1=1, axis2=2) > 0.5).all()
    # cost is invariant to a common translation and scales linearly with a common scaling
    c1, _ = oe.emd_forward(a + 3.0, b + 3.0)
    np.testing.assert_allclose(c1, cost, rtol=2e-3)


def test_emd_unequal_sizes_use_integer_multiplicity(oracle_mod):
    rng = np.random.default_rng(1)
    a = rng.random((1, 64, 3), dtype=np.float32)
    b = rng.random((1, 128, 3), dtype=np.float32)
    _, match = oe.emd_forward(a, b)               # n < m: multiL = m / n = 2 (emd.cuh:10-16)
    m = match.reshape(1, 128, 64)
    np.testing.assert_allclose(m.sum(1), 2.0, atol=1e-4)


def test_emd_grads_match_finite_differences_of_cost_at_fixed_match(oracle_mod):
    rng = np.random.default_rng(2)
    a = rng.random((1, 40, 3)).astype(np.float32)
    b = rng.random((1, 40, 3)).astype(np.float32)
    _, match = oe.emd_forward(a, b)
    g1, g2 = oe.grads(a, b, match)
    eps = 1e-3
    for (pt, c) in [(3, 0), (17, 2)]:
        ap = a.copy(); ap[0, pt, c] += eps
        am = a.copy(); am[0, pt, c] -= eps
        fd = (oe.matchcost(ap, b, match) - oe.matchcost(am, b, match)) / (2 * eps)
        assert abs(fd[0] - g1[0, pt, c]) < 2e-2 * max(1.0, abs(fd[0]))


def test_soft_correspondence_restatement_matches_reference(golden_dir):
    """oracle.emd.soft_correspondence (fp64) vs src_corr produced by the REAL reference SVDHead (fp32, CPU)."""
    g = np.load(f"{golden_dir}/svd_head.npz")
    got = oe.soft_correspondence(g["src_emb"], g["tgt_emb"], g["tgt"].transpose(0, 2, 1))
    assert got.shape == g["src_corr"].shape
    assert np.abs(got - g["src_corr"]).max() < 2e-6


def test_svd_head_restatement_matches_reference(oracle_mod, golden_dir):
    g = np.load(f"{golden_dir}/svd_head.npz")
    R, t = oe.svd_head_tail(g["src"].transpose(0, 2, 1), g["src_corr"])
    np.testing.assert_allclose(R, g["R"], atol=1e-5)
    np.testing.assert_allclose(t, g["t"], atol=1e-5)
    assert np.allclose(np.linalg.det(g["R"]), 1.0, atol=1e-5)     # includes the two mirrored items
