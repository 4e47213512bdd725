"""oracle/rpm.py (numpy restatement of RPMNet's matching tail) against the fixture the REAL reference produced
(tests/golden/rpm_tail.npz, make_golden.py gen_rpm).  CPU only."""
import numpy as np

from oracle import rpm


def test_square_distance_and_affinity(golden_dir):
    g = np.load(f"{golden_dir}/rpm_tail.npz")
    d = rpm.square_distance(g["feat_src"], g["feat_ref"])
    np.testing.assert_allclose(d, g["dist"], rtol=0, atol=2e-5)
    aff = -g["beta"][:, None, None] * (d - g["alpha"][:, None, None])
    np.testing.assert_allclose(aff, g["affinity"], rtol=0, atol=1e-4)


def test_sinkhorn(golden_dir):
    g = np.load(f"{golden_dir}/rpm_tail.npz")
    np.testing.assert_allclose(rpm.sinkhorn(g["affinity"], 5, True), g["log_perm"], rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(rpm.sinkhorn(g["affinity"], 3, False), g["log_noslack"], rtol=2e-6, atol=2e-5)
    perm, weighted, rs = rpm.match_tail(g["affinity"], g["xyz_ref"], 5, True)
    np.testing.assert_allclose(perm, g["perm"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(weighted, g["weighted"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(rs, g["rowsum"], rtol=1e-4, atol=1e-12)
    # slack semantics: every row / column of the slack-augmented matrix sums to <= 1
    assert (perm.sum(2) <= 1 + 1e-5).all() and (perm.sum(1) <= 1 + 1e-5).all()


def test_rigid_transform(golden_dir):
    g = np.load(f"{golden_dir}/rpm_tail.npz")
    np.testing.assert_allclose(rpm.compute_rigid_transform(g["xyz_src"], g["weighted"], g["rowsum"]), g["T"], atol=2e-5)
    T2 = rpm.compute_rigid_transform(g["a2"], g["b2"], g["w2"])
    np.testing.assert_allclose(T2, g["T2"], atol=2e-5)
    assert (np.linalg.det(T2[:, :, :3]) > 0).all()          # reflections are never returned
