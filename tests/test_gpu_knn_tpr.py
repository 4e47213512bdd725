"""GPU parity of the thread-per-row kNN kernel (learning3d_b200/csrc/knn_tpr.cu) — knn() of
utils/model_common_utils.py:3-9 for large batches: bit-exact against the CPU oracle and against the
warp-per-row-pair kernel of knn.cu on the same inputs, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _hooks():
    from learning3d_b200 import _C
    _C.lib().l3d_debug_force_slow_path(0)
    _C.lib().l3d_debug_knn_path(0)
    yield
    _C.lib().l3d_debug_force_slow_path(0)
    _C.lib().l3d_debug_knn_path(0)


def _knn(x_np, k, path, want_val=False):
    from learning3d_b200 import _C
    B, _, N = x_np.shape
    x = torch.from_numpy(x_np).to(_dev())
    idx = torch.full((B, N, k), -1, dtype=torch.int64, device=x.device)
    val = torch.full((B, N, k), float("nan"), dtype=torch.float32, device=x.device) if want_val else None
    _C.lib().l3d_debug_knn_path(path)
    _C.check(_C.lib().l3d_knn_expansion(_C.ptr(x), B, N, k, _C.ptr(idx), _C.ptr(val) if want_val else None, _C.stream()))
    _C.lib().l3d_debug_knn_path(0)
    torch.cuda.synchronize()
    return (idx.cpu().numpy(), val.cpu().numpy()) if want_val else idx.cpu().numpy()


def _cloud(rng, B, N, dist):
    if dist == "rand":
        return rng.random((B, 3, N), dtype=np.float32)
    x = rng.standard_normal((B, 3, N)).astype(np.float32)
    if dist == "sphere":   # ModelNet-style: centred, unit-sphere normalised
        x -= x.mean(-1, keepdims=True)
        x /= np.sqrt((x ** 2).sum(1, keepdims=True)).max(-1, keepdims=True)
    return x


@pytest.mark.parametrize("B,N,k", [
    (4, 1024, 20),     # the compile-time instantiation (N = 1024, k = 20) on a small batch
    (3, 1024, 16),     # N = 1024, run-time k
    (2, 128, 20),      # smallest cloud: groups of 2 candidates
    (3, 160, 7),       # padded groups (npad = 256), odd k (scalar index stores)
    (1, 2048, 24),     # largest cloud / largest k of this path
    (2, 1056, 20),     # npad = 1152: 9 candidate pairs per group
    (5, 992, 1),       # k = 1
    (9, 512, 24),      # unit ranges crossing cloud boundaries inside a CTA
])
@pytest.mark.parametrize("dist", ["rand", "randn", "sphere"])
def test_tpr_matches_oracle_and_warp_kernel(oracle_mod, B, N, k, dist):
    rng = np.random.default_rng(99 + N + k)
    x = _cloud(rng, B, N, dist)
    got = _knn(x, k, 2)
    assert np.array_equal(got, oracle_mod.knn_expansion(x, k))
    assert np.array_equal(got, _knn(x, k, 1))


def test_tpr_values_bit_exact(oracle_mod):
    rng = np.random.default_rng(5)
    x = rng.random((3, 3, 1024), dtype=np.float32)
    gi, gv = _knn(x, 20, 2, want_val=True)
    oi, ov = oracle_mod.knn_expansion(x, 20, want_val=True)
    assert np.array_equal(gi, oi) and np.array_equal(gv, ov)


def test_tpr_full_size_c2(oracle_mod):
    """BASELINE config C2 (B=32, N=1024, k=20): the automatic dispatch takes the thread-per-row kernel here."""
    torch.manual_seed(4321)
    x = torch.rand(32, 3, 1024).numpy()
    auto = _knn(x, 20, 0)
    assert np.array_equal(auto, oracle_mod.knn_expansion(x, 20, mt=True))
    assert np.array_equal(auto, _knn(x, 20, 1))


def test_tpr_second_block_and_overflow(oracle_mod):
    """More than 32 survivors (every point twice: the second sort-32 + merge), more than 64 (every point 64 times,
    all points identical: the exact k-round scan), and the lowest-index tie rule in both."""
    rng = np.random.default_rng(3)
    half = rng.random((2, 3, 512), dtype=np.float32)
    twice = np.concatenate([half, half], -1)                       # N = 1024, every key appears twice
    assert np.array_equal(_knn(twice, 20, 2), oracle_mod.knn_expansion(twice, 20))
    triple = np.concatenate([half[:, :, :320]] * 3 + [half[:, :, :64]], -1)   # N = 1024, mostly three copies
    assert np.array_equal(_knn(triple, 20, 2), oracle_mod.knn_expansion(triple, 20))
    base = rng.random((1, 3, 16), dtype=np.float32)
    x64 = np.tile(base, (1, 1, 64))
    assert np.array_equal(_knn(x64, 20, 2), oracle_mod.knn_expansion(x64, 20))
    ones = np.ones((2, 3, 256), np.float32)
    assert np.array_equal(_knn(ones, 9, 2), oracle_mod.knn_expansion(ones, 9))
    grid = np.stack(np.meshgrid(*[np.arange(8, dtype=np.float32)] * 3, indexing="ij"), 0).reshape(1, 3, 512)
    assert np.array_equal(_knn(grid, 24, 2), oracle_mod.knn_expansion(grid, 24))


def test_tpr_forced_slow_path(oracle_mod):
    from learning3d_b200 import _C
    rng = np.random.default_rng(12)
    x = rng.random((2, 3, 256), dtype=np.float32)
    _C.lib().l3d_debug_force_slow_path(1)
    slow = _knn(x, 20, 2)
    _C.lib().l3d_debug_force_slow_path(0)
    assert np.array_equal(slow, oracle_mod.knn_expansion(x, 20))


def test_tpr_fused_graph_feature(oracle_mod):
    """get_graph_feature() in the same launch (model_common_utils.py:132-155) on the thread-per-row kernel."""
    from learning3d_b200 import _C
    rng = np.random.default_rng(21)
    for (B, N, k) in [(3, 1024, 20), (2, 256, 12)]:
        x_np = rng.random((B, 3, N), dtype=np.float32)
        x = torch.from_numpy(x_np).to(_dev())
        idx = torch.empty((B, N, k), dtype=torch.int64, device=x.device)
        feat = torch.full((B, 6, N, k), float("nan"), dtype=torch.float32, device=x.device)
        _C.lib().l3d_debug_knn_path(2)
        _C.check(_C.lib().l3d_knn_graph_feature(_C.ptr(x), B, N, k, _C.ptr(idx), _C.ptr(feat), _C.stream()))
        _C.lib().l3d_debug_knn_path(0)
        torch.cuda.synchronize()
        want = oracle_mod.knn_expansion(x_np, k)
        assert np.array_equal(idx.cpu().numpy(), want)
        assert np.array_equal(feat.cpu().numpy(), oracle_mod.graph_feature(x_np, want))


def test_tpr_host_entry_point_c2(oracle_mod):
    """l3d_knn_expansion_host at C2 (uint16 indices on the wire, widened on the host) rides the same kernel."""
    from learning3d_b200 import _C
    import ctypes
    rng = np.random.default_rng(8)
    x = rng.random((32, 3, 1024), dtype=np.float32)
    out = np.empty((32, 1024, 20), np.int64)
    _C.check(_C.lib().l3d_knn_expansion_host(x.ctypes.data_as(ctypes.c_void_p), 32, 1024, 20,
                                             out.ctypes.data_as(ctypes.c_void_p)))
    assert np.array_equal(out, oracle_mod.knn_expansion(x, 20, mt=True))


@pytest.mark.parametrize("pinned", [True, False])
def test_host_entry_point_repeated_calls_same_buffers(oracle_mod, pinned):
    """Repeated l3d_knn_expansion_host calls on the same host buffers (pinned and pageable): new data in the same buffers
    must give new results (the widening workers and the staging buffers are reused from call to call)."""
    from learning3d_b200 import _C
    import ctypes
    rng = np.random.default_rng(17)
    B, N, k = 32, 1024, 20
    if pinned:
        xt = torch.empty(B, 3, N).pin_memory()
        ot = torch.empty(B, N, k, dtype=torch.int64).pin_memory()
        x, out = xt.numpy(), ot.numpy()
    else:
        x, out = np.empty((B, 3, N), np.float32), np.empty((B, N, k), np.int64)
    for rep in range(4):
        x[...] = rng.random((B, 3, N), dtype=np.float32)
        out[...] = -1
        _C.check(_C.lib().l3d_knn_expansion_host(x.ctypes.data_as(ctypes.c_void_p), B, N, k,
                                                 out.ctypes.data_as(ctypes.c_void_p)))
        assert np.array_equal(out, oracle_mod.knn_expansion(x, k, mt=True)), "call %d" % rep


@pytest.mark.parametrize("B,N,k", [
    (40, 1024, 20),    # 640 duo units on 148 CTAs: some duos run two units (their shared arrays are reused)
    (600, 128, 20),    # 1200 units of 64 rows, 2 per cloud: a CTA's range spans several rounds of resident clouds
    (2, 256, 7),       # duo kernel, odd k (scalar index stores)
    (3, 1088, 24),     # N % 64 == 0 but not a multiple of 128: padded groups in both candidate halves
])
def test_tpr_many_units_and_rounds(oracle_mod, B, N, k):
    rng = np.random.default_rng(B * 7 + N)
    x = rng.random((B, 3, N), dtype=np.float32)
    assert np.array_equal(_knn(x, k, 2), oracle_mod.knn_expansion(x, k, mt=True))
