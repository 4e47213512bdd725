"""CPU: the C-ABI library loads and exports every symbol include/l3d_b200.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "l3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(l3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from learning3d_b200 import _C
    assert os.path.exists(_C.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_C.LIB_PATH)
    declared = header_symbols()
    assert declared, "no symbols parsed from the header"
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, "declared in include/l3d_b200.h but not exported: %s" % missing
    # the ctypes signature table covers exactly the header
    assert sorted(_C.exported_symbols()) == declared


def test_version_and_error_strings_without_gpu():
    from learning3d_b200 import _C
    lib = _C.lib()
    assert lib.l3d_abi_version() >= 1
    assert lib.l3d_error_string(0) == b"ok"
    assert b"invalid" in lib.l3d_error_string(-1)
    assert b"not supported" in lib.l3d_error_string(-2)
    assert isinstance(_C.launch_count(), int)


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    from learning3d_b200.utils import knn
    with pytest.raises(RuntimeError, match="CUDA"):
        knn(torch.rand(1, 3, 32), 4)


def test_new_entry_points_validate_arguments_without_gpu():
    """Argument checks of the tensor-core entry points run before any CUDA call: exercised here on CPU."""
    from learning3d_b200 import _C
    import pytest
    import torch
    lib = _C.lib()
    null = _C.ptr(None)
    assert lib.l3d_soft_correspondence(null, null, null, 0, 8, 4, 4, null, null) == 0        # B == 0: nothing to do
    assert lib.l3d_soft_correspondence(null, null, null, 1, 8, 4, 4, null, null) == -1       # null pointers
    assert lib.l3d_soft_correspondence(null, null, null, -1, 8, 4, 4, null, null) == -1
    assert lib.l3d_knn_features(null, 0, 64, 16, 4, null, null, null) == 0
    assert lib.l3d_knn_features(null, 1, 64, 16, 4, null, null, null) == -1
    assert lib.l3d_knn_features(null, 1, 0, 16, 4, null, null, null) == -1                     # C < 1
    assert lib.l3d_knn_features_ws_bytes(2, 64, 100) >= 2 * 100 * 100 * 4 + 2 * 100 * 4
    assert lib.l3d_knn_features_ws_bytes(0, 64, 100) == 0
    assert lib.l3d_knn_graph_feature(null, 1, 16, 4, null, null, null) == -1                   # feat_dev missing
    from learning3d_b200.utils import get_graph_feature
    from learning3d_b200.utils.svd import soft_correspondence, SVDHead
    with pytest.raises(RuntimeError, match="CUDA"):
        soft_correspondence(torch.rand(1, 8, 4), torch.rand(1, 8, 4), torch.rand(1, 3, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        get_graph_feature(torch.rand(1, 64, 32), 4)                                           # feature-space graph
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            SVDHead(8, input_shape="bnc")(torch.rand(1, 8, 4), torch.rand(1, 8, 4), torch.rand(1, 4, 3), torch.rand(1, 4, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "learning3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
