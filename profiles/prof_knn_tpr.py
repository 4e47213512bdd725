"""One C2 knn() launch per kernel flavour for ncu (profiles/r02 capture of knn_tpr_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from learning3d_b200 import _C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
path = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.rand(B, 3, 1024, device="cuda:0")
idx = torch.empty(B, 1024, 20, dtype=torch.int64, device="cuda:0")
L = _C.lib()
L.l3d_debug_knn_path(path)
for _ in range(3):
    _C.check(L.l3d_knn_expansion(_C.ptr(x), B, 1024, 20, _C.ptr(idx), None, _C.stream()))
torch.cuda.synchronize()
