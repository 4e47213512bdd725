import sys, math, torch
sys.path.insert(0, ".")
from learning3d_b200 import _C
lib = _C.lib(); DEV = "cuda:0"
B, D, N = 32, 512, 1024
a = torch.randn(B, D, N, device=DEV); b = torch.randn(B, D, N, device=DEV); t = torch.rand(B, 3, N, device=DEV)
out = torch.empty(B, 3, N, device=DEV)
for _ in range(3):
    lib.l3d_soft_correspondence(_C.ptr(a), _C.ptr(b), _C.ptr(t), B, D, N, N, _C.ptr(out), _C.stream())
torch.cuda.synchronize()
