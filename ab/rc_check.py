import sys, torch
sys.path.insert(0, ".")
from learning3d_b200 import _C
lib = _C.lib()
B, K, M, P, k, N = 2, 64, 64, 20480, 20, 1024
wt = torch.randn(K, M, device="cuda"); x = torch.randn(B, K, P, device="cuda")
sc = torch.ones(M, device="cuda"); sh = torch.zeros(M, device="cuda")
h = torch.empty(B, M, P, device="cuda"); cat = torch.empty(B, 512, N, device="cuda")
rc = lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(sc), _C.ptr(sh), B, M, K, P, k, 1, _C.ptr(h), _C.ptr(cat), 512 * N, 0, _C.stream())
print("rc", rc, lib.l3d_error_string(rc))
