"""Error magnitudes of the CUDA EMD against the sequential oracle (run on the GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning3d_b200 import _C
from oracle import emd as oe
lib = _C.lib()
for (B, n, m) in [(2, 256, 256), (2, 1024, 1024), (1, 1500, 1500), (8, 1024, 1024)]:
    rng = np.random.default_rng(n + m)
    a = rng.random((B, n, 3), dtype=np.float32); b = rng.random((B, m, 3), dtype=np.float32)
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    cost = torch.empty(B, device="cuda"); match = torch.empty(B, n, m, device="cuda")
    ws = torch.empty(int(lib.l3d_emd_forward_ws_bytes(B, n, m)), dtype=torch.uint8, device="cuda")
    _C.check(lib.l3d_emd_forward(_C.ptr(ad), _C.ptr(bd), B, n, m, _C.ptr(cost), _C.ptr(match), _C.ptr(ws), _C.stream()))
    torch.cuda.synchronize()
    oc, om = oe.emd_forward(a, b)
    c = cost.cpu().numpy(); mm = match.cpu().numpy()
    print(B, n, m, "cost rel", np.abs(c - oc).max() / np.abs(oc).max(), "match abs", np.abs(mm - om).max(),
          "match mass", mm.reshape(B, m, n).sum(1).min(), mm.reshape(B, m, n).sum(1).max(), flush=True)
    g1 = torch.empty_like(ad); g2 = torch.empty_like(bd)
    ws2 = torch.empty(int(lib.l3d_emd_backward_ws_bytes(B, n, m)), dtype=torch.uint8, device="cuda")
    md = torch.from_numpy(om).cuda()
    _C.check(lib.l3d_emd_backward(_C.ptr(ad), _C.ptr(bd), _C.ptr(md), B, n, m, _C.ptr(g1), _C.ptr(g2), _C.ptr(ws2), _C.stream()))
    torch.cuda.synchronize()
    og1, og2 = oe.grads(a, b, om)
    print("   grads (same match) abs", np.abs(g1.cpu().numpy() - og1).max(), np.abs(g2.cpu().numpy() - og2).max(),
          "scale", np.abs(og1).max(), flush=True)
