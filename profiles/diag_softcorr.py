"""GPU diagnostic for the tcgen05 soft-correspondence kernel: raw score tile vs fp64 matmul, structured
inputs that expose operand-layout mistakes, then end-to-end src_corr error and timing.
Run: python profiles/diag_softcorr.py  (prints JSON-ish lines; exits non-zero on a pipeline timeout)."""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from learning3d_b200 import _C

DEV = "cuda:0"
lib = _C.lib()


def run(src_emb, tgt_emb, tgt, want_scores=True):
    B, D, Ns = src_emb.shape
    Nt = tgt_emb.shape[2]
    out = torch.full((B, 3, Ns), float("nan"), device=DEV)
    sc = torch.full((B, Ns, Nt), float("nan"), device=DEV) if want_scores else None
    if want_scores:
        rc = lib.l3d_debug_soft_correspondence_scores(_C.ptr(src_emb), _C.ptr(tgt_emb), _C.ptr(tgt), B, D, Ns, Nt,
                                                      _C.ptr(out), _C.ptr(sc), _C.stream())
    else:
        rc = lib.l3d_soft_correspondence(_C.ptr(src_emb), _C.ptr(tgt_emb), _C.ptr(tgt), B, D, Ns, Nt, _C.ptr(out),
                                         _C.stream())
    st = lib.l3d_soft_correspondence_status()
    return rc, st, out, sc


def ref(src_emb, tgt_emb, tgt):
    s = torch.matmul(src_emb.double().transpose(2, 1), tgt_emb.double())
    p = torch.softmax(s / math.sqrt(src_emb.shape[1]), dim=2)
    return s, torch.matmul(tgt.double(), p.transpose(2, 1))


def report(name, src_emb, tgt_emb, tgt):
    rc, st, out, sc = run(src_emb, tgt_emb, tgt)
    s_ref, o_ref = ref(src_emb, tgt_emb, tgt)
    es = (sc.double() - s_ref).abs().max().item()
    eo = (out.double() - o_ref).abs().max().item()
    scale = s_ref.abs().max().item()
    print("%s: rc=%d status=%d  max|S-Sref|=%.3e (|S|max %.3e)  max|corr-ref|=%.3e  nan_S=%d nan_out=%d" % (
        name, rc, st, es, scale, eo, int(torch.isnan(sc).sum()), int(torch.isnan(out).sum())), flush=True)
    return rc, st, es, eo, sc, s_ref


torch.manual_seed(0)
MODE = "tma"
if len(sys.argv) > 1 and sys.argv[1] == "generic":
    MODE = "generic"; lib.l3d_debug_soft_correspondence_force_generic(1)
if len(sys.argv) > 1 and sys.argv[1] == "tma1":
    MODE = "tma1"; lib.l3d_debug_soft_correspondence_force_generic(2)
print("operand pipeline:", MODE)
# E1: one tile, one K block
B, D, N = 1, 32, 128
a = torch.randn(B, D, N, device=DEV); b = torch.randn(B, D, N, device=DEV); t = torch.rand(B, 3, N, device=DEV)
rc, st, es, eo, sc, s_ref = report("E1 rand D32 N128", a, b, t)
if st != 0 or rc != 0:
    print("pipeline failure; stop"); sys.exit(2)
if es > 1e-3:
    print("S[0,:4,:8] =", sc[0, :4, :8].cpu().numpy()); print("ref        =", s_ref[0, :4, :8].float().cpu().numpy())
    if MODE == "tma":
        import numpy as np, ctypes
        buf = np.zeros(12288, dtype=np.float32)
        print("tiles rc", lib.l3d_debug_soft_correspondence_tiles(buf.ctypes.data_as(ctypes.c_void_p)))
        names = ["A_hi", "A_lo", "B_hi", "B_lo"]
        an = a[0].cpu().numpy()   # [D, N]
        offs = [0, 2048, 4096, 8192, 12288]
        for t_i, nm in enumerate(names):
            tl = buf[offs[t_i]:offs[t_i + 1]]
            print(nm, "nonzero", int((tl != 0).sum()), "absmax %.3e" % float(np.abs(tl).max()), "first8", tl[:8])
        # where did a[d=0, n=0..7] land?  expected (unswizzled) atom 0, row d=0, floats 0..7
        tl = buf[:2048]
        for (d, n) in [(0, 0), (0, 4), (1, 0), (1, 4), (8, 0), (0, 32), (0, 33)]:
            pos = np.nonzero(tl == an[d, n])[0]
            print("a[d=%d,n=%d]=%.6f found at float offsets %s" % (d, n, an[d, n], pos[:4]))
    # E2: one-hot channels: src_emb[d,i] = (d == i%32), tgt_emb[d,j] = (d == j%32)*(1+j)
    ii = torch.arange(N, device=DEV)
    a2 = torch.zeros(B, D, N, device=DEV); a2[0, ii % D, ii] = 1.0
    b2 = torch.zeros(B, D, N, device=DEV); b2[0, ii % D, ii] = (1.0 + ii).float()
    rc, st, out, sc2 = run(a2, b2, t)
    for i in (0, 1, 2, 3, 4, 5, 8, 9, 16, 31, 32, 33, 64, 127):
        nz = torch.nonzero(sc2[0, i]).flatten().cpu().tolist()
        print("onehot row %3d (expect j%%32==%2d): nonzero j=%s vals=%s" % (i, i % 32, nz[:12], [round(float(sc2[0, i, j]), 1) for j in nz[:12]]))
    # E3: only channel d0 nonzero in both -> S = outer product of channel rows; tells which K slots are read
    for d0 in (0, 1, 4, 7, 8, 31):
        a3 = torch.zeros(B, D, N, device=DEV); a3[0, d0] = 1.0
        b3 = torch.zeros(B, D, N, device=DEV); b3[0, d0] = 1.0
        rc, st, out, sc3 = run(a3, b3, t)
        print("channel %2d only: S min %.2f max %.2f mean %.3f" % (d0, sc3.min().item(), sc3.max().item(), sc3.mean().item()))
    sys.exit(1)

# larger shapes
for (B, D, Ns, Nt) in [(1, 64, 128, 256), (2, 96, 200, 333), (3, 80, 1000, 516), (2, 512, 1024, 1024)]:
    a = torch.randn(B, D, Ns, device=DEV); b = torch.randn(B, D, Nt, device=DEV); t = torch.rand(B, 3, Nt, device=DEV)
    report("rand B%d D%d Ns%d Nt%d" % (B, D, Ns, Nt), a, b, t)
# peaky softmax (trained-like embeddings)
B, D, N = 2, 512, 1024
a = 3 * torch.randn(B, D, N, device=DEV); b = a + 0.1 * torch.randn(B, D, N, device=DEV); t = torch.rand(B, 3, N, device=DEV)
report("peaky", a, b, t)

# timing at C3
B, D, N = 32, 512, 1024
a = torch.randn(B, D, N, device=DEV); b = torch.randn(B, D, N, device=DEV); t = torch.rand(B, 3, N, device=DEV)
out = torch.empty(B, 3, N, device=DEV)
def ours():
    lib.l3d_soft_correspondence(_C.ptr(a), _C.ptr(b), _C.ptr(t), B, D, N, N, _C.ptr(out), _C.stream())
def theirs():
    s = torch.matmul(a.transpose(2, 1).contiguous(), b) / math.sqrt(D)
    s = torch.softmax(s, dim=2)
    return torch.matmul(t, s.transpose(2, 1).contiguous())
def ours_single():
    lib.l3d_debug_soft_correspondence_force_generic(2); ours(); lib.l3d_debug_soft_correspondence_force_generic(0)
for name, fn in (("fused tcgen05 (auto: CTA pairs)", ours), ("fused tcgen05 (single CTA)", ours_single), ("torch fp32 (reference ops)", theirs)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("C3 B32 D512 N1024 %s: %.1f us  (%.1f TFLOP/s fp32-equivalent)" % (name, ms * 1e3, 2.0 * B * N * N * D / ms / 1e9), flush=True)
print("status", lib.l3d_soft_correspondence_status())
r = theirs(); ours(); torch.cuda.synchronize()
print("C3 max|ours - torch fp32| = %.3e" % (out - r).abs().max().item())
