"""A/B: EdgeConv layers 2-4 with the current library (3-D boxes forced / chunked 4-D maps) vs older builds, same
process; bit-equality of the two tensor-map flavours; attention passes in both flavours."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from learning3d_b200 import _C
cur = _C.lib()
P_ = ctypes.c_void_p
olds = {}
import os
for name in ("old", "083d85d", "d0c8e7c"):          # older builds, if present (make them with `git archive <rev>`)
    if not os.path.exists("ab/libl3d_%s.so" % name):
        continue
    lib = ctypes.CDLL("ab/libl3d_%s.so" % name)
    lib.l3d_conv1x1_bn_relu_maxk.restype = ctypes.c_int
    lib.l3d_conv1x1_bn_relu_maxk.argtypes = [P_, P_, P_, P_, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, P_, P_, ctypes.c_longlong, ctypes.c_int, P_]
    olds[name] = lib
dev = "cuda"
B, N, k = 32, 1024, 20
P = N * k
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
def mode(m): cur.l3d_debug_soft_correspondence_force_generic(m)
st = _C.stream()
torch.manual_seed(0)
for (K, M, hout) in ((64, 64, True), (64, 128, True), (128, 256, False)):
    wt = torch.randn(K, M, device=dev) * 0.1
    x = torch.randn(B, K, P, device=dev)
    sc = torch.rand(M, device=dev) + 0.5; sh = torch.randn(M, device=dev) * 0.1
    outs = {}
    def run(lib, h, cat):
        return lib.l3d_conv1x1_bn_relu_maxk(_C.ptr(wt), _C.ptr(x), _C.ptr(sc), _C.ptr(sh), B, M, K, P, k, 1, _C.ptr(h), _C.ptr(cat),
                                            512 * N, 0, st)
    res = []
    for m in (3, 0):
        mode(m)
        h = torch.zeros(B, M, P, device=dev) if hout else None
        cat = torch.zeros(B, 512, N, device=dev)
        assert run(cur, h, cat) == 0
        torch.cuda.synchronize()
        res.append((h, cat))
    print("K=%d M=%d 4-D == 3-D bitwise: h %s pool %s  status %d" % (
        K, M, "-" if not hout else torch.equal(res[0][0], res[1][0]), torch.equal(res[0][1], res[1][1]), cur.l3d_edgeconv_status()), flush=True)
    h, cat = res[0]
    line = "K=%d M=%d " % (K, M)
    for name, lib in olds.items():
        line += " %s %.1f" % (name, timeit(lambda: run(lib, h, cat)))
    mode(3); line += "  cur3d %.1f" % timeit(lambda: run(cur, h, cat))
    mode(0); line += "  cur4d %.1f" % timeit(lambda: run(cur, h, cat))
    print(line, flush=True)
    del x, h, cat, res

# linear 512 -> 512 and attention passes
Bq, D, Nn = 32, 512, 1024
xin = torch.randn(Bq, D, Nn, device=dev); wl = torch.randn(D, D, device=dev) * 0.05; bl = torch.randn(D, device=dev)
outl = [torch.empty(Bq, D, Nn, device=dev) for _ in range(2)]
def lin(o): return cur.l3d_linear_cm(_C.ptr(wl), _C.ptr(xin), _C.ptr(bl), _C.ptr(None), _C.ptr(None), Bq, D, D, Nn, 0, 0, _C.ptr(o), st)
mode(3); lin(outl[0]); t3 = timeit(lambda: lin(outl[0])); mode(0); lin(outl[1]); t4 = timeit(lambda: lin(outl[1]))
print("linear 512->512: 3d %.1f  4d %.1f  equal %s" % (t3, t4, torch.equal(outl[0], outl[1])), flush=True)
BH, Dk = 128, 128
q = torch.randn(BH, Dk, Nn, device=dev); kk = torch.randn(BH, Dk, Nn, device=dev)
P = _C.ptr
for precise in (1, 0):
    r = []
    for m in (3, 0):
        mode(m)
        stats = torch.empty(BH, Nn, 2, device=dev)
        f = lambda: cur.l3d_attention_stats(P(q), P(kk), BH, Dk, Nn, Nn, precise, P(stats), st)
        f(); r.append((timeit(f), stats[:, :, 0].clone()))
    print("stats precise=%d: 3d %.1f  4d %.1f  max equal %s" % (precise, r[0][0], r[1][0], torch.equal(r[0][1], r[1][1])), flush=True)
stats = torch.empty(BH, Nn, 2, device=dev)
mode(0); cur.l3d_attention_stats(P(q), P(kk), BH, Dk, Nn, Nn, 0, P(stats), st)
r = []
for m in (3, 0):
    mode(m)
    pt = torch.empty(BH, Nn, Nn, device=dev)
    f = lambda: cur.l3d_attention_probs_t(P(q), P(kk), P(stats), BH, Dk, Nn, Nn, 0, P(pt), st)
    f(); r.append((timeit(f), pt))
print("probs: 3d %.1f  4d %.1f  equal %s" % (r[0][0], r[1][0], torch.equal(r[0][1], r[1][1])), flush=True)
pt = r[1][1]; del r
vt = torch.randn(32, Nn, 512, device=dev); rs = stats[:, :, 1].contiguous()
r = []
for m in (3, 0):
    mode(m)
    ctx = torch.empty(BH, 128, Nn, device=dev)
    f = lambda: cur.l3d_linear_cm(P(vt), P(pt), P(None), P(None), P(rs), BH, 128, Nn, Nn, 0, 4, P(ctx), st)
    f(); r.append((timeit(f), ctx))
print("p.v: 3d %.1f  4d %.1f  equal %s" % (r[0][0], r[1][0], torch.equal(r[0][1], r[1][1])), flush=True)
# SVD-head front at C3
src = torch.randn(32, 512, 1024, device=dev); tgt = torch.randn(32, 512, 1024, device=dev); xyz = torch.randn(32, 3, 1024, device=dev)
r = []
for m in (3, 0):
    mode(m)
    out = torch.empty(32, 3, 1024, device=dev)
    f = lambda: cur.l3d_soft_correspondence(P(src), P(tgt), P(xyz), 32, 512, 1024, 1024, P(out), st)
    assert f() == 0; r.append((timeit(f), out))
print("softcorr C3: 3d %.1f  4d %.1f  equal %s  status %d" % (r[0][0], r[1][0], torch.equal(r[0][1], r[1][1]), cur.l3d_soft_correspondence_status()), flush=True)
mode(0)
