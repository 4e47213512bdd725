"""Model-level device timings at BASELINE.json's configs C2 / C3 (run on the B200 box):

    python profiles/time_models.py

  * DGCNN.forward (kNN graph + EdgeConv stack + conv5), eval, B=32 N=1024: the fused tensor-core stack, its
    per-launch breakdown, and the reference's layer sequence in torch (cuDNN fp32 and cuDNN TF32) on the same GPU;
  * DCP.forward (DGCNN-512 + Transformer + SVDHead, cycle=True) with per-component times;
  * when the staged reference package is present (oracle/_ref/learning3d): the reference's own DCP unmodified vs
    rebound with learning3d_b200.bind.
CUDA-event timing, warm-up 3, mean of 10-20.  Prints one JSON object per line.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEV = torch.device("cuda:0")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def report(name, us, **kw):
    print(json.dumps(dict(op=name, us=round(us, 1), **kw)), flush=True)


def torch_layers(net, feat, B, N):
    h, pooled = feat, []
    for i in range(1, 5):
        h = torch.relu(getattr(net, "bn%d" % i)(getattr(net, "conv%d" % i)(h)))
        pooled.append(h.max(dim=-1, keepdim=True)[0])
    return torch.relu(net.bn5(net.conv5(torch.cat(pooled, 1)))).view(B, -1, N)


def main():
    from learning3d_b200 import _C
    from learning3d_b200.models import DCP, DGCNN
    from learning3d_b200.models.dgcnn import _edge_cache
    from learning3d_b200.utils import knn
    from oracle import ref_torch
    lib = _C.lib()
    torch.manual_seed(0)
    B, N, k = 32, 1024, 20
    x = torch.rand(B, N, 3, device=DEV)
    xt = x.permute(0, 2, 1).contiguous()
    for emb in (512, 1024):
        net = DGCNN(emb_dims=emb).to(DEV).eval()
        with torch.no_grad():
            us = timeit(lambda: net(x))
            report("DGCNN-%d forward fused (knn + EdgeConv tcgen05) B32 N1024" % emb, us, clouds_per_s=B / (us * 1e-6))
            for tf32 in (False, True):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                us_t = timeit(lambda: torch_layers(net, ref_torch.get_graph_feature(xt, k=k), B, N), iters=5, warm=2)
                report("DGCNN-%d forward reference torch ops (tf32=%s)" % (emb, tf32), us_t, clouds_per_s=B / (us_t * 1e-6))
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    # per-launch breakdown of the fused stack (emb 512)
    net = DGCNN(emb_dims=512).to(DEV).eval()
    cache = _edge_cache(net, DEV)
    widths = [w.shape[1] for w in cache["wt"]]
    P = N * k
    idx = knn(xt, k)
    cat = torch.empty(B, 512, N, device=DEV)
    hs = [torch.empty(B, widths[0], P, device=DEV), torch.empty(B, widths[1], P, device=DEV),
          torch.empty(B, widths[2], P, device=DEV)]
    w1, s1, t1 = cache["l1_host"]
    st = _C.stream()
    report("knn", timeit(lambda: knn(xt, k)))
    report("edgeconv layer1 (6->64, SIMT, h1 + pool)", timeit(lambda: lib.l3d_edgeconv_layer1(
        _C.ptr(xt), _C.ptr(idx), _C._P(w1.data_ptr()), _C._P(s1.data_ptr()), _C._P(t1.data_ptr()), B, N, k, 64,
        _C.ptr(hs[0]), _C.ptr(cat), 512 * N, 0, st)))
    coff = 64
    for i in (1, 2, 3):
        hin = hs[i - 1]
        hout = hs[i] if i < 3 else None
        us = timeit(lambda: lib.l3d_conv1x1_bn_relu_maxk(
            _C.ptr(cache["wt"][i]), _C.ptr(hin), _C.ptr(cache["scale"][i]), _C.ptr(cache["shift"][i]), B, widths[i],
            widths[i - 1], P, k, 1, _C.ptr(hout), _C.ptr(cat), 512 * N, coff, st))
        fl = 2.0 * B * P * widths[i] * widths[i - 1]
        report("edgeconv layer%d (%d->%d) tcgen05" % (i + 1, widths[i - 1], widths[i]), us,
               fp32_equiv_tflops=fl / us / 1e6, tf32_issued_tflops=3 * fl / us / 1e6)
        coff += widths[i]
    out = torch.empty(B, 512, N, device=DEV)
    us = timeit(lambda: lib.l3d_conv1x1_bn_relu_maxk(
        _C.ptr(cache["wt"][4]), _C.ptr(cat), _C.ptr(cache["scale"][4]), _C.ptr(cache["shift"][4]), B, 512, 512, N, 1, 1,
        _C.ptr(out), _C.ptr(None), 0, 0, st))
    fl = 2.0 * B * N * 512 * 512
    report("conv5 (512->512) tcgen05", us, fp32_equiv_tflops=fl / us / 1e6, tf32_issued_tflops=3 * fl / us / 1e6)
    assert lib.l3d_edgeconv_status() == 0

    # DCP components
    dcp = DCP(feature_model=DGCNN(emb_dims=512), cycle=True).to(DEV).eval()
    src = x
    tpl = torch.rand(B, N, 3, device=DEV)
    with torch.no_grad():
        us = timeit(lambda: dcp(tpl, src), iters=10)
        report("DCP forward (ours) B32 N1024 emb512 cycle", us, pairs_per_s=B / (us * 1e-6))
        f1, f2 = dcp.emb_nn(src), dcp.emb_nn(tpl)
        report("  DCP.emb_nn x2", 2 * timeit(lambda: dcp.emb_nn(src), iters=10))
        report("  DCP.pointer (transformer)", timeit(lambda: dcp.pointer(f1, f2), iters=10))
        report("  DCP.head x2", 2 * timeit(lambda: dcp.head(f1, f2, src, tpl), iters=10))

    from oracle import ref_pkg
    if ref_pkg.reference_root() is not None:
        from learning3d_b200 import bind
        ref = ref_pkg.import_reference()
        rnet = ref.models.DCP(feature_model=ref.models.DGCNN(emb_dims=512), cycle=True).to(DEV).eval()
        with torch.no_grad():
            us_ref = timeit(lambda: rnet(tpl, src), iters=5, warm=2)
            report("reference DCP forward, unmodified torch ops on the B200 (fp32)", us_ref, pairs_per_s=B / (us_ref * 1e-6))
            report("  reference DCP.emb_nn x2", 2 * timeit(lambda: rnet.emb_nn(src), iters=5, warm=2))
            g1, g2 = rnet.emb_nn(src), rnet.emb_nn(tpl)
            report("  reference DCP.pointer", timeit(lambda: rnet.pointer(g1, g2), iters=5, warm=2))
            report("  reference DCP.head x2", 2 * timeit(lambda: rnet.head(g1, g2, src, tpl), iters=5, warm=2))
            bind.bind(ref)
            us_b = timeit(lambda: rnet(tpl, src), iters=10)
            bind.unbind(ref)
            report("reference DCP forward rebound to libl3d_b200.so", us_b, pairs_per_s=B / (us_b * 1e-6),
                   speedup_vs_unmodified=us_ref / us_b)


if __name__ == "__main__":
    main()
