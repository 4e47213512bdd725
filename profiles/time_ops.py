"""Per-op device timings of every hot-path kernel family at the BASELINE.json config shapes.

    python profiles/time_ops.py            # prints one JSON object per op (run on the B200 box)

CUDA-event timing on the launching stream, 20 warm-up + 200 timed iterations; operands are rotated
through a pool larger than L2 where the op's footprint is small.  Not a bench.py replacement: this
feeds the per-kernel table in DESIGN.md §4 / profiles/.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from learning3d_b200 import _C  # noqa: E402
from learning3d_b200.utils import knn, get_graph_feature, query_ball_point, farthest_point_sample, SVDHead  # noqa: E402
from learning3d_b200.utils import pointconv_util as pcu  # noqa: E402
from learning3d_b200.utils.lib import pointnet2_utils as pu  # noqa: E402
from learning3d_b200.losses import ChamferDistanceLoss, EMDLoss  # noqa: E402
from learning3d_b200.losses.cuda.chamfer_distance import ChamferDistanceFunction  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, iters=200, warm=20):
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
    print(json.dumps(dict(op=name, us=round(us, 2), **kw)), flush=True)


def main():
    torch.manual_seed(0)
    # C2: DGCNN graph
    x = torch.rand(32, 3, 1024, device=DEV)
    us = timeit(lambda: knn(x, 20))
    report("knn B32 N1024 k20", us, pairs_per_s=32 * 1024 * 1024 / (us * 1e-6))
    idx = knn(x, 20)
    lib = _C.lib()
    out = torch.empty(32, 6, 1024, 20, device=DEV)
    us = timeit(lambda: lib.l3d_graph_feature(_C.ptr(x), _C.ptr(idx), 32, 3, 1024, 20, _C.ptr(out), _C.stream()))
    report("graph_feature gather B32 C3 N1024 k20", us, GBps=(out.numel() * 4 + idx.numel() * 8) / us / 1e3)
    report("get_graph_feature (knn+gather)", timeit(lambda: get_graph_feature(x, 20)))
    # C1 / B=32 Chamfer
    for B in (4, 32):
        a = torch.rand(B, 1024, 3, device=DEV, requires_grad=True)
        b = torch.rand(B, 1024, 3, device=DEV, requires_grad=True)
        d1 = torch.empty(B, 1024, device=DEV); d2 = torch.empty(B, 1024, device=DEV)
        i1 = torch.empty(B, 1024, dtype=torch.int, device=DEV); i2 = torch.empty(B, 1024, dtype=torch.int, device=DEV)
        us = timeit(lambda: lib.l3d_chamfer_forward(_C.ptr(a), _C.ptr(b), B, 1024, 1024, _C.ptr(d1), _C.ptr(d2),
                                                    _C.ptr(i1), _C.ptr(i2), _C.stream()))
        report("chamfer_forward kernel B%d" % B, us, pairs_per_s=2 * B * 1024 * 1024 / (us * 1e-6))
        ga = torch.empty_like(a); gb = torch.empty_like(b)
        us = timeit(lambda: lib.l3d_chamfer_backward(_C.ptr(a), _C.ptr(b), B, 1024, 1024, _C.ptr(d1), _C.ptr(d2),
                                                     _C.ptr(i1), _C.ptr(i2), _C.ptr(ga), _C.ptr(gb), _C.stream()))
        report("chamfer_backward kernel B%d" % B, us)
        crit = ChamferDistanceLoss()

        def fb():
            a.grad = b.grad = None
            crit(a, b).backward()
        us = timeit(fb)
        report("ChamferDistanceLoss fwd+bwd (python API) B%d" % B, us, clouds_per_s=B / (us * 1e-6))
    # C4: FlowNet3D grouping
    pc = (torch.rand(16, 2048, 3, device=DEV) * 4 - 2).contiguous()
    report("pn2 FPS B16 2048->1024", timeit(lambda: pu.furthest_point_sample(pc, 1024), iters=20, warm=3))
    fps = pu.furthest_point_sample(pc, 1024)
    new_xyz = pu.gather_operation(pc.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    report("pn2 ball_query B16 N2048 S1024 r.5 ns16", timeit(lambda: pu.ball_query(0.5, 16, pc, new_xyz)))
    report("torch-sem query_ball_point B16 N2048 S1024", timeit(lambda: query_ball_point(0.5, 16, pc, new_xyz)))
    p1 = torch.rand(16, 256, 3, device=DEV); p2 = torch.rand(16, 256, 3, device=DEV)
    report("pn2 knn B16 256x256 k64", timeit(lambda: pu.knn(64, p1, p2)))
    feat = torch.rand(16, 128, 256, device=DEV)
    _, kidx = pu.knn(64, p1, p2)
    us = timeit(lambda: pu.grouping_operation(feat, kidx))
    report("pn2 grouping B16 C128 256x64", us, GBps=2 * 16 * 128 * 256 * 64 * 4 / us / 1e3)
    q = torch.rand(16, 2048, 3, device=DEV); kn = torch.rand(16, 1024, 3, device=DEV)
    report("pn2 three_nn B16 2048<-1024", timeit(lambda: pu.three_nn(q, kn)))
    report("torch-sem FPS B32 1024->512", timeit(lambda: pcu.farthest_point_sample(torch.rand(32, 1024, 3, device=DEV), 512), iters=20, warm=3))
    xyz32 = torch.rand(32, 1024, 3, device=DEV)
    report("pointconv knn_point B32 N1024 S512 ns32", timeit(lambda: pcu.knn_point(32, xyz32, xyz32[:, :512].contiguous())))
    report("compute_density B32 N1024", timeit(lambda: pcu.compute_density(xyz32, 0.1)))
    # C5: EMD
    e1 = torch.rand(8, 1024, 3, device=DEV, requires_grad=True); e2 = torch.rand(8, 1024, 3, device=DEV)
    emd = EMDLoss()
    report("EMD forward B8 N1024", timeit(lambda: emd(e1, e2), iters=20, warm=3))

    def efb():
        e1.grad = None
        emd(e1, e2).backward()
    report("EMD fwd+bwd B8 N1024", timeit(efb, iters=20, warm=3))
    # C3: SVD head tail
    src = torch.rand(32, 3, 1024, device=DEV); corr = torch.rand(32, 3, 1024, device=DEV)
    R = torch.empty(32, 3, 3, device=DEV); t = torch.empty(32, 3, device=DEV)
    report("svd_head_tail B32 N1024", timeit(lambda: lib.l3d_svd_head_tail(_C.ptr(src), _C.ptr(corr), 32, 1024,
                                                                            _C.ptr(R), _C.ptr(t), _C.stream())))
    # C3: SVD head front (soft correspondences, tcgen05) and the whole SVDHead.forward under no_grad
    from learning3d_b200.utils.svd import soft_correspondence
    es = torch.randn(32, 512, 1024, device=DEV); et = torch.randn(32, 512, 1024, device=DEV)
    tg = torch.rand(32, 3, 1024, device=DEV)
    us = timeit(lambda: soft_correspondence(es, et, tg), iters=50, warm=5)
    report("soft_correspondence B32 d512 N1024", us, fp32_equiv_tflops=2.0 * 32 * 1024 * 1024 * 512 / us / 1e6)
    head = SVDHead(512, input_shape="bnc").to(DEV)
    srcp = torch.rand(32, 1024, 3, device=DEV); tgtp = torch.rand(32, 1024, 3, device=DEV)
    with torch.no_grad():
        report("SVDHead.forward B32 d512 N1024 (no_grad: fused front + Kabsch tail)",
               timeit(lambda: head(es, et, srcp, tgtp), iters=50, warm=5))
    # feature-space graphs (PRNet DGCNN)
    for C in (64, 128):
        xf = torch.randn(32, C, 1024, device=DEV)
        report("knn features B32 C%d N1024 k20" % C, timeit(lambda: knn(xf, 20), iters=50, warm=5))


if __name__ == "__main__":
    main()
