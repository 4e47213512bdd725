"""How loose is the Cauchy-Schwarz exponent reference on DCP's transformer?  Runs the C3 forward (pretrained reference
checkpoint if oracle/_ref holds it, else random weights) with learning3d_b200.utils.transformer_fused.attention_cm
instrumented: per attention call the largest row bound (log2 units), the largest true row maximum and the device flag."""
import math
import sys

import torch

sys.path.insert(0, ".")


def run(net, tag, dev):
    from learning3d_b200 import _C
    from learning3d_b200.utils import transformer_fused as tf
    orig = tf.attention_cm
    rows = []

    def probe(attn, xq, xkv, residual):
        lib = _C.lib()
        B, d, Nq = xq.shape
        Nk = xkv.shape[2]
        h, dk = attn.h, attn.d_k
        q = tf.linear_cm(xq, attn.linears[0]); k = tf.linear_cm(xkv, attn.linears[1])
        stats = torch.empty((B * h, Nq, 2), device=xq.device)
        ws = torch.empty(B * h + 1, dtype=torch.int32, device=xq.device)
        _C.check(lib.l3d_attention_bounds(_C.ptr(q), _C.ptr(k), B * h, dk, Nq, Nk, _C.ptr(stats), _C.ptr(ws), _C.stream()))
        bound = stats[:, :, 0].max().item()
        smax, smin = -1e30, 1e30
        for b0 in range(0, B * h, 16):
            s = torch.einsum("bdq,bdk->bqk", q.view(B * h, dk, Nq)[b0:b0 + 16], k.view(B * h, dk, Nk)[b0:b0 + 16])
            rm = s.max(-1).values * (math.log2(math.e) / math.sqrt(dk))
            smax, smin = max(smax, rm.max().item()), min(smin, rm.min().item())
        rows.append((bound, smax, smin, int(ws[B * h].item())))
        return orig(attn, xq, xkv, residual)

    tf.attention_cm = probe
    try:
        tpl = torch.rand(32, 1024, 3, device=dev)
        tpl = tpl - tpl.mean(dim=1, keepdim=True)
        src = tpl @ torch.linalg.qr(torch.randn(32, 3, 3, device=dev))[0].transpose(1, 2) + 0.1
        with torch.no_grad():
            net(tpl, src)
    finally:
        tf.attention_cm = orig
    print("DCP forward, B=32, N=1024, %s" % tag)
    for i, (b, hi, lo, f) in enumerate(rows):
        print("  attention call %d: largest row bound %.3g; row maxima in [%.3g, %.3g] (log2 units); flag %d" % (i, b, lo, hi, f))


def main():
    from learning3d_b200.models import DCP, DGCNN
    from oracle import ref_pkg                                # profiling script: the staged reference checkpoint only
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = DCP(feature_model=DGCNN(emb_dims=512), cycle=False).to(dev).eval()
    run(net, "random initial weights (torch defaults)", dev)
    with torch.no_grad():                                     # a transformer with O(1) logits: 8x larger q / k projections
        for m in net.pointer.modules():
            if m.__class__.__name__ == "MultiHeadedAttention":
                for lin in m.linears[:2]:
                    lin.weight.mul_(8.0)
    run(net, "random weights, q / k projections scaled x8 (peaky attention)", dev)
    ck = ref_pkg.checkpoint("exp_dcp/models/best_model.t7")
    if ck is not None:
        net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=False), strict=False)
        run(net, "pretrained exp_dcp checkpoint (its pointer weights are denormal: ~6e-41, the attention is uniform)", dev)


if __name__ == "__main__":
    main()
