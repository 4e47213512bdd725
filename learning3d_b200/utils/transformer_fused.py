"""Eval-mode forward of DCP's transformer (utils/transformer.py:219-263) on the tcgen05 pipelines, activations kept
channel-major [B, d_model, N] from DGCNN to SVDHead (the reference transposes at entry and exit, :257-262).

    nn.Linear            -> l3d_linear_cm        (3xTF32 GEMM, bias / ReLU / residual in the epilogue)
    attention (:17-23)   -> l3d_attention_stats + l3d_attention_probs_t (scores on tcgen05, never written; the
                            probabilities leave transposed) + l3d_linear_cm(w_heads) for p.v
    LayerNorm (:128-137) -> l3d_layernorm_cm

Works on this package's Transformer and, through learning3d_b200.bind, on the reference's own module (same
attribute tree: model.encoder.layers[i].self_attn.linears[0..3], .feed_forward.w_1/.w_2, .sublayer[j].norm,
model.decoder.layers[i].src_attn, ...).  Forward only: under autograd the module's own torch forward runs.
"""
import torch

from .. import _C


def _wt(lin):
    """weight.t() [K, M] (the MN-major operand) and bias of an nn.Linear, cached on the module."""
    w, b = lin.weight, lin.bias
    key = (w.data_ptr(), w._version, str(w.device), None if b is None else (b.data_ptr(), b._version))
    c = lin.__dict__.get("_l3d_wt")
    if c is None or c[0] != key:
        with torch.no_grad():
            c = (key, w.detach().float().t().contiguous(), None if b is None else b.detach().float().contiguous())
        lin.__dict__["_l3d_wt"] = c
    return c[1], c[2]


def linear_cm(x, lin, relu=False, residual=None):
    """x [B, K, N] -> [B, M, N] = lin(x^T)^T (+ residual), one tcgen05 launch."""
    wt, bias = _wt(lin)
    B, K, N = x.shape
    M = wt.shape[1]
    out = torch.empty((B, M, N), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().l3d_linear_cm(_C.ptr(wt), _C.ptr(x), _C.ptr(bias), _C.ptr(residual), _C.ptr(None), B, M, K, N,
                                    1 if relu else 0, 0, _C.ptr(out), _C.stream()), "linear")
    return out


def linear_cm_t(x, lin):
    """x [B, K, N] -> lin(x^T) [B, N, M], i.e. the position-major result, written directly by the GEMM epilogue."""
    wt, bias = _wt(lin)
    B, K, N = x.shape
    M = wt.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().l3d_linear_cm_t(_C.ptr(wt), _C.ptr(x), _C.ptr(bias), B, M, K, N, 0, _C.ptr(out), _C.stream()),
             "linear (transposed output)")
    return out


def layernorm_cm(x, norm):
    B, D, N = x.shape
    out = torch.empty_like(x)
    _C.check(_C.lib().l3d_layernorm_cm(_C.ptr(x), _C.ptr(norm.a_2.detach()), _C.ptr(norm.b_2.detach()), float(norm.eps),
                                       B, D, N, _C.ptr(out), _C.stream()), "layernorm")
    return out


def attention_cm(attn, xq, xkv, residual):
    """MultiHeadedAttention.forward (transformer.py:175-194) on channel-major inputs: xq [B,d,Nq], xkv [B,d,Nk] ->
    residual + linears[3](concat_heads(softmax(q k^T / sqrt(d_k)) v))   [B,d,Nq]."""
    lib = _C.lib()
    B, d, Nq = xq.shape
    Nk = xkv.shape[2]
    h, dk = attn.h, attn.d_k
    q = linear_cm(xq, attn.linears[0])
    k = linear_cm(xkv, attn.linears[1])
    vt = linear_cm_t(xkv, attn.linears[2])                    # v^T [B, Nk, h*d_k]: heads side by side
    st = _C.stream()
    stats = torch.empty((B * h, Nq, 2), dtype=torch.float32, device=xq.device)
    # exponent reference of every row: the Cauchy-Schwarz bound |q_i| max_j |k_j| / sqrt(d_k) (two small launches);
    # the statistics pass proper (row maxima from ONE TF32 MMA pass) only runs if some bound is too loose — decided on
    # the device.  Pass 2: 3xTF32 scores, unnormalised probabilities 2^(s - m_i) written transposed + exact row sums.
    ws = torch.empty(B * h + 1, dtype=torch.int32, device=xq.device)
    _C.check(lib.l3d_attention_bounds(_C.ptr(q), _C.ptr(k), B * h, dk, Nq, Nk, _C.ptr(stats), _C.ptr(ws), st), "attention bounds")
    _C.check(lib.l3d_attention_stats_if(_C.ptr(q), _C.ptr(k), B * h, dk, Nq, Nk, _C._P(ws.data_ptr() + 4 * B * h),
                                        _C.ptr(stats), st), "attention stats")
    probs_t = torch.empty((B * h, Nk, Nq), dtype=torch.float32, device=xq.device)
    _C.check(lib.l3d_attention_probs_t(_C.ptr(q), _C.ptr(k), _C.ptr(stats), B * h, dk, Nq, Nk, 0, _C.ptr(probs_t), st),
             "attention probabilities")
    rowsum = stats[:, :, 1].contiguous()                        # [B*h, Nq]
    ctx = torch.empty((B, d, Nq), dtype=torch.float32, device=xq.device)      # = [B*h, d_k, Nq]
    _C.check(lib.l3d_linear_cm(_C.ptr(vt), _C.ptr(probs_t), _C.ptr(None), _C.ptr(None), _C.ptr(rowsum), B * h, dk, Nk,
                               Nq, 0, h, _C.ptr(ctx), st), "attention p.v")
    return linear_cm(ctx, attn.linears[3], residual=residual)


def _ff(ff, y, residual):
    return linear_cm(linear_cm(y, ff.w_1, relu=True), ff.w_2, residual=residual)


def _encode(enc, x):
    for layer in enc.layers:
        y = layernorm_cm(x, layer.sublayer[0].norm)
        x = attention_cm(layer.self_attn, y, y, x)
        x = _ff(layer.feed_forward, layernorm_cm(x, layer.sublayer[1].norm), x)
    return layernorm_cm(x, enc.norm)


def _decode(dec, x, memory):
    for layer in dec.layers:
        y = layernorm_cm(x, layer.sublayer[0].norm)
        x = attention_cm(layer.self_attn, y, y, x)
        x = attention_cm(layer.src_attn, layernorm_cm(x, layer.sublayer[1].norm), memory, x)
        x = _ff(layer.feed_forward, layernorm_cm(x, layer.sublayer[2].norm), x)
    return layernorm_cm(x, dec.norm)


def fused_ok(net, src, tgt):
    if net.training or not (src.is_cuda and tgt.is_cuda) or src.dtype != torch.float32 or tgt.dtype != torch.float32:
        return False
    if torch.is_grad_enabled() and (src.requires_grad or tgt.requires_grad or any(p.requires_grad for p in net.parameters())):
        return False
    if src.dim() != 3 or tgt.dim() != 3 or (src.shape[2] & 3) or (tgt.shape[2] & 3) or (src.shape[1] & 3):
        return False
    try:
        layers = list(net.model.encoder.layers) + list(net.model.decoder.layers)
        attns = [l.self_attn for l in layers] + [l.src_attn for l in net.model.decoder.layers]
    except AttributeError:
        return False
    if any(a.d_k != 128 or getattr(a, "dropout", None) is not None for a in attns):
        return False
    for m in (net.model.src_embed, net.model.tgt_embed, net.model.generator):
        if len(list(m.children())) != 0:
            return False
    return all(len(list(getattr(l.feed_forward, "norm", torch.nn.Sequential()).children())) == 0 for l in layers)


def transformer_forward(self, *input):
    """Transformer.forward (transformer.py:255-263): (src [B,C,N], tgt [B,C,N]) -> (src_embedding, tgt_embedding)."""
    src, tgt = input[0], input[1]
    if not fused_ok(self, src, tgt):
        return self._l3d_torch_forward(*input)
    src, tgt = src.contiguous(), tgt.contiguous()
    with _C.on_device(src.device):
        tgt_embedding = _decode(self.model.decoder, tgt, _encode(self.model.encoder, src))
        src_embedding = _decode(self.model.decoder, src, _encode(self.model.encoder, tgt))
    return src_embedding, tgt_embedding
