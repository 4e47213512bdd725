"""Drop-in for learning3d/utils/svd.py:5-59 (DCP's SVD head).

Same constructor, state_dict key (`reflect`) and forward contract.  The soft-correspondence front
half (score GEMM, softmax, src_corr GEMM: svd.py:23-28) is ONE fused
tcgen05 kernel (l3d_soft_correspondence: 3xTF32 score tiles in TMEM, online softmax, never writes the
[B,N,N] score matrix); under autograd the same forward is paired with a recompute-backward on tcgen05
(_SoftCorr: row statistics + dS once, then two tensor-core GEMMs).  The tail — centring, H, the per-item torch.svd + torch.det loop with its host-synchronising
branch (svd.py:38-51) and t — is ONE launch of l3d_svd_head_tail for the whole batch, and its backward
(the reference trains through torch.svd's autograd) ONE launch of l3d_svd_head_tail_backward.
"""
import math

import torch
import torch.nn as nn

from .. import _C


class _SVDTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, src_corr):
        B, _, N = src.shape
        R = torch.empty((B, 3, 3), dtype=torch.float32, device=src.device)
        t = torch.empty((B, 3), dtype=torch.float32, device=src.device)
        with _C.on_device(src.device):
            _C.check(_C.lib().l3d_svd_head_tail(_C.ptr(src), _C.ptr(src_corr), B, N, _C.ptr(R),
                                                _C.ptr(t), _C.stream()), "SVDHead")
        ctx.save_for_backward(src, src_corr)
        return R, t

    @staticmethod
    def backward(ctx, grad_R, grad_t):
        src, src_corr = ctx.saved_tensors
        B, _, N = src.shape
        grad_R = (torch.zeros((B, 3, 3), device=src.device) if grad_R is None else grad_R).contiguous().float()
        grad_t = (torch.zeros((B, 3), device=src.device) if grad_t is None else grad_t).contiguous().float()
        g_src = torch.empty_like(src)
        g_corr = torch.empty_like(src_corr)
        with _C.on_device(src.device):
            _C.check(_C.lib().l3d_svd_head_tail_backward(_C.ptr(src), _C.ptr(src_corr), _C.ptr(grad_R),
                                                         _C.ptr(grad_t), B, N, _C.ptr(g_src), _C.ptr(g_corr),
                                                         _C.stream()), "SVDHead backward")
        return g_src, g_corr


def soft_correspondence(src_embedding, tgt_embedding, tgt):
    """svd.py:23-28 fused, forward only: softmax(src_emb^T tgt_emb / sqrt(d_k)) applied to tgt.

    src_embedding [B,D,Ns], tgt_embedding [B,D,Nt], tgt [B,3,Nt] (CUDA fp32) -> src_corr [B,3,Ns].
    One tcgen05 kernel (3xTF32 score tiles in TMEM + online softmax); the [B,Ns,Nt] score matrix of the
    reference is never written.  No autograd: SVDHead.forward uses it when no gradient is required.
    """
    src_embedding = _C.require_cuda(src_embedding, "src_embedding")
    tgt_embedding = _C.require_cuda(tgt_embedding, "tgt_embedding")
    tgt = _C.require_cuda(tgt, "tgt")
    B, D, Ns = src_embedding.shape
    Nt = tgt_embedding.shape[2]
    if tgt_embedding.shape[0] != B or tgt_embedding.shape[1] != D or tuple(tgt.shape) != (B, 3, Nt):
        raise ValueError("soft_correspondence: inconsistent shapes %s %s %s" % (
            tuple(src_embedding.shape), tuple(tgt_embedding.shape), tuple(tgt.shape)))
    out = torch.empty((B, 3, Ns), dtype=torch.float32, device=src_embedding.device)
    with _C.on_device(src_embedding.device):
        _C.check(_C.lib().l3d_soft_correspondence(_C.ptr(src_embedding), _C.ptr(tgt_embedding), _C.ptr(tgt),
                                                  B, D, Ns, Nt, _C.ptr(out), _C.stream()), "soft_correspondence")
    return out


class _SoftCorr(torch.autograd.Function):
    """Differentiable soft correspondences (svd.py:23-28) without the score matrix in the forward: the fused tcgen05
    kernel computes src_corr; the backward recomputes the probabilities on tcgen05 from the row statistics, writes
    dS (plain and transposed) once and finishes with two tensor-core GEMMs — no cuBLAS, no saved [B,N,N] tensors."""

    @staticmethod
    def forward(ctx, src_embedding, tgt_embedding, tgt):
        corr = soft_correspondence(src_embedding, tgt_embedding, tgt)
        ctx.save_for_backward(src_embedding, tgt_embedding, tgt, corr)
        return corr

    @staticmethod
    def backward(ctx, grad_corr):
        src_emb, tgt_emb, tgt, corr = ctx.saved_tensors
        lib = _C.lib()
        B, D, Ns = src_emb.shape
        Nt = tgt_emb.shape[2]
        dev = src_emb.device
        g = grad_corr.contiguous().float()
        with _C.on_device(dev):
            st = _C.stream()
            stats = torch.empty((B, Ns, 2), dtype=torch.float32, device=dev)
            _C.check(lib.l3d_attention_stats(_C.ptr(src_emb), _C.ptr(tgt_emb), B, D, Ns, Nt, 1, _C.ptr(stats), st),
                     "softcorr stats")
            need_s, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            ds = torch.empty((B, Ns, Nt), dtype=torch.float32, device=dev) if need_t else None
            ds_t = torch.empty((B, Nt, Ns), dtype=torch.float32, device=dev) if need_s else None
            _C.check(lib.l3d_soft_correspondence_dscores(_C.ptr(src_emb), _C.ptr(tgt_emb), _C.ptr(tgt), _C.ptr(stats),
                                                         _C.ptr(g), _C.ptr(corr), B, D, Ns, Nt, _C.ptr(ds), _C.ptr(ds_t),
                                                         st), "softcorr dscores")
            g_src = g_tgt = None
            if need_s:          # d src_emb [B,D,Ns] = tgt_emb [D x Nt] . dS^T [Nt x Ns]
                wt = tgt_emb.transpose(1, 2).contiguous()
                g_src = torch.empty_like(src_emb)
                _C.check(lib.l3d_linear_cm(_C.ptr(wt), _C.ptr(ds_t), _C.ptr(None), _C.ptr(None), _C.ptr(None), B, D, Nt, Ns,
                                           0, 1, _C.ptr(g_src), st), "softcorr d src_emb")
            if need_t:          # d tgt_emb [B,D,Nt] = src_emb [D x Ns] . dS [Ns x Nt]
                wt = src_emb.transpose(1, 2).contiguous()
                g_tgt = torch.empty_like(tgt_emb)
                _C.check(lib.l3d_linear_cm(_C.ptr(wt), _C.ptr(ds), _C.ptr(None), _C.ptr(None), _C.ptr(None), B, D, Ns, Nt,
                                           0, 1, _C.ptr(g_tgt), st), "softcorr d tgt_emb")
        return g_src, g_tgt, None


def _softcorr_trainable(src_embedding, tgt_embedding, tgt):
    """The tensor-core backward needs TMA-friendly shapes and does not produce d tgt (the cloud itself)."""
    D, Ns, Nt = src_embedding.shape[1], src_embedding.shape[2], tgt_embedding.shape[2]
    return (not tgt.requires_grad) and D % 128 == 0 and Ns % 4 == 0 and Nt % 4 == 0


def svd_head_tail(src, src_corr):
    """src, src_corr [B,3,N] (CUDA fp32) -> R [B,3,3], t [B,3]; differentiable."""
    return _SVDTail.apply(_C.require_cuda(src, "src"), _C.require_cuda(src_corr, "src_corr"))


class SVDHead(nn.Module):
    def __init__(self, emb_dims, input_shape="bnc"):
        super(SVDHead, self).__init__()
        self.emb_dims = emb_dims
        self.reflect = nn.Parameter(torch.eye(3), requires_grad=False)
        self.reflect[2, 2] = -1
        self.input_shape = input_shape

    def forward(self, *input):
        src_embedding, tgt_embedding, src, tgt = input[0], input[1], input[2], input[3]
        if self.input_shape == "bnc":
            src = src.permute(0, 2, 1)
            tgt = tgt.permute(0, 2, 1)
        needs_grad = torch.is_grad_enabled() and any(
            t.requires_grad for t in (src_embedding, tgt_embedding, src, tgt))
        if not needs_grad:
            # inference: scores, softmax and the correspondence GEMM are one fused tcgen05 kernel
            src_corr = soft_correspondence(src_embedding, tgt_embedding, tgt)
        elif _softcorr_trainable(src_embedding, tgt_embedding, tgt):
            # training: same fused forward, recompute-backward on tcgen05 (no score matrix kept, no cuBLAS)
            src_corr = _SoftCorr.apply(src_embedding.contiguous(), tgt_embedding.contiguous(), tgt.contiguous())
        else:
            # training: autograd needs the score matrix; same torch ops as the reference (svd.py:23-28)
            d_k = src_embedding.size(1)
            scores = torch.matmul(src_embedding.transpose(2, 1).contiguous(), tgt_embedding) / math.sqrt(d_k)
            scores = torch.softmax(scores, dim=2)
            src_corr = torch.matmul(tgt, scores.transpose(2, 1).contiguous())
        return svd_head_tail(src, src_corr)
