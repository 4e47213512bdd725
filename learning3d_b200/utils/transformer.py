"""DCP's "pointer" network: a 1-block encoder/decoder Transformer (4 heads, ff 1024) with the parameter
names of learning3d/utils/transformer.py:219-243 so reference checkpoints load
(`model.encoder.layers.0.self_attn.linears.0.weight`, `...sublayer.0.norm.a_2`, ...).

SURVEY.md §8f rank 2.  In eval mode / no-grad the whole forward runs on the tcgen05 pipelines with channel-major
activations (utils/transformer_fused.py: linear layers, attention scores + p.v, LayerNorm).  Under autograd the
torch layers below run; there attention goes through torch's fused scaled_dot_product_attention, so the
[B, heads, N, N] score tensor of the reference (transformer.py:17-23) is not materialised either.
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Norm(nn.Module):
    """LayerNorm as the reference writes it (:128-137): unbiased std, eps added to the std."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(-1, keepdim=True)
        return self.a_2 * (x - mean) / (x.std(-1, keepdim=True) + self.eps) + self.b_2


class _Residual(nn.Module):
    def __init__(self, size):
        super().__init__()
        self.norm = _Norm(size)

    def forward(self, x, fn):
        return x + fn(self.norm(x))


class MultiHeadedAttention(nn.Module):
    def __init__(self, h, d_model, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        self.d_k, self.h = d_model // h, h
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])

    def forward(self, query, key, value, mask=None):
        B = query.size(0)
        q, k, v = [lin(t).view(B, -1, self.h, self.d_k).transpose(1, 2)
                   for lin, t in zip(self.linears, (query, key, value))]
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)      # softmax(qk^T / sqrt(d_k)) v
        return self.linears[3](out.transpose(1, 2).reshape(B, -1, self.h * self.d_k))


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.norm = nn.Sequential()
        self.w_2 = nn.Linear(d_ff, d_model)

    def forward(self, x):
        return self.w_2(F.relu(self.w_1(x)))


class _EncoderLayer(nn.Module):
    def __init__(self, size, heads, d_ff):
        super().__init__()
        self.size = size
        self.self_attn = MultiHeadedAttention(heads, size)
        self.feed_forward = PositionwiseFeedForward(size, d_ff)
        self.sublayer = nn.ModuleList([_Residual(size), _Residual(size)])

    def forward(self, x):
        x = self.sublayer[0](x, lambda y: self.self_attn(y, y, y))
        return self.sublayer[1](x, self.feed_forward)


class _DecoderLayer(nn.Module):
    def __init__(self, size, heads, d_ff):
        super().__init__()
        self.size = size
        self.self_attn = MultiHeadedAttention(heads, size)
        self.src_attn = MultiHeadedAttention(heads, size)
        self.feed_forward = PositionwiseFeedForward(size, d_ff)
        self.sublayer = nn.ModuleList([_Residual(size) for _ in range(3)])

    def forward(self, x, memory):
        x = self.sublayer[0](x, lambda y: self.self_attn(y, y, y))
        x = self.sublayer[1](x, lambda y: self.src_attn(y, memory, memory))
        return self.sublayer[2](x, self.feed_forward)


class _Stack(nn.Module):
    def __init__(self, layer, n):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])
        self.norm = _Norm(layer.size)


class _Encoder(_Stack):
    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return self.norm(x)


class _Decoder(_Stack):
    def forward(self, x, memory):
        for layer in self.layers:
            x = layer(x, memory)
        return self.norm(x)


class _EncoderDecoder(nn.Module):
    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder
        self.src_embed, self.tgt_embed, self.generator = nn.Sequential(), nn.Sequential(), nn.Sequential()

    def forward(self, src, tgt):
        return self.decoder(tgt, self.encoder(src))


class Identity(nn.Module):
    def forward(self, *input):
        return input


class Transformer(nn.Module):
    """forward(src_emb [B,C,N], tgt_emb [B,C,N]) -> (src residual, tgt residual), both [B,C,N]."""

    def __init__(self, emb_dims, n_blocks, dropout, ff_dims, n_heads):
        super().__init__()
        self.emb_dims, self.N, self.dropout, self.ff_dims, self.n_heads = emb_dims, n_blocks, dropout, ff_dims, n_heads
        self.model = _EncoderDecoder(_Encoder(_EncoderLayer(emb_dims, n_heads, ff_dims), n_blocks),
                                     _Decoder(_DecoderLayer(emb_dims, n_heads, ff_dims), n_blocks))

    def _l3d_torch_forward(self, *input):
        """The reference's data flow (transformer.py:255-263) in torch ops: training / autograd path."""
        src = input[0].transpose(2, 1).contiguous()
        tgt = input[1].transpose(2, 1).contiguous()
        tgt_embedding = self.model(src, tgt).transpose(2, 1).contiguous()
        src_embedding = self.model(tgt, src).transpose(2, 1).contiguous()
        return src_embedding, tgt_embedding

    def forward(self, *input):
        from .transformer_fused import transformer_forward
        return transformer_forward(self, *input)
