"""Transformer pointer with the reference's module tree (reference model/transformer.py).

The classes below are parameter containers with the reference's attribute names, so the
state_dict keys are ``model.encoder.layers.0.self_attn.linears.0.weight`` etc. exactly as in
the reference's ``.t7`` files; every forward goes through the CUDA ops (vcr_net_b200/functional.py).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops


def clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class LayerNorm(nn.Module):
    """model/transformer.py:134-144: a_2*(x-mean)/(std_unbiased+eps)+b_2."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        return ops.layernorm(x.contiguous(), self.a_2, self.b_2, self.eps)


class SublayerConnection(nn.Module):
    """model/transformer.py:147-153: x + sublayer(norm(x))."""

    def __init__(self, size, dropout=None):
        super().__init__()
        self.norm = LayerNorm(size)

    def forward(self, x, sublayer):
        return x + sublayer(self.norm(x))


class MultiHeadedAttention(nn.Module):
    """model/transformer.py:188-224.  ``self.attn`` (the head-summed probability matrix the
    reference stores for plotting, :216-219) is NOT materialised -- documented deviation."""

    def __init__(self, h, d_model, is_src=False, overlap2=0.75, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        self.d_k = d_model // h
        self.h = h
        self.linears = clones(nn.Linear(d_model, d_model), 4)
        self.attn = None
        self.dropout = None
        self.is_src = is_src
        self.overlap2 = overlap2

    def forward(self, query, key, value, mask=None):
        if mask is not None:
            raise Exception("Not implemented: explicit attention masks are never used by VCR-Net")
        if key is not value:
            raise Exception("Not implemented: key and value must be the same tensor (as in VCR-Net)")
        xkv = None if key is query else key.contiguous()
        return Fn.mha_tok(self, query.contiguous(), xkv)


class PositionwiseFeedForward(nn.Module):
    """model/transformer.py:227-238."""

    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.norm = nn.Sequential()
        self.w_2 = nn.Linear(d_ff, d_model)
        self.dropout = None

    def forward(self, x):
        return Fn.ffn_tok(self, x.contiguous())


class EncoderLayer(nn.Module):
    def __init__(self, size, self_attn, feed_forward, dropout):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 2)
        self.size = size


class DecoderLayer(nn.Module):
    def __init__(self, size, self_attn, src_attn, feed_forward, dropout):
        super().__init__()
        self.size = size
        self.self_attn = self_attn
        self.src_attn = src_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 3)


class Encoder(nn.Module):
    def __init__(self, layer, N):
        super().__init__()
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)


class Decoder(nn.Module):
    def __init__(self, layer, N):
        super().__init__()
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)


class EncoderDecoder(nn.Module):
    """model/transformer.py:58-82."""

    def __init__(self, encoder, decoder, src_embed, tgt_embed, generator):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.src_embed = src_embed
        self.tgt_embed = tgt_embed
        self.generator = generator

    def forward(self, src, tgt, src_mask=None, tgt_mask=None, src_tgt=True):
        if src_mask is not None or tgt_mask is not None:
            raise Exception("Not implemented: masks")
        return Fn.encoder_decoder_tok(self, src.contiguous(), tgt.contiguous())


class Transformer(nn.Module):
    """model/transformer.py:241-272.  forward(src_emb [B,D,N], tgt_emb [B,D,N]) -> (src_p, tgt_p)."""

    def __init__(self, args):
        super().__init__()
        self.emb_dims = args.emb_dims
        self.N = args.n_blocks
        self.dropout = args.dropout
        self.ff_dims = args.ff_dims
        self.n_heads = args.n_heads
        self.overlap2 = float(args.overlap2)
        c = copy.deepcopy
        attn = MultiHeadedAttention(self.n_heads, self.emb_dims, is_src=False)
        if args.partial:
            src_attn = MultiHeadedAttention(self.n_heads, self.emb_dims, is_src=True, overlap2=self.overlap2)
        else:
            src_attn = MultiHeadedAttention(self.n_heads, self.emb_dims, is_src=False)
        ff = PositionwiseFeedForward(self.emb_dims, self.ff_dims, self.dropout)
        self.model = EncoderDecoder(Encoder(EncoderLayer(self.emb_dims, c(attn), c(ff), self.dropout), self.N),
                                    Decoder(DecoderLayer(self.emb_dims, c(attn), c(src_attn), c(ff), self.dropout),
                                            self.N),
                                    nn.Sequential(), nn.Sequential(), nn.Sequential())

    def forward_tokens(self, src_tok, tgt_tok, add_input=False, want_head=False, want_out=True):
        """want_head: also return (HeadPre src, HeadPre tgt) -- operand copies + squared norms of the outputs written by the
        final LayerNorm kernel -- or None when the path does not produce them.  want_out=False (with want_head): the fp32
        outputs are not stored (returned as None) when the head only needs the operand copies."""
        return Fn.transformer_tokens(self, src_tok, tgt_tok, add_input=add_input, want_head=want_head, want_out=want_out)

    def forward(self, *input):
        src_tok = ops.transpose_batched(input[0])
        tgt_tok = ops.transpose_batched(input[1])
        src_p, tgt_p = self.forward_tokens(src_tok, tgt_tok)
        return ops.transpose_batched(src_p), ops.transpose_batched(tgt_p)
