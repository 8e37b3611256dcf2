"""Bidirectional cross-view transformer of the 2D->3D lifting stage (stays in PyTorch).

State-dict compatible re-implementation of the modules the reference's
`Feature_Aligner` uses from `transformer/attention.py:196-275,336-396`
(`CrossAttention`, `BasicTransformerBlock`, `BidirectionTransformerBlock`,
`BidirectionTransformer`; GEGLU feed-forward `:81-108`).  Parameter names match
the reference so real checkpoints load; the unused LDM leftovers of the
reference file (`LinearAttention`, `SpatialSelfAttention`, `SpatialTransformer`,
`CheckpointFunction`) are not part of the model and are not reproduced.
Attention runs through `F.scaled_dot_product_attention` (fused kernels on B200).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class GEGLU(nn.Module):
    """x * gelu(gate) with one projection producing both halves (attention.py:81-88)."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, 2 * dim_out)

    def forward(self, x):
        value, gate = self.proj(x).chunk(2, dim=-1)
        return value * F.gelu(gate)


class FeedForward(nn.Module):
    """`net.0` = GEGLU (or Linear+GELU), `net.1` = Dropout, `net.2` = Linear (attention.py:91-108)."""

    def __init__(self, dim: int, dim_out: int | None = None, mult: int = 4, glu: bool = False, dropout: float = 0.0):
        super().__init__()
        inner = int(dim * mult)
        first = GEGLU(dim, inner) if glu else nn.Sequential(nn.Linear(dim, inner), nn.GELU())
        self.net = nn.Sequential(first, nn.Dropout(dropout), nn.Linear(inner, dim if dim_out is None else dim_out))

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):
    """Multi-head attention of x over `context` (self-attention when None); attention.py:196-237."""

    def __init__(self, query_dim: int, context_dim: int | None = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None, mask=None):
        src = x if context is None else context
        b, n, _ = x.shape
        split = lambda t: t.view(b, t.shape[1], self.heads, -1).transpose(1, 2)      # b h n d
        q, k, v = split(self.to_q(x)), split(self.to_k(src)), split(self.to_v(src))
        attn_mask = None
        if mask is not None:
            attn_mask = mask.reshape(b, 1, 1, -1).to(torch.bool)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask, scale=self.scale)
        return self.to_out(out.transpose(1, 2).reshape(b, n, -1))


class BasicTransformerBlock(nn.Module):
    """message = norm2(ff(cat[x, norm1(attn(x, ctx))])); return x + message (attention.py:240-258)."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, normalize=True):
        super().__init__()
        self.attn = CrossAttention(dim, context_dim, n_heads, d_head, dropout)
        self.ff = FeedForward(2 * dim, dim, glu=gated_ff, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim) if normalize else nn.Sequential()
        self.norm2 = nn.LayerNorm(dim) if normalize else nn.Sequential()

    def forward(self, x, context=None):
        message = self.norm1(self.attn(x, context))
        message = self.norm2(self.ff(torch.cat([x, message], dim=-1)))
        return x + message


class BidirectionTransformerBlock(nn.Module):
    """Self-attention on each view, then cross-attention in both directions (attention.py:260-275)."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, normalize=True):
        super().__init__()
        mk = lambda: BasicTransformerBlock(dim, n_heads, d_head, dropout, context_dim, gated_ff, normalize)
        self.attn_self_1, self.attn_self_2 = mk(), mk()
        self.attn_cross_1, self.attn_cross_2 = mk(), mk()

    def forward(self, x, context):
        x = self.attn_self_1(x)
        context = self.attn_self_2(context)
        return self.attn_cross_1(x, context), self.attn_cross_2(context, x)


class BidirectionTransformer(nn.Module):
    """GroupNorm(32) -> 1x1 proj -> tokens -> `depth` bidirectional blocks -> 1x1 proj + residual
    for both feature maps (attention.py:336-396)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, normalize=True):
        super().__init__()
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        if context_dim is not None:
            self.proj_context_in = nn.Conv2d(context_dim, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            BidirectionTransformerBlock(inner, n_heads, d_head, dropout, inner, normalize=normalize) for _ in range(depth))
        self.proj_out = nn.Conv2d(inner, in_channels, 1)
        self.proj_context_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context):
        (b, _, h, w), (_, _, hc, wc) = x.shape, context.shape
        tok = self.proj_in(self.norm(x)).flatten(2).transpose(1, 2)
        ctx = self.proj_context_in(self.norm(context)).flatten(2).transpose(1, 2)
        for block in self.transformer_blocks:
            tok, ctx = block(tok, ctx)
        tok = tok.transpose(1, 2).reshape(b, -1, h, w)
        ctx = ctx.transpose(1, 2).reshape(b, -1, hc, wc)
        return x + self.proj_out(tok), context + self.proj_context_out(ctx)
