"""models/attention.py of the reference: MultiheadAttention WITHOUT in-projections, q/k dim != v dim (:55-175),
forward = multi_head_attention_forward (:178-385).  Sequence-first API kept: query [L,B,E], key [S,B,E],
value [S,B,vdim]; returns (out [L,B,vdim], None) -- the head-averaged weights the reference computes are
discarded by every caller, so they are not produced."""
import torch
from torch import nn

from .. import ops


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None):
        super().__init__()
        self.dropout = float(dropout)             # on the attention probabilities, training only (attention.py:371)
        assert not add_bias_kv and not add_zero_attn
        self.embed_dim = embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.out_proj = nn.Linear(self.vdim, self.vdim)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def core(self, q16, k16, v16, mask_u8=None, q2=None, k2=None):
        """batch-first bf16 [B,L,E] operands -> bf16 [B,L,vdim] (before out_proj)."""
        return ops.attention(q16, k16, v16, self.num_heads, float(self.head_dim) ** -0.5, mask_u8=mask_u8, q2=q2, k2=k2,
                             drop_p=self.dropout if self.training else 0.0)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False, attn_mask=None):
        assert attn_mask is None, "attn_mask is never used on this path"
        to16 = lambda t: ops.cast_bf16(t.transpose(0, 1).contiguous()) if t.dtype != torch.bfloat16 else t.transpose(0, 1).contiguous()
        m = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask is not None else None
        o = self.core(to16(query), to16(key), to16(value), m)
        out = ops.linear(o, self.out_proj.weight, self.out_proj.bias, out_f32=True)
        return out.transpose(0, 1), None
