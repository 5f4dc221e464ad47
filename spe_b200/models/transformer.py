"""models/transformer.py of the reference (conditional-DETR transformer) over sm_100a kernels.

Same classes / constructor arguments / state_dict keys: MLP (:21-33), gen_sineembed_for_position (:35-49),
Transformer (:51-160), TransformerEncoder/Layer (:162-189, :253-310), TransformerDecoder (:192-250),
TransformerDecoderLayer (:313-466), build_transformer (:473-484).

Internal layout is batch-first token-major ([B,N,D]; the reference is sequence-first [N,B,D]): fp32 residual
streams + bf16 GEMM operands.  Only what the reference actually runs is implemented: post-norm layers, ReLU
FFN, return_intermediate decoder; dropout (main.py:73) in train() mode through un-fused routes (ops.dropout), fused epilogues otherwise.  Work the reference repeats is done once:
  * query_pos-only projections are batch invariant -> computed on [Q,D] and broadcast in the GEMM epilogue;
  * the decoder passes of forward_refine (transformer.py:147-155: same weights, same memory, different query embeddings) run as
    ONE pass over the concatenated queries [B, P*Q, D]: every token-wise GEMM / LayerNorm and the cross-attention see P*Q
    queries per image (queries never interact there), self-attention runs per pass on the [B*P, Q, D] view.  The memory-side
    cross-attention projections (k_content, v, k_pos) are thereby computed once, and the ~100 small
    launches of a decoder layer are issued once instead of once per pass;
  * the per-head [content | position] concat (transformer.py:408-414) is never materialised: the logits are
    two QK^T GEMMs accumulated into one S.
"""
import copy
import math
from typing import Optional

import torch
from torch import nn, Tensor

from .. import ops
from .attention import MultiheadAttention


class MLP(nn.Module):
    """Very simple multi-layer perceptron (transformer.py:21-33 / conditional_detr.py:626-638)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        """x bf16 [..., in] (fp32 is cast) -> fp32 [..., out]; ReLUs fused into the GEMM epilogues."""
        if x.dtype != torch.bfloat16:
            x = ops.cast_bf16(x)
        for i, layer in enumerate(self.layers):
            if i < self.num_layers - 1:
                x = ops.linear(x, layer.weight, layer.bias, act="relu")
            else:
                x = ops.linear(x, layer.weight, layer.bias, out_f32=True)
        return x


def gen_sineembed_for_position(pos_tensor, d_model=256):
    """transformer.py:35-49 (hard-coded /128 exponent): [..., 2] -> [..., d_model]."""
    return ops.query_sine_embed(pos_tensor, d_model)


class _PackedSelfAttention(nn.Module):
    """Parameter holder with nn.MultiheadAttention's state_dict keys (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.dropout = float(dropout)             # nn.MultiheadAttention(dropout=...): on the attention probabilities, training only
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


def _res_ffn(layer, s32, s16, p, training):
    """x + dropout(linear2(dropout(relu(linear1(x)))))   (transformer.py:285-286, :424-425)"""
    if training and p > 0.0:
        h = ops.dropout(ops.linear(s16, layer.linear1.weight, layer.linear1.bias, act="relu"), p)
        return s32 + ops.dropout(ops.linear(h, layer.linear2.weight, layer.linear2.bias, out_f32=True), p)
    return ops.ffn(s16, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias, residual=s32, act="relu")


def _res_proj(a16, proj, res32, p, training):
    """res + dropout(out_proj(a))   (transformer.py:283, :384, :421)"""
    if training and p > 0.0:
        return res32 + ops.dropout(ops.linear(a16, proj.weight, proj.bias, out_f32=True), p)
    return ops.linear(a16, proj.weight, proj.bias, residual=res32, out_f32=True)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.dropout = float(dropout)
        assert activation == "relu" and not normalize_before, "the reference runs post-norm ReLU layers only"
        self.self_attn = _PackedSelfAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before

    def forward_tokens(self, src32, src16, pos32, mask_u8):
        """forward_post (transformer.py:275-288) on token-major tensors."""
        sa = self.self_attn
        D = sa.embed_dim
        qk_in = ops.add_cast(src32, pos32)                                         # with_pos_embed
        qk = ops.linear(qk_in, sa.in_proj_weight[:2 * D], sa.in_proj_bias[:2 * D])   # q | k in one GEMM
        v = ops.linear(src16, sa.in_proj_weight[2 * D:], sa.in_proj_bias[2 * D:])
        a = ops.attention(qk[:, :, :D], qk[:, :, D:], v, sa.num_heads, float(sa.head_dim) ** -0.5, mask_u8=mask_u8,
                          drop_p=sa.dropout if self.training else 0.0)
        x = _res_proj(a, sa.out_proj, src32, self.dropout, self.training)
        s32, s16 = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, want_f32=True)
        x = _res_ffn(self, s32, s16, self.dropout, self.training)
        return ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, want_f32=True)


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers) if num_layers != 0 else nn.ModuleList()
        self.num_layers = num_layers
        self.norm = norm
        assert norm is None, "post-norm encoder has no final norm (transformer.py:61)"

    def forward_tokens(self, src32, src16, pos32, mask_u8):
        for layer in self.layers:
            src32, src16 = layer.forward_tokens(src32, src16, pos32, mask_u8)
        return src32, src16


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.dropout = float(dropout)
        assert activation == "relu" and not normalize_before
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj"):
            setattr(self, n, nn.Linear(d_model, d_model))
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout, vdim=d_model)
        for n in ("ca_qcontent_proj", "ca_qpos_proj", "ca_kcontent_proj", "ca_kpos_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            setattr(self, n, nn.Linear(d_model, d_model))
        self.cross_attn = MultiheadAttention(d_model * 2, nhead, dropout=dropout, vdim=d_model)
        self.nhead = nhead
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before

    def memory_side(self, memory16, pos16, is_first):
        """k_content(+k_pos at layer 0), v, k_pos of the encoder memory (transformer.py:391-406) -- pass invariant."""
        lin = ops.linear
        v = lin(memory16, self.ca_v_proj.weight, self.ca_v_proj.bias)
        if is_first:
            kpos32 = lin(pos16, self.ca_kpos_proj.weight, self.ca_kpos_proj.bias, out_f32=True)
            kpos16 = ops.cast_bf16(kpos32)
            kc = lin(memory16, self.ca_kcontent_proj.weight, self.ca_kcontent_proj.bias, residual=kpos32)
        else:
            kpos16 = lin(pos16, self.ca_kpos_proj.weight, self.ca_kpos_proj.bias)
            kc = lin(memory16, self.ca_kcontent_proj.weight, self.ca_kcontent_proj.bias)
        return kc, v, kpos16

    def forward_tokens(self, tgt32, tgt16, mem_side, mask_u8, qpos16, qsine16, is_first, passes=1):
        """forward_post (transformer.py:355-427).  tgt [B,P*Q,D] (P decoder passes side by side); qpos16 [P*Q,D] (batch
        invariant); qsine16 [B,P*Q,D]."""
        lin = ops.linear
        B, PQ, D = tgt16.shape
        # ---- self-attention: q = Wqc tgt + Wqp qpos, k likewise, v = Wv tgt  (:368-381)
        qp = lin(qpos16, self.sa_qpos_proj.weight, self.sa_qpos_proj.bias, out_f32=True)        # [Q,D] fp32
        kp = lin(qpos16, self.sa_kpos_proj.weight, self.sa_kpos_proj.bias, out_f32=True)
        q = lin(tgt16, self.sa_qcontent_proj.weight, self.sa_qcontent_proj.bias, residual=qp)
        k = lin(tgt16, self.sa_kcontent_proj.weight, self.sa_kcontent_proj.bias, residual=kp)
        v = lin(tgt16, self.sa_v_proj.weight, self.sa_v_proj.bias)
        if passes > 1:                                  # queries attend to each other within their own pass only
            Q = PQ // passes
            a = self.self_attn.core(q.view(B * passes, Q, D), k.view(B * passes, Q, D), v.view(B * passes, Q, D)).view(B, PQ, D)
        else:
            a = self.self_attn.core(q, k, v)
        x = _res_proj(a, self.self_attn.out_proj, tgt32, self.dropout, self.training)
        t32, t16 = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, want_f32=True)
        # ---- conditional cross-attention (:389-423)
        kc, vv, kpos16 = mem_side
        if is_first:
            qpp = lin(qpos16, self.ca_qpos_proj.weight, self.ca_qpos_proj.bias, out_f32=True)
            qc = lin(t16, self.ca_qcontent_proj.weight, self.ca_qcontent_proj.bias, residual=qpp)
        else:
            qc = lin(t16, self.ca_qcontent_proj.weight, self.ca_qcontent_proj.bias)
        qs = lin(qsine16, self.ca_qpos_sine_proj.weight, self.ca_qpos_sine_proj.bias)
        a = self.cross_attn.core(qc, kc, vv, mask_u8=mask_u8, q2=qs, k2=kpos16)       # scale = (2*dh)^-1/2 via embed_dim = 2*d_model
        x = _res_proj(a, self.cross_attn.out_proj, t32, self.dropout, self.training)
        t32, t16 = ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, want_f32=True)
        x = _res_ffn(self, t32, t16, self.dropout, self.training)
        return ops.layernorm(x, self.norm3.weight, self.norm3.bias, self.norm3.eps, want_f32=True)


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False, d_model=256):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        assert return_intermediate and norm is not None
        self.d_model = d_model
        self.query_scale = MLP(d_model, d_model, d_model, 2)
        self.ref_point_head = MLP(d_model, d_model, 2, 2)
        for layer_id in range(num_layers - 1):
            self.layers[layer_id + 1].ca_qpos_proj = None                      # transformer.py:203-204

    def forward_tokens(self, memory16, pos16, mask_u8, query_embeds, batch, mem_cache=None):
        """TransformerDecoder.forward (transformer.py:206-250) for P query-embedding tables at once (one decoder pass of the
        reference each) -> per pass: hs fp32 [L,B,Q,D] (+ .tokens16) and reference points [B,Q,2]."""
        if isinstance(query_embeds, torch.Tensor):
            query_embeds = [query_embeds]
        mem_cache = {} if mem_cache is None else mem_cache
        P = len(query_embeds)
        B, (Q, D) = batch, query_embeds[0].shape
        query_embed = torch.cat(list(query_embeds), 0) if P > 1 else query_embeds[0]     # [P*Q, D]
        PQ = P * Q
        qpos16 = ops.cast_bf16(query_embed)                                     # batch invariant (:127 .repeat)
        ref = self.ref_point_head(qpos16).sigmoid()                             # [P*Q,2]  (:216-217)
        sine = ops.query_sine_embed(ref, self.d_model)                          # [P*Q,D]
        out32 = torch.zeros((B, PQ, D), dtype=torch.float32, device=query_embed.device)     # tgt = 0 (:129)
        out16 = torch.zeros((B, PQ, D), dtype=torch.bfloat16, device=query_embed.device)
        inter32, inter16 = [], []
        for layer_id, layer in enumerate(self.layers):
            if layer_id == 0:
                qsine = sine.unsqueeze(0).expand(B, -1, -1)                     # pos_transformation = 1 (:223-224)
            else:
                qsine = sine.unsqueeze(0) * self.query_scale(out16)            # (:226-231)
            qsine16 = ops.cast_bf16(qsine.contiguous())
            if layer_id not in mem_cache:
                mem_cache[layer_id] = layer.memory_side(memory16, pos16, layer_id == 0)
            out32, out16 = layer.forward_tokens(out32, out16, mem_cache[layer_id], mask_u8, qpos16, qsine16, layer_id == 0, passes=P)
            n32, n16 = ops.layernorm(out32, self.norm.weight, self.norm.bias, self.norm.eps, want_f32=True)     # shared norm (:239)
            inter32.append(n32)
            inter16.append(n16)
        hs_list, ref_list = [], []
        for p in range(P):                                                      # un-interleave the passes (the stack is the copy)
            sl = slice(p * Q, (p + 1) * Q)
            hs = torch.stack([t[:, sl] for t in inter32])
            hs.tokens16 = torch.stack([t[:, sl] for t in inter16])
            hs_list.append(hs)
            ref_list.append(ref[sl].unsqueeze(0).expand(B, -1, -1))
        return hs_list, ref_list


class Transformer(nn.Module):
    def __init__(self, d_model=512, nhead=8, num_queries=300, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False, args=None, num_refines=1,
                 drloc=False):
        super().__init__()
        assert not drloc, "drloc branch is dead code in the reference (SURVEY §2)"
        encoder_layer = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.encoder = TransformerEncoder(encoder_layer, num_encoder_layers, None)
        decoder_layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.decoder = TransformerDecoder(decoder_layer, num_decoder_layers, nn.LayerNorm(d_model), return_intermediate=return_intermediate_dec,
                                          d_model=d_model)
        self.H = None
        self.W = None
        self._reset_parameters()
        self.d_model = d_model
        self.nhead = nhead
        self.drloc = drloc
        self.dec_layers = num_decoder_layers
        self.num_queries = num_queries
        self.num_refines = num_refines

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    @staticmethod
    def _tokens(t):
        """[B,D,h,w] -> token-major fp32 / bf16 (reuses the copies the producer attached, else transposes + casts)."""
        t32 = getattr(t, "tokens32", None)
        t16 = getattr(t, "tokens16", None)
        if t32 is None:
            t32 = t.flatten(2).transpose(1, 2).contiguous().float()
        if t16 is None:
            t16 = ops.cast_bf16(t32)
        return t32, t16

    def forward(self, src, mask, query_embed, pos_embed, queries_embed_refine=None):
        """forward_refine / forward_non_refine (transformer.py:94-160).  Returns (hs, references): lists over
        refine passes when queries_embed_refine is given, else a single pair."""
        B = src.shape[0]
        src32, src16 = self._tokens(src)
        pos32, pos16 = self._tokens(pos_embed)
        mask_u8 = mask.flatten(1).to(torch.uint8).contiguous()
        mem32, mem16 = self.encoder.forward_tokens(src32, src16, pos32, mask_u8)
        if queries_embed_refine is None:
            hs, refs = self.decoder.forward_tokens(mem16, pos16, mask_u8, [query_embed], B)
            return hs[0], refs[0]
        # forward_refine (:147-155): the num_refines + 1 decoder passes run side by side (see the module docstring)
        return self.decoder.forward_tokens(mem16, pos16, mask_u8, [query_embed] + [qe.weight for qe in queries_embed_refine], B)


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def build_transformer(args):
    return Transformer(d_model=args.hidden_dim, dropout=args.dropout, nhead=args.nheads, num_queries=args.num_queries,
                       dim_feedforward=args.dim_feedforward, num_encoder_layers=args.enc_layers, num_decoder_layers=args.dec_layers,
                       normalize_before=args.pre_norm, return_intermediate_dec=True)
