"""models/cait.py of the reference, hot classes only, as parameter-holding shells over sm_100a kernels.

Same class names, constructor arguments and state_dict keys as the reference:
  Attention_talking_head (cait.py:344-393), LayerScale_Block (:396-416), Multi_Class_Attention (:91-139),
  LayerScale_Block_CA_MultiClass (:311-328), PatchEmbedMine (:518-528), TSCAM_cait (:531-670),
  TSCAM_cait_two_branch (:674-831, the backbone the reference's launch scripts train).
Internally activations are token-major: fp32 residual stream [B,N,D], bf16 GEMM operands.
Dropout / DropPath / attention dropout (scripts/run_coco17.py:30-32) are honoured in train() mode through un-fused routes
(ops.dropout / ops.drop_path / attention-probability dropout); p = 0 and eval() keep the fused epilogues (BASELINE configs).
"""
from functools import partial

import torch
import torch.nn as nn

from .. import ops
from ..util.misc import NestedTensor

__all__ = ["Mlp", "PatchEmbedMine", "Attention_talking_head", "LayerScale_Block", "Multi_Class_Attention",
           "LayerScale_Block_CA_MultiClass", "TSCAM_cait", "TSCAM_cait_two_branch", "tscam_cait_xxs24", "tscam_cait_xxs36", "tscam_cait_s24", "tscam_cait_m36",
           "tscam_cait_xxs24_two_branch", "tscam_cait_xxs36_two_branch"]


def _no_drop(**rates):
    for k, v in rates.items():
        if v:
            raise NotImplementedError(f"spe_b200: {k}={v} is not implemented (the benchmark configs use 0; SURVEY.md H6)")


class Mlp(nn.Module):
    """timm Mlp holder (fc1 -> GELU -> fc2)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.drop = float(drop)
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)


class PatchEmbedMine(nn.Module):
    """Conv2d(k=s=patch) patch embedding (cait.py:518-528): holder for `proj`; compute = ops.PatchEmbedFn."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.img_size = img_size
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size[1] // patch_size) * (img_size[0] // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Attention_talking_head(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.attn_drop, self.proj_drop = float(attn_drop), float(proj_drop)
        assert qk_scale is None
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.proj_l = nn.Linear(num_heads, num_heads)
        self.proj_w = nn.Linear(num_heads, num_heads)
        self.attention_map = None      # never materialised (nothing on this path reads it)

    def get_attention_map(self):
        return self.attention_map

    def core(self, y16):
        """LN output bf16 [B,N,D] -> attention output bf16 [B,N,D] (before proj)."""
        qkv = ops.linear(y16, self.qkv.weight, self.qkv.bias)
        return ops.talking_heads_attention(qkv, self.proj_l.weight, self.proj_l.bias, self.proj_w.weight, self.proj_w.bias, self.num_heads,
                                           drop_p=self.attn_drop if self.training else 0.0)


class LayerScale_Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, Attention_block=Attention_talking_head, Mlp_block=Mlp, init_values=1e-4):
        super().__init__()
        self.drop, self.drop_path = float(drop), float(drop_path)
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp_block(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)

    def forward(self, x):
        """x fp32 [B,N,D] -> fp32 [B,N,D]   (cait.py:413-416)"""
        # layernorm_res hands x through so its residual-path gradient is added inside the LayerNorm backward kernel
        y, xr = ops.layernorm_res(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        o = self.attn.core(y)
        if self.training and (self.drop > 0.0 or self.drop_path > 0.0):
            # x + drop_path(gamma_1 * proj_drop(proj(o))) ; x + drop_path(gamma_2 * drop(fc2(drop(gelu(fc1(LN x))))))   (cait.py:391,414-415)
            z = ops.dropout(ops.linear(o, self.attn.proj.weight, self.attn.proj.bias, out_f32=True), self.drop)
            x = xr + ops.drop_path(self.gamma_1 * z, self.drop_path)
            y, xr = ops.layernorm_res(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
            h = ops.dropout(torch.nn.functional.gelu(ops.linear(y, self.mlp.fc1.weight, self.mlp.fc1.bias)), self.drop)
            z = ops.dropout(ops.linear(h, self.mlp.fc2.weight, self.mlp.fc2.bias, out_f32=True), self.drop)
            return xr + ops.drop_path(self.gamma_2 * z, self.drop_path)
        x = ops.linear(o, self.attn.proj.weight, self.attn.proj.bias, residual=xr, gamma=self.gamma_1)
        y, xr = ops.layernorm_res(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return ops.ffn(y, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, residual=xr, gamma=self.gamma_2, act="gelu")


class Multi_Class_Attention(nn.Module):
    """Class attention with 1+C class tokens as queries over all tokens (cait.py:91-139)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, num_classes=20):
        super().__init__()
        _no_drop(attn_drop=attn_drop, proj_drop=proj_drop)
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.num_classes = num_classes
        self.attention_map = None

    def get_attention_map(self):
        return self.attention_map

    def core(self, u16, want_map):
        T = self.num_classes + 1
        q = ops.linear(u16[:, :T].contiguous(), self.q.weight, self.q.bias)
        k = ops.linear(u16, self.k.weight, self.k.bias)
        v = ops.linear(u16, self.v.weight, self.v.bias)
        if want_map:
            # True: head-mean f32 [B,T,T+N] (all cait.py:658-667 consumes);  "probs": per-head bf16 [B,H,T,ld] (std_reweighting, :827)
            o, amap = ops.attention(q, k, v, self.num_heads, self.scale, want_mean=want_map)
            self.attention_map = amap
            return o
        return ops.attention(q, k, v, self.num_heads, self.scale)


class LayerScale_Block_CA_MultiClass(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, Attention_block=Multi_Class_Attention, Mlp_block=Mlp, init_values=1e-4,
                 num_classes=20):
        super().__init__()
        _no_drop(drop=drop, drop_path=drop_path)
        self.norm1 = norm_layer(dim)
        self.attn = Attention_block(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                    num_classes=num_classes)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp_block(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.num_classes = num_classes

    def forward(self, x, x_cls, want_map=False):
        """x fp32 [B,N,D] patch tokens, x_cls fp32 [B,1+C,D] -> new x_cls   (cait.py:322-328)"""
        u = torch.cat((x_cls, x), dim=1)
        u16 = ops.layernorm(u, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        o = self.attn.core(u16, want_map)
        x_cls = ops.linear(o, self.attn.proj.weight, self.attn.proj.bias, residual=x_cls.contiguous(), gamma=self.gamma_1)
        y = ops.layernorm(x_cls, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return ops.ffn(y, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, residual=x_cls, gamma=self.gamma_2, act="gelu")


class TSCAM_cait(nn.Module):
    """TSCAM_cait (cait.py:531-670): talking-heads CaiT trunk, norm_to_det tap, 2 class-attention blocks,
    per-class heads and the block-0 class-attention map as CAMs."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm,
                 global_pool=None, block_layers=LayerScale_Block, block_layers_token=LayerScale_Block_CA_MultiClass,
                 Patch_layer=PatchEmbedMine, act_layer=nn.GELU, Attention_block=Attention_talking_head, Mlp_block=Mlp, init_scale=1e-4,
                 Attention_block_token_only=Multi_Class_Attention, Mlp_block_token_only=Mlp, depth_token_only=2, mlp_ratio_clstk=4.0,
                 layer_to_det=23):
        super().__init__()
        self.drop_rate = float(drop_rate)                      # pos_drop (cait.py:449,625)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = Patch_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        self.blocks = nn.ModuleList([
            block_layers(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                         attn_drop=attn_drop_rate, drop_path=drop_path_rate, norm_layer=norm_layer, act_layer=act_layer,
                         Attention_block=Attention_block, Mlp_block=Mlp_block, init_values=init_scale) for _ in range(depth)])
        self.blocks_token_only = nn.ModuleList([
            block_layers_token(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio_clstk, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=0.0,
                               attn_drop=0.0, drop_path=0.0, norm_layer=norm_layer, act_layer=act_layer,
                               Attention_block=Attention_block_token_only, Mlp_block=Mlp_block_token_only, init_values=init_scale,
                               num_classes=num_classes) for _ in range(depth_token_only)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()    # unused by forward; kept for state_dict parity
        self.extra_cls_token = nn.Parameter(torch.zeros(1, self.num_classes, self.embed_dim))
        self.cls_head = nn.Linear(self.embed_dim, 1)
        self.cls_head_multi_cls = nn.Linear(self.embed_dim, self.num_classes)
        self.patch_size = patch_size
        self.norm_to_det = norm_layer(embed_dim)
        self.layer_to_det = layer_to_det
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        nn.init.trunc_normal_(self.extra_cls_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def finetune_det(self, img_size=[800, 1344], use_checkpoint=False):
        """re-grid pos_embed to the detection canvas (cait.py:572-586): one-time, at construction."""
        p = self.patch_size
        ph, pw = self.img_size[0] // p, self.img_size[1] // p
        nh, nw = img_size[0] // p, img_size[1] // p
        pe = self.pos_embed.data.transpose(1, 2).reshape(1, self.embed_dim, ph, pw)
        pe = nn.functional.interpolate(pe, size=(nh, nw), mode="bicubic", align_corners=False)
        self.pos_embed = nn.Parameter(pe.flatten(2).transpose(1, 2).contiguous())
        self.img_size = tuple(img_size)

    def forward(self, tensor_list: NestedTensor):
        x_img, _ = tensor_list.decompose()
        B, _, H, W = x_img.shape
        p = self.patch_size
        h, w = H // p, W // p
        C = self.num_classes
        ph, pw = self.img_size[0] // p, self.img_size[1] // p
        pos = ops.BicubicTokensFn.apply(self.pos_embed, ph, pw, h, w)                                   # cait.py:588-600,623
        x = ops.PatchEmbedFn.apply(x_img.float(), self.patch_embed.proj.weight, self.patch_embed.proj.bias, pos, p)   # :618,624
        x = ops.dropout(x, self.drop_rate, self.training)                                                # pos_drop (:625)
        x_feat = None
        for i, blk in enumerate(self.blocks):                                                            # :627-630
            x = blk(x)
            if i == self.layer_to_det:
                ops.grad_milestone(x, "detector_grads_done")       # backward: the transformer / head gradients are final from here on
                x_feat, x_feat16 = ops.layernorm(x, self.norm_to_det.weight, self.norm_to_det.bias, self.norm_to_det.eps, want_f32=True)
        cls = torch.cat((self.cls_token.expand(B, -1, -1), self.extra_cls_token.expand(B, -1, -1)), dim=1)     # :620-622
        for i, blk in enumerate(self.blocks_token_only):                                                 # :635-637
            cls = blk(x, cls, want_map=(i == 0))
        # :643-645 norm(cat(cls, x)) is per token; only the class tokens are consumed (:653-654)
        xa = ops.layernorm(cls, self.norm.weight, self.norm.bias, self.norm.eps)
        x_logits = ops.linear(xa[:, 1:1 + C].contiguous(), self.cls_head.weight, self.cls_head.bias, out_f32=True).squeeze(-1)
        x_cls_logits = ops.linear(xa[:, 0].contiguous(), self.cls_head_multi_cls.weight, self.cls_head_multi_cls.bias, out_f32=True)
        amap = self.blocks_token_only[0].attn.get_attention_map()                                        # head mean, [B,1+C,1+C+N]
        cams_cls = amap[:, 1:1 + C, 1 + C:].reshape(B, C, h, w)                                          # :658-667
        x_patch = x_feat.transpose(1, 2).reshape(B, self.embed_dim, h, w)                                # view of the token-major tensor
        x_patch.tokens32, x_patch.tokens16 = x_feat, x_feat16
        return {"x_logits": x_logits, "x_cls_logits": x_cls_logits, "cams_cls": cams_cls, "x_patch": x_patch}


class TSCAM_cait_two_branch(TSCAM_cait):
    """TSCAM_cait_two_branch (cait.py:674-831): the trunk keeps feeding the class-attention / CAM branch, while a copy of its
    last depth - layer_to_det talking-heads blocks (`blocks_det`, initialised from the trunk by init_blocks_det_weight,
    :724-726) continues from the tap after block layer_to_det - 1 (:776-777) and, through `norm_det`, feeds the detector.
    CAMs are the std-reweighted per-head class-attention maps of class-attention block 0 (:801-806, :827).
    state_dict keys = the reference's: everything of the trunk + blocks_det.* + norm_det.* (no norm_to_det)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm,
                 global_pool=None, block_layers=LayerScale_Block, block_layers_token=LayerScale_Block_CA_MultiClass,
                 Patch_layer=PatchEmbedMine, act_layer=nn.GELU, Attention_block=Attention_talking_head, Mlp_block=Mlp, init_scale=1e-4,
                 Attention_block_token_only=Multi_Class_Attention, Mlp_block_token_only=Mlp, depth_token_only=2, mlp_ratio_clstk=4.0,
                 layer_to_det=23, **kwargs):
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, num_classes=num_classes, embed_dim=embed_dim, depth=depth,
                         num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate,
                         attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer, global_pool=global_pool,
                         block_layers=block_layers, block_layers_token=block_layers_token, Patch_layer=Patch_layer, act_layer=act_layer,
                         Attention_block=Attention_block, Mlp_block=Mlp_block, init_scale=init_scale,
                         Attention_block_token_only=Attention_block_token_only, Mlp_block_token_only=Mlp_block_token_only,
                         depth_token_only=depth_token_only, mlp_ratio_clstk=mlp_ratio_clstk, layer_to_det=layer_to_det)
        del self.norm_to_det
        assert 1 <= layer_to_det <= depth, "two-branch tap: after block layer_to_det - 1 (cait.py:776-777)"
        self.blocks_det = nn.ModuleList([
            block_layers(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                         attn_drop=attn_drop_rate, drop_path=drop_path_rate, norm_layer=norm_layer, act_layer=act_layer,
                         Attention_block=Attention_block, Mlp_block=Mlp_block, init_values=init_scale) for _ in range(layer_to_det, depth)])
        self.norm_det = norm_layer(embed_dim)
        self.blocks_det.apply(self._init_weights)
        self.norm_det.apply(self._init_weights)

    def init_blocks_det_weight(self):
        """cait.py:724-726: the detection branch starts as a copy of the trunk's last blocks."""
        for i in range(1, 1 + len(self.blocks_det)):
            self.blocks_det[-i].load_state_dict(self.blocks[-i].state_dict(), strict=True)

    def forward(self, tensor_list: NestedTensor):
        x_img, _ = tensor_list.decompose()
        B, _, H, W = x_img.shape
        p = self.patch_size
        h, w = H // p, W // p
        C = self.num_classes
        ph, pw = self.img_size[0] // p, self.img_size[1] // p
        pos = ops.BicubicTokensFn.apply(self.pos_embed, ph, pw, h, w)                                   # cait.py:731-745, :767
        x = ops.PatchEmbedFn.apply(x_img.float(), self.patch_embed.proj.weight, self.patch_embed.proj.bias, pos, p)
        x = ops.dropout(x, self.drop_rate, self.training)                                                # pos_drop (:771)
        x_feat = None
        for i, blk in enumerate(self.blocks):                                                            # :773-777
            x = blk(x)
            if i + 1 == self.layer_to_det:
                x_feat = x
        for blk in self.blocks_det:                                                                      # :779-780
            x_feat = blk(x_feat)
        ops.grad_milestone(x_feat, "detector_grads_done")
        x_feat, x_feat16 = ops.layernorm(x_feat, self.norm_det.weight, self.norm_det.bias, self.norm_det.eps, want_f32=True)   # :782
        cls = torch.cat((self.cls_token.expand(B, -1, -1), self.extra_cls_token.expand(B, -1, -1)), dim=1)
        for i, blk in enumerate(self.blocks_token_only):                                                 # :787-789
            cls = blk(x, cls, want_map=("probs" if i == 0 else False))
        xa = ops.layernorm(cls, self.norm.weight, self.norm.bias, self.norm.eps)                         # :796 (class tokens only are consumed)
        x_logits = ops.linear(xa[:, 1:1 + C].contiguous(), self.cls_head.weight, self.cls_head.bias, out_f32=True).squeeze(-1)
        x_cls_logits = ops.linear(xa[:, 0].contiguous(), self.cls_head_multi_cls.weight, self.cls_head_multi_cls.bias, out_f32=True)
        probs = self.blocks_token_only[0].attn.get_attention_map()                                       # per head, bf16 [B,heads,1+C,ld]
        cams_cls = ops.cam_std_reweight(probs, 1, C, 1 + C, h * w).reshape(B, C, h, w)                   # :827-828
        x_patch = x_feat.transpose(1, 2).reshape(B, self.embed_dim, h, w)
        x_patch.tokens32, x_patch.tokens16 = x_feat, x_feat16
        return {"x_logits": x_logits, "x_cls_logits": x_cls_logits, "cams_cls": cams_cls, "x_patch": x_patch}


def _tscam(embed_dim, depth, num_heads, init_scale, **kw):
    kw.setdefault("img_size", 384)
    kw.pop("pretrained", None)
    return TSCAM_cait(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, qkv_bias=True,
                      norm_layer=partial(nn.LayerNorm, eps=1e-6), init_scale=init_scale, depth_token_only=2, **kw)


def tscam_cait_xxs24(**kw):
    """TSCAM_cait_XXS24 (cait.py:1384-1400 hyper-parameters)."""
    return _tscam(192, 24, 4, 1e-5, **kw)


def tscam_cait_xxs36(**kw):
    """TSCAM_cait_XXS36 (cait.py:1595-1613 hyper-parameters)."""
    return _tscam(192, 36, 4, 1e-5, **kw)


def tscam_cait_s24(**kw):
    """TSCAM_cait with the S24 hyper-parameters of cait.py:1860-1880."""
    return _tscam(384, 24, 8, 1e-5, **kw)


def tscam_cait_m36(**kw):
    """TSCAM_cait with the M36 hyper-parameters of cait.py:1905-1925."""
    return _tscam(768, 36, 16, 1e-6, **kw)


def _tscam_two_branch(embed_dim, depth, num_heads, init_scale, **kw):
    kw.setdefault("img_size", 384)
    kw.pop("pretrained", None)
    m = TSCAM_cait_two_branch(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(nn.LayerNorm, eps=1e-6), init_scale=init_scale, depth_token_only=2, **kw)
    m.init_blocks_det_weight()                       # cait.py:1663 does this once the (pretrained) trunk weights are in place
    return m


def tscam_cait_xxs24_two_branch(**kw):
    """TSCAM_cait_two_branch with the XXS24 hyper-parameters."""
    return _tscam_two_branch(192, 24, 4, 1e-5, **kw)


def tscam_cait_xxs36_two_branch(**kw):
    """TSCAM_cait_XXS36_Two_Branch (cait.py:1631-1664), the backbone of scripts/run_coco17.py:26 / run_voc0712.py:28."""
    return _tscam_two_branch(192, 36, 4, 1e-5, **kw)
