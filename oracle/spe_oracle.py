"""CPU oracle for the SPE hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 PyTorch (CPU, eager, autograd) *restatement* of the reference's
forward/criterion math for the path named in BASELINE.json.  It is the checker for the CUDA
product in `spe_b200/` and the CPU baseline that `bench.py --impl reference` times.  Nothing
under `spe_b200/` may import it; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline/reference legs do.

It is written functionally over a flat ``params`` dict whose keys are the reference's
``state_dict`` keys, so the very same tensors can be loaded into the real
reference (``tests/golden/make_golden.py`` does that in the build container, where
/root/reference exists) and into the CUDA modules.

Parity pinning: the reference ships no tests / golden vectors.  The oracle is
pinned instead against outputs of the *unmodified reference itself*, run in the build container
through ``oracle/ref_shim.py``; those outputs are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.  ``scipy.optimize.linear_sum_assignment`` (un-vendored, unpinned by
the reference; 1.18.1 in this image) is restated in ``oracle/lsap.c`` and pinned against the
installed scipy.

Reference lines followed (all under /root/reference):
  models/cait.py:344-416 (talking-heads block), :91-139,:311-328 (class-attention block),
  :518-528 (patch embed), :588-670 (TSCAM_cait forward), models/cait_backbone.py:87-109,
  models/position_encoding.py:37-57, models/transformer.py:21-49,:122-160,:206-250,:275-288,
  :355-427, models/attention.py:274-385, models/conditional_detr.py:68-116,:225-319,:399-494,
  :504-561, models/matcher.py:41-87, util/box_ops.py:18-74, util/misc.py:440-455,:477-480.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class SPEConfig:
    """Hyper-parameters of TSCAM_cait + conditional DETR (direct construction)."""
    embed_dim: int = 192          # D  (also the DETR d_model, cait_backbone.py:84-85)
    depth: int = 24               # talking-heads blocks
    num_heads: int = 4            # backbone heads
    img_classes: int = 2          # C_img: extra class tokens / cls_head_multi_cls width
    patch: int = 16
    layer_to_det: int = 23        # tap *after* block index == layer_to_det (cait.py:629-630)
    depth_token_only: int = 2
    mlp_ratio: float = 4.0
    pos_grid: Tuple[int, int] = (50, 84)   # pos_embed grid after finetune_det (cait.py:572-586)
    det_heads: int = 8            # DETR heads (main.py:75)
    ffn: int = 2048
    enc_layers: int = 6
    dec_layers: int = 6
    num_queries: int = 10
    det_classes: int = 3          # width of class_embed (= "num_classes" of ConditionalDETR_Refine)
    num_refines: int = 1
    ln_eps_backbone: float = 1e-6
    ln_eps_detr: float = 1e-5
    two_branch: bool = False      # TSCAM_cait_two_branch (cait.py:674-831): blocks_det branch, norm_det, std-reweighted CAMs;
                                  # the tap is then taken after block layer_to_det - 1 (cait.py:776-777)

    @property
    def d_model(self) -> int:
        return self.embed_dim


CFG1 = SPEConfig()  # BASELINE configs[0]: XXS24, 10 queries, 2 classes, 224x224
CFG2 = SPEConfig(embed_dim=384, depth=24, num_heads=8, img_classes=80, num_queries=300, det_classes=81)
CFG4 = SPEConfig(embed_dim=768, depth=36, num_heads=16, img_classes=80, num_queries=300, det_classes=81,
                 layer_to_det=35)


def tiny_config(**kw) -> SPEConfig:
    """A small configuration with every code path present; used for fast parity tests."""
    base = dict(embed_dim=128, depth=2, num_heads=2, img_classes=3, layer_to_det=1, pos_grid=(6, 7),
                det_heads=8, ffn=128, enc_layers=1, dec_layers=2, num_queries=12, det_classes=4)
    base.update(kw)
    return SPEConfig(**base)


# --------------------------------------------------------------------------------------------
# parameter recipe (deterministic, reference key names; App. B)
# --------------------------------------------------------------------------------------------
def param_shapes(cfg: SPEConfig) -> Dict[str, Tuple[int, ...]]:
    D, H, C = cfg.embed_dim, cfg.num_heads, cfg.img_classes
    Dh = int(D * cfg.mlp_ratio)
    s: Dict[str, Tuple[int, ...]] = {}
    bb = "backbone.0.body."
    s[bb + "cls_token"] = (1, 1, D)
    s[bb + "pos_embed"] = (1, cfg.pos_grid[0] * cfg.pos_grid[1], D)
    s[bb + "extra_cls_token"] = (1, C, D)
    s[bb + "patch_embed.proj.weight"] = (D, 3, cfg.patch, cfg.patch)
    s[bb + "patch_embed.proj.bias"] = (D,)

    def lin(prefix, out_f, in_f):
        s[prefix + ".weight"] = (out_f, in_f)
        s[prefix + ".bias"] = (out_f,)

    def ln(prefix, dim):
        s[prefix + ".weight"] = (dim,)
        s[prefix + ".bias"] = (dim,)

    for i in range(cfg.depth):
        p = f"{bb}blocks.{i}."
        s[p + "gamma_1"] = (D,)
        s[p + "gamma_2"] = (D,)
        ln(p + "norm1", D)
        ln(p + "norm2", D)
        lin(p + "attn.qkv", 3 * D, D)
        lin(p + "attn.proj", D, D)
        lin(p + "attn.proj_l", H, H)
        lin(p + "attn.proj_w", H, H)
        lin(p + "mlp.fc1", Dh, D)
        lin(p + "mlp.fc2", D, Dh)
    for i in range(cfg.depth_token_only):
        p = f"{bb}blocks_token_only.{i}."
        s[p + "gamma_1"] = (D,)
        s[p + "gamma_2"] = (D,)
        ln(p + "norm1", D)
        ln(p + "norm2", D)
        for n in ("q", "k", "v", "proj"):
            lin(p + "attn." + n, D, D)
        lin(p + "mlp.fc1", Dh, D)
        lin(p + "mlp.fc2", D, Dh)
    ln(bb + "norm", D)
    lin(bb + "head", C, D)
    lin(bb + "cls_head", 1, D)
    lin(bb + "cls_head_multi_cls", C, D)
    if cfg.two_branch:
        for j in range(cfg.depth - cfg.layer_to_det):                 # cait.py:707-713
            p = f"{bb}blocks_det.{j}."
            s[p + "gamma_1"] = (D,)
            s[p + "gamma_2"] = (D,)
            ln(p + "norm1", D)
            ln(p + "norm2", D)
            lin(p + "attn.qkv", 3 * D, D)
            lin(p + "attn.proj", D, D)
            lin(p + "attn.proj_l", H, H)
            lin(p + "attn.proj_w", H, H)
            lin(p + "mlp.fc1", Dh, D)
            lin(p + "mlp.fc2", D, Dh)
        ln(bb + "norm_det", D)
    else:
        ln(bb + "norm_to_det", D)

    Dd, Fd = cfg.d_model, cfg.ffn
    for i in range(cfg.enc_layers):
        p = f"transformer.encoder.layers.{i}."
        s[p + "self_attn.in_proj_weight"] = (3 * Dd, Dd)
        s[p + "self_attn.in_proj_bias"] = (3 * Dd,)
        lin(p + "self_attn.out_proj", Dd, Dd)
        lin(p + "linear1", Fd, Dd)
        lin(p + "linear2", Dd, Fd)
        ln(p + "norm1", Dd)
        ln(p + "norm2", Dd)
    for i in range(cfg.dec_layers):
        p = f"transformer.decoder.layers.{i}."
        names = ["sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj",
                 "self_attn.out_proj", "ca_qcontent_proj", "ca_kcontent_proj", "ca_kpos_proj",
                 "ca_v_proj", "ca_qpos_sine_proj", "cross_attn.out_proj"]
        if i == 0:
            names.insert(7, "ca_qpos_proj")
        for n in names:
            lin(p + n, Dd, Dd)
        lin(p + "linear1", Fd, Dd)
        lin(p + "linear2", Dd, Fd)
        for n in ("norm1", "norm2", "norm3"):
            ln(p + n, Dd)
    ln("transformer.decoder.norm", Dd)
    lin("transformer.decoder.query_scale.layers.0", Dd, Dd)
    lin("transformer.decoder.query_scale.layers.1", Dd, Dd)
    lin("transformer.decoder.ref_point_head.layers.0", Dd, Dd)
    lin("transformer.decoder.ref_point_head.layers.1", 2, Dd)
    for r in range(cfg.num_refines + 1):
        lin(f"class_embed.{r}", cfg.det_classes, Dd)
        lin(f"bbox_embed.{r}.layers.0", Dd, Dd)
        lin(f"bbox_embed.{r}.layers.1", Dd, Dd)
        lin(f"bbox_embed.{r}.layers.2", 4, Dd)
    s["query_embed.weight"] = (cfg.num_queries, Dd)
    for r in range(cfg.num_refines):
        s[f"queries_embed_refine.{r}.weight"] = (cfg.num_queries, Dd)
    return s


def make_params(cfg: SPEConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic random parameters with *signal on every path* (SURVEY §4 item 3).

    Not the reference initialiser (that needs timm); a recipe both sides can regenerate: one CPU
    generator, keys in sorted order.  LayerScale gammas are U(.05,.2), LN weights ~1, biases small,
    matrices N(0, fan_in^-1/2 * .7), class_embed bias = -log(99) + noise.
    """
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for k, shp in sorted(param_shapes(cfg).items()):
        if k.endswith("gamma_1") or k.endswith("gamma_2"):
            t = torch.rand(shp, generator=g) * 0.15 + 0.05
        elif ".norm" in k or k.endswith("norm.weight") or k.endswith("norm.bias") or "norm_to_det" in k:
            if k.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shp, generator=g)
            else:
                t = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("in_proj_bias") or k.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
            if k.startswith("class_embed"):
                t = t - math.log(99.0)
        elif len(shp) >= 2 and (k.endswith("weight") or k.endswith("in_proj_weight")):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if "embed.weight" in k or "embed_refine" in k:      # nn.Embedding tables
                t = torch.randn(shp, generator=g)
            elif ".proj_l." in k or ".proj_w." in k:                 # H x H head mixes: near identity
                t = torch.eye(shp[0]) + 0.3 * torch.randn(shp, generator=g) / math.sqrt(shp[0])
            else:
                t = torch.randn(shp, generator=g) * (0.7 / math.sqrt(fan_in))
        else:                                                    # cls_token / pos_embed / extra_cls_token
            t = 0.02 * torch.randn(shp, generator=g)
        out[k] = t.float().contiguous()
    return out


def make_inputs(cfg: SPEConfig, batch: int, height: int, width: int, seed: int = 0,
                max_gt: int = 3, repeat: int = 1, with_scores: bool = False):
    """Seeded synthetic images + targets.  ``repeat`` = hung_match_ratio exact copies."""
    g = torch.Generator().manual_seed(1000 + seed)
    images = torch.randn(batch, 3, height, width, generator=g)
    targets = []
    for _ in range(batch):
        n = int(torch.randint(1, max_gt + 1, (1,), generator=g))
        c = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        boxes = torch.cat([c, wh], 1)
        labels = torch.randint(1, cfg.det_classes, (n,), generator=g)
        img_label = torch.zeros(cfg.img_classes)
        img_label[(labels - 1).clamp(max=cfg.img_classes - 1)] = 1
        t = {"labels": labels.repeat_interleave(repeat), "boxes": boxes.repeat_interleave(repeat, 0),
             "img_label": img_label}
        if with_scores:
            t["scores"] = (torch.rand(n, generator=g) * 0.8 + 0.1).repeat_interleave(repeat)
        targets.append(t)
    return images, targets


# --------------------------------------------------------------------------------------------
# backbone  (models/cait.py)
# --------------------------------------------------------------------------------------------
def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p.get(name + ".bias"))


def _ln(p, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def talking_heads_attention(p, pre, x, H):
    """cait.py:374-393.  q scaled before QK^T; head mixes have biases; softmax between them."""
    B, N, D = x.shape
    dh = D // H
    qkv = _lin(p, pre + "qkv", x).reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * dh ** -0.5, qkv[1], qkv[2]
    s = q @ k.transpose(-2, -1)                                            # [B,H,N,N]
    s = torch.einsum("gh,bhij->bgij", p[pre + "proj_l.weight"], s) + p[pre + "proj_l.bias"].view(1, H, 1, 1)
    s = s.softmax(-1)
    s = torch.einsum("gh,bhij->bgij", p[pre + "proj_w.weight"], s) + p[pre + "proj_w.bias"].view(1, H, 1, 1)
    o = (s @ v).transpose(1, 2).reshape(B, N, D)
    return _lin(p, pre + "proj", o)


def mlp_gelu(p, pre, x):
    return _lin(p, pre + "fc2", F.gelu(_lin(p, pre + "fc1", x)))


def class_attention(p, pre, u, H, T):
    """cait.py:111-139: queries = first T tokens, keys/values = all tokens; returns (out, attn map)."""
    B, N, D = u.shape
    dh = D // H
    q = _lin(p, pre + "q", u[:, :T]).reshape(B, T, H, dh).permute(0, 2, 1, 3) * dh ** -0.5
    k = _lin(p, pre + "k", u).reshape(B, N, H, dh).permute(0, 2, 1, 3)
    v = _lin(p, pre + "v", u).reshape(B, N, H, dh).permute(0, 2, 1, 3)
    a = (q @ k.transpose(-2, -1)).softmax(-1)                              # [B,H,T,N]
    o = (a @ v).transpose(1, 2).reshape(B, T, D)
    return _lin(p, pre + "proj", o), a


def std_reweighting(cam: torch.Tensor) -> torch.Tensor:
    """TSCAM_cait_two_branch.std_reweighting (cait.py:801-806): cam [B,H,C,N] per-head class-attention maps -> [B,C,N];
    head weight = the head's (unbiased) std over the N keys, min-max normalised over the heads."""
    std = cam.std(dim=-1, keepdim=True)
    std = std - std.min(dim=1, keepdim=True)[0]
    std = std / std.max(dim=1, keepdim=True)[0]
    return (cam * std).sum(1)


def tscam_forward(p, cfg: SPEConfig, images: torch.Tensor):
    """TSCAM_cait.forward (cait.py:615-670) -> dict(x_logits, x_cls_logits, cams_cls, x_patch[B,D,h,w])."""
    bb = "backbone.0.body."
    B, _, Hi, Wi = images.shape
    h, w = Hi // cfg.patch, Wi // cfg.patch
    D, C, eps = cfg.embed_dim, cfg.img_classes, cfg.ln_eps_backbone
    x = F.conv2d(images, p[bb + "patch_embed.proj.weight"], p[bb + "patch_embed.proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)                                       # [B,N,D]
    pe = p[bb + "pos_embed"].transpose(1, 2).reshape(1, D, *cfg.pos_grid)
    pe = F.interpolate(pe, size=(h, w), mode="bicubic", align_corners=False).flatten(2).transpose(1, 2)
    x = x + pe
    x_feat = None
    for i in range(cfg.depth):
        pre = f"{bb}blocks.{i}."
        x = x + p[pre + "gamma_1"] * talking_heads_attention(p, pre + "attn.", _ln(p, pre + "norm1", x, eps), cfg.num_heads)
        x = x + p[pre + "gamma_2"] * mlp_gelu(p, pre + "mlp.", _ln(p, pre + "norm2", x, eps))
        if not cfg.two_branch and i == cfg.layer_to_det:
            x_feat = _ln(p, bb + "norm_to_det", x, eps)
        if cfg.two_branch and i + 1 == cfg.layer_to_det:                   # cait.py:776-777
            x_feat = x
    if cfg.two_branch:                                                     # cait.py:779-782
        for j in range(cfg.depth - cfg.layer_to_det):
            pre = f"{bb}blocks_det.{j}."
            x_feat = x_feat + p[pre + "gamma_1"] * talking_heads_attention(p, pre + "attn.", _ln(p, pre + "norm1", x_feat, eps), cfg.num_heads)
            x_feat = x_feat + p[pre + "gamma_2"] * mlp_gelu(p, pre + "mlp.", _ln(p, pre + "norm2", x_feat, eps))
        x_feat = _ln(p, bb + "norm_det", x_feat, eps)
    cls = torch.cat([p[bb + "cls_token"].expand(B, -1, -1), p[bb + "extra_cls_token"].expand(B, -1, -1)], 1)
    T = 1 + C
    amap0 = None
    for i in range(cfg.depth_token_only):
        pre = f"{bb}blocks_token_only.{i}."
        u = torch.cat([cls, x], 1)
        o, amap = class_attention(p, pre + "attn.", _ln(p, pre + "norm1", u, eps), cfg.num_heads, T)
        if i == 0:
            amap0 = amap
        cls = cls + p[pre + "gamma_1"] * o
        cls = cls + p[pre + "gamma_2"] * mlp_gelu(p, pre + "mlp.", _ln(p, pre + "norm2", cls, eps))
    xa = _ln(p, bb + "norm", torch.cat([cls, x], 1), eps)
    x_logits = _lin(p, bb + "cls_head", xa[:, 1:1 + C]).squeeze(-1)
    x_cls_logits = _lin(p, bb + "cls_head_multi_cls", xa[:, 0])
    if cfg.two_branch:
        cams = std_reweighting(amap0[:, :, 1:1 + C, 1 + C:]).reshape(B, C, h, w)      # cait.py:827-828
    else:
        cams = amap0.mean(1)[:, 1:1 + C, 1 + C:].reshape(B, C, h, w)
    x_patch = x_feat.transpose(1, 2).reshape(B, D, h, w)
    return {"x_logits": x_logits, "x_cls_logits": x_cls_logits, "cams_cls": cams, "x_patch": x_patch}


# --------------------------------------------------------------------------------------------
# position encodings
# --------------------------------------------------------------------------------------------
def sine_pos_2d(mask: torch.Tensor, d_model: int) -> torch.Tensor:
    """PositionEmbeddingSine(normalize=True) (position_encoding.py:37-57): mask [B,h,w] -> [B,D,h,w]."""
    npf = d_model // 2
    nm = ~mask
    y = nm.cumsum(1, dtype=torch.float32)
    x = nm.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    t = torch.arange(npf, dtype=torch.float32)
    t = 10000 ** (2 * (t // 2) / npf)
    px, py = x[..., None] / t, y[..., None] / t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), 4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), 4).flatten(3)
    return torch.cat((py, px), 3).permute(0, 3, 1, 2)


def query_sine_embed(ref_xy: torch.Tensor, d_model: int) -> torch.Tensor:
    """gen_sineembed_for_position (transformer.py:35-49); note the hard-coded /128. ref [...,2] -> [...,D]."""
    n = d_model // 2
    t = torch.arange(n, dtype=torch.float32)
    t = 10000 ** (2 * (t // 2) / 128)
    px = (ref_xy[..., 0] * (2 * math.pi))[..., None] / t
    py = (ref_xy[..., 1] * (2 * math.pi))[..., None] / t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), -1).flatten(-2)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), -1).flatten(-2)
    return torch.cat((py, px), -1)


# --------------------------------------------------------------------------------------------
# conditional-DETR transformer (batch-first restatement of the sequence-first reference)
# --------------------------------------------------------------------------------------------
def mha_core(q, k, v, H, key_padding_mask=None):
    """attention.py:274-378: q [B,L,E], k [B,S,E], v [B,S,Ev]; scale = (E/H)^-1/2; -inf key padding."""
    B, L, E = q.shape
    S, Ev = k.shape[1], v.shape[2]
    dq, dv = E // H, Ev // H
    qh = (q * dq ** -0.5).reshape(B, L, H, dq).transpose(1, 2)
    kh = k.reshape(B, S, H, dq).transpose(1, 2)
    vh = v.reshape(B, S, H, dv).transpose(1, 2)
    a = qh @ kh.transpose(-2, -1)
    if key_padding_mask is not None:
        a = a.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    a = a.softmax(-1)
    return (a @ vh).transpose(1, 2).reshape(B, L, Ev)


def encoder_layer(p, pre, src, pos, mask, cfg: SPEConfig):
    """TransformerEncoderLayer.forward_post (transformer.py:275-288) with nn.MultiheadAttention math."""
    Dd = cfg.d_model
    Wi, bi = p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"]
    qk_in = src + pos
    q = F.linear(qk_in, Wi[:Dd], bi[:Dd])
    k = F.linear(qk_in, Wi[Dd:2 * Dd], bi[Dd:2 * Dd])
    v = F.linear(src, Wi[2 * Dd:], bi[2 * Dd:])
    a = _lin(p, pre + "self_attn.out_proj", mha_core(q, k, v, cfg.det_heads, mask))
    src = _ln(p, pre + "norm1", src + a, cfg.ln_eps_detr)
    f = _lin(p, pre + "linear2", F.relu(_lin(p, pre + "linear1", src)))
    return _ln(p, pre + "norm2", src + f, cfg.ln_eps_detr)


def _mlp_relu(p, pre, x, n):
    for i in range(n):
        x = _lin(p, f"{pre}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def decoder_layer(p, pre, tgt, memory, pos, query_pos, qsine, mask, is_first, cfg: SPEConfig):
    """TransformerDecoderLayer.forward_post (transformer.py:355-427), batch-first."""
    H, Dd, eps = cfg.det_heads, cfg.d_model, cfg.ln_eps_detr
    dh = Dd // H
    q = _lin(p, pre + "sa_qcontent_proj", tgt) + _lin(p, pre + "sa_qpos_proj", query_pos)
    k = _lin(p, pre + "sa_kcontent_proj", tgt) + _lin(p, pre + "sa_kpos_proj", query_pos)
    v = _lin(p, pre + "sa_v_proj", tgt)
    t2 = _lin(p, pre + "self_attn.out_proj", mha_core(q, k, v, H))
    tgt = _ln(p, pre + "norm1", tgt + t2, eps)

    qc = _lin(p, pre + "ca_qcontent_proj", tgt)
    kc = _lin(p, pre + "ca_kcontent_proj", memory)
    v = _lin(p, pre + "ca_v_proj", memory)
    kp = _lin(p, pre + "ca_kpos_proj", pos)
    if is_first:
        qc = qc + _lin(p, pre + "ca_qpos_proj", query_pos)
        kc = kc + kp
    B, Q, _ = qc.shape
    S = kc.shape[1]
    qs = _lin(p, pre + "ca_qpos_sine_proj", qsine)
    qcat = torch.cat([qc.reshape(B, Q, H, dh), qs.reshape(B, Q, H, dh)], 3).reshape(B, Q, 2 * Dd)
    kcat = torch.cat([kc.reshape(B, S, H, dh), kp.reshape(B, S, H, dh)], 3).reshape(B, S, 2 * Dd)
    t2 = _lin(p, pre + "cross_attn.out_proj", mha_core(qcat, kcat, v, H, mask))
    tgt = _ln(p, pre + "norm2", tgt + t2, eps)
    f = _lin(p, pre + "linear2", F.relu(_lin(p, pre + "linear1", tgt)))
    return _ln(p, pre + "norm3", tgt + f, eps)


def decoder_forward(p, cfg: SPEConfig, memory, pos, mask, query_embed):
    """TransformerDecoder.forward (transformer.py:206-250) -> hs [L,B,Q,D], reference points [B,Q,2]."""
    B = memory.shape[0]
    qpos = query_embed.unsqueeze(0).expand(B, -1, -1)
    ref = _mlp_relu(p, "transformer.decoder.ref_point_head", qpos, 2).sigmoid()        # [B,Q,2]
    out = torch.zeros_like(qpos)
    inter = []
    for l in range(cfg.dec_layers):
        qs = query_sine_embed(ref, cfg.d_model)
        if l > 0:
            qs = qs * _mlp_relu(p, "transformer.decoder.query_scale", out, 2)
        out = decoder_layer(p, f"transformer.decoder.layers.{l}.", out, memory, pos, qpos, qs, mask, l == 0, cfg)
        inter.append(_ln(p, "transformer.decoder.norm", out, cfg.ln_eps_detr))
    return torch.stack(inter), ref


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def model_forward(p, cfg: SPEConfig, images: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """ConditionalDETR_Refine.forward (conditional_detr.py:68-116) -> {refine_idx: out dict}."""
    B, _, Hi, Wi = images.shape
    if mask is None:
        mask = torch.zeros(B, Hi, Wi, dtype=torch.bool)
    feats = tscam_forward(p, cfg, images)
    xp = feats["x_patch"]
    h, w = xp.shape[-2:]
    m = F.interpolate(mask[None].float(), size=(h, w)).to(torch.bool)[0]               # cait_backbone.py:92
    pos = sine_pos_2d(m, cfg.d_model)
    src = xp.flatten(2).transpose(1, 2)                                                # [B,N,D]
    posf = pos.flatten(2).transpose(1, 2)
    mf = m.flatten(1)
    mem = src
    for l in range(cfg.enc_layers):
        mem = encoder_layer(p, f"transformer.encoder.layers.{l}.", mem, posf, mf, cfg)
    embeds = [p["query_embed.weight"]] + [p[f"queries_embed_refine.{r}.weight"] for r in range(cfg.num_refines)]
    out = {}
    for r, qe in enumerate(embeds):
        hs, ref = decoder_forward(p, cfg, mem, posf, mf, qe)
        rb = inverse_sigmoid(ref)
        t = _mlp_relu(p, f"bbox_embed.{r}", hs, 3)
        t = torch.cat([t[..., :2] + rb, t[..., 2:]], -1)
        boxes = t.sigmoid()
        logits = _lin(p, f"class_embed.{r}", hs)
        o = {"pred_logits": logits[-1], "pred_boxes": boxes[-1], **feats, "x_patch_mask": m}
        o["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]
        out[r] = o
    return out


# --------------------------------------------------------------------------------------------
# box ops, matcher, criterion
# --------------------------------------------------------------------------------------------
def box_cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def box_iou(a, b):
    """util/box_ops.py:33-46 -> (iou [N,M], union [N,M])."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b - inter
    return inter / union, union


def generalized_box_iou(a, b):
    """util/box_ops.py:49-74."""
    assert (a[:, 2:] >= a[:, :2]).all() and (b[:, 2:] >= b[:, :2]).all()
    iou, union = box_iou(a, b)
    lt = torch.min(a[:, None, :2], b[:, :2])
    rb = torch.max(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


def match_cost(logits, boxes, tgt_ids, tgt_boxes, w_class=2.0, w_bbox=5.0, w_giou=2.0):
    """matcher.py:62-82 for ONE image: logits [Q,C], boxes [Q,4] -> cost [Q,G] (fp32, same op order)."""
    prob = logits.sigmoid()
    neg = (1 - 0.25) * (prob ** 2.0) * (-(1 - prob + 1e-8).log())
    posc = 0.25 * ((1 - prob) ** 2.0) * (-(prob + 1e-8).log())
    c_class = posc[:, tgt_ids] - neg[:, tgt_ids]
    c_bbox = torch.cdist(boxes, tgt_boxes, p=1)
    c_giou = -generalized_box_iou(box_cxcywh_to_xyxy(boxes), box_cxcywh_to_xyxy(tgt_boxes))
    return w_bbox * c_bbox + w_class * c_class + w_giou * c_giou


def lsap(cost) -> Tuple[torch.Tensor, torch.Tensor]:
    """linear_sum_assignment oracle: the installed scipy (what the reference calls, matcher.py:86)."""
    from scipy.optimize import linear_sum_assignment
    i, j = linear_sum_assignment(cost.detach().cpu().numpy())
    return torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)


@torch.no_grad()
def hungarian_match(pred_logits, pred_boxes, targets, weights=(2.0, 5.0, 2.0), solver=lsap):
    """HungarianMatcher.forward (matcher.py:41-87), per image (block diagonal only)."""
    out = []
    for b, t in enumerate(targets):
        c = match_cost(pred_logits[b], pred_boxes[b], t["labels"], t["boxes"], *weights)
        out.append(solver(c))
    return out


def focal_loss(logits, onehot, num_boxes, weights, alpha, gamma):
    """weighted_sigmoid_focal_loss (conditional_detr.py:468-494)."""
    prob = logits.sigmoid()
    ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
    p_t = (prob * onehot + (1 - prob) * (1 - onehot)).clamp(1e-5, 1 - 1e-5)
    loss = weights * ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * onehot + (1 - alpha) * (1 - onehot)) * loss
    return loss.mean(1).sum() / num_boxes


def _loss_labels(logits, targets, indices, num_boxes, alpha, gamma, log, refine):
    B, Q, C = logits.shape
    onehot = torch.zeros(B, Q, C)
    wts = torch.ones(B, Q, C)
    if refine:                                                      # conditional_detr.py:523-529
        for b, t in enumerate(targets):
            wts[b] = t["scores"].mean()
    src_l, tgt_l = [], []
    for b, (I, J) in enumerate(indices):
        lab = targets[b]["labels"][J]
        keep = lab < C                                              # label == C is the dropped no-object column
        onehot[b, I[keep], lab[keep]] = 1
        if refine:
            wts[b, I, :] = (targets[b]["scores"][J].unsqueeze(-1) * 3).clamp(max=1.0)
        src_l.append(logits[b, I])
        tgt_l.append(lab)
    out = {"loss_ce": focal_loss(logits, onehot, num_boxes, wts, alpha, gamma) * Q}
    if log:                                                          # util/misc.py:440-455
        src, tg = torch.cat(src_l), torch.cat(tgt_l)
        if tg.numel() == 0:
            acc = torch.zeros([])
        else:
            acc = (src.detach().argmax(-1) == tg).float().sum() * (100.0 / tg.numel())
        out["class_error"] = 100 - acc
    return out


def _loss_boxes(boxes, targets, indices, num_boxes, refine):
    src = torch.cat([boxes[b, I] for b, (I, _) in enumerate(indices)])
    tgt = torch.cat([targets[b]["boxes"][J] for b, (_, J) in enumerate(indices)])
    l1 = (src - tgt).abs()
    giou = 1 - torch.diag(generalized_box_iou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt)))
    if refine:                                                      # conditional_detr.py:548-559
        wt = torch.cat([targets[b]["scores"][J] for b, (_, J) in enumerate(indices)])
        l1 = l1 * wt.reshape(-1, 1)
        giou = giou * wt
    return {"loss_bbox": l1.sum() / num_boxes, "loss_giou": giou.sum() / num_boxes}


def _loss_cardinality(logits, targets):
    n = torch.as_tensor([len(t["labels"]) for t in targets], dtype=torch.float32)
    card = (logits.detach().argmax(-1) != logits.shape[-1] - 1).sum(1).float()
    return {"cardinality_error": (card - n).abs().mean()}


def _loss_img_label(out, targets):
    y = torch.stack([t["img_label"] for t in targets]).float()
    return {"img_label_logits": F.binary_cross_entropy_with_logits(out["x_logits"], y),
            "img_label_logits_tokens": F.binary_cross_entropy_with_logits(out["x_cls_logits"], y)}


def criterion_forward(out, targets, losses=("labels", "boxes", "cardinality"), alpha=0.25, gamma=2.0,
                      match_weights=(2.0, 5.0, 2.0), refine=False, world_size=1, solver=lsap,
                      return_indices=False):
    """SetCriterion.forward in eval mode (conditional_detr.py:399-466; the training-only GT jitter is
    RNG-driven and is fed pre-expanded through ``targets``).  ``refine`` selects
    SetCriterionRefine's score-weighted variants (:504-561)."""
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)
    res: Dict[str, torch.Tensor] = {}
    all_idx = []

    def one(o, suffix, log):
        idx = hungarian_match(o["pred_logits"], o["pred_boxes"], targets, match_weights, solver)
        all_idx.append(idx)
        for name in losses:
            if name == "labels":
                d = _loss_labels(o["pred_logits"], targets, idx, num_boxes, alpha, gamma, log, refine)
            elif name == "boxes":
                d = _loss_boxes(o["pred_boxes"], targets, idx, num_boxes, refine)
            elif name == "cardinality":
                d = _loss_cardinality(o["pred_logits"], targets)
            elif name == "image_label":
                if suffix:
                    continue
                d = _loss_img_label(o, targets)
            else:
                raise ValueError(name)
            res.update({k + suffix: v for k, v in d.items()})

    one(out, "", True)
    for i, aux in enumerate(out.get("aux_outputs", [])):
        one(aux, f"_{i}", False)
    return (res, all_idx) if return_indices else res


def default_weight_dict(cfg: SPEConfig, cls=2.0, bbox=5.0, giou=2.0, img=1.0, img_tok=1.0):
    """conditional_detr.py:765-778 (coefficients = main.py defaults)."""
    base = {"loss_ce": cls, "loss_bbox": bbox, "img_label_logits": img, "img_label_logits_tokens": img_tok,
            "loss_giou": giou}
    wd = dict(base)
    for i in range(cfg.dec_layers - 1):
        wd.update({f"{k}_{i}": v for k, v in base.items()})
    return wd


def total_loss(loss_dict, weight_dict):
    return sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)


def train_step(p, cfg, images, targets, losses=("labels", "boxes", "cardinality"), gamma=2.0, refine_idx=0):
    """fwd + criterion(out[refine_idx]) + bwd; returns (outputs, loss dict, grads dict)."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = model_forward(p, cfg, images)
    ld, idx = criterion_forward(out[refine_idx], targets, losses, gamma=gamma, return_indices=True)
    wd = default_weight_dict(cfg)
    loss = total_loss(ld, wd)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return out, ld, idx, grads, loss.detach()
