"""Direct construction of the detector from hyper-parameters (the reference's build(args)
hard-codes class counts and registered backbones; BASELINE's configs need direct construction)."""
from functools import partial

import torch
from torch import nn

from .models.cait import TSCAM_cait, TSCAM_cait_two_branch
from .models.cait_backbone import Backbone, Joiner
from .models.conditional_detr import ConditionalDETR_Refine, SetCriterion, SetCriterionRefine
from .models.matcher import HungarianMatcher
from .models.position_encoding import PositionEmbeddingSine
from .models.transformer import Transformer


def build_detector(cfg, device="cuda"):
    """cfg: any object with the attributes of oracle.spe_oracle.SPEConfig (embed_dim, depth, num_heads, img_classes,
    patch, layer_to_det, depth_token_only, mlp_ratio, pos_grid, det_heads, ffn, enc_layers, dec_layers, num_queries,
    det_classes, num_refines, ln_eps_backbone[, two_branch])."""
    img_size = (cfg.pos_grid[0] * cfg.patch, cfg.pos_grid[1] * cfg.patch)
    body_cls = TSCAM_cait_two_branch if getattr(cfg, "two_branch", False) else TSCAM_cait
    body = body_cls(img_size=img_size, patch_size=cfg.patch, embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads,
                      mlp_ratio=cfg.mlp_ratio, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=cfg.ln_eps_backbone), init_scale=1e-5,
                      depth_token_only=cfg.depth_token_only, num_classes=cfg.img_classes, layer_to_det=cfg.layer_to_det)
    body.img_size = img_size
    joiner = Joiner(Backbone(None, body=body), PositionEmbeddingSine(cfg.embed_dim // 2, normalize=True))
    joiner.num_channels = cfg.embed_dim
    tr = Transformer(d_model=cfg.embed_dim, dropout=0.0, nhead=cfg.det_heads, num_queries=cfg.num_queries, dim_feedforward=cfg.ffn,
                     num_encoder_layers=cfg.enc_layers, num_decoder_layers=cfg.dec_layers, normalize_before=False, return_intermediate_dec=True)
    model = ConditionalDETR_Refine(joiner, tr, num_classes=cfg.det_classes, num_queries=cfg.num_queries, aux_loss=True, num_refines=cfg.num_refines)
    return model.to(device)


def default_weight_dict(dec_layers, cls=2.0, bbox=5.0, giou=2.0, img=1.0, img_tok=1.0):
    base = {"loss_ce": cls, "loss_bbox": bbox, "img_label_logits": img, "img_label_logits_tokens": img_tok, "loss_giou": giou}
    wd = dict(base)
    for i in range(dec_layers - 1):
        wd.update({f"{k}_{i}": v for k, v in base.items()})
    return wd


def build_criterion(cfg, losses=("labels", "boxes", "cardinality"), gamma=2.0, refine=False, match_ratio=1, device="cuda"):
    matcher = HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, match_ratio=match_ratio)
    cls = SetCriterionRefine if refine else SetCriterion
    crit = cls(cfg.det_classes, matcher=matcher, weight_dict=default_weight_dict(cfg.dec_layers), focal_alpha=0.25, losses=list(losses),
               gamma=gamma, box_jitter=0.1)
    return crit.to(device)
