"""Import the UNMODIFIED reference (/root/reference) on a modern torch without timm.

TEST / BASELINE INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: there the shim only finds the staged
copy under baseline/_ref/ (bench.py --impl reference); no test marked gpu may depend on it.  Used by tests/golden/make_golden.py to produce the committed golden vectors that
pin oracle/spe_oracle.py, and by tests that are skipped when the reference is absent.
Recipe = SURVEY.md Appendix D: a fake `timm` (Mlp / PatchEmbed / DropPath / trunc_normal_ /
register_model / create_model) and the `_LinearWithBias` alias removed from torch.
"""
import os
import sys
import types

import torch
import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _find_root():
    """$SPE_REFERENCE_ROOT, else /root/reference (build container), else the verbatim copy baseline/stage_reference.py makes
    (git-ignored; the only form in which the reference reaches the GPU box, for bench.py --impl reference)."""
    env = os.environ.get("SPE_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isdir(os.path.join(cand, "models")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1 - self.p
        m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x.div(keep) * m


_REGISTRY = {}


def _install():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_spe_shim", False):
        return
    import torch.nn.modules.linear as L
    if not hasattr(L, "_LinearWithBias"):
        L._LinearWithBias = L.NonDynamicallyQuantizableLinear

    def register_model(fn):
        _REGISTRY[fn.__name__] = fn
        return fn

    def create_model(name, pretrained=False, **kw):
        kw = {k: v for k, v in kw.items() if v is not None}
        return _REGISTRY[name](pretrained=False, **kw)

    timm = types.ModuleType("timm")
    timm._spe_shim = True
    models = types.ModuleType("timm.models")
    registry = types.ModuleType("timm.models.registry")
    vit = types.ModuleType("timm.models.vision_transformer")
    layers = types.ModuleType("timm.models.layers")
    registry.register_model = register_model
    models.create_model = create_model
    timm.create_model = create_model
    vit.Mlp, vit.PatchEmbed = _Mlp, _PatchEmbed
    vit._cfg = lambda url="", **kw: dict(url=url, **kw)
    layers.DropPath = _DropPath
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.to_2tuple = lambda x: (x, x) if not isinstance(x, tuple) else x
    timm.models, models.registry, models.vision_transformer, models.layers = models, registry, vit, layers
    for n, m in [("timm", timm), ("timm.models", models), ("timm.models.registry", registry),
                 ("timm.models.vision_transformer", vit), ("timm.models.layers", layers)]:
        sys.modules[n] = m


def load_reference():
    """Returns a namespace with the reference's hot-path modules (models.*, util.*)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install()
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.cait = importlib.import_module("models.cait")
    ns.cait_backbone = importlib.import_module("models.cait_backbone")
    ns.position_encoding = importlib.import_module("models.position_encoding")
    ns.transformer = importlib.import_module("models.transformer")
    ns.conditional_detr = importlib.import_module("models.conditional_detr")
    ns.matcher = importlib.import_module("models.matcher")
    ns.box_ops = importlib.import_module("util.box_ops")
    ns.misc = importlib.import_module("util.misc")
    return ns


def build_reference_model(cfg, params=None):
    """Direct construction of Joiner(Backbone-like(TSCAM_cait), PositionEmbeddingSine) +
    Transformer + ConditionalDETR_Refine with dropout 0, from an oracle SPEConfig."""
    from functools import partial
    ref = load_reference()
    # img_size such that the initial pos_embed has pos_grid tokens
    img_size = (cfg.pos_grid[0] * cfg.patch, cfg.pos_grid[1] * cfg.patch)
    body_cls = ref.cait.TSCAM_cait_two_branch if getattr(cfg, "two_branch", False) else ref.cait.TSCAM_cait
    body = body_cls(img_size=img_size, patch_size=cfg.patch, embed_dim=cfg.embed_dim, depth=cfg.depth,
                               num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio, qkv_bias=True,
                               norm_layer=partial(nn.LayerNorm, eps=cfg.ln_eps_backbone), init_scale=1e-5,
                               depth_token_only=cfg.depth_token_only, num_classes=cfg.img_classes,
                               layer_to_det=cfg.layer_to_det)
    body.img_size = img_size

    class _Backbone(nn.Module):          # cait_backbone.Backbone minus timm.create_model(pretrained=True)
        def __init__(self, body):
            super().__init__()
            self.body = body
            self.num_channels = cfg.embed_dim

        forward = ref.cait_backbone.Backbone.forward

    pos = ref.position_encoding.PositionEmbeddingSine(cfg.d_model // 2, normalize=True)
    joiner = ref.cait_backbone.Joiner(_Backbone(body), pos)
    joiner.num_channels = cfg.embed_dim
    tr = ref.transformer.Transformer(d_model=cfg.d_model, dropout=0.0, nhead=cfg.det_heads, num_queries=cfg.num_queries,
                                     dim_feedforward=cfg.ffn, num_encoder_layers=cfg.enc_layers,
                                     num_decoder_layers=cfg.dec_layers, normalize_before=False,
                                     return_intermediate_dec=True)
    model = ref.conditional_detr.ConditionalDETR_Refine(joiner, tr, num_classes=cfg.det_classes,
                                                       num_queries=cfg.num_queries, aux_loss=True,
                                                       num_refines=cfg.num_refines)
    if params is not None:
        model.load_state_dict(params, strict=True)
    return ref, model


def build_reference_criterion(ref, cfg, weight_dict, losses, gamma=2.0, refine=False, match_ratio=1):
    matcher = ref.matcher.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, match_ratio=match_ratio)
    cls = ref.conditional_detr.SetCriterionRefine if refine else ref.conditional_detr.SetCriterion
    crit = cls(cfg.det_classes, matcher=matcher, weight_dict=weight_dict, focal_alpha=0.25, losses=list(losses),
               gamma=gamma, box_jitter=0.1)
    crit.eval()
    return crit
