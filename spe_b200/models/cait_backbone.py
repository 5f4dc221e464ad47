"""models/cait_backbone.py of the reference: Backbone (:67-94), Joiner (:98-109), build_backbone (:112-120)."""
import torch
import torch.nn.functional as F
from torch import nn

from ..util.misc import NestedTensor
from . import cait
from .position_encoding import build_position_encoding

# name -> factory; replaces timm.create_model(args.backbone, pretrained=True, ...) (cait_backbone.py:76-83).
# There is no network in this environment: weights are random-init, load a checkpoint with load_state_dict.
BACKBONES = {
    "TSCAM_cait_XXS24": (cait.tscam_cait_xxs24, 192),
    "TSCAM_cait_S24": (cait.tscam_cait_s24, 384),
    "TSCAM_cait_M36": (cait.tscam_cait_m36, 768),
    "TSCAM_cait_XXS36": (cait.tscam_cait_xxs36, 192),                            # cait.py:1595
    "TSCAM_cait_XXS36_Two_Branch": (cait.tscam_cait_xxs36_two_branch, 192),      # cait.py:1631; scripts/run_coco17.py:26, run_voc0712.py:28
    "TSCAM_cait_XXS24_Two_Branch": (cait.tscam_cait_xxs24_two_branch, 192),
}


class Backbone(nn.Module):
    def __init__(self, backbone, train_backbone=True, num_channels=None, return_interm_layers=False, args=None, body=None):
        super().__init__()
        if body is None:
            if args.dataset_file == "coco":
                num_classes = 90
            elif "voc" in args.dataset_file:
                num_classes = 20
            else:
                num_classes = getattr(args, "img_classes", 20)
            factory, num_channels = BACKBONES[backbone]
            body = factory(num_classes=num_classes, drop_rate=args.backbone_drop_rate, drop_path_rate=args.drop_path_rate,
                           attn_drop_rate=args.drop_attn_rate, layer_to_det=args.layer_to_det)
            args.hidden_dim = num_channels                      # cait_backbone.py:84-85
        self.body = body
        self.num_channels = num_channels if num_channels is not None else body.embed_dim

    def forward(self, tensor_list: NestedTensor):
        backbone_out = self.body(tensor_list)
        x = backbone_out["x_patch"]
        m = tensor_list.mask
        assert m is not None
        mask = F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]      # nearest, :92
        backbone_out["x_patch"] = NestedTensor(x, mask)
        return backbone_out


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)

    def forward(self, tensor_list: NestedTensor):
        backbone_out = self[0](tensor_list)
        x = backbone_out["x_patch"]
        pos = [self[1](x)]
        return backbone_out, pos


def build_backbone(args):
    backbone = Backbone(args.backbone, args.lr_backbone > 0, None, args.masks, args=args)
    backbone.body.finetune_det()
    position_embedding = build_position_encoding(args)
    model = Joiner(backbone, position_embedding)
    model.num_channels = backbone.num_channels
    return model
