"""models/position_encoding.py of the reference: PositionEmbeddingSine (:21-57) on the sm_100a kernel
spe_sine_pos_2d.  Output keeps the reference layout [B, D, h, w]; the token-major fp32/bf16 copies the
transformer consumes are attached as `.tokens32 / .tokens16`."""
import math

import torch
from torch import nn

from .. import ops
from ..util.misc import NestedTensor


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if not normalize or temperature != 10000 or (scale is not None and abs(scale - 2 * math.pi) > 1e-9):
            raise NotImplementedError("spe_b200 implements the configuration the reference builds: normalize=True, T=1e4, scale=2pi")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi

    def forward(self, tensor_list: NestedTensor):
        x, mask = tensor_list.tensors, tensor_list.mask
        assert mask is not None
        B, _, h, w = x.shape
        D = 2 * self.num_pos_feats
        pos32, pos16 = ops.sine_pos_2d(mask.to(torch.uint8), D)
        pos = pos32.transpose(1, 2).reshape(B, D, h, w)
        pos.tokens32, pos.tokens16 = pos32, pos16
        return pos


def build_position_encoding(args):
    N_steps = args.hidden_dim // 2
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(N_steps, normalize=True)
    raise ValueError(f"not supported {args.position_embedding}")
