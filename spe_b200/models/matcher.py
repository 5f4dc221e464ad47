"""models/matcher.py of the reference: HungarianMatcher (:20-87) on the GPU.

The cost matrix is built block-diagonally by a warp-per-query kernel (the reference builds the full
cross-batch [B*Q, sum G] matrix and keeps the diagonal) and the assignment is solved by a
batched shortest-augmenting-path kernel that reproduces scipy.optimize.linear_sum_assignment bit-exactly,
ties included -- no .cpu() round trip inside the criterion.  forward() keeps the reference
contract: a list of (idx_pred, idx_gt) CPU int64 tensor pairs."""
import torch
from torch import nn

from .. import criterion_ops as CO


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, match_ratio: int = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        self.match_ratio = match_ratio
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @property
    def weights(self):
        return (self.cost_class, self.cost_bbox, self.cost_giou)

    @torch.no_grad()
    def match_dense(self, outputs, packed):
        """device-side: i32 [B,Q] query -> gt index (or -1); no host synchronisation."""
        return CO.match(outputs["pred_logits"], outputs["pred_boxes"], packed, self.weights)

    @torch.no_grad()
    def forward(self, outputs, targets):
        packed = targets if isinstance(targets, CO.PackedTargets) else CO.pack_targets(targets, outputs["pred_logits"].device)
        return CO.indices_from_dense(self.match_dense(outputs, packed).cpu())


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou,
                            match_ratio=args.hung_match_ratio)
