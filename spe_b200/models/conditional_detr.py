"""models/conditional_detr.py of the reference: ConditionalDETR_Refine (:33-124), SetCriterion (:190-494),
SetCriterionRefine (:497-589), MLP (:626-638), build (:733-802) -- Python shells over libspe_b200.so.

Deviation kept explicit (SURVEY §8b note): forward returns a dict subclass holding {0: ..., 1: ...} that also
answers string keys from entry 0, so both engine.train_one_epoch (flat dict) and the *_refine loops work."""
import copy
import math

import torch
from torch import nn

from .. import criterion_ops as CO
from .. import ops
from ..util import box_ops
from ..util.misc import NestedTensor, inverse_sigmoid, nested_tensor_from_tensor_list
from .cait_backbone import build_backbone
from .matcher import build_matcher
from .transformer import MLP, build_transformer


class RefineOutputs(dict):
    """{refine_idx: out_dict}; string keys proxy to refine 0 (see module docstring)."""

    def __missing__(self, key):
        if isinstance(key, str) and 0 in self:
            return dict.__getitem__(self, 0)[key]
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (isinstance(key, str) and dict.__contains__(self, 0) and key in dict.__getitem__(self, 0))


class ConditionalDETR_Refine(nn.Module):
    def __init__(self, backbone, transformer, num_classes, num_queries, aux_loss=False, num_refines=1, drloc=False):
        super().__init__()
        self.num_queries = num_queries
        self.num_refines = num_refines
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.class_embed = nn.ModuleList([nn.Linear(hidden_dim, num_classes) for _ in range(num_refines + 1)])
        self.bbox_embed = nn.ModuleList([MLP(hidden_dim, hidden_dim, 4, 3) for _ in range(num_refines + 1)])
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.queries_embed_refine = nn.ModuleList([nn.Embedding(num_queries, hidden_dim) for _ in range(num_refines)])
        self.backbone = backbone
        self.aux_loss = aux_loss
        bias_value = -math.log((1 - 0.01) / 0.01)
        for class_embed in self.class_embed:
            class_embed.bias.data = torch.ones(num_classes) * bias_value
        for bbox_embed in self.bbox_embed:
            nn.init.constant_(bbox_embed.layers[-1].weight.data, 0)
            nn.init.constant_(bbox_embed.layers[-1].bias.data, 0)

    def forward(self, samples: NestedTensor):
        """conditional_detr.py:68-116."""
        if isinstance(samples, (list, torch.Tensor)):
            samples = nested_tensor_from_tensor_list(samples)
        if not samples.tensors.is_cuda:
            raise RuntimeError("spe_b200 runs on CUDA (sm_100a) only; move the samples to the GPU")
        _, _, H, W = samples.tensors.shape
        self.transformer.H = H // self.backbone[0].body.patch_size
        self.transformer.W = W // self.backbone[0].body.patch_size
        features, pos = self.backbone(samples)
        src, mask = features["x_patch"].decompose()
        assert mask is not None
        Hs, references = self.transformer(src, mask, self.query_embed.weight, pos[-1], queries_embed_refine=self.queries_embed_refine)
        out = RefineOutputs()
        for r in range(self.num_refines + 1):
            hs = Hs[r]                                        # fp32 [L,B,Q,D] (+ bf16 copy)
            hs16 = getattr(hs, "tokens16", None)
            if hs16 is None:
                hs16 = ops.cast_bf16(hs)
            ref_bs = inverse_sigmoid(references[r])           # [B,Q,2]
            tmp = self.bbox_embed[r](hs16)                    # fp32 [L,B,Q,4]
            coord = torch.cat([tmp[..., :2] + ref_bs, tmp[..., 2:]], -1).sigmoid()       # :104-106
            logits = ops.linear(hs16, self.class_embed[r].weight, self.class_embed[r].bias, out_f32=True)
            o = {"pred_logits": logits[-1], "pred_boxes": coord[-1], **features}
            if self.aux_loss:
                o["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], coord[:-1])]
            out[r] = o
        return out


class SetCriterion(nn.Module):
    """conditional_detr.py:190-494.  Matching + every loss run on the device with no host synchronisation; so does the GT
    jitter/repeat of training mode (:410-431; csrc/targets.cu) unless `device_jitter` is switched off."""
    refine = False

    def __init__(self, num_classes, matcher, weight_dict, focal_alpha, losses, gamma, box_jitter):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.losses = losses
        self.focal_alpha = focal_alpha
        self.gamma = gamma
        self.eos_coef = 0.1
        self.hung_match_ratio = getattr(matcher, "match_ratio", 1)
        self.box_jitter = box_jitter
        empty_weight = torch.ones(self.num_classes)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)

    def update_hung_match_ratio(self, ratio=5):
        assert hasattr(self.matcher, "match_ratio")
        self.matcher.match_ratio = ratio
        self.hung_match_ratio = ratio

    # ---- training-mode target expansion (:410-431), host/torch prep exactly as the reference ----
    def _jitter_repeat(self, targets):
        out = copy.deepcopy(targets)
        r = self.hung_match_ratio
        for t in out:
            reps = []
            for j in range(len(t["labels"])):
                box_j = t["boxes"][j].reshape(1, 4)
                scale = torch.cat([torch.empty((1000, 1), dtype=box_j.dtype, device=box_j.device).uniform_(1 - self.box_jitter, 1 + self.box_jitter)
                                   for _ in range(4)], dim=1)
                sb = scale * box_j
                a, b = box_ops.box_cxcywh_to_xyxy(sb), box_ops.box_cxcywh_to_xyxy(box_j)
                lt, rb = torch.max(a[:, :2], b[:, :2]), torch.min(a[:, 2:], b[:, 2:])
                wh = (rb - lt).clamp(min=0)
                inter = wh[:, 0] * wh[:, 1]
                iou = inter / ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter)
                keep = torch.where(iou > 0.7)[0]
                n = min(r - 1, keep.numel())
                rep = box_j.repeat(r, 1)
                rep[:n] = sb[keep[:n]]
                reps.append(rep)
            if reps:
                t["boxes"] = torch.cat(reps)
            t["labels"] = t["labels"].unsqueeze(1).repeat(1, r).reshape(-1)
            if "scores" in t:
                t["scores"] = t["scores"].unsqueeze(1).repeat(1, r).reshape(-1)
        return out

    def prepare_targets(self, targets):
        """the target list the losses see: jittered + repeated GT in training mode (:410-431), unchanged in eval mode.  Host loop with
        torch's RNG, exactly the reference's sequence of draws (used when device_jitter is off, and by callers that want the list)."""
        return self._jitter_repeat(targets) if self.training else targets

    # in training mode the jitter / repeat runs on the device (csrc/targets.cu), one launch instead of a Python loop over
    # images x boxes; `criterion.device_jitter = False` restores the reference's host loop and its torch RNG stream.
    device_jitter = True

    def jitter_rng(self, device):
        rngs = self.__dict__.setdefault("_jitter_rngs", {})
        key = str(device)
        if key not in rngs:
            rngs[key] = CO.JitterRng(device)
        return rngs[key]

    def prepare_packed(self, targets, device):
        """targets (list of dicts, or already packed) -> CO.PackedTargets on `device` as the losses see them."""
        if isinstance(targets, CO.PackedTargets):
            return targets
        if not self.training:
            return CO.pack_targets(targets, device)
        if not self.device_jitter or not torch.device(device).type == "cuda" or not any(len(t["labels"]) for t in targets):
            return CO.pack_targets(self.prepare_targets(targets), device)
        return CO.jitter_repeat(CO.pack_targets(targets, device), self.hung_match_ratio, self.box_jitter, self.jitter_rng(device))

    def loss_img_label(self, outputs, targets):
        if isinstance(targets, CO.PackedTargets):
            y = targets.img_label
        else:
            y = torch.stack([t["img_label"] for t in targets]).to(outputs["x_logits"].device).float()
        return {"img_label_logits": CO.BceLogitsFn.apply(outputs["x_logits"], y)[0],
                "img_label_logits_tokens": CO.BceLogitsFn.apply(outputs["x_cls_logits"], y)[0]}

    @torch.no_grad()
    def match_all_levels(self, outputs, packed):
        """Hungarian matching of the final + every auxiliary decoder level against pre-packed targets: i32 [L,B,Q] query -> gt map
        (device side, one cost launch + one batched assignment launch).  engine.TrainStep calls this for the two criteria of a
        step on two streams -- the assignment kernel is latency bound on a few dozen CTAs -- and hands the result to forward()."""
        if isinstance(outputs, RefineOutputs):
            outputs = outputs[0]
        levels = [outputs] + list(outputs.get("aux_outputs", []))
        return CO.match_levels(torch.stack([l["pred_logits"].detach() for l in levels]), torch.stack([l["pred_boxes"].detach() for l in levels]),
                               packed, self.matcher.weights)

    def forward(self, outputs, targets, _matches=None):
        """conditional_detr.py:399-466.  (_matches: optional result of match_all_levels for these outputs / targets.)"""
        if isinstance(outputs, RefineOutputs):
            outputs = outputs[0]
        dev = outputs["pred_logits"].device
        if isinstance(targets, CO.PackedTargets):
            tg = T = targets            # pre-packed (engine.TrainStep: static buffers; jitter/repeat already applied by prepare_targets)
        else:
            tg = T = self.prepare_packed(targets, dev)
        mw = self.matcher.weights
        det = tuple(l for l in self.losses if l in ("labels", "boxes", "cardinality"))
        losses = {}

        levels = [outputs] + list(outputs.get("aux_outputs", []))
        # all decoder levels are matched by ONE cost launch + ONE batched LSAP launch (L*B independent problems)
        r2g = _matches if _matches is not None else CO.match_levels(torch.stack([l["pred_logits"].detach() for l in levels]),
                                                                    torch.stack([l["pred_boxes"].detach() for l in levels]), T, mw)

        def level(i, o, suffix, log):
            res = CO.set_losses(o["pred_logits"], o["pred_boxes"], T, mw, self.focal_alpha, self.gamma, refine=self.refine, losses=det, log=log,
                                r2g=r2g[i])
            res.pop("_r2g")
            losses.update({k + suffix: v for k, v in res.items()})

        level(0, outputs, "", True)
        if "image_label" in self.losses:
            losses.update(self.loss_img_label(outputs, tg))
        for i in range(1, len(levels)):
            level(i, levels[i], f"_{i - 1}", False)
        return losses


class SetCriterionRefine(SetCriterion):
    """conditional_detr.py:497-589: focal / L1 / GIoU weighted by the pseudo-label scores."""
    refine = True


class PostProcess(nn.Module):
    """conditional_detr.py:592-623 (eval-time top-k box decoding; host-side torch, not on the training path)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, keep_queries=100):
        out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(out_logits) == len(target_sizes) and target_sizes.shape[1] == 2
        prob = out_logits.sigmoid()
        k = min(keep_queries, prob.shape[1] * prob.shape[2])
        topk_values, topk_indexes = torch.topk(prob.view(out_logits.shape[0], -1), k, dim=1)
        topk_boxes = topk_indexes // out_logits.shape[2]
        labels = topk_indexes % out_logits.shape[2]
        boxes = box_ops.box_cxcywh_to_xyxy(out_bbox).clamp(min=0)            # :611 clamps the corners at 0
        boxes = torch.gather(boxes, 1, topk_boxes.unsqueeze(-1).repeat(1, 1, 4))
        img_h, img_w = target_sizes.unbind(1)
        boxes = boxes * torch.stack([img_w, img_h, img_w, img_h], dim=1)[:, None, :]
        return [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(topk_values, labels, boxes)]


class PostProcessRefine(nn.Module):
    """conditional_detr.py:641-677: per image and per class PRESENT in targets[i]['labels'], the best-scoring query's (score, box,
    class) -- the pseudo labels of the refine stage (engine.py:122, :271-308).  Host-side glue on [B,Q,C] tensors (no_grad); the
    reference's double Python loop is replaced by one gather per image, results identical (classes in ascending order)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, targets=None):
        out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(out_logits) == len(target_sizes) and target_sizes.shape[1] == 2
        prob = out_logits.sigmoid()
        top_values, top_indexes = torch.max(prob, dim=1)                                        # [B,C]
        top_boxes = torch.gather(out_bbox, 1, top_indexes.unsqueeze(-1).repeat(1, 1, 4))        # [B,C,4]
        C = out_logits.shape[2]
        results = []
        for ii in range(len(targets)):
            present = torch.zeros(C, dtype=torch.bool, device=out_logits.device)
            lab = targets[ii]["labels"].to(out_logits.device).long()
            present[lab[(lab >= 0) & (lab < C)]] = True
            cls = torch.nonzero(present).flatten()
            results.append({"scores": top_values[ii][cls], "labels": cls, "boxes": top_boxes[ii][cls].reshape(-1, 4)})
        return results


class PostProcessRefineMulti(nn.Module):
    """conditional_detr.py:680-715: like PostProcessRefine, but every query whose class probability reaches half of the class's best
    probability is kept (queries in ascending order within a class)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, targets=None):
        out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(out_logits) == len(target_sizes) and target_sizes.shape[1] == 2
        prob = out_logits.sigmoid()
        top_values, _ = torch.max(prob, dim=1)
        keep = prob >= 0.5 * top_values.unsqueeze(1).expand_as(prob)                            # [B,Q,C]
        C = out_logits.shape[2]
        results = []
        for ii in range(len(targets)):
            present = torch.zeros(C, dtype=torch.bool, device=out_logits.device)
            lab = targets[ii]["labels"].to(out_logits.device).long()
            present[lab[(lab >= 0) & (lab < C)]] = True
            cq = torch.nonzero((keep[ii] & present.unsqueeze(0)).t())                          # rows (class, query), class-major ascending
            results.append({"scores": prob[ii, cq[:, 1], cq[:, 0]], "labels": cq[:, 0], "boxes": out_bbox[ii, cq[:, 1]]})
        return results


def build(args):
    """conditional_detr.py:733-802: (model, criterion, criterion_refine, postprocessors, refine_postprocessors)."""
    num_classes = 21 if args.dataset_file != "coco" else 91
    num_classes = getattr(args, "det_classes", num_classes)
    backbone = build_backbone(args)
    transformer = build_transformer(args)
    model = ConditionalDETR_Refine(backbone, transformer, num_classes=num_classes, num_queries=args.num_queries, aux_loss=args.aux_loss,
                                   num_refines=args.num_refines)
    matcher, matcher_refine = build_matcher(args), build_matcher(args)
    weight_dict = {"loss_ce": args.cls_loss_coef, "loss_bbox": args.bbox_loss_coef, "img_label_logits": args.img_label_loss_coef,
                   "img_label_logits_tokens": args.img_label_tokens_loss_coef, "loss_giou": args.giou_loss_coef}
    if args.aux_loss:
        aux = {}
        for i in range(args.dec_layers - 1):
            aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    criterion = SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, focal_alpha=args.focal_alpha,
                             losses=["labels", "boxes", "cardinality", "image_label"], gamma=args.focal_gamma, box_jitter=args.box_jitter)
    criterion_refine = SetCriterionRefine(num_classes, matcher=matcher_refine, weight_dict=weight_dict, focal_alpha=args.focal_alpha,
                                          losses=["labels", "boxes", "cardinality"], gamma=args.focal_gamma, box_jitter=args.box_jitter)
    device = torch.device(args.device)
    criterion.to(device)
    criterion_refine.to(device)
    return model, criterion, criterion_refine, {"bbox": PostProcess()}, {"bbox": PostProcessRefine()}       # :788-790
