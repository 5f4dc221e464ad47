"""One iteration of the reference's refine training loop (engine.train_one_epoch_refine, engine.py:113-165) composed from the
device-side pieces of this package -- the integration proof that the rows of SURVEY section 8 fit together:

    outputs       = model(samples)                                              # {0: ..., 1: ...}
    pseudo_label  = get_pseudo_label_multi_boxes(outputs[0], samples, targets)  # CAM -> boxes on the device   (engine.py:117, N1)
    targets      += pseudo_label                                                # engine.py:119-120
    refine labels = PostProcessRefine(outputs[r]) per refine stage              # engine.py:122, :272-308
    loss_dict     = criterion(outputs[0], targets) + criterion_refine(outputs[r], refine labels[r])   (device-side jitter, N2)
    weights       = epoch schedule (engine.py:134-142)
    zero_grad, backward, clip_grad_norm_, optimizer.step                        # engine.py:161-165 (FlatAdamW, N4)

The data loader, logging (`utils.MetricLogger`, `reduce_dict`) and the epoch loop stay the caller's (out of scope).  The pseudo labels
depend on the step's own outputs, so this iteration is launched eagerly (engine.TrainStep captures the fixed-target step)."""
import copy

import torch

from . import pseudo_labels as PL


def refine_weight_dict(weight_dict, num_refines, header="ref"):
    """engine.get_refine_weight_dict (engine.py:260-268)"""
    wd = copy.deepcopy(weight_dict)
    for rf in range(1, num_refines + 1):
        for k in list(weight_dict.keys()):
            wd["%s_%d_%s" % (header, rf, k)] = weight_dict[k]
    return wd


@torch.no_grad()
def refinements_pseudo_label(outputs, targets, num_refines, postprocessors):
    """engine.get_refinements_pseudo_label / output_to_pseudo_label (engine.py:272-308)"""
    out = {}
    sizes = torch.stack([t["orig_size"] for t in targets], dim=0)
    for k, v in outputs.items():
        if k == num_refines:
            break
        res = postprocessors["bbox"](v, sizes, targets)
        labels = []
        for t, r in zip(targets, res):
            p = dict(t)
            p.update({"labels": r["labels"].detach().clone(), "boxes": r["boxes"].detach().clone(), "scores": r["scores"].detach().clone()})
            labels.append(p)
        out[k + 1] = labels
    return out


def refine_iteration(model, criterion, criterion_refine, postprocessors, optimizer, samples, targets, args, epoch, header="ref"):
    """-> (losses, loss_dict).  `optimizer`: spe_b200.optim.FlatAdamW (clip inside) or any torch optimizer (then clip_max_norm of args
    is applied with torch.nn.utils.clip_grad_norm_)."""
    tens = samples.tensors if hasattr(samples, "tensors") else samples
    outputs = model(samples)
    pseudo = PL.get_pseudo_label_multi_boxes(outputs[0], tens, targets, args)
    targets = [dict(t, **p) for t, p in zip(targets, pseudo)]
    refine = refinements_pseudo_label(outputs, targets, args.num_refines, postprocessors)
    loss_dict = criterion(outputs[0], targets)
    for rf, out in outputs.items():
        if rf == 0:
            continue
        for k, v in criterion_refine(out, refine[rf]).items():
            loss_dict["%s_%d_%s" % (header, rf, k)] = v
    wd = refine_weight_dict(criterion.weight_dict, args.num_refines, header)
    if epoch < 7:                                               # engine.py:134-137: image-level losses only
        for k in wd:
            if not ("img_label" in k or "drloc" in k):
                wd[k] = 0.0
    if epoch < 15:                                              # engine.py:139-142: refine heads switched on later
        for k in wd:
            if header in k:
                wd[k] = 0.0
    losses = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    optimizer.zero_grad()
    losses.backward()
    if hasattr(optimizer, "gbuf"):
        optimizer.step()                                        # clip_grad_norm_(max_norm) + AdamW, 4 launches
    else:
        if getattr(args, "clip_max_norm", 0) > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), args.clip_max_norm)
        optimizer.step()
    return losses.detach(), {k: v.detach() for k, v in loss_dict.items()}
