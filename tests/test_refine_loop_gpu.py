"""Integration: iterations of the reference's refine training loop (engine.py:113-165) built from this package's pieces --
build_model(args) with the script arguments, device-side CAM pseudo labels (N1), PostProcessRefine pseudo labels, both criteria in
train() mode with device-side GT jitter (N2), flat clip + AdamW (N4)."""
import argparse
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("epoch", [3, 20])
def test_refine_training_iterations(golden_dir, epoch):
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.models import build_model
    from spe_b200.optim import FlatAdamW
    from spe_b200.refine_loop import refine_iteration
    fix = json.load(open(os.path.join(golden_dir, "build_args.json")))
    d = dict(fix["cases"]["run_coco17"]["args"])
    d["device"] = "cuda"
    args = argparse.Namespace(**d)
    torch.manual_seed(0)
    model, criterion, criterion_refine, postprocessors, refine_postprocessors = build_model(args)
    dev = torch.device("cuda")
    model.to(dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma_" in n:
                p.fill_(0.1)                                       # LayerScale 1e-5 would freeze the backbone for a 6-step test
    model.train(); criterion.train(); criterion_refine.train()
    if getattr(args, "hungarian_multi", False):
        criterion.update_hung_match_ratio(args.hung_match_ratio)
        criterion_refine.update_hung_match_ratio(args.hung_match_ratio)
    gbuf = FlatGradBuffer(model.parameters())
    opt = FlatAdamW(model, gbuf, lr=args.lr, lr_backbone=args.lr_backbone, lr_cls_head=getattr(args, "lr_cls_head", args.lr_backbone),
                    weight_decay=args.weight_decay, clip_max_norm=args.clip_max_norm)
    g = torch.Generator().manual_seed(1)
    B, H, W = 2, 128, 160
    images = torch.randn(B, 3, H, W, generator=g).to(dev)
    targets = []
    for b in range(B):
        il = torch.zeros(90)
        il[torch.randint(0, 80, (2 + b,), generator=g)] = 1
        targets.append({"img_label": il.to(dev), "orig_size": torch.tensor([H, W], device=dev), "size": torch.tensor([H, W], device=dev)})
    p0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    vals = []
    for it in range(6):
        losses, ld = refine_iteration(model, criterion, criterion_refine, refine_postprocessors, opt, images, [dict(t) for t in targets], args, epoch)
        assert torch.isfinite(losses), (it, losses)
        vals.append(float(losses))
        for k in ("loss_ce", "loss_bbox", "loss_giou", "img_label_logits", "img_label_logits_tokens", "ref_1_loss_ce", "ref_1_loss_giou"):
            assert k in ld and torch.isfinite(ld[k]), k
    moved = {n: float((p.detach() - p0[n]).abs().max()) for n, p in model.named_parameters()}
    if epoch < 7:     # only the image-level heads (backbone side) receive gradient: the detector stays put apart from weight decay
        assert max(v for n, v in moved.items() if n.startswith("backbone")) > 1e-5
    else:
        assert max(v for n, v in moved.items() if n.startswith("transformer")) > 1e-5
        assert max(v for n, v in moved.items() if n.startswith("backbone")) > 1e-6
    assert int(opt.state[3]) == 6
    print("refine loop epoch %d: losses %s" % (epoch, ["%.4f" % v for v in vals]))
