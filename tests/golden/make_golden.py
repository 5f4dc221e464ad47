"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in the build container.

    python tests/golden/make_golden.py

Parameters and inputs come from the deterministic recipes in oracle/spe_oracle.py (make_params /
make_inputs), are loaded into the reference's own modules (`load_state_dict(strict=True)`) through
oracle/ref_shim.py, and the reference's forward + SetCriterion (eval mode) + backward
are run on CPU fp32.  What is stored: outputs, every loss, the matcher indices of every decoder level,
and per-parameter gradient (sum, L2) fingerprints plus a few full gradients.  The GPU box has no
reference tree: tests only read the .pt files written here.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, spe_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FULL_GRADS = ["backbone.0.body.blocks.0.attn.proj_l.weight", "backbone.0.body.blocks.0.attn.proj_w.bias",
              "backbone.0.body.blocks.1.gamma_1", "transformer.decoder.layers.0.ca_qpos_proj.weight",
              "class_embed.0.bias", "bbox_embed.0.layers.2.weight", "query_embed.weight"]


def run_case(name, cfg, batch, height, width, seed, losses, gamma, repeat, refine_idx=0, max_gt=3):
    torch.manual_seed(0)
    params = O.make_params(cfg, seed)
    images, targets = O.make_inputs(cfg, batch, height, width, seed=seed, max_gt=max_gt, repeat=repeat,
                                    with_scores=refine_idx > 0)
    ref, model = ref_shim.build_reference_model(cfg, params)
    model.train()
    wd = O.default_weight_dict(cfg)
    crit = ref_shim.build_reference_criterion(ref, cfg, wd, losses, gamma=gamma, refine=refine_idx > 0)
    out = model(images)
    o = out[refine_idx]
    ld = crit(o, targets)
    loss = sum(ld[k] * wd[k] for k in ld if k in wd)
    loss.backward()
    idx = [crit.matcher({"pred_logits": o["pred_logits"], "pred_boxes": o["pred_boxes"]}, targets)]
    for aux in o["aux_outputs"]:
        idx.append(crit.matcher(aux, targets))
    g = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in model.named_parameters()}
    gold = {
        "meta": dict(name=name, cfg=cfg.__dict__, batch=batch, height=height, width=width, seed=seed,
                     losses=list(losses), gamma=gamma, repeat=repeat, refine_idx=refine_idx, max_gt=max_gt,
                     torch=torch.__version__),
        "outputs": {r: {"pred_logits": out[r]["pred_logits"].detach(), "pred_boxes": out[r]["pred_boxes"].detach(),
                        "aux_logits": torch.stack([a["pred_logits"] for a in out[r]["aux_outputs"]]).detach(),
                        "aux_boxes": torch.stack([a["pred_boxes"] for a in out[r]["aux_outputs"]]).detach()}
                    for r in out},
        "x_logits": out[0]["x_logits"].detach(), "x_cls_logits": out[0]["x_cls_logits"].detach(),
        "cams_cls": out[0]["cams_cls"].detach(),
        "x_patch_sum": out[0]["x_patch"].tensors.detach().sum(1),
        "losses": {k: v.detach() for k, v in ld.items()},
        "total_loss": loss.detach(),
        "indices": idx,
        "grad_fingerprint": {k: torch.stack([v.sum(), v.norm()]) for k, v in g.items()},
        "grads": {k: g[k].clone() for k in FULL_GRADS if k in g},
    }
    path = os.path.join(HERE, name + ".pt")
    torch.save(gold, path)
    print(name, "loss", float(loss), "->", path, os.path.getsize(path) // 1024, "KiB")


def lsap_vectors():
    """scipy known-answer vectors for the assignment kernel (ties, both orientations, cfg5 shape)."""
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(7)
    cases = []
    shapes = [(1, 1), (1, 9), (9, 1), (5, 5), (12, 30), (30, 12), (300, 20), (300, 50), (64, 300), (300, 300), (300, 1000)]
    for nr, nc in shapes:
        for kind in ("normal", "int", "dupcol"):
            if kind == "normal":
                c = torch.randn(nr, nc, generator=g)
            elif kind == "int":
                c = torch.randint(0, 4, (nr, nc), generator=g).float()
            else:
                c = torch.randn(nr, max(1, (nc + 4) // 5), generator=g).repeat_interleave(5, 1)[:, :nc].contiguous()
            i, j = linear_sum_assignment(c.numpy())
            cases.append({"shape": (nr, nc), "kind": kind, "seed_cost_sum": float(c.double().sum()),
                          "rows": torch.as_tensor(i), "cols": torch.as_tensor(j)})
    torch.save({"generator_seed": 7, "shapes": shapes, "cases": cases}, os.path.join(HERE, "lsap_scipy.pt"))
    print("lsap_scipy.pt", len(cases), "cases")


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    if "--only-cfg2" in sys.argv:
        run_case("cfg2_s24_640", O.CFG2, 1, 640, 640, 0, ("labels", "boxes", "cardinality"), 2.0, repeat=5, max_gt=6)
        sys.exit(0)
    tiny = O.tiny_config()
    run_case("tiny_det", tiny, 2, 48, 64, 0, ("labels", "boxes", "cardinality"), 2.0, repeat=2)
    run_case("tiny_refine", tiny, 2, 48, 64, 1, ("labels", "boxes", "cardinality"), 0.5, repeat=1, refine_idx=1)
    # TSCAM_cait_two_branch (the scripts' backbone, cait.py:674-831): 3 trunk blocks, tap after block 1, 2 blocks_det
    run_case("tiny_two_branch", O.tiny_config(depth=3, layer_to_det=1, two_branch=True, num_heads=4), 2, 48, 64, 2,
             ("labels", "boxes", "cardinality"), 2.0, repeat=1)
    # CaiT-M36 head geometry (16 heads x 48, BASELINE configs[3]) on a 2-block trunk: exercises the H = 16 talking-heads kernels
    run_case("tiny_h16", O.tiny_config(embed_dim=768, num_heads=16, pos_grid=(4, 5)), 2, 48, 64, 4,
             ("labels", "boxes", "cardinality"), 2.0, repeat=1)
    run_case("cfg1_xxs24_224", O.CFG1, 1, 224, 224, 0, ("labels", "boxes", "cardinality"), 2.0, repeat=1, max_gt=2)
    # BASELINE configs[1], the benchmarked configuration (TSCAM-S24, 300 queries, 81 logits, 3x640x640), batch 1: the backbone
    # attention runs at its real shape (H = 8, N = 1600) -- ~1 min of CPU for the reference's forward + backward
    if "--no-cfg2" not in sys.argv:
        run_case("cfg2_s24_640", O.CFG2, 1, 640, 640, 0, ("labels", "boxes", "cardinality"), 2.0, repeat=5, max_gt=6)
    lsap_vectors()
