"""Pins oracle/spe_oracle.py against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import spe_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _run(gold):
    m = gold["meta"]
    cfg = O.SPEConfig(**m["cfg"])
    params = O.make_params(cfg, m["seed"])
    images, targets = O.make_inputs(cfg, m["batch"], m["height"], m["width"], seed=m["seed"], max_gt=m["max_gt"],
                                    repeat=m["repeat"], with_scores=m["refine_idx"] > 0)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = O.model_forward(p, cfg, images)
    ld, idx = O.criterion_forward(out[m["refine_idx"]], targets, m["losses"], gamma=m["gamma"],
                                  refine=m["refine_idx"] > 0, return_indices=True)
    loss = O.total_loss(ld, O.default_weight_dict(cfg))
    loss.backward()
    return cfg, p, out, ld, idx, loss


@pytest.mark.parametrize("name", ["tiny_det", "tiny_refine", "tiny_two_branch", "tiny_h16", "cfg1_xxs24_224", "cfg2_s24_640"])
def test_oracle_matches_reference_golden(golden_dir, name):
    gold = _load(golden_dir, name)
    cfg, p, out, ld, idx, loss = _run(gold)
    for r, g in gold["outputs"].items():
        torch.testing.assert_close(out[r]["pred_logits"], g["pred_logits"], rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(out[r]["pred_boxes"], g["pred_boxes"], rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(torch.stack([a["pred_logits"] for a in out[r]["aux_outputs"]]), g["aux_logits"],
                                   rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(torch.stack([a["pred_boxes"] for a in out[r]["aux_outputs"]]), g["aux_boxes"],
                                   rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(out[0]["x_logits"], gold["x_logits"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(out[0]["x_cls_logits"], gold["x_cls_logits"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(out[0]["cams_cls"], gold["cams_cls"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out[0]["x_patch"].sum(1), gold["x_patch_sum"], rtol=1e-4, atol=1e-4)
    assert set(ld) == set(gold["losses"])
    for k, v in gold["losses"].items():
        torch.testing.assert_close(ld[k], v, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss, gold["total_loss"], rtol=1e-4, atol=1e-5)
    # matcher indices: bit exact, every decoder level
    assert len(idx) == len(gold["indices"])
    for lvl_o, lvl_g in zip(idx, gold["indices"]):
        for (i, j), (gi, gj) in zip(lvl_o, lvl_g):
            assert torch.equal(i, gi) and torch.equal(j, gj)
    # gradients: fingerprints of every parameter + a few full tensors
    for k, fp in gold["grad_fingerprint"].items():
        gr = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        mine = torch.stack([gr.sum(), gr.norm()])
        torch.testing.assert_close(mine, fp, rtol=2e-3, atol=2e-5, msg=lambda s: f"{k}: {s}")
    for k, g in gold["grads"].items():
        gr = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        torch.testing.assert_close(gr, g, rtol=1e-3, atol=1e-6)


def test_param_recipe_covers_reference_state_dict(golden_dir):
    gold = _load(golden_dir, "tiny_det")
    cfg = O.SPEConfig(**gold["meta"]["cfg"])
    assert set(O.param_shapes(cfg)) == set(gold["grad_fingerprint"])
