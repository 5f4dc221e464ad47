"""PostProcessRefine / PostProcessRefineMulti (conditional_detr.py:641-715; the pseudo-label step of the refine training loop) against a
literal restatement of the reference's loops.  Host-side torch glue: runs on CPU."""
import torch

from spe_b200.models.conditional_detr import PostProcessRefine, PostProcessRefineMulti


def _ref_refine(out_logits, out_bbox, targets):
    prob = out_logits.sigmoid()
    top_values, top_indexes = torch.max(prob, dim=1)
    top_boxes = torch.gather(out_bbox, 1, top_indexes.unsqueeze(-1).repeat(1, 1, 4))
    res = []
    for ii in range(len(targets)):
        l, s, b = [], [], []
        for cc in range(out_logits.shape[2]):
            if cc in targets[ii]["labels"]:
                l.append(cc); s.append(top_values[ii][cc].reshape(-1)); b.append(top_boxes[ii][cc].reshape(1, -1))
        res.append({"scores": torch.cat(s), "labels": torch.tensor(l), "boxes": torch.cat(b)})
    return res


def _ref_multi(out_logits, out_bbox, targets):
    prob = out_logits.sigmoid()
    top_values, _ = torch.max(prob, dim=1)
    keep_idx = prob >= 0.5 * top_values.unsqueeze(1).expand_as(prob)
    res = []
    for ii in range(len(targets)):
        l, s, b = [], [], []
        for cc in range(out_logits.shape[2]):
            if cc in targets[ii]["labels"]:
                k = keep_idx[ii, :, cc].nonzero(as_tuple=False).reshape(-1)
                s.append(prob[ii, k, cc]); b.append(out_bbox[ii, k]); l += [cc] * k.shape[0]
        res.append({"scores": torch.cat(s), "labels": torch.tensor(l), "boxes": torch.cat(b)})
    return res


def test_postprocess_refine_matches_reference_loops():
    g = torch.Generator().manual_seed(3)
    B, Q, C = 3, 40, 21
    logits, boxes = torch.randn(B, Q, C, generator=g), torch.rand(B, Q, 4, generator=g)
    targets = [{"labels": torch.tensor([3, 7, 7, 20])}, {"labels": torch.tensor([1])}, {"labels": torch.tensor([5, 2, 19, 0])}]
    sizes = torch.ones(B, 2)
    for mod, ref in ((PostProcessRefine(), _ref_refine), (PostProcessRefineMulti(), _ref_multi)):
        got = mod({"pred_logits": logits, "pred_boxes": boxes}, sizes, targets)
        exp = ref(logits, boxes, targets)
        for a, b in zip(got, exp):
            assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"])


def test_postprocess_classes_match_the_reference_classes():
    """PostProcess against the reference's own class (models/conditional_detr.py:592-623),
    imported through oracle/ref_shim.py -- skipped where neither /root/reference nor the staged copy baseline/_ref exists."""
    import pytest
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    from spe_b200.models.conditional_detr import PostProcess
    ref = ref_shim.load_reference().conditional_detr
    g = torch.Generator().manual_seed(4)
    B, Q, C = 3, 120, 21
    logits, boxes = torch.randn(B, Q, C, generator=g), torch.rand(B, Q, 4, generator=g) * 0.5 + 0.1
    out = {"pred_logits": logits, "pred_boxes": boxes}
    sizes = torch.tensor([[480., 640.], [333., 500.], [800., 1333.]])
    targets = [{"labels": torch.tensor([3, 7, 7, 20])}, {"labels": torch.tensor([1])}, {"labels": torch.tensor([5, 2, 19, 0])}]
    for a, b in zip(PostProcess()(out, sizes), ref.PostProcess()(out, sizes)):
        assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"])
    # PostProcessRefine / PostProcessRefineMulti of the reference call .get_device(): CUDA only -> tests/test_postprocess_gpu.py
