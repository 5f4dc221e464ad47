"""PostProcessRefine / PostProcessRefineMulti against the reference's own classes (models/conditional_detr.py:641-715; CUDA only there:
they call .get_device()), imported through oracle/ref_shim.py from the staged copy baseline/_ref -- skipped where it does not exist."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_refine_postprocessors_match_the_reference_classes():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not staged (baseline/stage_reference.py)")
    from spe_b200.models.conditional_detr import PostProcess, PostProcessRefine, PostProcessRefineMulti
    ref = ref_shim.load_reference().conditional_detr
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(4)
    B, Q, C = 3, 120, 21
    logits, boxes = torch.randn(B, Q, C, generator=g).to(dev), (torch.rand(B, Q, 4, generator=g) * 0.5 + 0.1).to(dev)
    out = {"pred_logits": logits, "pred_boxes": boxes}
    sizes = torch.tensor([[480., 640.], [333., 500.], [800., 1333.]], device=dev)
    targets = [{"labels": torch.tensor([3, 7, 7, 20], device=dev)}, {"labels": torch.tensor([1], device=dev)}, {"labels": torch.tensor([5, 2, 19, 0], device=dev)}]
    for a, b in zip(PostProcess()(out, sizes), ref.PostProcess()(out, sizes)):
        assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"])
    for mine, theirs in ((PostProcessRefine(), ref.PostProcessRefine()), (PostProcessRefineMulti(), ref.PostProcessRefineMulti())):
        for a, b in zip(mine(out, sizes, targets), theirs(out, sizes, targets)):
            assert torch.equal(a["labels"].long(), torch.as_tensor(b["labels"]).long().reshape(-1))
            assert torch.equal(a["scores"], torch.as_tensor(b["scores"]).reshape(-1)) and torch.equal(a["boxes"], b["boxes"].reshape(-1, 4))
