"""SURVEY N1 on the GPU: csrc/cam_boxes.cu through spe_b200.pseudo_labels against the committed cv2 golden vectors (bit-exact integer
boxes and fp32 normalised boxes), the reference-shaped get_pseudo_label call, and -- where cv2 is importable -- fresh random maps
against the cv2-based oracle."""
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gold():
    return torch.load(os.path.join(HERE, "golden", "cam_boxes.pt"), weights_only=False)


def test_cam_boxes_match_cv2_golden():
    from spe_b200 import pseudo_labels as PL
    dev = torch.device("cuda")
    n = 0
    for case in _gold()["cases"]:
        pairs = torch.nonzero(case["img_label"] > 0)
        boxes, xyxy = PL.cam_boxes(case["cams"].to(dev), pairs, case["image_size"], case["cam_thr"], return_xyxy=True)
        want_xyxy = torch.cat(case["xyxy"]).to(torch.int32)
        want = torch.cat(case["boxes"])
        assert torch.equal(xyxy.cpu(), want_xyxy), (case["image_size"], case["cam_thr"], xyxy.cpu().tolist(), want_xyxy.tolist())
        assert torch.equal(boxes.cpu(), want)
        n += len(want)
    assert n >= 50


def test_get_pseudo_label_signature_and_order():
    from spe_b200 import pseudo_labels as PL
    dev = torch.device("cuda")
    case = _gold()["cases"][0]
    H, W = case["image_size"]
    outputs = {"cams_cls": case["cams"].to(dev)}
    samples = SimpleNamespace(tensors=torch.zeros(case["cams"].shape[0], 3, H, W, device=dev))
    targets = [{"img_label": case["img_label"][b].to(dev)} for b in range(case["cams"].shape[0])]
    args = SimpleNamespace(cam_thr=case["cam_thr"], num_classes=case["cams"].shape[1])
    got = PL.get_pseudo_label(outputs, samples, targets, args)
    assert len(got) == len(case["boxes"])
    for g, wb, wl in zip(got, case["boxes"], case["labels"]):
        assert g["boxes"].is_cuda and torch.equal(g["boxes"].cpu(), wb) and torch.equal(g["labels"].cpu(), wl)


def test_cam_boxes_random_maps_vs_cv2_oracle():
    pytest.importorskip("cv2")
    import numpy as np
    import cv2
    from oracle import cam_boxes as OC
    from spe_b200 import pseudo_labels as PL
    dev = torch.device("cuda")
    rng = np.random.default_rng(123)
    for (h, w, H, W, sig) in [(40, 40, 640, 640, 2.0), (40, 40, 640, 640, 0.8), (14, 14, 224, 224, 1.0), (32, 24, 512, 384, 1.5), (20, 20, 333, 500, 0.6)]:
        B, C = 3, 8
        cams = torch.from_numpy(np.stack([[cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), sig) for _ in range(C)] for _ in range(B)]))
        lab = torch.ones(B, C)
        for thr in (0.2, 0.5):
            pl, raw = OC.pseudo_labels(cams, lab, (H, W), cam_thr=thr)
            boxes, xyxy = PL.cam_boxes(cams.to(dev), torch.nonzero(lab > 0), (H, W), thr, return_xyxy=True)
            assert torch.equal(xyxy.cpu(), torch.cat(raw).to(torch.int32)), (h, w, H, W, sig, thr)
            assert torch.equal(boxes.cpu(), torch.cat([p["boxes"] for p in pl]))
