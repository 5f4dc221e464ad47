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


def test_multi_boxes_match_cv2_golden():
    """engine.get_pseudo_label_multi_boxes / cams_deit.get_multi_bboxes: all contours (outer borders at any nesting level and hole
    borders) above the area ratio, by decreasing area -- integer boxes, counts and normalised boxes bit-exact."""
    from spe_b200 import pseudo_labels as PL
    dev = torch.device("cuda")
    n = multi = 0
    for case in _gold()["cases"]:
        pairs = torch.nonzero(case["img_label"] > 0)
        for ratio, want in case["multi"].items():
            boxes, counts, xyxy = PL.cam_boxes_multi(case["cams"].to(dev), pairs, case["image_size"], case["cam_thr"], ratio, max_boxes=8, return_xyxy=True)
            assert counts.cpu().tolist() == [len(x) for x in want["xyxy"]], (case["image_size"], ratio, counts.cpu().tolist(), [len(x) for x in want["xyxy"]])
            for k, w in enumerate(want["xyxy"]):
                assert torch.equal(xyxy[k, :len(w)].cpu(), w.to(torch.int32)), (case["image_size"], ratio, k, xyxy[k].cpu().tolist(), w.tolist())
                multi += len(w) > 1
                n += 1
            # the reference-shaped call
            H, W = case["image_size"]
            got = PL.get_pseudo_label_multi_boxes({"cams_cls": case["cams"].to(dev)}, SimpleNamespace(tensors=torch.zeros(case["cams"].shape[0], 3, H, W, device=dev)),
                                                  [{"img_label": case["img_label"][b]} for b in range(case["cams"].shape[0])],
                                                  SimpleNamespace(cam_thr=case["cam_thr"], multi_box_ratio=ratio, num_classes=case["cams"].shape[1]))
            for g, wb, wl in zip(got, want["boxes"], want["labels"]):
                assert torch.equal(g["boxes"].cpu(), wb) and torch.equal(g["labels"].cpu(), wl)
    assert n >= 100 and multi >= 20, (n, multi)


def test_multi_boxes_rough_maps_vs_cv2_oracle():
    """rough maps (many components, holes, islands in holes): the contour SET above the ratio equals cv2's; order is checked where the
    areas are distinct."""
    pytest.importorskip("cv2")
    import numpy as np
    import cv2
    from oracle import cam_boxes as OC
    from spe_b200 import pseudo_labels as PL
    dev = torch.device("cuda")
    rng = np.random.default_rng(321)
    nbox = 0
    for (h, w, H, W, sig) in [(40, 40, 320, 320, 0.7), (24, 24, 96, 96, 0.5), (30, 20, 128, 192, 0.6), (16, 16, 64, 64, 0.4)]:
        B, C = 2, 8
        cams = torch.from_numpy(np.stack([[np.abs(cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), sig)) for _ in range(C)] for _ in range(B)]))
        lab = torch.ones(B, C)
        for thr, ratio in ((0.2, 0.5), (0.3, 0.05), (0.1, 0.0)):
            _, raw, areas = OC.pseudo_labels_multi(cams, lab, (H, W), cam_thr=thr, area_ratio=ratio)
            K = 64
            boxes, counts, xyxy = PL.cam_boxes_multi(cams.to(dev), torch.nonzero(lab > 0), (H, W), thr, ratio, max_boxes=K, return_xyxy=True)
            for k, (w_, a_) in enumerate(zip(raw, areas)):
                if len(w_) > K:
                    continue
                got = xyxy[k, :int(counts[k])].cpu()
                assert int(counts[k]) == len(w_), (h, w, thr, ratio, k, int(counts[k]), len(w_))
                if len(set(a_)) == len(a_):
                    assert torch.equal(got, w_.to(torch.int32)), (h, w, thr, ratio, k)
                else:
                    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, w_.tolist())), (h, w, thr, ratio, k)
                nbox += len(w_)
    assert nbox > 300
